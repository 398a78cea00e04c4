"""sgemm_tc3x at the three shapes whose default instantiation differs, for ncu (2 launches each):
  4096^3 (pair kernel, 256 x 256 tiles; TC3X_TUNE3=3 keeps the launch non-cooperative for ncu's replay), 1024^3 (128 x 128 tiles),
  128 x 4096 x 4096 (128 x 128 tiles, B_lo computed in shared memory).
  ncu --set full --clock-control none --import-source on -k regex:'sgemm_tc3x_kernel|split_lo_kernel' -o gpurun_out/r2_tc_shapes python tools/ncu_tc_shapes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402
import bench  # noqa: E402

ctx = w.Context(0)
t3 = int(os.environ.get("TC3X_TUNE3", "3"))
for (M, N, K), tune in (((4096, 4096, 4096), (0, 0, 0, t3)), ((1024, 1024, 1024), (0, 0, 0, 0)), ((128, 4096, 4096), (0, 0, 0, 0))):
    sets = bench.make_sets(ctx, M, N, K, 2, 100)
    k = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=tune))
    for a, b, c in sets:
        ctx.launch(k, a, b, c)
    ctx.sync()
    print((M, N, K), k.geometry(), flush=True)
    k.free()
    bench.free_sets(sets)
ctx.close()
