"""One SGEMM_TC3X launch each at M = 256 and M = 1024 (N = K = 4096) for an ncu capture of the small-M regime:
  ncu --set full --clock-control none --import-source on -k regex:sgemm_tc3x_kernel -c 4 -o gpurun_out/skinny python tools/ncu_skinny.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
N = K = 4096
for M in (256, 1024):
    sets = bench.make_sets(ctx, M, N, K, 2, 100)
    k = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
    for a, b, c in sets:
        ctx.launch(k, a, b, c)
    ctx.sync()
    k.free(); bench.free_sets(sets)
