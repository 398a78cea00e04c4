"""Launches every shipped hot-path kernel at its BASELINE shape a few times, for ncu (round-2 profiles):

  ncu --set full --clock-control none --import-source on -k regex:'sgemm_tc3x_kernel|split_lo_kernel|sgemm_simt_kernel|gemv_stream_kernel' \
      -o gpurun_out/r2_full python tools/ncu_all.py

Launch order (2 launches each, different operand sets; default geometry, i.e. what b200mm_kernel_get ships without autotune):
  split_lo_kernel + sgemm_tc3x_kernel 4096^3  |  sgemm_simt_kernel 4096^3  |  gemv_stream_kernel<GemvF32> cfg3
  |  gemv_stream_kernel<GemvS8> cfg4  |  gemv_stream_kernel<GemvS8, grouped> cfg4 (group_k = 128)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402
import bench  # noqa: E402

ctx = w.Context(0)
reps = int(os.environ.get("REPS", "2"))
M = N = K = 4096
sets = bench.make_sets(ctx, M, N, K, reps, 100)
# TC3X_TUNE = "t0,t1,t2,t3": e.g. "0,0,0,3" keeps the whole operand split in the pre-pass, so the pair kernel is launched without the
# cooperative attribute (ncu's kernel replay refused the cluster + cooperative launch: LaunchFailed after 2 passes); "513,0,0,0" = 1-CTA kernel
tc_tune = tuple(int(v) for v in os.environ.get("TC3X_TUNE", "0,0,0,0").split(","))
for kid in (w.KernelId.SGEMM_TC3X, w.KernelId.SGEMM_SIMT):
    k = ctx.kernel(kid, M, N, K, w.KernelParams(tune=tc_tune if kid == w.KernelId.SGEMM_TC3X else (0, 0, 0, 0)))
    for a, b, c in sets:
        ctx.launch(k, a, b, c)
    ctx.sync()
    print(kid.name, k.geometry(), flush=True)
    k.free()
bench.free_sets(sets)
for kid, Kv, Nv, quant, gk in ((w.KernelId.GEMV_F32, 4096, 16384, False, 0), (w.KernelId.QGEMV_SINT8, 4096, 14336, True, 0),
                               (w.KernelId.QGEMV_SINT8, 4096, 14336, True, 128)):
    gsets = bench.make_sets(ctx, 1, Nv, Kv, reps, 900, quant=quant, group_k=gk)
    k = ctx.kernel(kid, 1, Nv, Kv, w.KernelParams(absmax=0.0 if gk else 2.0, batch=1, group_k=gk))
    for a, b, c in gsets:
        ctx.launch(k, a, b, c)
    ctx.sync()
    print(kid.name, gk, k.geometry(), flush=True)
    k.free(); bench.free_sets(gsets)
ctx.close()
