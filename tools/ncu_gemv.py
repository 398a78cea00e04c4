"""Launches each streaming-GEMV instantiation a few times (default geometry, no autotune) for an ncu capture:
  ncu --set full --clock-control none --import-source on -k regex:gemv_stream_kernel -c 9 -o gpurun_out/gemv python tools/ncu_gemv.py
Launch order: fp32 cfg3 x3, sint8 cfg4 x3, sint8 cfg4 with per-group scales (group_k = 128) x3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
for kid, K, N, quant, gk in ((w.KernelId.GEMV_F32, 4096, 16384, False, 0), (w.KernelId.QGEMV_SINT8, 4096, 14336, True, 0),
                             (w.KernelId.QGEMV_SINT8, 4096, 14336, True, 128)):
    sets = bench.make_sets(ctx, 1, N, K, 3, 900, quant=quant, group_k=gk)
    k = ctx.kernel(kid, 1, N, K, w.KernelParams(absmax=2.0, batch=1, group_k=gk))
    for a, b, c in sets:
        ctx.launch(k, a, b, c)
    ctx.sync()
    print(kid.name, gk, k.geometry())
    k.free(); bench.free_sets(sets)
