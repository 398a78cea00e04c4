"""Sweeps the sint8 GEMV (with / without per-group scales) at cfg4 (1x4096 . 4096x14336): variant x split count,
back-to-back PDL launches over 8 rotated weight sets.  Usage: bench_grouped.py [group_k ...]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
K, N = 4096, 14336
gks = [int(a) for a in sys.argv[1:]] or [0]
variants = [int(v) for v in os.environ.get("VARIANTS", "0,4,11,12,1,6,7").split(",")]
for gk in gks:
    sets = bench.make_sets(ctx, 1, N, K, 8, 600, quant=True, group_k=gk)
    for variant in variants:
        best = None
        for splits in (0, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16):
            try:
                k = ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(absmax=2.0, batch=1, group_k=gk, tune=(variant, splits, 0, 0)))
            except Exception as e:
                print(gk, variant, splits, "ERR", e); continue
            ms = min(bench.time_back_to_back(ctx, k, sets, 400, 40) for _ in range(2))
            b = K * N + 4 * K + 4 * N + (4 * (K // gk) * N if gk else 0)
            print(f"group_k={gk:5d} variant={variant:2d} splits={splits:2d}  {ms*1e3:7.2f} us  {b/ms/1e6:7.0f} GB/s  geometry={k.geometry()}", flush=True)
            if splits and (best is None or ms < best[0]): best = (ms, splits)
            k.free()
        print(f"  -> variant {variant}: best {best[0]*1e3:.2f} us at splits={best[1]}", flush=True)
    bench.free_sets(sets)
