"""SGEMM_TC3X at small M (the row panels of the pipelined host-buffer path, and skinny GEMMs in general):
step time (split_lo + GEMM, back to back) and GEMM-only time per launch, N = K = 4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
N = K = 4096
for M in (128, 256, 512, 1024, 2048, 4096):
    sets = bench.make_sets(ctx, M, N, K, 3, 100)
    for name, tune in (("default", (0, 0, 0, 0)), ("pure stream-K", (0, 1, 0, 0)), ("BN=128", (128, 0, 0, 0)), ("BK=32", (0, 0, 32, 0))):
        try:
            k = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=tune))
        except Exception as e:
            print(M, name, "ERR", e); continue
        tot, per = bench.time_kernel_steps(ctx, k, sets, 20, 5)
        flop = 2.0 * M * N * K
        print(f"M={M:5d} {name:14s} step {tot/20*1e3:8.1f} us ({flop/(tot/20*1e-3)/1e12:6.1f} TFLOP/s)   gemm kernel {np.mean(per)*1e3:8.1f} us ({flop/(np.mean(per)*1e-3)/1e12:6.1f} TFLOP/s)  grid {k.geometry()[0]}", flush=True)
        k.free()
    bench.free_sets(sets)
