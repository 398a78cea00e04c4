"""Times the SIMT SGEMM at 4096^3 with and without the split-K remainder schedule."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgpu_mm_b200 as w

ctx = w.Context(0)
M = N = K = int(os.environ.get("SIZE", "4096"))
bufs = []
for s in range(3):
    a = ctx.buffer(M * K * 4); a.fill_weights(1 + s, M * K)
    b = ctx.buffer(K * N * 4); b.fill_weights(11 + s, K * N)
    c = ctx.buffer(M * N * 4)
    bufs.append((a, b, c))
SEQ = int(w.Flags.SEQUENTIAL_K)
for name, flags, tune in (("default (split-K remainder, 2 launches)", 0, (0, 0)), ("sequential-K (1 launch)", SEQ, (0, 0)),
                          ("default, group 8", 0, (0, 8)), ("default, group 32", 0, (0, 32)), ("sequential-K, m-fastest", SEQ, (0, 4096))):
    kern = ctx.kernel(w.KernelId.SGEMM_SIMT, M, N, K, w.KernelParams(flags=flags, tune=tune + (0, 0)))
    for i in range(3):
        ctx.launch(kern, *bufs[i % 3])
    ctx.sync()
    ctx.timer_begin()
    n = 10
    for i in range(n):
        ctx.launch(kern, *bufs[i % 3])
    ms = ctx.timer_end() / n
    print(f"{name}: grid {kern.geometry()[0]} {ms:.4f} ms  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s", flush=True)
    kern.free()
ctx.close()
