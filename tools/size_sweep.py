"""sgemm_tc3x across square sizes: default schedule vs forced 1-CTA / 2-CTA, and cuBLAS TF32 (one MMA per product, i.e. a third of the
tensor work, and not FP32-accurate) / cuBLAS fp32 on the same box for scale.  Per-step time incl. the split pre-pass."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import wgpu_mm_b200 as w

ctx = w.Context(0)
def t_ours(S, tune0):
    a = ctx.buffer(S * S * 4); a.fill_weights(1, S * S); b = ctx.buffer(S * S * 4); b.fill_weights(2, S * S); c = ctx.buffer(S * S * 4)
    k = ctx.kernel(w.KernelId.SGEMM_TC3X, S, S, S, w.KernelParams(tune=(tune0, 0, 0, 0)))
    n = max(3, int(2e-2 / (2.0 * S ** 3 / 200e12)))
    for _ in range(3): ctx.launch(k, a, b, c)
    ctx.sync(); best = 1e9
    for r in range(3):
        ctx.timer_begin()
        for _ in range(n): ctx.launch(k, a, b, c)
        best = min(best, ctx.timer_end() / n)
    g = k.geometry()[0][0]
    k.free(); [x.free() for x in (a, b, c)]
    return 2.0 * S ** 3 / best / 1e9, g
def t_torch(S, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    x = torch.randn(S, S, device="cuda"); y = torch.randn(S, S, device="cuda")
    n = max(3, int(2e-2 / (2.0 * S ** 3 / 200e12)))
    for _ in range(3): torch.matmul(x, y)
    torch.cuda.synchronize(); best = 1e9
    for r in range(3):
        e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        for _ in range(n): torch.matmul(x, y)
        e1.record(); e1.synchronize(); best = min(best, e0.elapsed_time(e1) / n)
    return 2.0 * S ** 3 / best / 1e9
print(f"{'size':>6} {'default':>14} {'1-CTA':>8} {'2-CTA':>8} {'cuBLAS tf32':>12} {'cuBLAS fp32':>12}   (TFLOP/s, FP32-equivalent)")
for S in (512, 1024, 1536, 2048, 3072, 4096, 5120, 6144, 8192):
    d, g = t_ours(S, 0); s1, _ = t_ours(S, 513); s2, _ = t_ours(S, 512)
    print(f"{S:6d} {d:8.1f} (g{g:3d}) {s1:8.1f} {s2:8.1f} {t_torch(S, True):12.1f} {t_torch(S, False):12.1f}", flush=True)
