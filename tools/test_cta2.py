"""Bring-up check of the 2-CTA (cta_group::2) sgemm_tc3x kernel against the 1-CTA kernel and FP64 (tune[0] = 512 / 513)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402

ctx = w.Context(0)
shapes = [(256, 256, 256), (256, 512, 1024), (512, 256, 64), (1024, 1024, 1024), (300, 520, 260), (4096, 4096, 4096), (2304, 4096, 512)]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    a = ctx.buffer(M * K * 4); a.fill_weights(1, M * K)
    b = ctx.buffer(K * N * 4); b.fill_weights(2, K * N)
    outs, times = [], []
    for tune0, bk in ((513, 0), (512, 6 if os.environ.get("STG") else 0)) + (((512, 32),) if os.environ.get("BK32") else ()):
        c = ctx.buffer_from(np.full(M * N, 7.0, dtype=np.float32))
        k = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=(tune0, 0, bk, 0)))
        ctx.launch(k, a, b, c)
        outs.append(c.read(np.float32).reshape(M, N))
        for _ in range(3):
            ctx.launch(k, a, b, c)
        ctx.sync(); ctx.timer_begin()
        for _ in range(5):
            ctx.launch(k, a, b, c)
        times.append(ctx.timer_end() / 5)
        print(f"   tune {tune0} bk {bk}: geometry {k.geometry()}  {times[-1] * 1e3:.1f} us", flush=True)
        k.free(); c.free()
    A = a.read(np.float32).reshape(M, K).astype(np.float64)
    B = b.read(np.float32).reshape(K, N).astype(np.float64)
    rows = sorted({0, 1, 127, 128, 129, 255, M // 2, M - 1})
    ref = A[rows] @ B
    e1 = np.abs(outs[0][rows] - ref).max() / np.abs(ref).max()
    e2 = np.abs(outs[1][rows] - ref).max() / np.abs(ref).max()
    d = np.abs(outs[0] - outs[1]).max()
    print(f"{M}x{N}x{K}: 1-CTA {times[0] * 1e3:8.1f} us  2-CTA {times[1] * 1e3:8.1f} us  rel_f64 {e1:.2e} / {e2:.2e}  max|1cta-2cta| {d:.3e}  "
          f"unwritten {int((outs[1] == 7.0).sum())}  {'OK' if e2 < 5e-6 else 'FAIL'}", flush=True)
    a.free(); b.free()
