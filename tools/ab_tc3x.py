"""A/B timing of sgemm_tc3x variants, interleaved round-robin so that clock / power drift hits all variants alike.
python tools/ab_tc3x.py MxNxK [rounds]   variants: 1-CTA (tune 513), 2-CTA BK=16 (512), 2-CTA BK=32 (512, tune[2]=32)"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402

ctx = w.Context(0)
M, N, K = (int(v) for v in sys.argv[1].split("x"))
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nsets = 3 if M * K + K * N + M * N <= 3 * 4096 * 4096 else 1
sets = []
for i in range(nsets):
    a = ctx.buffer(M * K * 4); a.fill_weights(1 + 10 * i, M * K)
    b = ctx.buffer(K * N * 4); b.fill_weights(2 + 10 * i, K * N)
    c = ctx.buffer(M * N * 4)
    sets.append((a, b, c))
variants = {"1cta": (513, 0), "2cta_tmast": (512, 0), "2cta_stg": (512, 6), "2cta_bk32": (512, 32)}  # tmast = TMA-store epilogue (default), stg = st.global epilogue
kerns = {n: ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=(t0, 0, t2, 0))) for n, (t0, t2) in variants.items()}
iters = max(3, int(os.environ.get("ITERS", "0")) or int(2e-2 / (2.0 * M * N * K / 250e12)) or 3)
res = {n: [] for n in variants}
for n, k in kerns.items():
    for i in range(3):
        ctx.launch(k, *sets[i % nsets])
ctx.sync()
names = list(kerns)
for r in range(rounds):
    for n in names[r % len(names):] + names[:r % len(names)]:  # rotate the starting variant: no variant always runs "first" or "hottest"
        k = kerns[n]
        ctx.timer_begin()
        for i in range(iters):
            ctx.launch(k, *sets[i % nsets])
        res[n].append(ctx.timer_end() / iters)
flop = 2.0 * M * N * K
for n, v in res.items():
    v = np.array(v)
    print(f"{M}x{N}x{K} {n:10s}: median {np.median(v) * 1e3:9.1f} us ({flop / np.median(v) / 1e9:6.1f} TFLOP/s)  best {v.min() * 1e3:9.1f} us ({flop / v.min() / 1e9:6.1f})  rounds {rounds} x {iters} launches")
