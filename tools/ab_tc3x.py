"""A/B timing of sgemm_tc3x variants, interleaved round-robin so that clock / power drift hits all variants alike.
python tools/ab_tc3x.py MxNxK [rounds]   variants: 1-CTA (tune 513), 2-CTA BK=16 (512), 2-CTA BK=32 (512, tune[2]=32), and where the lo
operands come from (tune[3]: 5 = computed in shared memory, 1 / 4 = only B, 2 / 3 = pre-pass, 0 = by shape); VARIANTS=a,b,... selects a subset."""
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
# (tune[0], tune[2], tune[3]); default epilogue = TMA store, stg = st.global epilogue; split2 = A_lo and B_lo computed in shared memory,
# split1 = B_lo only (A by row bands), no suffix = lo operands from the split_lo pre-pass (A by row bands; _r1: everything in the pre-pass)
variants = {"1cta": (513, 0, 2), "2cta": (512, 0, 2), "2cta_split1": (512, 0, 4), "2cta_split2": (512, 0, 5), "2cta_r1": (512, 0, 3),
            "1cta_split1": (513, 0, 4), "1cta_split2": (513, 0, 5), "2cta_stg": (512, 6, 2), "2cta_bk32": (512, 32, 2)}
if os.environ.get("VARIANTS"):
    variants = {n: variants[n] for n in os.environ["VARIANTS"].split(",")}
kerns = {n: ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=(t0, 0, t2, t3))) for n, (t0, t2, t3) in variants.items()}
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
    print(f"{M}x{N}x{K} {n:13s}: median {np.median(v) * 1e3:9.1f} us ({flop / np.median(v) / 1e9:6.1f} TFLOP/s)  best {v.min() * 1e3:9.1f} us ({flop / v.min() / 1e9:6.1f})  rounds {rounds} x {iters} launches")
