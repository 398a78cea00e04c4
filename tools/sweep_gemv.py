"""GPU sweep of the streaming GEMV variants / split counts (kernel-only GB/s, weight sets rotated > L2)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgpu_mm_b200 as w  # noqa: E402

ctx = w.Context(0)


def run(kid, K, N, nsets, variant, splits, quant, iters=40):
    a = ctx.buffer(K * 4); a.fill_weights(1, K)
    bs = []
    for s in range(nsets):
        b = ctx.buffer(K * N if quant else K * N * 4)
        b.fill_weights(10 + s, (K * N // 4) if quant else K * N)  # random bits are fine for bandwidth
        bs.append(b)
    c = ctx.buffer(N * 4)
    kern = ctx.kernel(kid, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(variant, splits, 0, 0)))
    for i in range(8):
        ctx.launch(kern, a, bs[i % nsets], c)
    ctx.sync()
    ctx.timer_begin()  # plain back-to-back launches: no per-launch events in between
    for i in range(iters * 4):
        ctx.launch(kern, a, bs[i % nsets], c)
    tot = ctx.timer_end() / 4
    kern.profile(True)
    for i in range(iters):
        ctx.launch(kern, a, bs[i % nsets], c)
    per = kern.profile_read(256)
    g, blk = kern.geometry()
    kern.free()
    for b in bs + [a, c]:
        b.free()
    bytes_ = (K * N if quant else 4 * K * N) + 4 * K + 4 * N
    ms = float(np.median(per))
    b2b = tot / iters
    return bytes_ / b2b / 1e6, ms * 1e3, b2b * 1e3, g, blk


CASES = (("gemv_f32 4096x16384", w.KernelId.GEMV_F32, 4096, 16384, 4, False),
         ("qgemv_s8 4096x14336", w.KernelId.QGEMV_SINT8, 4096, 14336, 8, True))
if os.environ.get("ONLY_Q"):
    CASES = CASES[1:]
for name, kid, K, N, nsets, quant in CASES:
    print(name, flush=True)
    variants = list(range(int(os.environ.get("NVARIANTS", "4"))))
    if os.environ.get("VARIANTS"):
        variants = [int(v) for v in os.environ["VARIANTS"].split(",")]
    for variant in variants:
        for splits in (0, 2, 3, 4, 5, 6, 8):
            try:
                gbps, us, b2b, g, blk = run(kid, K, N, nsets, variant, splits, quant)
            except w.B200mmError as e:
                print(f"  variant {variant} splits {splits}: {e}")
                continue
            print(f"  variant {variant} splits {splits:2d} grid {g} block {blk[0]}: {gbps:7.0f} GB/s (back-to-back {b2b:6.2f} us; event-pair kernel {us:6.1f} us)", flush=True)
ctx.close()
