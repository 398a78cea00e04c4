import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import wgpu_mm_b200 as w
ctx = w.Context(0)
def run(K, N, splits, variant=0, panels=0):
    nsets = max(2, (500 << 20) // (K * N))
    Ws = []
    for i in range(nsets):
        b = ctx.buffer(K * N); b.fill_weights(10 + i, K * N // 4); Ws.append(b)
    x = ctx.buffer(K * 4); x.fill_weights(1, K); y = ctx.buffer(N * 4)
    k = ctx.kernel(w.KernelId.QGEMV_SINT8, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(variant, splits, 0, panels)))
    for i in range(40): ctx.launch(k, x, Ws[i % nsets], y)
    ctx.sync(); best = 1e9
    for r in range(3):
        ctx.timer_begin()
        for i in range(400): ctx.launch(k, x, Ws[i % nsets], y)
        best = min(best, ctx.timer_end() / 400)
    g = k.geometry()[0]
    by = K * N + 4 * K + 4 * N
    print(f"K={K} N={N} grid={g} ctas={g[0]*g[1]}: {best*1e3:6.2f} us  {by/best/1e6:6.0f} GB/s  frac {by/best/1e6/6550:.3f}   fixed = {best*1e3 - by/7.0e6:.2f} us over bytes/7TB/s", flush=True)
    k.free()
    for b in Ws + [x, y]: b.free()
for K, N, sp in ((4096, 14336, 0), (4096, 18944, 4), (4096, 9472, 8), (4096, 37888, 2), (4096, 28672, 2), (4096, 14336, 8), (8192, 9472, 8)):
    run(K, N, sp)
print("balanced ragged panels (variant 21 = 128-column panels, tune[3] = panel count):")
for sp, panels in ((2, 148), (1, 296), (2, 112), (2, 144), (2, 152), (1, 224), (4, 112)):
    run(4096, 14336, sp, 21, panels)
