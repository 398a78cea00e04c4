import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np
import wgpu_mm_b200 as w
ctx = w.Context(0)
for quant, K, N in ((True, 4096, 1792), (True, 4096, 3584), (True, 4096, 7168), (True, 2048, 2048), (True, 8192, 1024),
                    (False, 4096, 2048), (False, 4096, 1024), (False, 2048, 512), (False, 4096, 4096)):
    eb = 1 if quant else 4
    nsets = max(2, min(32, (200 << 20) // (K * N * eb)))
    Ws = []
    for i in range(nsets):
        b = ctx.buffer(K * N * eb); b.fill_weights(10 + i, K * N * eb // 4); Ws.append(b)
    x = ctx.buffer(K * 4); x.fill_weights(1, K); y = ctx.buffer(N * 4)
    for flags in (0, int(w.Flags.AUTOTUNE)):
        k = ctx.kernel(w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32, 1, N, K, w.KernelParams(absmax=2.0, batch=1, flags=flags))
        for i in range(40): ctx.launch(k, x, Ws[i % nsets], y)
        ctx.sync(); best = 1e9
        for r in range(3):
            ctx.timer_begin()
            for i in range(400): ctx.launch(k, x, Ws[i % nsets], y)
            best = min(best, ctx.timer_end() / 400)
        print(f"{'sint8' if quant else 'fp32 '} K={K} N={N} {'autotuned' if flags else 'default  '} grid={k.geometry()[0]} block={k.geometry()[1][0]}: {best*1e3:6.2f} us", flush=True)
        k.free()
    for b in Ws + [x, y]: b.free()
