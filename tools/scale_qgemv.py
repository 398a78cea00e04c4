"""How does the sint8 / fp32 GEMV time scale with K (fixed overhead vs streaming rate)?"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import wgpu_mm_b200 as w
ctx = w.Context(0)
N = 14336
for quant, pdl_off in ((True, 0), (True, 1), (False, 0), (False, 1)):
    for K in (1024, 4096, 16384):
        nbytes = K * N if quant else 4 * K * N
        nsets = max(2, int(600e6 // nbytes) + 1)
        a = ctx.buffer(K * 4); a.fill_weights(1, K)
        bs = []
        for s in range(nsets):
            b = ctx.buffer(nbytes); b.fill_weights(10 + s, nbytes // 4); bs.append(b)
        c = ctx.buffer(N * 4)
        kern = ctx.kernel(w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(0, 0, pdl_off, 0)))
        for i in range(10):
            ctx.launch(kern, a, bs[i % nsets], c)
        ctx.sync()
        iters = 200
        ctx.timer_begin()
        for i in range(iters):
            ctx.launch(kern, a, bs[i % nsets], c)
        us = ctx.timer_end() / iters * 1e3
        print(f"quant={quant} pdl={"off" if pdl_off else "on"} K={K:6d} bytes={nbytes / 1e6:8.1f} MB grid={kern.geometry()[0]} {us:8.2f} us  {nbytes / us / 1e3:7.0f} GB/s", flush=True)
        kern.free()
        for b in bs + [a, c]:
            b.free()
ctx.close()
