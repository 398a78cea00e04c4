"""Timeline of back-to-back GEMV launches from per-CTA globaltimer stamps (b200mm_debug_gemv_trace): where the fixed
per-launch cost of the microsecond-scale kernels goes (CTA start, first loads, PDL wait, x staged, streaming done, cluster
barrier, exit), relative to the end of the previous launch.  python tools/trace_gemv.py [f32|s8] [K N]"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402

quant = (sys.argv[1] if len(sys.argv) > 1 else "s8") == "s8"
K, N = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else ((4096, 14336) if quant else (4096, 16384))
ctx = w.Context(0)
nsets = max(2, (400 << 20) // (K * N * (1 if quant else 4)))
Ws = []
for i in range(nsets):
    b = ctx.buffer(K * N * (1 if quant else 4)); b.fill_weights(10 + i, K * N // (4 if quant else 1)); Ws.append(b)
x = ctx.buffer(K * 4); x.fill_weights(1, K)
y = ctx.buffer(N * 4)
kern = ctx.kernel(w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32, 1, N, K, w.KernelParams(absmax=2.0, batch=1, tune=(int(os.environ.get("VARIANT", "0")), int(os.environ.get("SPLITS", "0")), 0, 0)))
(g, blk) = kern.geometry()
ctas = g[0] * g[1] * g[2]
slots = 8
tb = ctx.buffer(slots * ctas * 8 * 8)
tb.write(np.zeros(slots * ctas * 8, dtype=np.uint64))
lib = w.lib()
for i in range(50):
    ctx.launch(kern, x, Ws[i % nsets], y)
ctx.sync()
lib.b200mm_debug_gemv_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
lib.b200mm_debug_gemv_trace(kern.handle, C.c_void_p(tb.ptr), slots)
ctx.timer_begin()
for i in range(slots):
    ctx.launch(kern, x, Ws[i % nsets], y)
ms = ctx.timer_end()
lib.b200mm_debug_gemv_trace(kern.handle, None, 0)
t = tb.read(np.uint64).reshape(slots, ctas, 8).astype(np.int64)
print(f"{'sint8' if quant else 'fp32'} K={K} N={N} grid={g} block={blk[0]}: {ms / slots * 1e3:.2f} us per launch (event pair around {slots} launches)")
names = ["cta start", "loads issued", "pdl wait done", "x staged", "stream done", "partial ready", "y stored", "exit"]
for s in range(2, slots):
    prev_end = t[s - 1, :, 7].max()
    base = prev_end
    row = []
    for j in range(8):
        col = t[s, :, j]
        col = col[col > 0]
        if len(col) == 0:
            row.append(f"{names[j]}: -")
            continue
        row.append(f"{names[j]}: {np.median(col - base) / 1e3:+.2f} [{(col.min() - base) / 1e3:+.2f},{(col.max() - base) / 1e3:+.2f}]")
    print(f"launch {s} (us after the previous launch's last exit; median [min,max] over CTAs):\n   " + "\n   ".join(row))
    print(f"   launch-to-launch period (last exit to last exit): {(t[s, :, 7].max() - prev_end) / 1e3:.2f} us")
