import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np
import wgpu_mm_b200 as w
ctx = w.Context(0)
M = N = K = 4096
kern = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
hs = []
arrs = []
for n in (M * K, K * N, M * N):
    h = C.c_void_p(); w._lib.check(w.lib().b200mm_host_alloc(n * 4, C.byref(h))); hs.append(h)
    arrs.append(np.ctypeslib.as_array((C.c_float * n).from_address(h.value)))
arrs[0][:] = 0.01; arrs[1][:] = 0.02
dA, dB, dC = ctx.buffer(M * K * 4), ctx.buffer(K * N * 4), ctx.buffer(M * N * 4)
for i in range(6):
    t = time.perf_counter(); ctx.mm_host(kern, arrs[0], arrs[1], arrs[2], dA, dB, dC); print(f"call {i}: {(time.perf_counter() - t) * 1e3:.3f} ms", file=sys.stderr)
