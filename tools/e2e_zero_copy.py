"""Experiment: b200mm_mm_host at 4096^3 with C staged in HBM + D2H copies (default) vs epilogue stores straight into the pinned
host buffer (B200MM_HOST_ZEROCOPY_C=1).  Prints ms per call for both and whether the two results are bit-identical."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
M = N = K = 4096
sets = bench.make_sets(ctx, M, N, K, 2, 100)
kern = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K)
hs = [C.c_void_p() for _ in range(4)]
for h in hs:
    w._lib.check(w.lib().b200mm_host_alloc(M * N * 4, C.byref(h)))
npA, npB, npC0, npC1 = (np.ctypeslib.as_array((C.c_float * (M * N)).from_address(h.value)) for h in hs)
sets[0][0].read_into(npA); sets[0][1].read_into(npB)
dA, dB, dC = sets[1]
res = {}
for mode, out in (("0", npC0), ("1", npC1), ("0", npC0), ("1", npC1)):
    os.environ["B200MM_HOST_ZEROCOPY_C"] = mode
    out[:] = -1.0
    for _ in range(2):
        ctx.mm_host(kern, npA, npB, out, dA, dB, dC)
    t0 = time.perf_counter()
    for _ in range(6):
        ctx.mm_host(kern, npA, npB, out, dA, dB, dC)
    ms = (time.perf_counter() - t0) / 6 * 1e3
    print(f"zero_copy_c={mode}: {ms:.3f} ms per call = {2.0*M*N*K/ms/1e9:.1f} TFLOP/s end to end", flush=True)
print("bit-identical:", bool(np.array_equal(npC0, npC1)), " any unwritten:", bool((npC1 == -1.0).any()))
