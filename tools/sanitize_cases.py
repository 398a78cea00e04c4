"""Small launches of every round-2 code path for compute-sanitizer (memcheck / racecheck / synccheck):
   compute-sanitizer --tool memcheck python tools/sanitize_cases.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import wgpu_mm_b200 as w  # noqa: E402
from wgpu_mm_b200.quant import sint8_quantize_grouped  # noqa: E402

ctx = w.Context(0)
rng = np.random.default_rng(0)


def gemm(M, N, K, tune, kid=w.KernelId.SGEMM_TC3X):
    A = (rng.random((M, K), dtype=np.float32) - 0.5) * 0.4
    B = (rng.random((K, N), dtype=np.float32) - 0.5) * 0.4
    dA, dB, dC = ctx.buffer_from(A), ctx.buffer_from(B), ctx.buffer_from(np.full(M * N, 7.0, dtype=np.float32))
    k = ctx.kernel(kid, M, N, K, w.KernelParams(tune=tune))
    ctx.launch(k, dA, dB, dC)
    got = dC.read(np.float32).reshape(M, N)
    err = np.abs(got - A.astype(np.float64) @ B.astype(np.float64)).max()
    print(f"gemm {kid.name} {M}x{N}x{K} tune={tune}: max abs err {err:.2e}", flush=True)
    assert err < 1e-4
    k.free()
    for b in (dA, dB, dC):
        b.free()


def gemv(K, N, quant, tune=(0, 0, 0, 0), group_k=0, M=1):
    x = (rng.random((M, K), dtype=np.float32) - 0.5) * 0.4
    if quant:
        Wf = (rng.random((K, N), dtype=np.float32) - 0.5) * 0.4
        B = sint8_quantize_grouped(Wf, K, N, group_k) if group_k else w.quant.sint8_quantize(Wf, K, N)[0]
    else:
        B = (rng.random((K, N), dtype=np.float32) - 0.5) * 0.4
    dx, dB, dy = ctx.buffer_from(x), ctx.buffer_from(B), ctx.buffer_from(np.full(M * N, 7.0, dtype=np.float32))
    k = ctx.kernel(w.KernelId.QGEMV_SINT8 if quant else w.KernelId.GEMV_F32, M, N, K, w.KernelParams(absmax=0.4 if not group_k else 0.0, batch=1, tune=tune, group_k=group_k))
    for _ in range(2):
        ctx.launch(k, dx, dB, dy)
    got = dy.read(np.float32)
    assert not (got == 7.0).any()
    print(f"gemv {'s8' if quant else 'f32'} {M}x{K}x{N} tune={tune} group_k={group_k}: ok", flush=True)
    k.free()
    for b in (dx, dB, dy):
        b.free()


gemm(256, 256, 256, (512, 0, 0, 0))          # pair kernel, one tile
gemm(300, 520, 260, (512, 0, 0, 0))          # pair kernel, ragged, second CTA of a pair partly out of range
gemm(512, 768, 1024, (512, 0, 0, 0))         # pair kernel, k-split schedule
gemm(4224, 256, 512, (513, 0, 0, 0))         # 1-CTA kernel, 3 row bands: in-kernel A split + cooperative launch
gemm(4352, 512, 256, (512, 0, 0, 0))         # pair kernel with the in-kernel A split
gemm(512, 768, 1024, (512, 0, 0, 4))         # pair kernel, only B_lo computed in shared memory (A by row bands)
gemm(4352, 512, 256, (513, 0, 0, 4))         # 1-CTA kernel, the same
gemm(512, 768, 1024, (512, 0, 0, 5))         # pair kernel, A_lo and B_lo computed in shared memory (no pre-pass)
gemm(4224, 256, 512, (513, 0, 0, 5))         # 1-CTA kernel, the same
gemm(128, 4096, 1024, (0, 0, 0, 0))          # skinny M: the shape rule picks B_lo in shared memory
gemm(130, 66, 34, (0, 0, 0, 0))              # padded staging path (N % 4, K % 4 != 0)
gemm(256, 256, 64, (0, 0, 0, 0), w.KernelId.SGEMM_SIMT)  # split-K with the inverted ownership
gemv(1024, 1024, False)
gemv(1024, 1024, True)
gemv(1000, 1040, True, (13, 2, 0, 17))        # balanced ragged panels
gemv(512, 4096, True, (21, 4, 0, 74))
gemv(1024, 512, True, group_k=64)
gemv(1024, 512, True, group_k=32)
gemv(1024, 512, True, group_k=128, M=3)
gemv(512, 1024, False, M=13)
gemv(512, 1024, True, M=6)
print("sanitize_cases: done")
