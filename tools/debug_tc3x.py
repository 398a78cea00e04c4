"""GPU bring-up aid for the tcgen05 kernel: runs small shapes, prints error structure, saves arrays for offline analysis."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
import wgpu_mm_b200 as w  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
ctx = w.Context(0)
print(ctx.device_info(), flush=True)
save = {}
for (M, N, K, bn, flags) in [(128, 256, 8, 256, 1), (128, 256, 32, 256, 1), (128, 256, 32, 256, 0), (128, 128, 32, 128, 0), (128, 256, 64, 256, 0),
                             (256, 512, 128, 256, 0), (512, 512, 512, 128, 0), (4096, 4096, 4096, 256, 0)]:
    if K % 4:
        continue
    A = oracle.generate_weight_data(1, M, K)
    B = oracle.generate_weight_data(2, K, N)
    kern = ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=(bn, 0, 0, 0), flags=flags))
    dA, dB = ctx.buffer_from(A), ctx.buffer_from(B)
    dC = ctx.buffer_from(np.full(M * N, 123.25, dtype=np.float32))
    ctx.launch(kern, dA, dB, dC)
    got = dC.read(np.float32).reshape(M, N)
    if M <= 512:
        ref = oracle.mm_f64(A, B)
        err = np.abs(got - ref)
    else:
        rows = np.arange(0, M, 97)
        ref = oracle.mm_f64_rows(A, B, rows)
        err = np.abs(got[rows] - ref)
    print(f"M={M} N={N} K={K} bn={bn} flags={flags}: max|err|={err.max():.3e} max|ref|={np.abs(ref).max():.3e} "
          f"unwritten={(got == 123.25).sum()} nan={np.isnan(got).sum()}", flush=True)
    if err.max() > 1e-3 and M <= 256:
        # where is it wrong?  rows x 32-col blocks
        blk = err.reshape(err.shape[0] // 32, 32, err.shape[1] // 32, 32).max(axis=(1, 3))
        print("  err by 32x32 block:\n", np.array2string(blk, precision=2, max_line_width=200))
        save[f"got_{M}_{N}_{K}_{bn}_{flags}"] = got
        save[f"A_{M}_{N}_{K}"] = A
        save[f"B_{M}_{N}_{K}"] = B
    for b in (dA, dB, dC):
        b.free()
    kern.free()
if save:
    np.savez_compressed(os.path.join(out_dir, "tc3x_debug.npz"), **save)
ctx.close()
