"""Where does computing the lo tiles in shared memory (Tc3xCfg::SPLIT) pay?  Per-step time (pre-pass included) of the three forms,
interleaved: split2 = A_lo and B_lo in the kernel (tune[3] = 5), split1 = B_lo only (4), pre = split_lo pre-pass (2).
profiles/r2_split_shapes.log; the default rule in setup_tc3x follows it."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
shapes = [(512, 512, 512), (1024, 1024, 1024), (1536, 1536, 1536), (2048, 2048, 2048), (3072, 3072, 3072), (4096, 4096, 4096), (8192, 8192, 8192),
          (16, 4096, 4096), (64, 4096, 4096), (128, 4096, 4096), (256, 4096, 4096), (512, 4096, 4096), (1024, 4096, 4096), (2048, 4096, 4096),
          (128, 14336, 4096), (256, 16384, 4096), (4096, 1024, 4096), (16384, 2048, 4096)]
for (M, N, K) in shapes:
    nsets = 3 if (M * K + K * N + M * N) * 4 < 400e6 else 1
    sets = bench.make_sets(ctx, M, N, K, nsets, 100)
    kerns = {n: ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=(0, 0, 0, t3))) for n, t3 in (("split2", 5), ("split1", 4), ("pre", 2))}
    iters = max(5, min(50, int(3e-3 / (2.0 * M * N * K / 200e12 + 10e-6))))
    res = {n: [] for n in kerns}
    for r in range(4):
        for n, k in kerns.items():
            for i in range(2): ctx.launch(k, *sets[i % nsets])
            ctx.timer_begin()
            for i in range(iters): ctx.launch(k, *sets[i % nsets])
            res[n].append(ctx.timer_end() / iters)
    flop = 2.0 * M * N * K
    print(f"{M:6d}x{N:6d}x{K:5d} grid {kerns['split2'].geometry()[0][0]:4d} " + "  ".join(f"{n} {np.median(v)*1e3:8.1f} us ({flop/np.median(v)/1e9:6.1f} TF)" for n, v in res.items()), flush=True)
    for k in kerns.values(): k.free()
    bench.free_sets(sets)
