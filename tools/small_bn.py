"""Small SGEMMs: 128 x 256 tiles (tune[0] = 513) against 128 x 128 tiles (tune[0] = 128: twice the tiles, so fewer k-slices per tile
to reach one CTA per SM and less fix-up traffic) and what the default rule picks, per step, interleaved.  profiles/r2_small_bn.log.
python tools/small_bn.py [MxNxK ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import wgpu_mm_b200 as w
import bench

ctx = w.Context(0)
shapes = [(n, n, n) for n in (256, 512, 768, 1024, 1280, 1536, 2048, 2560)] if len(sys.argv) < 2 else [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for (M, N, K) in shapes:
    sets = bench.make_sets(ctx, M, N, K, 3, 100)
    kerns = {name: ctx.kernel(w.KernelId.SGEMM_TC3X, M, N, K, w.KernelParams(tune=t)) for name, t in (("default", (0, 0, 0, 0)), ("bn256", (513, 0, 0, 0)), ("bn128", (128, 0, 0, 2)), ("bn128_split1", (128, 0, 0, 1)), ("streamk", (513, 1, 0, 0)), ("pair", (512, 0, 0, 0)), ("pair_streamk", (512, 1, 0, 0)), ("simt", None)) if t}
    kerns["simt"] = ctx.kernel(w.KernelId.SGEMM_SIMT, M, N, K)
    res = {k: [] for k in kerns}
    for r in range(5):
        for name, k in kerns.items():
            for i in range(3): ctx.launch(k, *sets[i % 3])
            ctx.timer_begin()
            for i in range(30): ctx.launch(k, *sets[i % 3])
            res[name].append(ctx.timer_end() / 30)
    flop = 2.0 * M * N * K
    print(f"{M:5d}x{N:5d}x{K:5d} " + "  ".join(f"{name} grid {kerns[name].geometry()[0][0]:4d} {np.median(v)*1e3:7.1f} us ({flop/np.median(v)/1e9:6.1f} TF)" for name, v in res.items()), flush=True)
    for k in kerns.values(): k.free()
    bench.free_sets(sets)
