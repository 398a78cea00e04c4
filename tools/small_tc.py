import sys; sys.path.insert(0, "/root/repo")
import wgpu_mm_b200 as w
ctx = w.Context(0)
for S in (512, 1024, 2048):
    a = ctx.buffer(S*S*4); a.fill_weights(1, S*S); b = ctx.buffer(S*S*4); b.fill_weights(2, S*S); c = ctx.buffer(S*S*4)
    k = ctx.kernel(w.KernelId.SGEMM_TC3X, S, S, S)
    for _ in range(5): ctx.launch(k, a, b, c)
    ctx.sync()
    k.free()
ctx.close()
