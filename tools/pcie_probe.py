"""PCIe probe for the end-to-end (host-buffer) path: H2D alone, D2H alone, both directions at once (pinned memory)."""
import torch, time
torch.cuda.init()
MB = 1 << 20
hA = torch.empty(128 * MB, dtype=torch.uint8).pin_memory()
hC = torch.empty(64 * MB, dtype=torch.uint8).pin_memory()
dA = torch.empty(128 * MB, dtype=torch.uint8, device="cuda")
dC = torch.empty(64 * MB, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(n):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best * 1e3

def h2d():
    with torch.cuda.stream(s1): dA.copy_(hA, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): hC.copy_(dC, non_blocking=True)
def both():
    h2d(); d2h()
def h2d_chunks(n=32):
    c = 128 * MB // n
    with torch.cuda.stream(s1):
        for i in range(n): dA[i * c:(i + 1) * c].copy_(hA[i * c:(i + 1) * c], non_blocking=True)
a, b, c, d = t(h2d), t(d2h), t(both), t(h2d_chunks)
print(f"H2D 128 MiB: {a:.3f} ms ({128*MB/a/1e6:.1f} GB/s)   D2H 64 MiB: {b:.3f} ms ({64*MB/b/1e6:.1f} GB/s)   both at once: {c:.3f} ms   H2D in 32 chunks: {d:.3f} ms")
