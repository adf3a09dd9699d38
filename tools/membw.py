"""Measured HBM rates on this GPU for pure-write, pure-read and copy streams (context for the per-kernel rooflines)."""
import torch

DEV = "cuda"
n = 1 << 28  # 1 GiB of fp32
a = torch.empty(n, device=DEV)
b = torch.empty(n, device=DEV)


def t(fn, reps=5):
    fn()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best * 1e-3


w = t(lambda: a.zero_())
r = t(lambda: a.sum())
c = t(lambda: b.copy_(a))
m = t(lambda: torch.add(a, 1.0, out=b))
print(f"memset {4 * n / w / 1e9:.0f} GB/s | read (sum) {4 * n / r / 1e9:.0f} GB/s | copy {8 * n / c / 1e9:.0f} GB/s (r+w) | add-scalar {8 * n / m / 1e9:.0f} GB/s (r+w)")
for sz in (28.7e6, 57e6, 115e6):
    k = int(sz // 4)
    x, y = torch.empty(k, device=DEV), torch.empty(k, device=DEV)
    big = torch.empty(1 << 27, device=DEV)

    def f():
        big.zero_()  # flush L2
    f()
    ts = []
    for _ in range(5):
        f()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        y.copy_(x)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"copy of {sz / 1e6:.1f} MB (L2 flushed before): {min(ts) * 1e3:.1f} us = {2 * sz / (min(ts) * 1e-3) / 1e9:.0f} GB/s")
