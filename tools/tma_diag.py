"""Diagnostic: error of the TMA/TF32 GEMMs against float64 references with different operand roundings."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa
from tmae_b200 import ops
DEV = "cuda"
ops.set_precision("tf32")

def rn(t):
    i = t.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()
def rz(t):
    i = t.float().contiguous().view(torch.int32)
    return (i & ~0x1FFF).view(torch.float32).double()
def bf(t):
    return t.to(torch.bfloat16).double()
def report(name, got, fn):
    got = got.double().cpu()
    out = []
    for tag, f in (("exact", lambda t: t.double()), ("tf32rn", rn), ("tf32rz", rz), ("bf16", bf)):
        ref = fn(f)
        e = (got - ref).abs()
        out.append(f"{tag}: max {e.max():.2e} mean {e.mean():.2e}")
    print(f"{name:28s} |ref|max {fn(lambda t: t.double()).abs().max():.2f}  " + " | ".join(out))

for m, n, k in [(128, 64, 64), (1000, 128, 128), (2049, 256, 512), (333, 256, 128), (65, 48, 128), (5000, 512, 256)]:
    g = torch.Generator().manual_seed(m)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5, torch.randn(n, generator=g)
    dy = torch.randn(m, n, generator=g)
    for tma in (1, 0):
        ops.set_option("tma", tma)
        tag = f"{'tma' if tma else 'stg'} {m}x{n}x{k}"
        y = ops.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV))
        report(tag + " fwd", y, lambda f: f(x) @ f(w).T + b.double())
        dx = ops.linear_bwd_data(dy.to(DEV), w.to(DEV))
        report(tag + " dx", dx, lambda f: f(dy) @ f(w))
        dw, db = torch.empty(n, k, device=DEV), torch.empty(n, device=DEV)
        ops.linear_bwd_weight(dy.to(DEV), x.to(DEV), dw, db)
        report(tag + " dw", dw, lambda f: f(dy).T @ f(x))
