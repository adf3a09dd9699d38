"""Timeline of CTA 0 of one TMA GEMM launch (diagnostics):  python tools/gemm_trace.py m n k [mode]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa: E402,F401
from tmae_b200 import ops  # noqa: E402
from tmae_b200._lib import lib  # noqa: E402

m, n, k = (int(v) for v in sys.argv[1:4])
mode = sys.argv[4] if len(sys.argv) > 4 else "nt"
ops.set_precision("tf32")
DEV = "cuda"
xs = [torch.randn(m, k, device=DEV) for _ in range(3)]
dys = [torch.randn(m, n, device=DEV) for _ in range(3)]
w, b = torch.randn(n, k, device=DEV), torch.randn(n, device=DEV)
fn = (lambda i: ops.linear_fwd(xs[i], w, b)) if mode == "nt" else (lambda i: ops.linear_bwd_data(dys[i], w))
for i in range(3):
    fn(i)
cap = 4096
buf = torch.zeros(cap, dtype=torch.int64, device=DEV)
lib().debug_set_trace(buf.data_ptr(), cap)
fn(0)
torch.cuda.synchronize()
lib().debug_set_trace(None, 0)
t = buf.cpu().tolist()
names = {1: "P issue", 2: "M acc-free", 3: "M stage-full", 4: "E wait", 5: "E start", 6: "E done"}
ev = []
for v in t:
    if v == 0:
        continue
    v &= (1 << 64) - 1
    ev.append((v & 0xffffffffff, (v >> 56) & 0xff, (v >> 40) & 0xffff))
ev.sort()
t0 = ev[0][0]
role = {1: "P ", 2: " M", 3: " M"}
for c, e, i in ev:
    print(f"{(c - t0) / 1.9e3:8.2f} us  {names[e]:14s} {i}")
