"""Runs a few launches of the kernels under study (for ncu):  python tools/prof_target.py nn|nt|attn|attn1|attn1s"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa: E402,F401
from tmae_b200 import ops  # noqa: E402

DEV = "cuda"
what = sys.argv[1]
ops.set_precision("tf32")
if what == "nn":
    m, n, k = 50000, 256, 256
    dy, w = torch.randn(m, n, device=DEV), torch.randn(n, k, device=DEV)
    for _ in range(4):
        ops.linear_bwd_data(dy, w)
elif what == "nt128":
    m, n, k = 56000, 128, 128
    xs = [torch.randn(m, k, device=DEV) for _ in range(4)]
    w, b = torch.randn(n, k, device=DEV), torch.randn(n, device=DEV)
    for i in range(8):
        ops.linear_fwd(xs[i % 4], w, b)
elif what == "nt":
    m, n, k = 50000, 256, 256
    x, w, b = torch.randn(m, k, device=DEV), torch.randn(n, k, device=DEV), torch.randn(n, device=DEV)
    for _ in range(4):
        ops.linear_fwd(x, w, b)
else:
    M, C, g, B = {"attn": (50000, 256, 234, 4), "attn1": (56000, 128, 468, 4), "attn1s": (14000, 128, 468, 4), "attn3": (28000, 256, 117, 4)}[what]
    ops.set_option("attn_tc", 1)
    rng = np.random.default_rng(0)
    cells = np.unique(np.clip((rng.normal(0, g / 5, (M * 3, 2)) + g / 2).astype(np.int64), 0, g - 1) @ np.array([g, 1]))
    per = min(M // B, cells.shape[0])
    cells = np.sort(rng.choice(cells, per, replace=False))
    c = np.concatenate([np.stack([np.full(per, b), cells // g, cells % g], 1) for b in range(B)])
    coords = torch.tensor(c, dtype=torch.int32, device=DEV)
    P = ops.window_partition(coords, B, g, g, [(16, 0, 16), (32, 16, 32), (64, 32, 100000)])
    m = coords.shape[0]
    q, k, v = (torch.randn(m, C, device=DEV) for _ in range(3))
    tau = torch.ones(1, device=DEV)
    args = (P.tok_a[0], P.cnt_a[0], P.tok_a[0], P.cnt_a[0], P.n_win[0:1], ops.small_end(P, 0), ops.mid_end(P, 0), min(P.wcap, m), tau, 0.01, 8)
    for _ in range(3):
        o, lse = ops.window_attention_fwd(q, k, v, *args, False)
        ops.window_attention_bwd(torch.randn_like(o), q, k, v, o, lse, *args, torch.zeros(1, device=DEV), False)
torch.cuda.synchronize()
