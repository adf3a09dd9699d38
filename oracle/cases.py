"""TEST INFRASTRUCTURE ONLY -- shared seeded cases for oracle / parity tests.

`fill_params` gives every parameter and buffer a deterministic, non-trivial value derived from
its NAME, so tier 1 (reference), tier 2 (restatement) and the CUDA modules get identical weights
without shipping a 47 MB state_dict: the module trees share parameter names by construction.
"""
import zlib

import numpy as np
import torch

SMALL = dict(range=[-15.36, -15.36, -5.0, 15.36, 15.36, 3.0], voxel=[0.32, 0.32, 8.0], grid=[96, 96, 1])


def fill_params(module, seed=0):
    sd = module.state_dict()
    for name, t in sd.items():
        g = torch.Generator().manual_seed(seed * 1000003 + zlib.crc32(name.encode()))
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            t.zero_()
        elif leaf == "tau":
            t.copy_(0.5 + torch.rand(t.shape, generator=g))
        elif leaf == "running_var":
            t.copy_(0.5 + torch.rand(t.shape, generator=g))
        elif leaf == "running_mean":
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1 and leaf == "weight":  # BN / LN scale
            t.copy_(1.0 + 0.1 * torch.randn(t.shape, generator=g))
        elif t.dim() == 1:  # biases
            t.copy_(0.1 * torch.randn(t.shape, generator=g))
        else:
            fan_in = t[0].numel()
            if t.dim() == 4 and "deblocks" in name:  # ConvTranspose2d (Cin, Cout, k, k)
                fan_in = t.shape[0]
            t.copy_(torch.randn(t.shape, generator=g) / np.sqrt(fan_in))
    module.load_state_dict(sd)
    return module


def small_points(seed, n_keep=4000, batch_size=2, kind="once"):
    """Synthetic scan pairs cropped to the 96x96 test grid; a few points are left outside the
    x/y range and outside the z range on purpose."""
    import tmae_b200  # noqa: F401  (package alias)
    from tmae_b200 import synth
    rng = np.random.default_rng(seed)
    cur, prev = [], []
    for b in range(batch_size):
        for dst, p in zip((cur, prev), synth.scan_pair(seed + b, 60000, kind)):
            m = (np.abs(p[:, 0]) < 15.9) & (np.abs(p[:, 1]) < 15.9)
            p = p[m]
            p = p[rng.permutation(p.shape[0])[:n_keep]]
            dst.append(p)
    return synth.collate(cur), synth.collate(prev)


def fixed_mask(voxel_coords, batch_size, ratio, seed):
    """A seeded 75 % mask per sample with the reference's keep count int(L * (1 - ratio))."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for b in range(batch_size):
        L = int((voxel_coords[:, 0] == b).sum())
        keep = int(L * (1 - ratio))
        ids = torch.argsort(torch.rand(L, generator=g))[:keep]
        m = torch.ones(L)
        m[ids] = 0
        out.append(m)
    return torch.cat(out)
