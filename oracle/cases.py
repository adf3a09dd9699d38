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


def raw_samples(seed, n_points=3000, kind="once"):
    """Four raw samples for the batch-assembly tests (tmae_b200.synth.raw_scan_pair): moving vehicle; all-zero poses
    (static vehicle: no transformation, once_utils.py:5-10); zero previous pose with a non-zero current pose (only the
    second map applies); and frame_id == frame_id_prev (alignment skipped, once_temporal_dataset.py:169).  Ragged sizes;
    a few hand-placed points sit exactly on the ego radius and on the crop bounds (both tests are strict / closed)."""
    import tmae_b200  # noqa: F401
    from tmae_b200 import synth
    rng32 = np.asarray(synth.SHAPES[kind]["range"], np.float32)
    out = []
    for i, n in enumerate([n_points, n_points // 2 + 7, n_points // 3 + 1, n_points // 4 + 3]):
        s = synth.raw_scan_pair(seed + i, n, kind, static=(i == 1))
        if i == 2:
            s["pose_prev"] = np.zeros(7)
            s["pose"][4:] = [1.5, -0.7, 0.1]   # keeps the mapped points inside the crop
        if i == 3:
            s["frame_id_prev"] = s["frame_id"]
        for key in ("points", "points_prev"):
            p = s[key]
            p[0, :2] = [2.0, 0.5]             # |x| == r: not an ego point (strict <)
            p[1, :2] = [-1.999, 1.999]        # ego point
            p[2, :2] = [rng32[0], rng32[4]]   # on the crop bounds: kept (closed interval)
            p[3, :2] = [np.nextafter(rng32[3], np.float32(np.inf)), 0.0]   # one ulp outside
            p[4, :2] = [np.nan, 0.0]          # NaN fails the range test
        out.append(s)
    return out


def bev_input(seed, B=2, C=128, Y=40, X=40):
    """A seeded BEV map shaped like `spatial_features`: non-negative (it follows a ReLU) and ~70 % empty."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, Y, X, generator=g).clamp_min(0)
    occ = (torch.rand(B, 1, Y, X, generator=g) < 0.3).float()
    return x * occ
