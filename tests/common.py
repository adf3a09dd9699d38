"""Shared helpers for oracle and parity tests."""
import os

import torch

from oracle import cases, restated

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(kind):
    return torch.load(os.path.join(GOLDEN, f"small_{kind}.pt"), weights_only=False)


def golden_inputs(g):
    m = g["meta"]
    return cases.small_points(m["seed"], m["n_keep"], m["batch"])


def run_tier2(kind, pts, ptsp, B, mask_seed=None, train=True, device="cpu"):
    """Runs the tier-2 oracle; returns (vfe, bb, batch_dict after vfe (detached copy), batch_dict after bb)."""
    S = cases.SMALL
    vfe, bb = restated.build(kind, S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.to(device), bb.to(device)
    vfe.train(train), bb.train(train)
    bd = dict(points=torch.from_numpy(pts).to(device), points_prev=torch.from_numpy(ptsp).to(device), batch_size=B)
    bd = vfe(bd)
    after_vfe = {k: v for k, v in bd.items()}
    if kind == "pretrain":
        bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), B, 0.75, mask_seed).to(device)
    bb.trace = []
    bd = bb(bd)
    return vfe, bb, after_vfe, bd


def assert_close(a, b, rtol=1e-5, atol=1e-5, what=""):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} off, max err {err.max().item():.3e} (|ref| max {b.abs().max().item():.3e})"


def assert_equal_int(a, b, what=""):
    a, b = torch.as_tensor(a).long().cpu(), torch.as_tensor(b).long().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    assert torch.equal(a, b), f"{what}: {int((a != b).sum())} of {a.numel()} differ"
