"""BASELINE.json configurations at FULL size on the GPU: parity against the tier-2 oracle where it finishes in seconds
(one ONCE scan pair, one Waymo-shaped scan pair) and size-independent properties where it does not (120k-point
finetune batches, the 0.1 m stress grid)."""
import numpy as np
import pytest
import torch

from common import assert_close, assert_equal_int
from oracle import cases, restated

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200
    from tmae_b200 import ops, synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
FEAT_TOL = dict(rtol=2e-4, atol=2e-4)


def _pair(kind, shape_name, n_points, batch, npf, seed=1000, train=True, mask_seed=3):
    shape = synth.SHAPES[shape_name]
    grid = synth.grid_size(shape).tolist()
    pts, ptsp = synth.batch(seed, batch, n_points, shape_name)
    vfe, bb = tmae_b200.build_model(kind, grid, shape["voxel"], shape["range"], num_point_features=npf)
    ovfe, obb = restated.build(kind, grid, shape["voxel"], shape["range"], num_point_features=npf)
    for m in (vfe, bb, ovfe, obb):
        cases.fill_params(m)
        m.train(train)
    vfe.to(DEV), bb.to(DEV)
    bd = vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=batch))
    obd = ovfe(dict(points=torch.from_numpy(pts), points_prev=torch.from_numpy(ptsp), batch_size=batch))
    for sfx in ("", "_prev"):
        assert_equal_int(bd["voxel_coords" + sfx], obd["voxel_coords" + sfx], "voxel_coords" + sfx)
        assert_close(bd["voxel_features" + sfx], obd["voxel_features" + sfx].detach(), 1e-5, 1e-5, "voxel_features" + sfx)
    if kind == "pretrain":
        mask = cases.fixed_mask(obd["voxel_coords"], batch, 0.75, mask_seed)
        bd["voxel_mae_mask_in"], obd["voxel_mae_mask_in"] = mask.to(DEV), mask
    bd, obd = bb(bd), obb(obd)
    return vfe, bb, bd, ovfe, obb, obd


@pytest.mark.parametrize("shape_name,n_points,npf", [("once", 60000, 5), ("waymo", 180000, 6)])
def test_full_scan_pair_pretrain_parity(shape_name, n_points, npf):
    """configs[1] / configs[4] shapes, one scan pair: voxel coordinates and strided-conv sites bit-exact, features and the
    Chamfer loss against the fp32 oracle (fp32 parity mode)."""
    vfe, bb, bd, ovfe, obb, obd = _pair("pretrain", shape_name, n_points, 1, npf)
    grid = synth.grid_size(synth.SHAPES[shape_name]).tolist()
    for k, sp in bd["multi_scale_3d_features"].items():
        o = obd["multi_scale_3d_features"][k]
        assert_equal_int(sp.indices, o.indices, k + " indices")
        assert_close(sp.features, o.features.detach(), what=k, **FEAT_TOL)
    sf = bd["spatial_features"].detach()
    assert tuple(sf.shape) == (1, 128, grid[1], grid[0])
    assert_close(sf, obd["spatial_features"].detach(), what="spatial_features", **FEAT_TOL)
    loss, _ = bb.get_loss()
    oloss, _ = obb.get_loss()
    assert abs(loss.item() - oloss.item()) <= 1e-4 * abs(oloss.item())
    loss.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in list(vfe.parameters()) + list(bb.parameters()))


def test_finetune_dense_scans_properties():
    """configs[3] shape (120k-point scans, finetune forward, tensor-core mode), batch 2: Siamese-batched == per-frame
    encoding, every level of the 5-level partition is exercised, outputs finite."""
    shape = synth.ONCE
    grid = synth.grid_size(shape).tolist()
    pts, ptsp = synth.batch(2000, 2, 120000)
    outs = []
    ops.set_precision("tf32")
    try:
        for batched in (True, False):
            vfe, bb = tmae_b200.build_model("finetune", grid, shape["voxel"], shape["range"])
            cases.fill_params(vfe), cases.fill_params(bb)
            vfe.to(DEV).eval(), bb.to(DEV).eval()
            bb.siamese_batched = batched
            with torch.no_grad():
                bd = bb(vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=2)))
            outs.append(bd["spatial_features"].float())
            if batched:
                plans, tparts = bb.last_plan
                lb = plans[2].stages[0].part.level_base[0].cpu().tolist()
                assert len(lb) == 6 and lb[-1] == int(plans[2].stages[0].part.n_win[0])
                assert all(b > a for a, b in zip(lb[:-1], lb[1:])), f"a partition level is empty at this density: {lb}"
    finally:
        ops.set_precision("fp32")
    assert torch.isfinite(outs[0]).all()
    assert_close(outs[0], outs[1], 1e-4, 1e-4, "Siamese-batched vs per-frame, 120k-point scans")


def test_stress_grid_partition_properties():
    """0.1 m voxels (1498^2 grid, SURVEY 8d secondary stress for rows A1-A6): voxelise + two-shift window partition;
    the token table is a permutation of the voxels, counts and level boundaries are consistent."""
    pts, _ = synth.batch(1000, 2, 60000)
    grid = [1498, 1498, 1]
    v = ops.voxelize(torch.from_numpy(pts).to(DEV), synth.ONCE["range"], [0.1, 0.1, 8.0], grid, 2)
    nk, nv = v["counts"][:2].tolist()
    vc = v["voxel_coords"][:nv]
    idx = torch.stack((vc[:, 0], vc[:, 2], vc[:, 3]), 1).int()
    levels = [(8, 0, 8), (16, 8, 16), (32, 16, 32), (48, 32, 48), (64, 48, 100000)]
    P = ops.window_partition(idx, 2, grid[0], grid[1], levels)
    assert int(P.status.item()) == 0
    for s in range(2):
        nw = int(P.n_win[s])
        cnt = P.cnt_a[s, :nw].cpu().long()
        assert int(cnt.sum()) == nv and int(cnt.min()) >= 1 and int(cnt.max()) <= 64
        tok = P.tok_a[s].view(-1, 64)[:nw].cpu().long()
        valid = torch.arange(64)[None, :] < cnt[:, None]
        rows = tok[valid]
        assert torch.equal(rows.sort().values, torch.arange(nv)), "token table is a permutation of the voxels"
        win, slot = P.win_a[s, :nv].cpu().long(), P.slot_a[s, :nv].cpu().long()
        assert torch.equal(tok[win, slot], torch.arange(nv)), "tok[win[m], slot[m]] == m"
        lb = P.level_base[s].cpu().tolist()
        assert lb[0] == 0 and lb[-1] == nw
        for l, (t, lo, hi) in enumerate(levels):
            c = cnt[lb[l]:lb[l + 1]]
            assert c.numel() == 0 or (int(c.min()) >= max(lo, 1) and int(c.max()) < hi and int(c.max()) <= t)
        # reference arithmetic of get_window_coors (sst_utils.py:23-48) on the host for the window id of every voxel
        shift = 8 if s == 0 else 4
        wx, wy = (idx[:, 2].cpu().long() + shift) // 8, (idx[:, 1].cpu().long() + shift) // 8
        nwx, nwy = int(np.ceil(grid[0] / 8) + 1), int(np.ceil(grid[1] / 8) + 1)
        bwi = idx[:, 0].cpu().long() * nwx * nwy + wx * nwy + wy
        # compacted window ids are the ranks of the distinct reference ids inside each level
        order = torch.unique(bwi, sorted=True)
        assert order.numel() == nw
