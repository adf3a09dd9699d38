"""CPU tests of the oracle itself: tier 2 (oracle/restated.py) against the committed golden tensors
made from the reference's own modules (tests/golden/make_golden.py), tier 1 vs tier 2 live where
/root/reference exists, and hand-computable micro-cases / properties."""
import os

import numpy as np
import pytest
import torch
from hypothesis import given, settings, strategies as st

from common import GOLDEN, assert_close, assert_equal_int, golden_inputs, load_golden, run_tier2
from oracle import cases, ref_loader, restated

ROW = 3


@pytest.fixture(scope="module", params=["pretrain", "finetune"])
def run(request):
    kind = request.param
    g = load_golden(kind)
    pts, ptsp = golden_inputs(g)
    vfe, bb, av, bd = run_tier2(kind, pts, ptsp, g["meta"]["batch"], g["meta"]["mask_seed"])
    return kind, g, vfe, bb, av, bd


def test_vfe_against_golden(run):
    kind, g, vfe, bb, av, bd = run
    for sfx in ("", "_prev"):
        assert_equal_int(av["voxel_coords" + sfx], g["voxel_coords" + sfx], "voxel_coords" + sfx)
        assert_close(av["voxel_features" + sfx][::ROW], g["voxel_features" + sfx], what="voxel_features" + sfx)
        if kind == "pretrain":
            assert_equal_int(av["point_inverse_indices" + sfx], g["point_inverse_indices" + sfx])
            assert_equal_int(av["point_coords" + sfx], g["point_coords" + sfx])
            assert av["points" + sfx].shape[0] == g["n_points" + sfx]


def test_partition_against_golden(run):
    kind, g, vfe, bb, av, bd = run
    S = cases.SMALL
    info = restated.sst_input(av["voxel_features_prev"].detach(), av["voxel_coords_prev"],
                              [S["grid"][0], S["grid"][1], 1], bb.sst_blocks[0].pre_cfg)
    assert_equal_int(info["voxel_keep_inds"], g["part_keep"])
    for s in range(2):
        assert_equal_int(info[f"batch_win_inds_shift{s}"], g[f"part_bwi{s}"])
        assert_equal_int(info[f"voxel_drop_level_shift{s}"], g[f"part_lvl{s}"])
        assert_equal_int(info[f"coors_in_win_shift{s}"], g[f"part_ciw{s}"])
        t = info[f"flat2win_inds_shift{s}"]
        assert sorted(k for k in t if not isinstance(k, str)) == sorted(g[f"part_f2w{s}"])
        for dl, (inds, pos) in g[f"part_f2w{s}"].items():
            assert_equal_int(t[dl][0], inds), assert_equal_int(t[dl][1][0], pos)
    a, b = restated.sst_input_temporal(av["voxel_features"].detach(), av["voxel_coords"],
                                       av["voxel_features_prev"].detach(), av["voxel_coords_prev"],
                                       [S["grid"][0], S["grid"][1], 1], bb.wca_blocks[0].pre_cfg)
    for tag, inf in (("cur", a), ("prv", b)):
        for s in range(2):
            assert_equal_int(inf[f"voxel_keep_inds_shift{s}"], g[f"tpart_{tag}_keep{s}"])
            assert_equal_int(inf[f"voxel_drop_level_shift{s}"], g[f"tpart_{tag}_lvl{s}"])
            for dl, (inds, pos) in g[f"tpart_{tag}_f2w{s}"].items():
                assert_equal_int(inf[f"flat2win_inds_shift{s}"][dl][0], inds)


def test_backbone_against_golden(run):
    kind, g, vfe, bb, av, bd = run
    for k, v in bd["multi_scale_3d_features"].items():
        assert_equal_int(v.indices, g[k + "_indices"], k)
        assert_close(v.features[::ROW], g[k + "_features"], rtol=1e-4, atol=1e-4, what=k)
    sf = bd["spatial_features"].detach()
    assert abs(sf.double().sum().item() - g["spatial_sum"]) <= 1e-5 * g["spatial_abs_sum"]
    c = av["voxel_coords"].long()
    assert_close(sf.permute(0, 2, 3, 1)[c[:, 0], c[:, 2], c[:, 3]][::ROW], g["spatial_at_voxels"], rtol=1e-4, atol=1e-4)
    if kind == "pretrain":
        r = bb.forward_ret_dict
        assert_close(r["gt_points"][::ROW], g["gt_points"], what="gt_points")
        assert_close(r["pred_points"][::ROW], g["pred_points"], rtol=1e-4, atol=1e-4, what="pred_points")
        loss, _ = bb.get_loss()
        assert abs(loss.item() - g["loss"]) <= 1e-5 * abs(g["loss"])
        loss.backward()
        for m, pre in ((vfe, "vfe."), (bb, "backbone_3d.")):
            for k, p in m.named_parameters():
                ref = g["grad_abs_sum"][k]
                assert abs(p.grad.double().abs().sum().item() - ref) <= 2e-4 * ref + 1e-7, k


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present (tier 1 unavailable)")
@pytest.mark.parametrize("kind", ["pretrain", "finetune"])
def test_tier1_equals_tier2_live(kind):
    """The reference's own modules vs the restatement on a DIFFERENT seed than the goldens."""
    ns = ref_loader.load()
    S = cases.SMALL
    pts, ptsp = cases.small_points(23, 900, 2)
    v1, b1 = ref_loader.build(kind, S["grid"], S["voxel"], S["range"])
    cases.fill_params(v1), cases.fill_params(b1)
    bd1 = v1(dict(points=torch.from_numpy(pts), points_prev=torch.from_numpy(ptsp), batch_size=2))
    mask = cases.fixed_mask(bd1["voxel_coords"], 2, 0.75, 9)
    off = [0]

    def rm(N, L, ratio, device):
        m = mask[off[0]:off[0] + L][None]
        off[0] += L
        return m
    orig = ns.common_utils.random_masking
    ns.common_utils.random_masking = rm
    try:
        bd1 = b1(bd1)
    finally:
        ns.common_utils.random_masking = orig
    v2, b2, av, bd2 = run_tier2(kind, pts, ptsp, 2, 9)
    assert_equal_int(bd1["voxel_coords"], bd2["voxel_coords"])
    assert_close(bd1["spatial_features"].detach(), bd2["spatial_features"].detach(), 1e-5, 1e-5)
    for k in bd1["multi_scale_3d_features"]:
        a, b = bd1["multi_scale_3d_features"][k], bd2["multi_scale_3d_features"][k]
        assert_equal_int(a.indices, b.indices)
        assert_close(a.features.detach(), b.features.detach(), 1e-5, 1e-5)
    if kind == "pretrain":
        assert_close(b1.get_loss()[0].detach(), b2.get_loss()[0].detach(), 1e-6, 0)


def test_micro_cases_from_reference():
    m = torch.load(os.path.join(GOLDEN, "micro.pt"), weights_only=False)
    assert_equal_int(restated.stable_rank(m["ingroup_in"]), m["ingroup_out"])
    assert m["ingroup_out"].tolist() == [0, 0, 1, 2, 0, 1, 3, 0, 1, 1]
    gi = restated.group_inner_inds(m["group_inv"], 3, 4)
    assert gi.tolist() == [[0, 5, 0, 5], [3, 3, 3, 3], [1, 2, 4, 6]]
    assert_close(m["group_pts"][gi], m["group_out"], 0, 0)
    for s in (0, 1):
        bwi, ciw = restated.window_coords(m["win_coords"], [96, 96, 1], [8, 8, 1], bool(s))
        assert_equal_int(bwi, m[f"win_bwi_{s}"]), assert_equal_int(ciw, m[f"win_ciw_{s}"])


def test_reference_docstring_example():
    """SiamWCA.py:690-706: windows [1,2,3,2] vs prev [1,3,3,2,3,2,1,3], buckets 1/2/4 -> cur levels [1,1,2,1]."""
    di = {0: {"max_tokens": 1, "drop_range": [0, 2]}, 1: {"max_tokens": 2, "drop_range": [2, 4]},
          2: {"max_tokens": 4, "drop_range": [4, 100]}}
    cur = torch.tensor([1, 2, 3, 2])
    prv = torch.tensor([1, 3, 3, 2, 3, 2, 1, 3])
    k, l, kp, lp = restated.drop_temporal(cur, prv, di)
    assert l.tolist() == [1, 1, 2, 1] and lp.tolist() == [1, 2, 2, 1, 2, 1, 1, 2]
    assert k.all() and kp.all()


def test_chamfer_known_answer():
    x = torch.tensor([[[0., 0, 0], [1, 0, 0]], [[5., 5, 5], [5, 5, 6]]])
    y = torch.tensor([[[0., 0, 1], [1, 0, 0], [3, 0, 0]], [[0., 0, 0], [0, 0, 0], [0, 0, 0]]])
    w = torch.tensor([1., 0.])
    # cloud 0: x->y mins 1, 0 -> mean .5 ; y->x mins 1, 0, 4 -> mean 5/3 ; sum(w) = 1
    assert abs(restated.chamfer(x, y, w).item() - (0.5 + 5 / 3)) < 1e-6
    assert restated.chamfer(x, y, torch.zeros(2)).item() == 0.0


def test_truncation_rule():
    """z in (lo - vs, lo) truncates to cell 0 and is KEPT (common_utils.py:74-75)."""
    p = torch.tensor([[0, 0.1, 0.1, -12.9, 0.5], [0, 0.1, 0.1, -13.1, 0.5], [0, 0.1, 0.1, 2.99, 0.5],
                      [0, 0.1, 0.1, 3.0, 0.5], [0, -15.5, 0.1, 0.0, 0.5]])
    S = cases.SMALL
    keep, c = restated.in_range_coords(p, S["range"], S["voxel"], S["grid"])
    assert keep.tolist() == [True, False, True, False, True]


@settings(max_examples=60, deadline=None)
@given(st.lists(st.integers(0, 12), min_size=0, max_size=200))
def test_stable_rank_property(groups):
    g = torch.tensor(groups, dtype=torch.long)
    r = restated.stable_rank(g).tolist()
    seen = {}
    for i, x in enumerate(groups):
        assert r[i] == seen.get(x, 0)
        seen[x] = seen.get(x, 0) + 1


@settings(max_examples=25, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 400))
def test_partition_invariants(seed, m):
    """Assert table of SURVEY.md section 4: compact window ids have no gaps, window2flat writes every
    row once, flat2window/window2flat round-trips, every voxel gets a level."""
    rng = np.random.default_rng(seed)
    cells = rng.choice(2 * 40 * 40, size=min(m, 3200), replace=False)
    cells.sort()
    coords = torch.tensor(np.stack([cells // 1600, np.zeros_like(cells), (cells % 1600) // 40, cells % 40], 1))
    pre = restated.model_cfg("finetune")["BACKBONE_3D"]["SST_BLOCK_LIST"][0]["PREPROCESS"]
    feat = torch.randn(coords.shape[0], 8)
    info = restated.sst_input(feat, coords, [40, 40, 1], pre)
    assert info["voxel_keep_inds"].shape[0] == coords.shape[0]  # F5: nothing is ever dropped
    for s in range(2):
        t = info[f"flat2win_inds_shift{s}"]
        back = restated.window2flat(restated.flat2window(feat, t), t)
        assert torch.equal(back, feat)
        for dl in (k for k in t if not isinstance(k, str)):
            T = t["batching_info"][dl]["max_tokens"]
            win = t[dl][0] // T
            assert win.unique().numel() == int(win.max()) + 1
            assert (t[dl][0] % T < T).all() and t[dl][0].unique().numel() == t[dl][0].numel()


def test_fp32_oracle_distance_from_float64_gradients():
    """Documents the noise floor of gradient parity: the fp32 tier-2 oracle against the same oracle in float64
    (tests/golden/small_pretrain_grad64.pt).  The worst tensor is a few 1e-3 of its scale away, the median ~2e-4 --
    the GPU tests therefore pin gradients to the float64 fixture, not to the fp32 run."""
    g = load_golden("pretrain")
    pts, ptsp = golden_inputs(g)
    vfe, bb, av, bd = run_tier2("pretrain", pts, ptsp, g["meta"]["batch"], g["meta"]["mask_seed"])
    loss, _ = bb.get_loss()
    loss.backward()
    g64 = torch.load(os.path.join(GOLDEN, "small_pretrain_grad64.pt"), weights_only=False)
    assert abs(loss.item() - g64["loss"]) <= 1e-6 * abs(g64["loss"])
    dev = []
    for pre, m in (("vfe.", vfe), ("backbone_3d.", bb)):
        for k, p in m.named_parameters():
            scale, sample = g64["grads"][pre + k]
            dev.append((p.grad.flatten()[::g64["stride"]].double() - sample.double()).abs().max().item() / (scale + 1e-12))
    dev.sort()
    assert dev[-1] < 2e-2 and dev[len(dev) // 2] < 1e-3
    assert dev[-1] > 1e-4, "the fp32 oracle is closer to float64 than documented: tighten the GPU gradient tolerances"


def test_point_order_invariance():
    """SURVEY 8c pin (4): shuffling the points of a scan leaves voxel_coords bit-identical (unique(dim=0) sorts) and moves
    voxel_features only by fp32 summation order (measured on the reference's own modules: 1.4e-6) -- the noise floor every
    fp32 tolerance in tests/ sits above."""
    S = cases.SMALL
    pts, ptsp = cases.small_points(23, 1200, 2)
    vfe, _ = restated.build("pretrain", S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe)
    vfe.train()
    a = vfe(dict(points=torch.from_numpy(pts), points_prev=torch.from_numpy(ptsp), batch_size=2))
    g = torch.Generator().manual_seed(5)
    p1, p2 = torch.randperm(pts.shape[0], generator=g), torch.randperm(ptsp.shape[0], generator=g)
    cases.fill_params(vfe)   # running statistics back to their start values
    b = vfe(dict(points=torch.from_numpy(pts)[p1], points_prev=torch.from_numpy(ptsp)[p2], batch_size=2))
    for sfx in ("", "_prev"):
        assert torch.equal(a["voxel_coords" + sfx], b["voxel_coords" + sfx])
        err = (a["voxel_features" + sfx] - b["voxel_features" + sfx]).abs().max().item()
        assert err < 1e-5, f"voxel_features{sfx} moved by {err:.2e} under a point shuffle"
