"""GPU parity tests of the drop-in modules (vfe -> backbone_3d.forward -> get_loss) against the tier-2 oracle
and the committed golden tensors made from the reference's own modules (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from common import GOLDEN, assert_close, assert_equal_int, golden_inputs, load_golden, run_tier2
from oracle import cases

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200
    from tmae_b200 import ops
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
S = cases.SMALL
ROW = 3
# fp32 parity mode.  The reference's own fp32 path moves by ~1e-6 under a point shuffle and ~6e-6 (BEV features)
# under its nondeterministic slot order; sums run in a different order on the GPU, so features
# after 18 encoder layers + 13 BatchNorms are compared at rtol 1e-4 + atol 1e-4, single ops at 1e-5 (test_gpu_ops).
FEAT_TOL = dict(rtol=1e-4, atol=1e-4)


def run_product(kind, pts, ptsp, B, mask_seed=None, train=True):
    vfe, bb = tmae_b200.build_model(kind, S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.to(DEV), bb.to(DEV)
    vfe.train(train), bb.train(train)
    bb.debug_refs = True
    bd = dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=B)
    bd = vfe(bd)
    after_vfe = dict(bd)
    if kind == "pretrain":
        bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), B, 0.75, mask_seed).to(DEV)
    bd = bb(bd)
    return vfe, bb, after_vfe, bd


@pytest.fixture(scope="module", params=["pretrain", "finetune"])
def both(request):
    kind = request.param
    g = load_golden(kind)
    pts, ptsp = golden_inputs(g)
    B, ms = g["meta"]["batch"], g["meta"]["mask_seed"]
    prod = run_product(kind, pts, ptsp, B, ms)
    orac = run_tier2(kind, pts, ptsp, B, ms)
    return kind, g, prod, orac


def test_vfe_matches_oracle_and_golden(both):
    kind, g, (vfe, bb, av, bd), (ovfe, obb, oav, obd) = both
    for sfx in ("", "_prev"):
        assert_equal_int(av["voxel_coords" + sfx], oav["voxel_coords" + sfx])
        assert_equal_int(av["voxel_coords" + sfx], g["voxel_coords" + sfx])
        assert_close(av["voxel_features" + sfx], oav["voxel_features" + sfx].detach(), 1e-5, 1e-5, "voxel_features" + sfx)
        assert_close(av["voxel_features" + sfx][::ROW], g["voxel_features" + sfx], 1e-5, 1e-5, "golden voxel_features" + sfx)
        if kind == "pretrain":
            assert_equal_int(av["point_inverse_indices" + sfx], g["point_inverse_indices" + sfx])
            assert_equal_int(av["point_coords" + sfx], g["point_coords" + sfx])
            assert_close(av["points" + sfx], oav["points" + sfx], 0, 0)
        else:
            assert "points" + sfx not in av and "point_coords" + sfx not in av
    # BatchNorm running statistics were updated like the oracle's
    for k, v in vfe.state_dict().items():
        if "running" in k:
            assert_close(v, ovfe.state_dict()[k], 1e-5, 1e-6, k)


def test_partition_tables_match_golden(both):
    """Reference-format tables of the prev frame's stage-1 partition and of the temporal stage-1 partition."""
    kind, g, (vfe, bb, av, bd), _ = both
    plans, tparts = bb.last_plan
    if kind == "finetune":
        P, m = plans[1].stages[0].part, plans[1].stages[0].m
        bwi, lvl, f2w = (t.cpu() for t in P.ref_a)
        assert_equal_int(torch.arange(m), g["part_keep"])
        for s in range(2):
            assert_equal_int(bwi[s, :m], g[f"part_bwi{s}"]), assert_equal_int(lvl[s, :m], g[f"part_lvl{s}"])
            ciw = g[f"part_ciw{s}"].long()
            assert_equal_int(P.posidx_a[s, :m].cpu(), ciw[:, 1] * 8 + ciw[:, 2])
            for dl, (inds, pos) in g[f"part_f2w{s}"].items():
                assert_equal_int(f2w[s, :m][pos.long()], inds, f"flat2win level {dl}")
        T = tparts[0]
        for tag, ref, win in (("cur", T.ref_a, T.win_a), ("prv", T.ref_b, T.win_b)):
            bwi, lvl, f2w = (t.cpu() for t in ref)
            for s in range(2):
                keep = torch.where(win[s].cpu() >= 0)[0]
                assert_equal_int(keep, g[f"tpart_{tag}_keep{s}"])
                assert_equal_int(lvl[s][keep], g[f"tpart_{tag}_lvl{s}"])
                lv = g[f"tpart_{tag}_lvl{s}"].long()
                for dl, (inds, pos) in g[f"tpart_{tag}_f2w{s}"].items():
                    assert_equal_int(f2w[s][keep][lv == dl], inds, f"temporal flat2win level {dl}")


def test_backbone_matches_oracle_and_golden(both):
    kind, g, (vfe, bb, av, bd), (ovfe, obb, oav, obd) = both
    for k, sp in bd["multi_scale_3d_features"].items():
        o = obd["multi_scale_3d_features"][k]
        assert_equal_int(sp.indices, o.indices, k + " indices"), assert_equal_int(sp.indices, g[k + "_indices"])
        assert sp.spatial_shape == o.spatial_shape
        assert_close(sp.features, o.features.detach(), what=k, **FEAT_TOL)
        assert_close(sp.features[::ROW], g[k + "_features"], what="golden " + k, **FEAT_TOL)
    assert bd["multi_scale_3d_strides"] == obd["multi_scale_3d_strides"]
    assert bd["spatial_features_stride"] == obd["spatial_features_stride"]
    sf = bd["spatial_features"].detach()
    assert tuple(sf.shape) == (g["meta"]["batch"], 128, 96, 96)
    assert_close(sf, obd["spatial_features"].detach(), what="spatial_features", **FEAT_TOL)
    assert abs(sf.double().sum().item() - g["spatial_sum"]) <= 1e-4 * g["spatial_abs_sum"]
    if kind == "pretrain":
        assert_close(bd["voxel_mae_mask"], obd["voxel_mae_mask"], 0, 0)
        assert_close(bd["voxel_features"], obd["voxel_features"].detach(), what="pyramid voxel features", **FEAT_TOL)
        r, orr = bb.forward_ret_dict, obb.forward_ret_dict
        assert_close(r["pred_points"], orr["pred_points"].detach(), what="pred_points", **FEAT_TOL)
        assert_close(bb.gt_points(), orr["gt_points"], 1e-6, 1e-6, "gt_points")
        assert_close(bb.gt_points()[::ROW], g["gt_points"], 1e-6, 1e-6, "golden gt_points")


def test_pretrain_loss_and_gradients(both):
    kind, g, (vfe, bb, av, bd), (ovfe, obb, oav, obd) = both
    if kind != "pretrain":
        pytest.skip("loss exists only in pretraining")
    loss, _ = bb.get_loss()
    oloss, _ = obb.get_loss()
    assert abs(loss.item() - oloss.item()) <= 1e-4 * abs(oloss.item())
    assert abs(loss.item() - g["loss"]) <= 1e-4 * abs(g["loss"])
    loss.backward()
    oloss.backward()
    g64 = torch.load(os.path.join(GOLDEN, "small_pretrain_grad64.pt"), weights_only=False)
    assert abs(loss.item() - g64["loss"]) <= 1e-6 * abs(g64["loss"])
    worst, worst32, rel = [], [], []
    for m, om, pre in ((vfe, ovfe, "vfe."), (bb, obb, "backbone_3d.")):
        op = dict(om.named_parameters())
        for k, p in m.named_parameters():
            assert p.grad is not None, k
            scale, sample = g64["grads"][pre + k]
            st = g64["stride"]
            e_p = (p.grad.flatten()[::st].cpu().double() - sample.double()).abs().max().item() / (scale + 1e-12)
            ref = op[k].grad
            e_o = (ref.flatten()[::st].double() - sample.double()).abs().max().item() / (scale + 1e-12)
            worst.append((e_p, pre + k))
            rel.append((e_p - 2 * e_o, e_p, e_o, pre + k))
            worst32.append(((p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12), pre + k))
            ga = p.grad.double().abs().sum().item()
            assert abs(ga - g["grad_abs_sum"][k]) <= 5e-3 * g["grad_abs_sum"][k] + 1e-7, k
    worst.sort(reverse=True), worst32.sort(reverse=True), rel.sort(reverse=True)
    # Gradient parity is pinned to the oracle evaluated in FLOAT64 (tests/golden/make_grad64.py): on this case the
    # fp32 oracle itself sits up to 6.2e-3 of a tensor's scale (median 2e-4) away from exact arithmetic -- 18 encoder
    # layers + 13 BatchNorms amplify rounding (tests/test_oracle.py::test_fp32_oracle_distance_from_float64_gradients)
    # -- so fp32-vs-fp32 agreement below that only measures how similar the op ORDER is.  Against float64: the median
    # tensor within 3e-4 of its scale (the fp32 oracle's median tensor: 1-2e-4), no tensor further than the fp32 oracle's own worst
    # (5e-3; the temperature scalars, sums of ~10^5 signed terms with heavy cancellation, 1e-2).
    oracle_worst = max(e_o for _, _, e_o, _ in rel)
    not_tau = [w for w in worst if not w[1].endswith(".tau")]
    assert not_tau[0][0] < max(5e-3, oracle_worst), not_tau[:5]
    assert worst[0][0] < 1e-2, worst[:5]
    oracle_median = sorted(e_o for _, _, e_o, _ in rel)[len(rel) // 2]   # 0.9e-4 on the sampled entries, 2e-4 on whole tensors
    assert worst[len(worst) // 2][0] < max(3e-4, 2 * oracle_median), (worst[len(worst) // 2], oracle_median)
    # and against the fp32 oracle: within the oracle's own distance from float64 (x3)
    assert worst32[0][0] < 2e-2, worst32[:5]


def test_siamese_batched_equals_separate():
    """Both frames through the shared SST blocks as one row set (default) vs. one pass per frame like the reference:
    same features, loss, gradients and BatchNorm running statistics."""
    pts, ptsp = cases.small_points(77, 900, 2)
    outs = []
    for batched in (True, False):
        vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
        cases.fill_params(vfe), cases.fill_params(bb)
        vfe.to(DEV), bb.to(DEV)
        bb.siamese_batched = batched
        bd = vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=2))
        bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), 2, 0.75, 5).to(DEV)
        bd = bb(bd)
        loss, _ = bb.get_loss()
        loss.backward()
        outs.append((bd["spatial_features"].detach(), loss.detach(), {k: p.grad for k, p in bb.named_parameters()},
                     {k: v for k, v in bb.state_dict().items() if "running" in k or "num_batches" in k}))
    (sf_a, l_a, g_a, st_a), (sf_b, l_b, g_b, st_b) = outs
    assert_close(sf_a, sf_b, 1e-5, 1e-5, "spatial_features")
    assert abs(l_a.item() - l_b.item()) <= 1e-5 * abs(l_b.item())
    for k in g_b:
        scale = g_b[k].abs().max().item() + 1e-12
        assert (g_a[k] - g_b[k]).abs().max().item() / scale < 1e-3, k
    for k in st_b:
        assert_close(st_a[k].float(), st_b[k].float(), 1e-5, 1e-6, k)


def test_side_stream_prepass_is_equivalent():
    """batch_dict["side_stream"]: the coordinate-only pre-pass on the library's side stream gives the same integers
    bit for bit and the same loss / features as the single-stream path, over several steps with the main stream busy."""
    pts, ptsp = cases.small_points(79, 900, 2)
    res = []
    for use_side in (False, True):
        vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
        cases.fill_params(vfe), cases.fill_params(bb)
        vfe.to(DEV), bb.to(DEV)
        bb.mask_generator = torch.Generator(device=DEV).manual_seed(7)
        side = ops.side_stream(DEV) if use_side else None
        outs = []
        for it in range(3):
            with torch.cuda.stream(side) if use_side else torch.cuda.stream(torch.cuda.current_stream()):
                p, pp = torch.from_numpy(pts).to(DEV, non_blocking=True), torch.from_numpy(ptsp).to(DEV, non_blocking=True)
            bd = dict(points=p, points_prev=pp, batch_size=2)
            if use_side:
                bd["side_stream"] = side
            bd = bb(vfe(bd))
            loss, _ = bb.get_loss()
            loss.backward()
            outs.append((bd["voxel_coords"].clone(), bd["voxel_mae_mask"].clone(), bd["spatial_features"].detach().clone(), loss.detach().clone()))
        torch.cuda.synchronize()
        res.append(outs)
    for (c0, m0, s0, l0), (c1, m1, s1, l1) in zip(*res):
        assert_equal_int(c0, c1, "voxel_coords"), assert_close(m0, m1, 0, 0, "mask")
        assert_close(s0, s1, 1e-5, 1e-5, "spatial_features")
        assert abs(l0.item() - l1.item()) <= 1e-5 * abs(l0.item())


def test_side_stream_tables_are_all_recorded(monkeypatch):
    """Every table the pre-pass allocates on the side stream is marked as used by the main stream (Tensor.record_stream), however
    deep it sits in the plan objects.  Regression: the window-partition tables of the pre-training path sat one level below the walk's
    old depth limit, returned to the side stream's pool as soon as the plan was dropped, and were handed to the NEXT step's pre-pass
    while this step's backward was still reading them (an illegal address / silently wrong gradients, but only with the host running
    ahead)."""
    recorded = set()
    orig = torch.Tensor.record_stream

    def spy(self, stream):
        recorded.add(self.untyped_storage().data_ptr())
        return orig(self, stream)
    monkeypatch.setattr(torch.Tensor, "record_stream", spy)
    pts, ptsp = cases.small_points(81, 900, 2)
    vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.to(DEV), bb.to(DEV)
    bb.mask_generator = torch.Generator(device=DEV).manual_seed(7)
    side = ops.side_stream(DEV)
    bd = dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=2, side_stream=side)
    bd = bb(vfe(bd))
    torch.cuda.synchronize()

    found = []

    def walk(obj, path, seen):
        if obj is None or isinstance(obj, (str, bytes, int, float, bool, torch.nn.Module)) or id(obj) in seen:
            return
        seen.add(id(obj))
        if isinstance(obj, torch.Tensor):
            if obj.is_cuda:
                found.append((path, obj))
        elif isinstance(obj, dict):
            for k, v in obj.items():
                walk(v, f"{path}.{k}", seen)
        elif isinstance(obj, (list, tuple)):
            for i, v in enumerate(obj):
                walk(v, f"{path}[{i}]", seen)
        elif hasattr(obj, "__slots__"):
            for k in obj.__slots__:
                walk(getattr(obj, k, None), f"{path}.{k}", seen)
        elif hasattr(obj, "__dict__"):
            for k, v in vars(obj).items():
                walk(v, f"{path}.{k}", seen)
    walk(bb.last_plan, "plan", set())
    assert len(found) > 50, "the walk must reach the partition tables"
    missing = [path for path, t in found if t.untyped_storage().data_ptr() not in recorded]
    assert not missing, f"side-stream tensors without record_stream: {missing[:8]}"
    names = " ".join(path for path, _ in found)
    assert ".part.tok_a" in names and ".tok_b" in names and ".subm" in names


def test_fused_decoder_batchnorm_matches_torch():
    """Throughput mode: the decoder's BatchNorm2d + ReLU + concat on the library's bf16 kernels against the same
    decoder with torch's BatchNorm2d / ReLU / cat under autocast (both bf16): features, loss, gradients and
    BatchNorm running statistics agree to bf16 rounding."""
    pts, ptsp = cases.small_points(78, 900, 2)
    outs = []
    for fused in (True, False):
        vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
        cases.fill_params(vfe), cases.fill_params(bb)
        vfe.to(DEV), bb.to(DEV)
        bb.decoder_autocast, bb.fused_decoder_bn = torch.bfloat16, fused
        bd = vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=2))
        bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), 2, 0.75, 5).to(DEV)
        bd = bb(bd)
        loss, _ = bb.get_loss()
        loss.backward()
        outs.append((bd["spatial_features"].detach().float(), loss.detach(), {k: p.grad for k, p in bb.named_parameters()},
                     {k: v for k, v in bb.state_dict().items() if "decoder" in k and ("running" in k or "num_batches" in k)}))
    (sf_a, l_a, g_a, st_a), (sf_b, l_b, g_b, st_b) = outs
    assert sf_a.shape == sf_b.shape
    assert_close(sf_a, sf_b, 3e-2, 3e-2, "spatial_features (bf16 decoder)")
    assert abs(l_a.item() - l_b.item()) <= 2e-2 * abs(l_b.item())
    for k in g_b:
        scale = g_b[k].abs().max().item() + 1e-12
        assert (g_a[k].float() - g_b[k].float()).abs().max().item() / scale < 8e-2, k
    for k in st_b:
        assert_close(st_a[k].float(), st_b[k].float(), 1e-2, 1e-3, k)


def test_eval_mode_forward():
    """BatchNorm in eval mode (running statistics) through the same kernels."""
    pts, ptsp = cases.small_points(31, 800, 2)
    vfe, bb, av, bd = run_product("finetune", pts, ptsp, 2, train=False)
    ovfe, obb, oav, obd = run_tier2("finetune", pts, ptsp, 2, train=False)
    assert_close(av["voxel_features"], oav["voxel_features"].detach(), 1e-5, 1e-5)
    assert_close(bd["spatial_features"], obd["spatial_features"].detach(), **FEAT_TOL)


def test_dynvfe_single_frame():
    pts, _ = cases.small_points(41, 700, 2)
    from oracle import restated
    cfg = tmae_b200.config.model_cfg("pretrain")["VFE"]
    v = tmae_b200.DynVFE(cfg, 4, S["voxel"], S["range"], S["grid"]).to(DEV)
    o = restated.DynVFE(cfg, 4, S["voxel"], S["range"], S["grid"])
    cases.fill_params(v), cases.fill_params(o)
    bd = v(dict(points=torch.from_numpy(pts).to(DEV), batch_size=2))
    obd = o(dict(points=torch.from_numpy(pts), batch_size=2))
    assert_equal_int(bd["voxel_coords"], obd["voxel_coords"])
    assert_close(bd["pillar_features"], obd["pillar_features"].detach(), 1e-5, 1e-5)
    assert bd["voxel_features"] is bd["pillar_features"]


def test_random_mask_statistics():
    """mask_voxels: per sample exactly int(L * 0.25) voxels stay visible (common_utils.py:54)."""
    pts, ptsp = cases.small_points(51, 900, 3)
    vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
    vfe.to(DEV), bb.to(DEV)
    bd = vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=3))
    mask, n_vis = bb.mask_voxels(bd["voxel_coords"], bd["voxels_per_sample"])
    b = bd["voxel_coords"][:, 0]
    for s, L in enumerate(bd["voxels_per_sample"]):
        assert int((mask[b == s] == 0).sum()) == int(L * 0.25)
    assert n_vis == int((mask == 0).sum())
    mask2, _ = bb.mask_voxels(bd["voxel_coords"], bd["voxels_per_sample"])
    assert not torch.equal(mask, mask2)


def test_cpu_tensors_are_rejected():
    with pytest.raises(RuntimeError):
        ops.add_pos(torch.zeros(4, 128), torch.zeros(4, dtype=torch.uint8), torch.zeros(64, 128))
