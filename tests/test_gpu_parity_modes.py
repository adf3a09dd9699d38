"""End-to-end parity of the THROUGHPUT modes (what bench.py measures) against the fp32 oracle: vfe -> backbone_3d.forward
-> Chamfer loss -> backward with `ops.set_precision("tf32" | "bf16")` and the bf16 cuDNN decoder, exactly the
configuration of bench.py (`--precision`, `--decoder bf16`), on the golden case (pinned to the reference's own modules by
tests/golden/make_golden.py), on one full ONCE-shaped scan pair and on one Waymo-shaped scan pair.

Tolerances.  `north_star` states rtol 1e-3 for reduced precision; SURVEY 7.3-6 notes that a pure rtol of 1e-3 is tighter
than one bf16 ulp (2^-8 = 3.9e-3), so the comparison is |a - b| <= atol * max|b| + rtol * |b| per tensor, with
  tf32 mode (fp32 storage, 10-bit operands; the decoder still runs bf16):  features rtol 1e-3, atol 1e-2 ; loss 2e-3 relative
  bf16 mode (bf16 storage, 8-bit mantissa through 18 encoder layers + 13 BatchNorms):  features rtol 1e-3, atol 4e-2 ;
             loss 1e-2 relative
The atol figures are set against a measured yardstick: the REFERENCE's own mixed-precision noise.  The reference trains
under fp16 autocast (tools/train_utils/train_utils.py:73-77); the tier-2 oracle run under torch.autocast on the same
inputs and weights moves `spatial_features` by 1.4e-2 of its scale (bf16 autocast, CPU) -- measured again here on the
GPU for fp16 and bf16 and written to gpurun_out/parity_modes.json next to the product's own distances.  Gradients are
compared with the float64 oracle gradients (tests/golden/small_pretrain_grad64.pt) relative to each tensor's scale.
Every run also asserts (tmae_dispatch_counts) that the tensor-core kernels -- not an fp32 fallback -- served the GEMMs.
"""
import json
import os

import pytest
import torch

from common import GOLDEN, assert_equal_int, golden_inputs, load_golden, run_tier2
from oracle import cases, restated

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200
    from tmae_b200 import ops, synth
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
S = cases.SMALL
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# mode -> feature rtol, feature atol relative to max|ref|, loss relative, gradient error (max over a tensor's sampled entries, relative to
# the tensor's scale, against the float64 oracle): worst tensor / median tensor.  Measured on B200 (gpurun_out/parity_modes.json, copied
# to profiles/r02_parity_modes.json) next to the REFERENCE's own mixed-precision noise on the same case:
#                          features (max err / scale)   loss rel   gradient worst / median
#   oracle, fp16 autocast       1.4e-3 .. 2.6e-3         3e-5          0.11 / 0.028      <- what the reference trains with
#   oracle, bf16 autocast       1.1e-2 .. 1.8e-2         1.5e-4        0.61 / 0.097
#   product, tf32 mode          1.0e-3 .. 5.5e-3         3e-5          0.17 / 0.049      (encoder 1e-3; the bf16 cuDNN decoder adds the rest)
#   product, bf16 mode          0.8e-2 .. 2.0e-2         3e-4          0.59 / 0.124
TOL = {
    "tf32": dict(rtol=1e-3, atol=1e-2, loss=2e-3, grad=0.3, grad_median=0.08),
    "bf16": dict(rtol=1e-3, atol=4e-2, loss=1e-2, grad=0.9, grad_median=0.2),
}
_report = {}


def _dist(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    scale = b.abs().max().item() + 1e-30
    return dict(max_over_scale=(a - b).abs().max().item() / scale, rms_rel=((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30)).item())


def _check(a, b, tol, what, rep):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, f"{what}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    rep[what] = _dist(a, b)
    err = (a - b).abs()
    lim = tol["atol"] * b.abs().max() + tol["rtol"] * b.abs()
    bad = err > lim
    assert not bad.any(), f"{what}: {int(bad.sum())}/{bad.numel()} outside rtol {tol['rtol']} + atol {tol['atol']} x scale; max err / scale = {rep[what]['max_over_scale']:.3e}"


def _save_report():
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_modes.json"), "w") as f:
            json.dump(_report, f, indent=1, sort_keys=True)


def _product(kind, mode, grid, voxel, rng, pts, ptsp, B, mask, npf=5, train=True):
    vfe, bb = tmae_b200.build_model(kind, grid, voxel, rng, num_point_features=npf)
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.to(DEV).train(train), bb.to(DEV).train(train)
    bb.decoder_autocast = torch.bfloat16   # bench.py:--decoder bf16
    before = ops.dispatch_counts()
    try:
        ops.set_precision(mode)
    except NotImplementedError as e:
        pytest.skip(str(e))
    try:
        bd = vfe(dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=B))
        if kind == "pretrain":
            bd["voxel_mae_mask_in"] = mask.to(DEV)
        bd = bb(bd)
        loss = None
        if kind == "pretrain":
            loss, _ = bb.get_loss()
            loss.backward()
    finally:
        ops.set_precision("fp32")
    after = ops.dispatch_counts()
    assert after["tma"] > before["tma"], "no TMA-fed tcgen05 GEMM ran in a tensor-core mode"
    assert after["simt_in_tc_mode"] == before["simt_in_tc_mode"], "a GEMM of the hot path fell back to the fp32 FFMA kernel"
    return vfe, bb, bd, loss


def _grad_errors(modules, g64):
    """Per parameter tensor: max |grad - float64 oracle grad| over the sampled entries, relative to the tensor's scale (max |grad64|);
    the temperature scalars share ONE scale, the largest temperature gradient of the model (several of them are sums of ~1e5 signed
    terms that cancel to 1e-2 of that: relative to themselves their error is meaningless in any arithmetic)."""
    st = g64["stride"]
    tau_scale = max(sc for k, (sc, _) in g64["grads"].items() if k.endswith(".tau"))
    errs = []
    for m, pre in modules:
        for k, p in m.named_parameters():
            assert p.grad is not None and torch.isfinite(p.grad).all(), k
            scale, sample = g64["grads"][pre + k]
            if k.endswith(".tau"):
                scale = tau_scale
            errs.append(((p.grad.float().flatten()[::st].cpu().double() - sample.double()).abs().max().item() / (scale + 1e-12), pre + k))
    errs.sort(reverse=True)
    return errs


def _oracle_autocast(kind, pts, ptsp, B, mask_seed, dtype, backward=False):
    """The tier-2 oracle under torch.autocast on the GPU (stock torch ops): the reference's own mixed-precision noise.  backward: with
    the reference's loss scaling for fp16 (GradScaler, tools/train_utils/train_utils.py:88-89; a fixed 2^10 here)."""
    vfe, bb = restated.build(kind, S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.to(DEV), bb.to(DEV)
    bd = dict(points=torch.from_numpy(pts).to(DEV), points_prev=torch.from_numpy(ptsp).to(DEV), batch_size=B)
    with torch.autocast("cuda", dtype=dtype):
        bd = vfe(bd)
        bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), B, 0.75, mask_seed).to(DEV)
        bd = bb(bd)
        loss, _ = bb.get_loss()
    if backward:
        sc = 1024.0 if dtype == torch.float16 else 1.0
        (loss * sc).backward()
        for p in list(vfe.parameters()) + list(bb.parameters()):
            if p.grad is not None:
                p.grad.div_(sc)
        return bd, loss, vfe, bb
    return bd, loss


@pytest.fixture(scope="module")
def golden_case():
    g = load_golden("pretrain")
    pts, ptsp = golden_inputs(g)
    B, ms = g["meta"]["batch"], g["meta"]["mask_seed"]
    ovfe, obb, oav, obd = run_tier2("pretrain", pts, ptsp, B, ms)
    oloss, _ = obb.get_loss()
    mask = cases.fixed_mask(oav["voxel_coords"], B, 0.75, ms)
    return g, pts, ptsp, B, ms, mask, (ovfe, obb, oav, obd, oloss)


def test_reference_mixed_precision_noise(golden_case):
    """Yardstick: distance of the oracle under fp16 / bf16 autocast from the fp32 oracle (same inputs, weights, mask)."""
    g, pts, ptsp, B, ms, mask, (ovfe, obb, oav, obd, oloss) = golden_case
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        try:
            bd, loss, avfe, abb = _oracle_autocast("pretrain", pts, ptsp, B, ms, dt, backward=True)
        except Exception as e:   # stock-torch op without an autocast rule on this build: record, the yardstick is informative only
            _report[f"oracle_autocast_{name}"] = {"error": repr(e)[:200]}
            continue
        rep = {"loss_rel": abs(loss.item() - oloss.item()) / abs(oloss.item()),
               "spatial_features": _dist(bd["spatial_features"].float(), obd["spatial_features"])}
        g64 = torch.load(os.path.join(GOLDEN, "small_pretrain_grad64.pt"), weights_only=False)
        errs = _grad_errors(((avfe, "vfe."), (abb, "backbone_3d.")), g64)
        rep["grad_worst"], rep["grad_median"] = errs[:5], errs[len(errs) // 2][0]
        for k, sp in bd["multi_scale_3d_features"].items():
            rep[k] = _dist(sp.features.float(), obd["multi_scale_3d_features"][k].features)
        _report[f"oracle_autocast_{name}"] = rep
    _save_report()


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_golden_case_in_throughput_mode(golden_case, mode):
    g, pts, ptsp, B, ms, mask, (ovfe, obb, oav, obd, oloss) = golden_case
    tol = TOL[mode]
    vfe, bb, bd, loss = _product("pretrain", mode, S["grid"], S["voxel"], S["range"], pts, ptsp, B, mask)
    rep = _report.setdefault(f"golden_{mode}", {})
    assert_equal_int(bd["voxel_coords"], oav["voxel_coords"], "voxel_coords")
    assert_equal_int(bd["voxel_coords"], g["voxel_coords"], "golden voxel_coords")
    for k, sp in bd["multi_scale_3d_features"].items():
        o = obd["multi_scale_3d_features"][k]
        assert_equal_int(sp.indices, o.indices, k + " indices")
        _check(sp.features.float(), o.features, tol, k, rep)
        _check(sp.features.float()[::3], g[k + "_features"], tol, "golden " + k, rep)
    _check(bd["spatial_features"].float(), obd["spatial_features"], tol, "spatial_features", rep)
    _check(bd["voxel_features"].float(), obd["voxel_features"], tol, "pyramid voxel features", rep)
    _check(bb.forward_ret_dict["pred_points"].float(), obb.forward_ret_dict["pred_points"], tol, "pred_points", rep)
    rep["loss_rel"] = abs(loss.item() - oloss.item()) / abs(oloss.item())
    assert rep["loss_rel"] <= tol["loss"], rep["loss_rel"]
    assert abs(loss.item() - g["loss"]) <= tol["loss"] * abs(g["loss"])
    # gradients against the float64 oracle (the fp32 oracle itself is up to 6e-3 of a tensor's scale away from it)
    g64 = torch.load(os.path.join(GOLDEN, "small_pretrain_grad64.pt"), weights_only=False)
    errs = _grad_errors(((vfe, "vfe."), (bb, "backbone_3d.")), g64)
    rep["grad_worst"] = errs[:5]
    rep["grad_median"] = errs[len(errs) // 2][0]
    _save_report()
    assert errs[0][0] <= tol["grad"], errs[:5]
    assert rep["grad_median"] <= tol["grad_median"], rep["grad_median"]


def test_bf16_mode_forward_follows_fused_optimizer_steps(golden_case):
    """Training in the bench configuration (bf16 mode, torch.optim.AdamW(fused=True)): the forward must see every optimizer step.
    Regression: the fused optimizer does not move Tensor._version, so the bf16 weight shadows (keyed on it) stayed at their step-1
    values for the rest of the run -- same timing, silently wrong training.  After three steps the trained modules and a FRESH pair
    loaded from their state_dict must produce bit-identical eval outputs, and the loss must have moved."""
    g, pts, ptsp, B, ms, mask, _ = golden_case
    P = torch.from_numpy(pts).to(DEV), torch.from_numpy(ptsp).to(DEV)

    def build(state=None):
        vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
        if state is None:
            cases.fill_params(vfe), cases.fill_params(bb)
        else:
            vfe.load_state_dict(state[0]), bb.load_state_dict(state[1])
        vfe.to(DEV), bb.to(DEV)
        bb.decoder_autocast = torch.bfloat16
        return vfe, bb

    def run(vfe, bb):
        bd = vfe(dict(points=P[0], points_prev=P[1], batch_size=B))
        bd["voxel_mae_mask_in"] = mask.to(DEV)
        bd = bb(bd)
        return bd, bb.get_loss()[0]

    ops.set_precision("bf16")
    try:
        vfe, bb = build()
        vfe.train(), bb.train()
        opt = torch.optim.AdamW(list(vfe.parameters()) + list(bb.parameters()), lr=2e-3, weight_decay=0.01, fused=True)
        losses = []
        for _ in range(3):
            _, loss = run(vfe, bb)
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            losses.append(loss.item())
        assert abs(losses[2] - losses[0]) > 1e-3 * abs(losses[0]), f"the loss did not move over three optimizer steps: {losses}"
        vfe.eval(), bb.eval()
        with torch.no_grad():
            bd1, l1 = run(vfe, bb)
        state = ({k: v.detach().cpu().clone() for k, v in vfe.state_dict().items()}, {k: v.detach().cpu().clone() for k, v in bb.state_dict().items()})
        vfe2, bb2 = build(state)
        vfe2.eval(), bb2.eval()
        with torch.no_grad():
            bd2, l2 = run(vfe2, bb2)
        assert torch.equal(bd1["spatial_features"], bd2["spatial_features"]), "trained modules and a fresh copy of their weights disagree: stale derived weight operands"
        assert l1.item() == l2.item()
    finally:
        ops.set_precision("fp32")


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("shape_name,n_points,npf", [("once", 60000, 5), ("waymo", 180000, 6)])
def test_full_scan_pair_in_throughput_mode(shape_name, n_points, npf, mode):
    """configs[1] / configs[4] shapes at full size, one scan pair, against the fp32 oracle on the host."""
    shape = synth.SHAPES[shape_name]
    grid = synth.grid_size(shape).tolist()
    pts, ptsp = synth.batch(1000, 1, n_points, shape_name)
    ovfe, obb = restated.build("pretrain", grid, shape["voxel"], shape["range"], num_point_features=npf)
    cases.fill_params(ovfe), cases.fill_params(obb)
    obd = ovfe(dict(points=torch.from_numpy(pts), points_prev=torch.from_numpy(ptsp), batch_size=1))
    mask = cases.fixed_mask(obd["voxel_coords"], 1, 0.75, 3)
    obd["voxel_mae_mask_in"] = mask
    obd = obb(obd)
    oloss, _ = obb.get_loss()
    tol = TOL[mode]
    vfe, bb, bd, loss = _product("pretrain", mode, grid, shape["voxel"], shape["range"], pts, ptsp, 1, mask, npf)
    rep = _report.setdefault(f"full_{shape_name}_{mode}", {})
    assert_equal_int(bd["voxel_coords"], obd["voxel_coords"], "voxel_coords")
    for k, sp in bd["multi_scale_3d_features"].items():
        o = obd["multi_scale_3d_features"][k]
        assert_equal_int(sp.indices, o.indices, k + " indices")
        _check(sp.features.float(), o.features, tol, k, rep)
    _check(bd["spatial_features"].float(), obd["spatial_features"], tol, "spatial_features", rep)
    _check(bb.forward_ret_dict["pred_points"].float(), obb.forward_ret_dict["pred_points"], tol, "pred_points", rep)
    rep["loss_rel"] = abs(loss.item() - oloss.item()) / abs(oloss.item())
    _save_report()
    assert rep["loss_rel"] <= tol["loss"], rep["loss_rel"]
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in list(vfe.parameters()) + list(bb.parameters()))


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_finetune_forward_in_throughput_mode(mode):
    """configs[3] path (finetune forward, 5 drop levels, temporal cross-attention on every voxel) on the golden case."""
    g = load_golden("finetune")
    pts, ptsp = golden_inputs(g)
    B = g["meta"]["batch"]
    ovfe, obb, oav, obd = run_tier2("finetune", pts, ptsp, B, train=False)
    tol = TOL[mode]
    vfe, bb, bd, _ = _product("finetune", mode, S["grid"], S["voxel"], S["range"], pts, ptsp, B, None, train=False)
    rep = _report.setdefault(f"finetune_{mode}", {})
    for k, sp in bd["multi_scale_3d_features"].items():
        _check(sp.features.float(), obd["multi_scale_3d_features"][k].features, tol, k, rep)
    _check(bd["spatial_features"].float(), obd["spatial_features"], tol, "spatial_features", rep)
    _save_report()
