"""GPU tests of the bf16-STORAGE kernels (gemm_bf16.cu, rowops_bf16.cu, layer_bf16.cu) against float64 references
computed on the SAME bf16-rounded operands: products of bf16 values are exact in fp32 and the accumulator is fp32 (TMEM),
so a result differs from float64 only by the accumulation order and by the final rounding of the OUTPUT to bf16
(half an ulp = 2^-9 = 2e-3 relative) -- tolerance rtol 4e-3 + a small atol for bf16 outputs, 1e-4-class for fp32 outputs
(weight gradients, statistics).  Every GEMM test asserts through tmae_dispatch_counts that the tcgen05 kernel ran."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from common import assert_close
from oracle import restated

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200  # noqa: F401
    from tmae_b200 import ops

DEV = "cuda"
BF = torch.bfloat16
RT, AT = 4e-3, 2e-3


def rb(t):
    """fp32 -> bf16-rounded values as float64."""
    return t.to(BF).double()


def dev_bf(t):
    return t.to(BF).to(DEV).contiguous()


@pytest.fixture(autouse=True)
def _counts():
    before = ops.dispatch_counts()
    yield
    after = ops.dispatch_counts()
    assert after["simt_in_tc_mode"] == before["simt_in_tc_mode"]


def test_casts_and_shadows():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1000, 37, generator=g)
    y = ops.cast_bf16(x.to(DEV))
    assert torch.equal(y.cpu(), x.to(BF))
    assert torch.equal(ops.cast_f32(y).cpu(), x.to(BF).float())
    ps = [torch.nn.Parameter(torch.randn(n, generator=g).to(DEV)) for n in (5, 64, 1000, 12345)]
    sh = ops.WeightShadows()
    sh.register(ps)
    sh.refresh()
    for p in ps:
        assert torch.equal(sh.get(p), p.detach().to(BF))
    with torch.no_grad():
        ps[2].add_(1.0)          # in-place update bumps _version: exactly that shadow is refreshed
    assert torch.equal(sh.get(ps[2]), ps[2].detach().to(BF))
    with torch.no_grad():
        for p in ps:
            p.mul_(0.5)
    sh.refresh()
    for p in ps:
        assert torch.equal(sh.get(p), p.detach().to(BF))


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (1, 128, 64), (1000, 128, 128), (333, 256, 128), (2049, 256, 512), (4100, 512, 256),
                                   (5000, 384, 128), (777, 768, 256), (130, 64, 96)])
def test_bf16_linear_fwd_bwd(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5, torch.randn(n, generator=g)
    xd, wd, bd = dev_bf(x), dev_bf(w), b.to(DEV)
    before = ops.dispatch_counts()["tma"]
    lin = rb(x) @ rb(w).T + b.double()
    for act, f in ((ops.ACT_NONE, lambda t: t), (ops.ACT_GELU, F.gelu), (ops.ACT_RELU, F.relu)):
        y, pre = ops.bf16_linear_fwd(xd, wd, bd, act=act, want_preact=True)
        assert y.dtype == BF and pre.dtype == BF
        assert_close(pre.float(), lin, RT, AT, f"preact m={m}")
        assert_close(y.float(), f(lin), RT, AT, f"linear act={act}")
    # GELU with the derivative in place of the pre-activation copy, and the backward that multiplies by it
    lin_g = lin.clone().requires_grad_()
    F.gelu(lin_g).sum().backward()
    y, der = ops.bf16_linear_fwd(xd, wd, bd, act=ops.ACT_GELU_DERIV, want_preact=True)
    assert_close(y.float(), F.gelu(lin), RT, AT, "gelu (derivative-saving form)")
    assert_close(der.float(), lin_g.grad, RT, AT, "gelu'")
    y0 = ops.bf16_linear_fwd(xd, wd, None)
    y1 = ops.bf16_linear_fwd(xd, wd, bd, out=y0.clone(), accumulate=True)
    assert_close(y1.float(), y0.double().cpu() + lin, RT, 2 * AT, "C +=")
    dy = torch.randn(m, n, generator=g)
    dyd = dev_bf(dy)
    dx = ops.bf16_linear_bwd_data(dyd, wd)
    assert_close(dx.float(), rb(dy) @ rb(w), RT, AT, "dx")
    pre = torch.randn(m, k, generator=g)
    p = rb(pre).requires_grad_()
    F.gelu(p).backward(rb(dy) @ rb(w))
    dxg = ops.bf16_linear_bwd_data(dyd, wd, gelu_pre=dev_bf(pre))
    assert_close(dxg.float(), p.grad, RT, AT, "dx * gelu'")
    der = dev_bf(torch.rand(m, k, generator=g) * 1.2 - 0.1)
    dxd = ops.bf16_linear_bwd_data(dyd, wd, gelu_pre=der, pre_is_derivative=True)
    assert_close(dxd.float(), (rb(dy) @ rb(w)) * der.double().cpu(), RT, AT, "dx * saved gelu'")
    dx2 = ops.bf16_linear_bwd_data(dyd, wd, dx=dx.clone(), accumulate=True)
    assert_close(dx2.float(), dx.double().cpu() + rb(dy) @ rb(w), RT, 2 * AT, "dx accumulate")
    dw = ops.bf16_linear_bwd_weight(dyd, xd)
    assert dw.dtype == torch.float32
    assert_close(dw, rb(dy).T @ rb(x), 1e-4, 1e-4 * max(1, m) ** .5, "dw")
    assert_close(ops.bf16_colsum(dyd), rb(dy).sum(0), 1e-5, 1e-4 * max(1, m) ** .5, "colsum")
    if k % 64 == 0:   # the same GEMM with one-hot extra columns: dW and the binned column sums (position-table gradient) at once
        pi = torch.randint(0, 64, (m,), generator=g, dtype=torch.uint8)
        oh = ops.onehot64_bf16(pi.to(DEV))
        assert torch.equal(oh.float().cpu(), F.one_hot(pi.long(), 64).float())
        dw2, dt = ops.bf16_linear_bwd_weight(dyd, xd, onehot=oh)
        assert_close(dw2, rb(dy).T @ rb(x), 1e-4, 1e-4 * max(1, m) ** .5, "dw (one-hot form)")
        ref = torch.zeros(64, n, dtype=torch.float64).index_add_(0, pi.long(), rb(dy))
        assert_close(dt, ref.T, 1e-4, 1e-4 * max(1, m) ** .5, "dtable^T")
    assert ops.dispatch_counts()["tma"] > before


@pytest.mark.parametrize("m,c,parts,hd", [(3000, 128, 3, 16), (1700, 256, 3, 32), (900, 128, 1, 16), (5, 256, 2, 32), (129, 128, 2, 16)])
def test_bf16_qkv_projection(m, c, parts, hd):
    """Packed projection + position table + per-head normalisation of the q / k columns, against the reference op order
    (x + pos) W^T + b -> F.normalize (cosine_msa.py:57-62,151-152) in float64."""
    gen = torch.Generator().manual_seed(m + c)
    n = parts * c
    n_pos = min(2, parts) * c if parts != 1 else c
    norm_cols = n_pos                       # q and k columns (cross kv: only k)
    x = torch.randn(m, c, generator=gen)
    w = torch.randn(n, c, generator=gen) / c ** .5
    b = torch.randn(n, generator=gen) * 0.1
    lut = torch.randn(64, c, generator=gen)
    pi = torch.randint(0, 64, (m,), generator=gen, dtype=torch.uint8)
    table, _ = ops.pos_table(lut.to(DEV), w.to(DEV), b.to(DEV), n_pos)
    y, inv = ops.bf16_qkv_fwd(dev_bf(x), dev_bf(w), table, pi.to(DEV), norm_cols, hd)
    ref = rb(x) @ rb(w).T + table.double().cpu()[pi.long()]
    H = norm_cols // hd
    heads = ref[:, :norm_cols].reshape(m, H, hd)
    nrm = heads.norm(dim=-1).clamp_min(1e-12)
    ref[:, :norm_cols] = (heads / nrm[..., None]).reshape(m, norm_cols)
    assert_close(y.float(), ref, RT, AT, "qkv")
    assert_close(inv, 1 / nrm, 1e-4, 1e-6, "1 / |.|")
    # the form the fused layer uses: the table row arrives through the MMA as a one-hot second A operand, table rounded to bf16
    if c % 64 == 0:
        wcat = ops.bf16_qkv_wcat(lut.to(DEV), w.to(DEV), b.to(DEV), n_pos)
        assert torch.equal(wcat[:, :c].cpu(), w.to(BF)), "[W | table]: weight part"
        assert_close(wcat[:, c:].float().T, table.double().cpu(), 1e-2, 1e-2, "[W | table]: table part")
        y2, inv2 = ops.bf16_qkv_fwd_onehot(dev_bf(x), ops.onehot64_bf16(pi.to(DEV)), wcat, norm_cols, hd)
        ref2 = rb(x) @ rb(w).T + wcat[:, c:].double().cpu().T[pi.long()]
        h2 = ref2[:, :norm_cols].reshape(m, H, hd)
        n2 = h2.norm(dim=-1).clamp_min(1e-12)
        ref2[:, :norm_cols] = (h2 / n2[..., None]).reshape(m, norm_cols)
        assert_close(y2.float(), ref2, RT, AT, "qkv (one-hot form)")
        assert_close(inv2, 1 / n2, 1e-4, 1e-6, "1 / |.| (one-hot form)")
        assert_close(y2.float(), y.double().cpu(), 2e-2, 2e-2, "one-hot form vs table-epilogue form")
    # sanity against the un-rounded reference formulation
    full = (x.double() + torch.cat([lut.double()[pi.long()]] * 1, 1)) @ w.double()[:n_pos].T + b.double()[:n_pos]
    fh = full.reshape(m, H, hd)
    fh = fh / fh.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    assert (y.float().cpu()[:, :norm_cols].double() - fh.reshape(m, norm_cols)).abs().max() < 3e-2


def test_shadows_follow_a_fused_optimizer_step():
    """torch.optim.AdamW(fused=True) updates parameters WITHOUT moving Tensor._version: the bf16 shadows and the [W | table] operands must
    still follow (global optimizer post-step hook -> ops._weights_epoch), and so after ops.weights_changed() for `p.data` arithmetic."""
    g = torch.Generator().manual_seed(3)
    w = torch.nn.Parameter((torch.randn(384, 128, generator=g) / 11).to(DEV))
    b = torch.nn.Parameter((torch.randn(384, generator=g) * 0.1).to(DEV))
    lut = torch.randn(64, 128, generator=g).to(DEV)
    sh, qo = ops.WeightShadows(), ops.QkvOperands()
    assert torch.equal(sh.get(w), w.detach().to(BF))
    first = qo.get(w, b, lut).clone()
    opt = torch.optim.AdamW([w, b], lr=0.05, fused=True)
    w.grad, b.grad = torch.randn_like(w), torch.randn_like(b)
    v0 = w._version
    opt.step()
    assert torch.equal(sh.get(w), w.detach().to(BF)), f"stale bf16 shadow after a fused optimizer step (_version {v0} -> {w._version})"
    got = qo.get(w, b, lut)
    assert torch.equal(got, ops.bf16_qkv_wcat(lut, w.detach(), b.detach(), 256)) and not torch.equal(got, first)
    w.data.mul_(2.0)                      # invisible to the version counter
    ops.weights_changed()
    assert torch.equal(sh.get(w), w.detach().to(BF))
    assert torch.equal(qo.get(w, b, lut), ops.bf16_qkv_wcat(lut, w.detach(), b.detach(), 256))


def test_derived_weight_caches_drop_dead_models():
    """The bf16 shadows and [W | table] operands hold their parameters by weak reference: a dropped model takes its copies with it and
    is never refreshed again."""
    import gc
    sh, qo = ops.WeightShadows(), ops.QkvOperands()
    keep_w = torch.nn.Parameter(torch.randn(64, 32, device=DEV))
    w = torch.nn.Parameter(torch.randn(384, 128, device=DEV) / 11)
    b = torch.nn.Parameter(torch.randn(384, device=DEV) * 0.1)
    lut = torch.randn(64, 128, device=DEV)
    sh.get(w), sh.get(keep_w), qo.get(w, b, lut)
    assert len(sh.items) == 2 and len(qo.items) == 1
    del w, b
    gc.collect()
    sh.refresh(force=True), qo.refresh(force=True)
    assert len(sh.items) == 1 and len(qo.items) == 0
    assert torch.equal(sh.get(keep_w), keep_w.detach().to(BF))


def test_qkv_operands_cache_rebuilds_all_stale_in_one_launch():
    """ops.QkvOperands: the per-layer [W | table^T] operands equal the single-layer builder, are rebuilt when a master changes in place
    (an optimizer step) and not otherwise."""
    g = torch.Generator().manual_seed(5)
    qo = ops.QkvOperands()
    layers = []
    for c in (128, 256, 128):
        w = (torch.randn(3 * c, c, generator=g) / c ** .5).to(DEV)
        b = (torch.randn(3 * c, generator=g) * 0.1).to(DEV)
        lut = torch.randn(64, c, generator=g).to(DEV)
        layers.append((w, b, lut))
    first = [qo.get(*l) for l in layers]
    for (w, b, lut), got in zip(layers, first):
        assert torch.equal(got, ops.bf16_qkv_wcat(lut, w, b, 2 * w.shape[1]))
    calls = ops.launch_count()
    again = [qo.get(*l) for l in layers]
    assert ops.launch_count() == calls and all(a.data_ptr() == b_.data_ptr() for a, b_ in zip(first, again)), "unchanged masters: no launch"
    for w, b, _ in layers:
        w.mul_(1.5), b.add_(0.25)
    calls = ops.launch_count()
    third = [qo.get(*l) for l in layers]
    assert ops.launch_count() == calls + 1, "all stale operands in ONE launch"
    for (w, b, lut), got in zip(layers, third):
        assert torch.equal(got, ops.bf16_qkv_wcat(lut, w, b, 2 * w.shape[1]))


@pytest.mark.parametrize("m,n,k,masked", [(1000, 128, 128, False), (4097, 256, 256, False), (777, 128, 256, True), (300, 256, 512, True), (1, 128, 128, False)])
def test_bf16_linear_residual_layernorm(m, n, k, masked):
    g = torch.Generator().manual_seed(m + n)
    a, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5, torch.randn(n, generator=g) * 0.1
    res = torch.randn(m, n, generator=g)
    gamma, beta = 1 + 0.1 * torch.randn(n, generator=g), 0.1 * torch.randn(n, generator=g)
    mask = (torch.rand(m, generator=g) < 0.6).to(torch.uint8) if masked else None
    y, v, mean, rstd = ops.bf16_linear_ln_fwd(dev_bf(a), dev_bf(w), b.to(DEV), dev_bf(res), mask.to(DEV) if masked else None, gamma.to(DEV),
                                              beta.to(DEV), 1e-5)
    branch = rb(a) @ rb(w).T + b.double()
    if masked:
        branch = branch * mask.double()[:, None]
    vref = rb(res) + branch
    yref = F.layer_norm(vref, (n,), gamma.double(), beta.double(), 1e-5)
    assert_close(v.float(), vref, RT, AT, "pre-norm sum")
    assert_close(y.float(), yref, RT, 2 * AT, "LayerNorm output")
    assert_close(mean[:m], vref.mean(1), 1e-4, 1e-5, "mean")
    assert_close(rstd[:m], 1 / (vref.var(1, unbiased=False) + 1e-5).sqrt(), 1e-3, 1e-5, "rstd")
    # backward from the saved (bf16) pre-norm sum
    dy = torch.randn(m, n, generator=g)
    vs = v.float().cpu().double().requires_grad_()
    gd = gamma.double().requires_grad_()
    bd_ = beta.double().requires_grad_()
    F.layer_norm(vs, (n,), gd, bd_, 1e-5).backward(rb(dy))
    dv, dres, dg, db, dc = ops.bf16_layernorm_bwd(dev_bf(dy), v, mask.to(DEV) if masked else None, gamma.to(DEV), mean, rstd, want_dres=masked,
                                                  want_colsum=True)
    assert_close(dv.float(), vs.grad, RT, AT, "dv")
    s = m ** .5
    assert_close(dg, gd.grad, 2e-3, 2e-3 * s, "dgamma"), assert_close(db, bd_.grad, 1e-4, 1e-4 * s, "dbeta")
    if masked:
        assert_close(dres.float(), vs.grad * mask.double()[:, None], RT, AT, "dres")
        assert_close(dc, (vs.grad * mask.double()[:, None]).sum(0), 5e-3, 5e-3 * s, "colsum(dres)")
    else:
        assert_close(dc, vs.grad.sum(0), 5e-3, 5e-3 * s, "colsum(dv)")


@pytest.mark.parametrize("rows,n", [(1, 128), (5000, 384), (777, 768), (3, 256)])
def test_bf16_binned_colsum(rows, n):
    g = torch.Generator().manual_seed(rows)
    dy = torch.randn(rows, n, generator=g)
    pi = torch.randint(0, 64, (rows,), generator=g, dtype=torch.uint8)
    t = ops.bf16_binned_colsum(dev_bf(dy), pi.to(DEV))
    ref = torch.zeros(64, n, dtype=torch.float64).index_add_(0, pi.long(), rb(dy))
    assert_close(t, ref, 1e-4, 1e-4 * rows ** .5)


def _coords(seed, m, B, g):
    rng = np.random.default_rng(seed)
    cells = np.sort(rng.choice(B * g * g, size=min(m, B * g * g), replace=False))
    return torch.tensor(np.stack([cells // (g * g), (cells % (g * g)) // g, cells % g], 1), dtype=torch.int32)


@pytest.mark.parametrize("seed,m,B,g,cin,cout", [(0, 1500, 2, 96, 128, 128), (1, 700, 2, 47, 128, 256), (2, 3000, 1, 90, 256, 256)])
def test_bf16_sparse_conv(seed, m, B, g, cin, cout):
    c = _coords(seed, m, B, g)
    m = c.shape[0]
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(m, cin, generator=gen)
    for subm in (True, False):
        conv = restated.SparseConv(cin, cout, 3, 1 if subm else 2, 1, subm).double()
        w32 = conv.weight.detach().float()
        with torch.no_grad():
            conv.weight.copy_(rb(w32))
        xr = rb(x).requires_grad_()
        out = conv(restated.SparseTensor(xr, c, [g, g], B))
        if subm:
            table = ops.subm_table(c.to(DEV), B, g, g)
            table_t, flip, rows_out = table, True, m
        else:
            idx_out, n_out, table, table_t, _ = ops.strided_table(c.to(DEV), B, g, g)
            rows_out = int(n_out)
            table, flip = table[:rows_out], False
        y = ops.bf16_sparse_conv_fwd(dev_bf(x), table, dev_bf(w32), rows_out)
        scale = (9 * cin) ** .5 * 0.1
        assert_close(y.float(), out.features.detach(), RT, AT * scale, f"sparse conv fwd subm={subm}")
        dy = torch.randn(rows_out, cout, generator=gen)
        out.features.backward(rb(dy))
        dx = ops.bf16_sparse_conv_fwd(dev_bf(dy), table_t, ops.transpose_taps_bf16(w32.to(DEV), flip), m)
        assert_close(dx.float(), xr.grad, RT, AT * scale, f"sparse conv dx subm={subm}")
        dw = ops.bf16_sparse_conv_bwd_weight(dev_bf(dy), dev_bf(x), table, w32.shape)
        assert_close(dw, conv.weight.grad, 1e-4, 1e-4 * m ** .5, f"sparse conv dw subm={subm}")


def test_b16_dense_moves():
    c = _coords(3, 900, 2, 40)
    x = torch.randn(c.shape[0], 128)
    d = ops.densify_nhwc_b16(dev_bf(x), c.to(DEV), 2, 40, 40)
    ref = torch.zeros(2, 40, 40, 128, dtype=BF)
    ref[c[:, 0].long(), c[:, 1].long(), c[:, 2].long()] = x.to(BF)
    assert torch.equal(d.cpu(), ref)
    assert torch.equal(ops.gather_nhwc_b16(d, c.to(DEV)).cpu(), x.to(BF))


@pytest.mark.parametrize("C,cross", [(128, False), (256, False), (128, True), (256, True)])
@pytest.mark.parametrize("impl", [0, 1])
def test_bf16_encoder_layer_matches_fp32_layer(C, cross, impl):
    """One whole encoder layer, forward and backward, in the bf16-storage mode against the SAME layer in the fp32 parity mode
    (itself pinned to the oracle by test_gpu_ops.py / test_gpu_e2e.py): outputs within bf16 rounding of the layer's scale,
    every parameter gradient within 3 % of its scale.  impl 0 = cast bridge to the fp32-I/O attention kernels, 1 = tcgen05."""
    from tmae_b200 import backbone, config
    from tmae_b200.plan import _levels
    torch.manual_seed(C + cross)
    H, FF, B, g = 8, 2 * C, 2, 64
    cfg = config.model_cfg("pretrain")["BACKBONE_3D"]["SST_BLOCK_LIST"][0]
    levels = _levels(cfg["PREPROCESS"])
    layer = backbone.EncoderLayer(C, H, FF, cfg["ENCODER"]["LAYER_CFG"], cross).to(DEV)
    with torch.no_grad():
        at = layer.win_attn.cross_attn if cross else layer.win_attn.self_attn
        at.in_proj_bias.normal_(0, 0.1), at.tau.fill_(0.4)
    lut = backbone.pos_embed_table(C, 10000, normalize=False).to(DEV)
    ca = _coords(0, 2500 if not cross else 700, B, g)
    cb = _coords(5, 2200, B, g) if cross else None
    P = ops.window_partition(ca.to(DEV), B, g, g, levels, coords_b=cb.to(DEV) if cross else None)
    if cross:
        P.keep_a = (P.win_a >= 0).to(torch.uint8)
        P.keep_b = (P.win_b >= 0).to(torch.uint8)
    x = torch.randn(ca.shape[0], C, device=DEV)
    xp = torch.randn(cb.shape[0], C, device=DEV) if cross else None
    dy = torch.randn(ca.shape[0], C, device=DEV)
    outs = {}
    for mode in ("fp32", "bf16"):
        ops.set_precision(mode)
        if mode == "bf16":
            if impl == 1 and not ops.attention_tc_available():
                ops.set_precision("fp32")
                pytest.skip("tcgen05 attention kernel not built")
            ops.set_attention_impl(impl)
        try:
            layer.zero_grad()
            xi = (x.to(BF) if mode == "bf16" else x.clone()).requires_grad_()
            xpi = None if not cross else (xp.to(BF) if mode == "bf16" else xp.clone()).requires_grad_()
            y = layer.forward_cross(xi, xpi, P, 1, lut) if cross else layer.forward_self(xi, P, 1, lut)
            y.backward(dy.to(y.dtype))
            outs[mode] = (y.detach().float(), xi.grad.float(), None if not cross else xpi.grad.float(), {k: p.grad.clone() for k, p in layer.named_parameters()})
        finally:
            ops.set_precision("fp32")
    (y0, dx0, dk0, g0), (y1, dx1, dk1, g1) = outs["fp32"], outs["bf16"]

    def rel(a, b):
        return (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
    assert rel(y1, y0) < 2e-2, ("y", rel(y1, y0))
    assert rel(dx1, dx0) < 4e-2, ("dx", rel(dx1, dx0))
    if cross:
        assert rel(dk1, dk0) < 4e-2, ("dx_kv", rel(dk1, dk0))
    for k in g0:
        assert rel(g1[k], g0[k]) < (8e-2 if k.endswith("tau") else 3e-2), (k, rel(g1[k], g0[k]))


def _attn_ref64(q, k, v, P, shift, cross, tau, H):
    """float64 window attention on GIVEN unit vectors (no re-normalisation): per window softmax(q k^T / tau) v over the valid keys."""
    ma, C = q.shape
    hd = C // H
    nw = int(P.n_win[shift])
    qt, qc = P.tok_a[shift].cpu().view(-1, 64), P.cnt_a[shift].cpu()
    kt, kc = (P.tok_b[shift].cpu().view(-1, 64), P.cnt_b[shift].cpu()) if cross else (qt, qc)
    o = torch.zeros(ma, C, dtype=torch.float64)
    lse = torch.zeros(ma, H, dtype=torch.float64)
    for w in range(nw):
        qi, ki = qt[w, :qc[w]].long(), kt[w, :kc[w]].long()
        if len(qi) == 0 or len(ki) == 0:
            continue
        Q, K, V = q[qi].view(-1, H, hd), k[ki].view(-1, H, hd), v[ki].view(-1, H, hd)
        S = torch.einsum("qhd,khd->hqk", Q, K) / tau
        o[qi] = torch.einsum("hqk,khd->qhd", S.softmax(-1), V).reshape(len(qi), C)
        lse[qi] = torch.logsumexp(S, -1).T
    return o, lse


@pytest.mark.parametrize("small_warps", [1, 0])
@pytest.mark.parametrize("C,cross", [(128, False), (256, False), (128, True), (256, True)])
def test_bf16_window_attention_tcgen05(C, cross, small_warps):
    """small_warps = 1 (default): <= 16-token windows on the warp kernels, 17..64-token windows on tcgen05 tiles; 0: all on tcgen05."""
    ops.set_option("attn_small_warps", small_warps)
    try:
        _attention_tc_case(C, cross)
    finally:
        ops.set_option("attn_small_warps", 1)


def _attention_tc_case(C, cross):
    from tmae_b200 import config
    from tmae_b200.plan import _levels
    H, B, g = 8, 2, 64
    hd = C // H
    levels = _levels(config.model_cfg("pretrain")["BACKBONE_3D"]["SST_BLOCK_LIST"][0]["PREPROCESS"])
    ca = _coords(0, 2500 if not cross else 600, B, g)
    cb = _coords(5, 2200, B, g) if cross else None
    P = ops.window_partition(ca.to(DEV), B, g, g, levels, coords_b=cb.to(DEV) if cross else None)
    gen = torch.Generator().manual_seed(C)
    ma, mb = ca.shape[0], (cb.shape[0] if cross else ca.shape[0])

    def unit(t):
        t = t.view(-1, H, hd)
        return (t / t.norm(dim=-1, keepdim=True)).reshape(-1, C)
    # packed layouts as the layers use them: self (m, 3C) = [q | k | v]; cross q (mq, C), kv (mkv, 2C)
    q, k, v = unit(torch.randn(ma, C, generator=gen)), unit(torch.randn(mb, C, generator=gen)), torch.randn(mb, C, generator=gen)
    if cross:
        qd = dev_bf(q)
        kvd = dev_bf(torch.cat([k, v], 1))
        kd, vd = kvd[:, :C], kvd[:, C:]
    else:
        qkv = dev_bf(torch.cat([q, k, v], 1))
        qd, kd, vd = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    tau = torch.tensor([[[0.37]]])
    for shift in (0, 1):
        qt, qc = P.tok_a[shift], P.cnt_a[shift]
        kt, kc = (P.tok_b[shift], P.cnt_b[shift]) if cross else (qt, qc)
        o, lse = ops.bf16_window_attention_fwd(qd, kd, vd, qt, qc, kt, kc, P.n_win[shift:shift + 1], ops.small_end(P, shift), ops.mid_end(P, shift),
                                               min(P.wcap, ma), tau.to(DEV), 0.01, H, zero_out=True)
        qh, kh, vh = rb(q).requires_grad_(), rb(k).requires_grad_(), rb(v).requires_grad_()
        tau64 = torch.tensor(0.37, dtype=torch.float64, requires_grad=True)
        oref, lref = _attn_ref64(qh, kh, vh, P, shift, cross, tau64, H)
        assert_close(o.float(), oref.detach(), RT, AT * 2, f"attention output shift {shift}")     # P is rounded to bf16 before P V
        assert_close(lse, lref.detach(), 1e-4, 1e-4, f"log-sum-exp shift {shift}")
        # backward: gradients wrt the un-normalised projections, through d(x / |x|) = (g - u (u . g)) / |x| per head with the given 1 / |x|
        do = torch.randn(ma, C, generator=gen)
        oref.backward(rb(do))
        inv_q, inv_k = 0.5 + torch.rand(ma, H, generator=gen), 0.5 + torch.rand(mb, H, generator=gen)

        def through_norm(g, u, inv):
            g, u = g.view(-1, H, hd), u.detach().view(-1, H, hd)
            return ((g - u * (u * g).sum(-1, keepdim=True)) * inv.double()[..., None]).reshape(-1, C)
        dq_ref, dk_ref, dv_ref = through_norm(qh.grad, qh, inv_q), through_norm(kh.grad, kh, inv_k), vh.grad
        dtau = torch.zeros(1, device=DEV)
        if cross:
            iq, ik = inv_q.to(DEV), inv_k.to(DEV)
        else:   # self layers keep [q heads | k heads] per row
            both = torch.cat([inv_q, inv_k], 1).to(DEV)
            iq, ik = both[:, :H], both[:, H:]
        dq, dk, dv = ops.bf16_window_attention_bwd(dev_bf(do), qd, kd, vd, o, lse, iq, ik, qt, qc, kt, kc, P.n_win[shift:shift + 1], ops.small_end(P, shift),
                                                   ops.mid_end(P, shift), min(P.wcap, ma), tau.to(DEV), 0.01, H, dtau, zero=True)
        assert dq.stride(0) == qd.stride(0) and dk.stride(0) == kd.stride(0), "gradients mirror the packed projection layout"
        gs = max(dq_ref.abs().max().item(), 1e-6)
        assert_close(dq.float(), dq_ref, 2e-2, 4e-3 * gs, "dq"), assert_close(dk.float(), dk_ref, 2e-2, 4e-3 * dk_ref.abs().max().item(), "dk")
        assert_close(dv.float(), dv_ref, 2e-2, 4e-3 * dv_ref.abs().max().item(), "dv")
        assert_close(dtau, tau64.grad.reshape(1), 2e-2, 2e-2 * abs(tau64.grad.item()), "dtau")
