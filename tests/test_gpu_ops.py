"""GPU parity tests, op by op, through the C ABI (python -m pytest tests -m gpu).

Each CUDA entry point is compared with the tier-2 oracle (oracle/restated.py) or, for plain dense algebra,
with float64 torch on the same seeded inputs.  Integer outputs must be bit-exact; fp32 outputs must be within
rtol 1e-5 + atol 1e-5 (parity mode) unless a test states and justifies another bound.
"""
import numpy as np
import pytest
import torch

from common import assert_close, assert_equal_int
from oracle import cases, restated

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200  # noqa: F401
    from tmae_b200 import ops, synth
    from tmae_b200._lib import lib
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False

DEV = "cuda"
S = cases.SMALL


def test_library_loads_on_sm100():
    L = lib()
    assert L.version() == 1
    L.device_check()


@pytest.mark.parametrize("n", [0, 1, 255, 4096, 4097, 1_000_003])
def test_exclusive_scan(n):
    L = lib()
    g = torch.Generator().manual_seed(n)
    x = torch.randint(0, 5, (max(n, 1),), generator=g, dtype=torch.int32).to(DEV)
    out = torch.empty_like(x)
    tot = torch.full((1,), -1, dtype=torch.int32, device=DEV)
    scratch = torch.empty(int(L.scan_scratch_elems(n)), dtype=torch.int32, device=DEV)
    L.exclusive_scan_i32(x.data_ptr(), out.data_ptr(), n, tot.data_ptr(), scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
    ref = torch.cumsum(x[:n].long(), 0) - x[:n].long()
    assert_equal_int(out[:n], ref)
    assert int(tot) == int(x[:n].sum())


def _voxelize_both(pts, B, shape=S):
    v = ops.voxelize(torch.from_numpy(pts).to(DEV), shape["range"], shape["voxel"], shape["grid"], B)
    c = v["counts"].cpu()
    nk, nv = int(c[0]), int(c[1])
    ref = restated.voxelize(torch.from_numpy(pts), shape["range"], shape["voxel"], shape["grid"])
    return v, nk, nv, c, ref


@pytest.mark.parametrize("seed,n_keep,B", [(7, 1500, 2), (8, 4000, 3), (9, 50, 1)])
def test_voxelize_small(seed, n_keep, B):
    pts, _ = cases.small_points(seed, n_keep, B)
    v, nk, nv, c, (rp, rc, rinv, rvc, rmean) = _voxelize_both(pts, B)
    assert nk == rp.shape[0] and nv == rvc.shape[0]
    assert_close(v["points"][:nk], rp, 0, 0, "kept points")
    assert_equal_int(v["point_coords"][:nk], rc, "point_coords")
    assert_equal_int(v["inverse"][:nk], rinv, "inverse")
    assert_equal_int(v["voxel_coords"][:nv], rvc, "voxel_coords")
    assert_close(v["voxel_mean"][:nv], rmean, 1e-6, 1e-6, "voxel mean")
    # CSR: ascending point rows per voxel, counts match
    off, order = v["voxel_offset"].cpu().long(), v["pt_order"].cpu().long()
    cnt = torch.bincount(rinv, minlength=nv)
    assert_equal_int(off[1:nv + 1] - off[:nv], cnt)
    assert_equal_int(rinv[order[:nk]], torch.repeat_interleave(torch.arange(nv), cnt))
    seg_start = torch.zeros(nk, dtype=torch.bool)
    seg_start[off[:nv]] = True
    assert ((order[1:nk] > order[:nk - 1]) | seg_start[1:]).all()
    # first voxel row of each sample
    starts = [int((rvc[:, 0] < b).sum()) for b in range(B)]
    assert c[2:].tolist() == starts


def test_voxelize_edge_cases():
    # every point outside -> zero voxels; truncation rule for z; a point on the upper bound is dropped
    far = np.array([[0, 100.0, 0, 0, .5], [0, 0, 0, 9.0, .5]], np.float32)
    v = ops.voxelize(torch.from_numpy(far).to(DEV), S["range"], S["voxel"], S["grid"], 1)
    assert v["counts"][:2].tolist() == [0, 0]
    p = np.array([[0, 0.1, 0.1, -12.9, .5], [0, 0.1, 0.1, -13.1, .5], [0, 0.1, 0.1, 2.99, .5], [0, 0.1, 0.1, 3.0, .5],
                  [0, -15.5, 0.1, 0.0, .5], [0, 15.36, 0.0, 0.0, .5]], np.float32)
    v = ops.voxelize(torch.from_numpy(p).to(DEV), S["range"], S["voxel"], S["grid"], 1)
    keep, _ = restated.in_range_coords(torch.from_numpy(p), S["range"], S["voxel"], S["grid"])
    assert int(v["counts"][0]) == int(keep.sum()) == 3
    # empty input
    v = ops.voxelize(torch.zeros(0, 5, device=DEV), S["range"], S["voxel"], S["grid"], 2)
    assert v["counts"][:2].tolist() == [0, 0]


def test_voxelize_full_size_properties():
    """BASELINE-sized input (2 x 60k points, 468^2 grid, also the 0.1 m stress grid): size-independent properties."""
    from tmae_b200 import synth
    pts, _ = synth.batch(1000, 2, 60000)
    for voxel, grid in ((synth.ONCE["voxel"], synth.grid_size(synth.ONCE)), ([0.1, 0.1, 8.0], [1498, 1498, 1])):
        v = ops.voxelize(torch.from_numpy(pts).to(DEV), synth.ONCE["range"], voxel, grid, 2)
        nk, nv = v["counts"][:2].tolist()
        vc = v["voxel_coords"][:nv].cpu()
        key = ((vc[:, 0] * grid[2] + vc[:, 1]) * grid[1] + vc[:, 2]) * grid[0] + vc[:, 3]
        assert (key[1:] > key[:-1]).all(), "voxel rows strictly ascending (= unique + sorted)"
        pc, inv = v["point_coords"][:nk].cpu(), v["inverse"][:nk].cpu()
        assert torch.equal(vc[inv], pc), "every point's voxel row holds its own coordinates"
        assert inv.unique().numel() == nv
        # shuffling the points leaves the voxel set bit-identical and the means equal up to summation order
        perm = np.random.default_rng(0).permutation(pts.shape[0])
        v2 = ops.voxelize(torch.from_numpy(pts[perm]).to(DEV), synth.ONCE["range"], voxel, grid, 2)
        assert v2["counts"][:2].tolist() == [nk, nv]
        assert torch.equal(v2["voxel_coords"][:nv].cpu(), vc)
        assert_close(v2["voxel_mean"][:nv], v["voxel_mean"][:nv], 1e-5, 1e-5, "mean under permutation")
    ref = restated.voxelize(torch.from_numpy(pts), synth.ONCE["range"], synth.ONCE["voxel"], synth.grid_size(synth.ONCE))
    v = ops.voxelize(torch.from_numpy(pts).to(DEV), synth.ONCE["range"], synth.ONCE["voxel"], synth.grid_size(synth.ONCE), 2)
    nk, nv = v["counts"][:2].tolist()
    assert_equal_int(v["voxel_coords"][:nv], ref[3]), assert_equal_int(v["inverse"][:nk], ref[2])
    assert_close(v["voxel_mean"][:nv], ref[4], 1e-6, 1e-6)


def test_vfe_point_features():
    pts, _ = cases.small_points(3, 1200, 2)
    v, nk, nv, c, (rp, rc, rinv, rvc, rmean) = _voxelize_both(pts, 2)
    x = ops.vfe_point_features(v["points"][:nk], v["point_coords"][:nk], v["inverse"][:nk], v["voxel_mean"][:nv], S["range"], S["voxel"])
    ref = restated.vfe_point_features(rp, rc, rinv, rmean, S["range"], S["voxel"])
    assert_close(x, ref, 1e-6, 1e-6, "vfe point features")


# ------------------------------------------------------------------------------------ partition
def _coords(seed, m, B, g):
    rng = np.random.default_rng(seed)
    cells = np.sort(rng.choice(B * g * g, size=min(m, B * g * g), replace=False))
    return torch.tensor(np.stack([cells // (g * g), (cells % (g * g)) // g, cells % g], 1), dtype=torch.int32)


def _levels(kind):
    pre = restated.model_cfg(kind)["BACKBONE_3D"]["SST_BLOCK_LIST"][0]["PREPROCESS"]
    di = {int(k): v for k, v in pre["DROP_INFO"]["train"].items()}
    return pre, di, [(di[k]["max_tokens"], di[k]["drop_range"][0], di[k]["drop_range"][1]) for k in sorted(di)]


def _check_ref_tables(P, ref, info, di, m):
    """P.ref_* vs the oracle's SSTInputLayer tables (single frame, nothing dropped)."""
    bwi, lvl, f2w = (t.cpu() for t in ref)
    for s in range(2):
        assert_equal_int(bwi[s, :m], info[f"batch_win_inds_shift{s}"], f"batch_win_inds s{s}")
        assert_equal_int(lvl[s, :m], info[f"voxel_drop_level_shift{s}"], f"drop level s{s}")
        t = info[f"flat2win_inds_shift{s}"]
        for dl in (k for k in t if not isinstance(k, str)):
            inds, (pos,) = t[dl]
            assert_equal_int(f2w[s, :m][pos], inds, f"flat2win s{s} level {dl}")
        ciw = info[f"coors_in_win_shift{s}"]
        assert_equal_int(P.posidx_a[s, :m].cpu(), ciw[:, 1] * 8 + ciw[:, 2], "in-window coords")


@pytest.mark.parametrize("kind", ["pretrain", "finetune"])
@pytest.mark.parametrize("seed,m,B,g", [(0, 3000, 2, 96), (1, 40, 1, 20), (2, 9000, 3, 117), (3, 1, 1, 8)])
def test_window_partition_single(kind, seed, m, B, g):
    pre, di, levels = _levels(kind)
    c = _coords(seed, m, B, g)
    m = c.shape[0]
    P = ops.window_partition(c.to(DEV), B, g, g, levels, want_ref=True)
    assert int(P.status) == 0
    coords4 = torch.stack([c[:, 0], torch.zeros_like(c[:, 0]), c[:, 1], c[:, 2]], 1).long()
    info = restated.sst_input(torch.zeros(m, 4), coords4, [g, g, 1], pre)
    assert info["voxel_keep_inds"].shape[0] == m
    _check_ref_tables(P, P.ref_a, info, di, m)
    # internal tables: slot = stable rank, token table inverts (window, slot) -> voxel
    win, slot, tok, cnt = P.win_a.cpu().long(), P.slot_a.cpu().long(), P.tok_a.cpu().long(), P.cnt_a.cpu().long()
    for s in range(2):
        assert_equal_int(slot[s, :m], restated.stable_rank(info[f"batch_win_inds_shift{s}"]), "slot = canonical stable rank")
        nw = int(P.n_win[s])
        assert nw == info[f"batch_win_inds_shift{s}"].unique().numel()
        assert_equal_int(tok[s][win[s, :m] * 64 + slot[s, :m]], torch.arange(m), "token table")
        assert_equal_int(cnt[s, :nw], torch.bincount(win[s, :m], minlength=nw), "window counts")
        lb = P.level_base[s].cpu().tolist()
        assert lb[0] == 0 and lb[-1] == nw and all(a <= b for a, b in zip(lb, lb[1:]))


@pytest.mark.parametrize("kind", ["pretrain", "finetune"])
@pytest.mark.parametrize("seed,ma,mb,B,g", [(0, 800, 3000, 2, 96), (1, 3000, 700, 2, 48), (2, 5, 9, 1, 16)])
def test_window_partition_temporal(kind, seed, ma, mb, B, g):
    pre, di, levels = _levels(kind)
    ca, cb = _coords(seed, ma, B, g), _coords(seed + 100, mb, B, g)
    P = ops.window_partition(ca.to(DEV), B, g, g, levels, coords_b=cb.to(DEV), want_ref=True)
    assert int(P.status) == 0

    def c4(c):
        return torch.stack([c[:, 0], torch.zeros_like(c[:, 0]), c[:, 1], c[:, 2]], 1).long()
    a, b = restated.sst_input_temporal(torch.zeros(ca.shape[0], 4), c4(ca), torch.zeros(cb.shape[0], 4), c4(cb), [g, g, 1], pre)
    for inf, ref, win, m in ((a, P.ref_a, P.win_a, ca.shape[0]), (b, P.ref_b, P.win_b, cb.shape[0])):
        bwi, lvl, f2w = (t.cpu() for t in ref)
        for s in range(2):
            keep = torch.where(win[s, :m].cpu() >= 0)[0]
            assert_equal_int(keep, inf[f"voxel_keep_inds_shift{s}"], "temporal keep")
            assert_equal_int(lvl[s, :m][keep], inf[f"voxel_drop_level_shift{s}"], "temporal level")
            assert_equal_int(bwi[s, :m][keep], inf[f"batch_win_inds_shift{s}"], "temporal bwi")
            t = inf[f"flat2win_inds_shift{s}"]
            for dl in (k for k in t if not isinstance(k, str)):
                inds, (pos,) = t[dl]
                assert_equal_int(f2w[s, :m][keep][pos], inds, "temporal flat2win")
    # paired windows hold the same compact id in both frames
    for s in range(2):
        nw = int(P.n_win[s])
        wa, wb = P.win_a[s, :ca.shape[0]].cpu(), P.win_b[s, :cb.shape[0]].cpu()
        assert set(wa[wa >= 0].tolist()) == set(wb[wb >= 0].tolist()) == set(range(nw))


def test_window_partition_rejects_unsorted_and_overflow():
    _, _, levels = _levels("pretrain")
    c = _coords(0, 500, 1, 40)
    P = ops.window_partition(c.flip(0).contiguous().to(DEV), 1, 40, 40, levels)
    assert int(P.status) & 1
    full = _coords(0, 40 * 40, 1, 40)  # dense: 64 voxels per window, top level takes them all
    assert int(ops.window_partition(full.to(DEV), 1, 40, 40, levels).status) == 0
    assert int(ops.window_partition(full.to(DEV), 1, 40, 40, [(16, 0, 16), (32, 16, 100000)]).status) & 8


# ------------------------------------------------------------------------------------ dense algebra
@pytest.mark.parametrize("m,n,k", [(1, 64, 10), (777, 64, 10), (60001, 64, 10), (4099, 128, 11), (1000, 128, 64), (333, 384, 128), (2049, 256, 512), (65, 48, 128)])
def test_linear_fwd_bwd(m, n, k):
    g = torch.Generator().manual_seed(m + n + k)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5, torch.randn(n, generator=g)
    r = torch.randn(m, n, generator=g)
    xd, wd, bd, rd = (t.to(DEV) for t in (x, w, b, r))
    for act, f in ((ops.ACT_NONE, lambda t: t), (ops.ACT_GELU, torch.nn.functional.gelu), (ops.ACT_RELU, torch.relu)):
        y = ops.linear_fwd(xd, wd, bd, residual=rd, act=act)
        ref = f(x.double() @ w.double().T + b.double()) + r.double()
        assert_close(y, ref, 1e-5, 1e-5, f"linear act={act}")
    dy = torch.randn(m, n, generator=g)
    dx = ops.linear_bwd_data(dy.to(DEV), wd)
    assert_close(dx, dy.double() @ w.double(), 1e-5, 1e-5, "dx")
    dx2 = ops.linear_bwd_data(dy.to(DEV), wd, dx=dx.clone(), accumulate=True)
    assert_close(dx2, 2 * (dy.double() @ w.double()), 1e-5, 1e-5, "dx accumulate")
    dw, db = torch.empty_like(wd), torch.empty_like(bd)
    ops.linear_bwd_weight(dy.to(DEV), xd, dw, db)
    assert_close(dw, dy.double().T @ x.double(), 1e-5, 1e-4 * max(1, m) ** .5, "dw")  # fp32 sum over m rows
    assert_close(db, dy.double().sum(0), 1e-5, 1e-4 * max(1, m) ** .5, "db")
    pre = torch.randn(m, n, generator=g)
    p = pre.double().requires_grad_()
    torch.nn.functional.gelu(p).backward(dy.double())
    assert_close(ops.gelu_bwd(dy.to(DEV), pre.to(DEV)), p.grad, 1e-5, 1e-6, "gelu'")


@pytest.mark.parametrize("m,n,k,with_bias", [(1, 64, 10, False), (777, 64, 10, False), (5000, 128, 16, True), (240001, 64, 10, False)])
def test_linear_fwd_thin_k(m, n, k, with_bias):
    """The VFE's first layer (k = 10, no bias, no activation) takes the thin-reduction kernel in both precision modes."""
    g = torch.Generator().manual_seed(m + n)
    x, w = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5
    b = torch.randn(n, generator=g) if with_bias else None
    ref = x.double() @ w.double().T + (b.double() if with_bias else 0)
    for prec in ("fp32", "tf32"):
        ops.set_precision(prec)
        try:
            y = ops.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV) if with_bias else None)
        finally:
            ops.set_precision("fp32")
        assert_close(y, ref, 1e-5, 1e-5, f"thin-k linear ({prec})")


@pytest.mark.parametrize("rows,C", [(1, 128), (1000, 128), (4097, 256)])
def test_add_layernorm(rows, C):
    g = torch.Generator().manual_seed(rows)
    x, r = torch.randn(rows, C, generator=g), torch.randn(rows, C, generator=g)
    ga, be = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    mask = (torch.rand(rows, generator=g) > 0.4).to(torch.uint8)
    dy = torch.randn(rows, C, generator=g)
    for use_mask in (False, True):
        xd, rd = x.double().requires_grad_(), r.double().requires_grad_()
        gd, bd = ga.double().requires_grad_(), be.double().requires_grad_()
        v = xd + (rd * mask[:, None].double() if use_mask else rd)
        ref = torch.nn.functional.layer_norm(v, (C,), gd, bd, 1e-5)
        ref.backward(dy.double())
        mk = mask.to(DEV) if use_mask else None
        y, mean, rstd = ops.add_layernorm_fwd(x.to(DEV), r.to(DEV), mk, ga.to(DEV), be.to(DEV), 1e-5)
        assert_close(y, ref.detach(), 1e-5, 1e-5, "LN fwd")
        dg, db = torch.empty(C, device=DEV), torch.empty(C, device=DEV)
        dc = torch.full((C,), 7.0, device=DEV)   # overwritten: column sums of dres (masked case) or dv = the producer's bias gradient
        dv, dres = ops.add_layernorm_bwd(dy.to(DEV), x.to(DEV), r.to(DEV), mk, ga.to(DEV), mean, rstd, dg, db, want_dres=use_mask, dcolsum=dc)
        assert_close(dv, xd.grad, 1e-4, 1e-5, "LN dx")
        assert_close(dc, (rd.grad if use_mask else xd.grad).sum(0), 1e-4, 1e-4 * rows ** .5, "LN fused column sum")
        if use_mask:
            assert_close(dres, rd.grad, 1e-4, 1e-5, "LN dres (masked)")
        assert_close(dg, gd.grad, 1e-4, 1e-4 * rows ** .5, "LN dgamma")
        assert_close(db, bd.grad, 1e-4, 1e-4 * rows ** .5, "LN dbeta")


@pytest.mark.parametrize("rows,C,relu", [(5, 64, True), (3001, 128, True), (1500, 256, False)])
def test_batchnorm_rows(rows, C, relu):
    g = torch.Generator().manual_seed(rows + C)
    x = torch.randn(rows, C, generator=g) * 2 + 0.5
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).double()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, generator=g)), bn.bias.copy_(0.1 * torch.randn(C, generator=g))
    rm, rv = bn.running_mean.clone().float().to(DEV), bn.running_var.clone().float().to(DEV)
    ga, be = bn.weight.detach().float().to(DEV), bn.bias.detach().float().to(DEV)
    dy = torch.randn(rows, C, generator=g)
    xd = x.double().requires_grad_()
    ref = bn(xd)
    ref = torch.relu(ref) if relu else ref
    ref.backward(dy.double())
    y, mean, rstd = ops.bn_train_fwd(x.to(DEV), ga, be, rm, rv, 0.01, 1e-3, relu)
    assert_close(y, ref.detach(), 1e-5, 1e-5, "BN fwd")
    assert_close(rm, bn.running_mean, 1e-5, 1e-6, "running mean"), assert_close(rv, bn.running_var, 1e-5, 1e-6, "running var")
    dx, dg, db = ops.bn_bwd(dy.to(DEV), x.to(DEV), be, mean, rstd, ga, relu, True)
    assert_close(dx, xd.grad, 1e-4, 1e-5, "BN dx")
    assert_close(dg, bn.weight.grad, 1e-4, 1e-4 * rows ** .5, "BN dgamma"), assert_close(db, bn.bias.grad, 1e-4, 1e-4 * rows ** .5, "BN dbeta")
    # eval mode
    bn.eval()
    xe = x.double().requires_grad_()
    re = torch.relu(bn(xe)) if relu else bn(xe)
    re.backward(dy.double())
    m_e, r_e = bn.running_mean.float().to(DEV), torch.rsqrt(bn.running_var + 1e-3).float().to(DEV)
    ye = ops.bn_apply(x.to(DEV), m_e, r_e, ga, be, relu)
    assert_close(ye, re.detach(), 1e-5, 1e-5, "BN eval fwd")
    dxe, _, _ = ops.bn_bwd(dy.to(DEV), x.to(DEV), be, m_e, r_e, ga, relu, False)
    assert_close(dxe, xe.grad, 1e-4, 1e-5, "BN eval dx")


@pytest.mark.parametrize("rows,C", [(7, 128), (4099, 128), (2000, 64)])
def test_batchnorm_bf16_strided(rows, C):
    """BatchNorm2d + ReLU of the decoder on a channels-last bf16 map: output written into a column slice of a wider
    buffer (replaces torch.cat), gradient read from a column slice.  Reference: float64 BatchNorm on the same
    bf16-rounded input; tolerance = bf16 output rounding (rtol 2^-8)."""
    g = torch.Generator().manual_seed(rows * 3 + C)
    x = (torch.randn(rows, C, generator=g) * 1.5 + 0.3).bfloat16()
    bn = torch.nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).double()
    with torch.no_grad():
        bn.weight.copy_(1 + 0.1 * torch.randn(C, generator=g)), bn.bias.copy_(0.1 * torch.randn(C, generator=g))
    rm, rv = bn.running_mean.clone().float().to(DEV), bn.running_var.clone().float().to(DEV)
    ga, be = bn.weight.detach().float().to(DEV), bn.bias.detach().float().to(DEV)
    xd = x.double().requires_grad_()
    ref = torch.relu(bn(xd))
    wide = torch.full((rows, 3 * C), 7.0, dtype=torch.bfloat16, device=DEV)
    mean, rstd = ops.bn_bf16_fwd(x.to(DEV), ga, be, rm, rv, 0.01, 1e-3, True, True, wide[:, C:2 * C])
    assert_close(wide[:, C:2 * C].float(), ref.detach(), 2 ** -8, 1e-3, "BN bf16 fwd")
    assert (wide[:, :C] == 7).all() and (wide[:, 2 * C:] == 7).all(), "neighbouring column slices were touched"
    assert_close(rm, bn.running_mean, 1e-4, 1e-5, "running mean"), assert_close(rv, bn.running_var, 1e-4, 1e-5, "running var")
    dwide = torch.randn(rows, 3 * C, generator=g).bfloat16()
    ref.backward(dwide[:, C:2 * C].double())
    dx, dg, db = ops.bn_bf16_bwd(dwide.to(DEV)[:, C:2 * C], x.to(DEV), mean, rstd, ga, be, True, True)
    assert_close(dx.float(), xd.grad, 2 ** -7, 2e-3, "BN bf16 dx")
    assert_close(dg, bn.weight.grad, 1e-3, 1e-3 * rows ** .5, "BN bf16 dgamma"), assert_close(db, bn.bias.grad, 1e-3, 1e-3 * rows ** .5, "BN bf16 dbeta")
    # eval mode: given statistics
    bn.eval()
    re = torch.relu(bn(x.double()))
    m_e, r_e = bn.running_mean.float().to(DEV), torch.rsqrt(bn.running_var + 1e-3).float().to(DEV)
    out = torch.empty(rows, C, dtype=torch.bfloat16, device=DEV)
    ops.bn_bf16_fwd(x.to(DEV), ga, be, None, None, 0.01, 1e-3, True, False, out, m_e, r_e)
    assert_close(out.float(), re, 2 ** -8, 1e-3, "BN bf16 eval")


def test_segment_max_and_rows():
    pts, _ = cases.small_points(5, 900, 2)
    v, nk, nv, c, (rp, rc, rinv, rvc, rmean) = _voxelize_both(pts, 2)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(nk, 128, generator=g)
    out, arg = ops.segment_max_fwd(x.to(DEV), v["voxel_offset"], v["pt_order"], nv)
    assert_close(out, restated.segment_max(x, rinv, nv), 0, 0, "segment max")
    assert_close(x[arg.cpu().long(), torch.arange(128)[None]], out, 0, 0, "argmax rows")
    d = torch.randn(nv, 128, generator=g)
    dx = ops.segment_max_bwd(d.to(DEV), arg, nk)
    xr = x.double().requires_grad_()
    restated.segment_max(xr, rinv, nv).backward(d.double())
    assert_close(dx, xr.grad, 1e-6, 1e-6, "segment max backward (no ties in random data)")
    # sparse rows <-> dense map
    idx = rvc[:, [0, 2, 3]].int()
    f = torch.randn(nv, 128, generator=g)
    dense = ops.densify_nhwc(f.to(DEV), idx.to(DEV), 2, 96, 96)
    ref = restated.SparseTensor(f, idx, [96, 96], 2).dense()
    assert_close(dense.permute(0, 3, 1, 2), ref, 0, 0, "densify")
    assert_close(ops.gather_nhwc(dense, idx.to(DEV)), f, 0, 0, "gather back")
    lut = torch.randn(64, 128, generator=g)
    pi = torch.randint(0, 64, (nv,), generator=g).to(torch.uint8)
    assert_close(ops.add_pos(f.to(DEV), pi.to(DEV), lut.to(DEV)), f + lut[pi.long()], 0, 0, "add_pos")


# ------------------------------------------------------------------------------------ sparse conv
@pytest.mark.parametrize("seed,m,B,g,cin,cout", [(0, 1500, 2, 96, 128, 128), (1, 700, 2, 47, 128, 256), (2, 30, 1, 9, 256, 256)])
def test_sparse_conv(seed, m, B, g, cin, cout):
    c = _coords(seed, m, B, g)
    m = c.shape[0]
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(m, cin, generator=gen)
    for subm in (True, False):
        conv = restated.SparseConv(cin, cout, 3, 1 if subm else 2, 1, subm).double()
        xr = x.double().requires_grad_()
        ref = conv(restated.SparseTensor(xr, c, [g, g], B))
        w = conv.weight.detach().float().to(DEV)
        if subm:
            table = ops.subm_table(c.to(DEV), B, g, g)
            table_t, flip, rows_out = table, True, m
        else:
            idx_out, n_out, table, table_t, (yo, xo) = ops.strided_table(c.to(DEV), B, g, g)
            rows_out = int(n_out)
            assert [yo, xo] == ref.spatial_shape
            assert_equal_int(idx_out[:rows_out], ref.indices, "strided output sites (lexicographic)")
            table, flip = table[:rows_out], False
        y = ops.sparse_conv_fwd(x.to(DEV), table, w, rows_out)
        assert_close(y, ref.features.detach(), 1e-5, 1e-5, f"sparse conv fwd subm={subm}")
        dy = torch.randn(rows_out, cout, generator=gen)
        ref.features.backward(dy.double())
        dx = ops.sparse_conv_fwd(dy.to(DEV), table_t, ops.transpose_taps(w, flip), m)
        assert_close(dx, xr.grad, 1e-5, 1e-5, f"sparse conv dx subm={subm}")
        dw = ops.sparse_conv_bwd_weight(dy.to(DEV), x.to(DEV), table, w.shape)
        assert_close(dw, conv.weight.grad, 1e-5, 1e-4 * m ** .5, f"sparse conv dw subm={subm}")


# ------------------------------------------------------------------------------------ attention
def _attn_ref(mha, q_in, k_in, v_in, wq, wk, slotq, slotk, nw):
    """Oracle CosineMHA on zero-padded (64, windows, C) tensors built from the partition tables."""
    C = q_in.shape[1]
    Q, K, V = (torch.zeros(64, nw, C, dtype=torch.float64) for _ in range(3))
    Q[slotq, wq], K[slotk, wk], V[slotk, wk] = q_in, k_in, v_in
    pad = torch.ones(nw, 64, dtype=torch.bool)
    pad[wk, slotk] = False
    out = mha(Q, K, V, pad)
    return out[slotq, wq]


@pytest.mark.parametrize("tc", [0, 1])
@pytest.mark.parametrize("C,cross", [(128, False), (256, False), (128, True), (256, True)])
def test_window_attention_core(C, cross, tc):
    """tc = 1: windows above 16 tokens run on mma.sync TF32 (operands rounded to 10 mantissa bits, fp32 accumulate):
    tolerance 3e-3 instead of the fp32 path's 1e-5 / 1e-4."""
    ops.set_precision("tf32" if tc else "fp32")
    try:
        _attention_core_case(C, cross, 3e-3 if tc else None)
    finally:
        ops.set_precision("fp32")


def _attention_core_case(C, cross, tol):
    H, B, g = 8, 2, 64
    _, _, levels = _levels("pretrain")
    ca = _coords(0, 2500 if not cross else 600, B, g)
    cb = _coords(5, 2200, B, g) if cross else None
    P = ops.window_partition(ca.to(DEV), B, g, g, levels, coords_b=cb.to(DEV) if cross else None)
    gen = torch.Generator().manual_seed(C)
    ma, mb = ca.shape[0], (cb.shape[0] if cross else ca.shape[0])
    q = torch.randn(ma, C, generator=gen)
    k = torch.randn(mb, C, generator=gen)
    v = torch.randn(mb, C, generator=gen)
    tau = torch.tensor([[[0.37]]])
    shift = 1
    nw = int(P.n_win[shift])
    qt, qc = P.tok_a[shift], P.cnt_a[shift]
    kt, kc = (P.tok_b[shift], P.cnt_b[shift]) if cross else (qt, qc)
    o, lse = ops.window_attention_fwd(q.to(DEV), k.to(DEV), v.to(DEV), qt, qc, kt, kc, P.n_win[shift:shift + 1], ops.small_end(P, shift), ops.mid_end(P, shift),
                                      min(P.wcap, ma), tau.to(DEV), 0.01, H, zero_out=cross)
    # oracle: identity projections so that the core is isolated
    mha = restated.CosineMHA(C, H).double()
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.eye(C).repeat(3, 1)), mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(torch.eye(C)), mha.out_proj.bias.zero_(), mha.tau.copy_(tau)
    wa, sa = P.win_a[shift, :ma].cpu().long(), P.slot_a[shift, :ma].cpu().long()
    wb, sb = (P.win_b[shift, :mb].cpu().long(), P.slot_b[shift, :mb].cpu().long()) if cross else (wa, sa)
    ka, kb = wa >= 0, wb >= 0
    qd, kd, vd = q.double().requires_grad_(), k.double().requires_grad_(), v.double().requires_grad_()
    ref = _attn_ref(mha, qd[ka], kd[kb], vd[kb], wa[ka], wb[kb], sa[ka], sb[kb], nw)
    assert_close(o.cpu()[ka], ref.detach(), tol or 1e-5, tol or 1e-5, "attention output")
    if cross:
        assert (o.cpu()[~ka] == 0).all()
    do = torch.randn(ma, C, generator=gen)
    ref.backward(do.double()[ka])
    dtau = torch.zeros(1, 1, 1, device=DEV)
    dq, dk, dv = ops.window_attention_bwd(do.to(DEV), q.to(DEV), k.to(DEV), v.to(DEV), o, lse, qt, qc, kt, kc, P.n_win[shift:shift + 1],
                                          ops.small_end(P, shift), ops.mid_end(P, shift), min(P.wcap, ma), tau.to(DEV), 0.01, H, dtau, zero=cross)
    gt = (tol or 1e-4, (tol or 1e-5) * 3)
    assert_close(dq, qd.grad, *gt, "dq"), assert_close(dk, kd.grad, *gt, "dk"), assert_close(dv, vd.grad, *gt, "dv")
    assert_close(dtau, mha.tau.grad, tol or 1e-4, (tol or 1e-4) * mha.tau.grad.abs().item(), "dtau")


def test_window_attention_packed_strides():
    """q, k, v as column blocks of one packed (rows, 3C) projection output (row pitch 3C), gradients written into the
    same packed layout: identical results to separate contiguous tensors."""
    C, H, B, g = 128, 8, 2, 64
    _, _, levels = _levels("pretrain")
    ca = _coords(3, 1800, B, g)
    P = ops.window_partition(ca.to(DEV), B, g, g, levels)
    m = ca.shape[0]
    gen = torch.Generator().manual_seed(5)
    qkv = torch.randn(m, 3 * C, generator=gen).to(DEV)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    tau = torch.tensor([[[0.4]]], device=DEV)
    args = (P.tok_a[0], P.cnt_a[0], P.tok_a[0], P.cnt_a[0], P.n_win[0:1], ops.small_end(P, 0), ops.mid_end(P, 0), min(P.wcap, m), tau, 0.01, H)
    for tc in (0, 1):
        ops.set_precision("tf32" if tc else "fp32")
        try:
            o1, l1 = ops.window_attention_fwd(q, k, v, *args, False)
            o2, l2 = ops.window_attention_fwd(q.contiguous(), k.contiguous(), v.contiguous(), *args, False)
            assert torch.equal(o1, o2) and torch.equal(l1, l2)
            do = torch.randn(m, C, generator=gen).to(DEV)
            t1, t2 = torch.zeros(1, device=DEV), torch.zeros(1, device=DEV)
            g1 = ops.window_attention_bwd(do, q, k, v, o1, l1, *args, t1, False)
            g2 = ops.window_attention_bwd(do, q.contiguous(), k.contiguous(), v.contiguous(), o2, l2, *args, t2, False)
            assert g1[0].stride(0) == 3 * C, "gradients of a packed projection come back packed"
            for a, b in zip(g1, g2):
                assert torch.equal(a, b)
        finally:
            ops.set_precision("fp32")


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
@pytest.mark.parametrize("m,c,parts", [(3000, 128, 3), (1700, 256, 3), (900, 128, 1), (5, 256, 2)])
def test_packed_projection_with_position_table(m, c, parts, prec):
    """y = x W^T + (pos_lut W_pos^T + b)[posidx] against (x + pos) W^T + b in float64, and its backward pieces:
    binned column sum, bias / position-term weight gradient."""
    gen = torch.Generator().manual_seed(m + c)
    n = parts * c
    n_pos = min(2, parts) * c if parts != 1 else c
    x = torch.randn(m, c, generator=gen)
    w = torch.randn(n, c, generator=gen) / c ** .5
    b = torch.randn(n, generator=gen) * 0.1
    lut = torch.randn(64, c, generator=gen)
    pi = torch.randint(0, 64, (m,), generator=gen, dtype=torch.uint8)
    xd, wd, ld = x.double(), w.double(), lut.double()
    pos = ld[pi.long()]
    ref = xd @ wd.T + b.double()
    ref[:, :n_pos] += pos @ wd[:n_pos].T
    ops.set_precision(prec)
    try:
        table, table_t = ops.pos_table(lut.to(DEV), w.to(DEV), b.to(DEV), n_pos)
        tref = b.double()[None, :].repeat(64, 1)
        tref[:, :n_pos] += ld @ wd[:n_pos].T
        assert_close(table, tref, 1e-5, 1e-5, "position table"), assert_close(table_t, tref.T, 1e-5, 1e-5, "position table (transposed)")
        y = ops.linear_fwd_lut(x.to(DEV), w.to(DEV), table, pi.to(DEV))
        onehot = ops.onehot64(pi.to(DEV))
        assert torch.equal(onehot.cpu(), torch.nn.functional.one_hot(pi.long(), 64).float())
        y2 = ops.linear_fwd_dual(x.to(DEV), w.to(DEV), onehot, table_t)
        dy_ = torch.randn(m, n, generator=gen)
        dtt = torch.empty(n, 64, device=DEV)
        ops.linear_bwd_weight(dy_.to(DEV), onehot, dtt)
    finally:
        ops.set_precision("fp32")
    tol = 1e-5 if prec == "fp32" else 3e-3
    assert_close(y, ref, 1e-5, 3e-5, "packed projection, table-bias epilogue (fp32 SIMT)")
    assert_close(y2, ref, tol, tol * 3, f"packed projection, dual-source GEMM ({prec})")
    dref_t = torch.zeros(64, n, dtype=torch.float64).index_add_(0, pi.long(), dy_.double()).T
    assert_close(dtt, dref_t, tol * 2, tol * 2 * m ** .5, f"dy^T onehot ({prec})")
    dw1 = torch.zeros(n, c, device=DEV)
    db1 = ops.pos_table_bwd(dtt, lut.to(DEV), dw1, n_pos, transposed=True)
    assert_close(db1, dy_.double().sum(0), tol * 2, tol * 2 * m ** .5, "bias gradient from the transposed table gradient")
    dy = torch.randn(m, n, generator=gen)
    dt = ops.binned_colsum(dy.to(DEV), pi.to(DEV))
    dref = torch.zeros(64, n, dtype=torch.float64).index_add_(0, pi.long(), dy.double())
    assert_close(dt, dref, 1e-5, 1e-4, "binned column sum")
    dw0 = torch.randn(n, c, generator=gen)
    dw = dw0.clone().to(DEV)
    db = ops.pos_table_bwd(dt, lut.to(DEV), dw, n_pos)
    assert_close(db, dy.double().sum(0), 1e-5, 1e-4 * m ** .5, "bias gradient")
    wref = dw0.double()
    wref[:n_pos] += dref[:, :n_pos].T @ ld
    assert_close(dw, wref, 1e-5, 1e-4 * m ** .5, "position term of the weight gradient")


# ------------------------------------------------------------------------------------ loss
def test_gt_group_and_chamfer():
    pts, _ = cases.small_points(4, 2500, 2)
    v, nk, nv, c, (rp, rc, rinv, rvc, rmean) = _voxelize_both(pts, 2)
    K = 64
    gt, inds = ops.gt_group(v["points"][:nk], v["voxel_offset"], v["pt_order"], v["voxel_coords"][:nv], S["range"], S["voxel"], nv, K, True)
    rinds = restated.group_inner_inds(rinv, nv, K)
    assert_equal_int(inds, rinds, "group_inner_inds (canonical)")
    centers = (rvc[:, 1:].flip(-1).float() + 0.5) * torch.tensor(S["voxel"]) + torch.tensor(S["range"][:3])
    rgt = rp[:, 1:4][rinds] - centers[:, None]
    assert_close(gt, rgt, 1e-6, 1e-6, "gt points")
    gen = torch.Generator().manual_seed(1)
    pred = torch.randn(nv, 16, 3, generator=gen) * 0.3
    w = (torch.rand(nv, generator=gen) < 0.75).float()
    for fused in (False, True):
        ctx = (v["points"][:nk], v["voxel_offset"], v["pt_order"], v["voxel_coords"][:nv].contiguous(), S["range"], S["voxel"], K)
        loss, state = ops.chamfer_fwd(pred.to(DEV), None if fused else gt, w.to(DEV), ctx if fused else None)
        pr = pred.double().requires_grad_()
        ref = restated.chamfer(pr, rgt.double(), w.double())
        assert_close(loss, ref.detach(), 1e-5, 1e-7, f"chamfer loss fused={fused}")
        ref.backward()
        gl = torch.tensor(1.0, device=DEV)
        dp = ops.chamfer_bwd(gl, pred.to(DEV), None if fused else gt, w.to(DEV), state, ctx if fused else None)
        assert_close(dp, pr.grad, 1e-4, 1e-8, f"chamfer dpred fused={fused}")
    loss0, _ = ops.chamfer_fwd(pred.to(DEV), gt, torch.zeros(nv, device=DEV))
    assert float(loss0) == 0.0


@pytest.mark.parametrize("n,groups,K", [(1, 1, 4), (5000, 37, 8), (60000, 14000, 64), (777, 2000, 3)])
def test_sst_ops_one_to_one_seam(n, groups, K):
    """tmae_ingroup_inds / tmae_group_inner_inds take the reference's own tensors (sst_ops_api.cpp:6-8) and return the canonical
    result of sst_ops_gpu.cu:14-39 (what a serial run of the reference kernels yields), bit for bit: unsorted group ids, groups
    larger than K (first K indices kept), groups smaller than K (cyclic padding), absent groups (row left at -1)."""
    from oracle import shims
    g = torch.Generator().manual_seed(n + groups)
    gid = torch.randint(0, groups, (n,), generator=g)
    ref = torch.zeros_like(gid) - 1
    shims.ingroup_inds_wrapper(gid, ref)
    out = ops.get_inner_win_inds(gid.to(DEV))
    assert_equal_int(out, ref, "ingroup_inds")
    m = int(gid.max()) + 1
    gref = torch.full((m, K), -1, dtype=torch.long)
    shims.group_inner_inds_wrapper(gid, gref)
    pts = torch.randn(n, 3, generator=g)
    got = ops.group_inner_inds(pts.to(DEV), gid.to(DEV), K)
    assert torch.equal(got.cpu(), pts[gref]), "group_inner_inds"
    got2 = ops.group_inner_inds(pts.to(DEV), gid.to(DEV), K, n_groups=m)   # no host sync form
    assert torch.equal(got2.cpu(), pts[gref])
