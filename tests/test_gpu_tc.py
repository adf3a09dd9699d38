"""GPU tests of the TMA-fed tcgen05 kind::tf32 GEMM family (TMAE_PREC_TF32, fp32 storage) against float64 references
computed on the SAME TF32-rounded operands (10 explicit mantissa bits): only the rounding mode of the copy engine /
tensor core and the accumulation order differ, bound rtol 1e-3 + atol 1e-3 (one TF32 ulp per operand; fp32 accumulate
in TMEM).  Every test also asserts through tmae_dispatch_counts that the TMA kernel -- not the fp32 FFMA kernel --
served the calls."""
import numpy as np
import pytest
import torch

from common import assert_close
from oracle import restated

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200  # noqa: F401
    from tmae_b200 import ops

DEV = "cuda"
TOL = 1e-4


def bf(t):
    return t.to(torch.bfloat16).double()


def tf32(t):
    """fp32 -> TF32 (10 explicit mantissa bits), round to nearest; what the TMA-fed kind::tf32 GEMMs consume."""
    i = t.float().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32).double()


@pytest.fixture(autouse=True, params=["tma_tf32"])
def _tc_mode(request):
    ops.set_precision("tf32")
    before = ops.dispatch_counts()
    yield request.param
    after = ops.dispatch_counts()
    ops.set_precision("fp32")
    assert after["tma"] > before["tma"], "the TMA-fed tcgen05 kernel did not run"
    assert after["simt_in_tc_mode"] == before["simt_in_tc_mode"], "a tensor-core-mode GEMM fell back to the fp32 FFMA kernel"


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (1, 128, 64), (1000, 128, 128), (333, 256, 128), (2049, 256, 512), (4100, 512, 256),
                                   (65, 48, 128), (5000, 128, 256)])
def test_tc_linear_fwd_bwd(m, n, k, _tc_mode):
    # bf16 staging: exact products of rounded operands, only summation order differs -> 1e-4.  TF32 through TMA: the
    # rounding (nearest vs truncation of the 13 dropped bits) is done by the copy engine / tensor core, so allow one
    # TF32 ulp (2^-11 relative per operand) on top: 1e-3.
    TOL = 1e-3 if _tc_mode == "tma_tf32" else 1e-4
    bf = tf32 if _tc_mode == "tma_tf32" else globals()["bf"]
    g = torch.Generator().manual_seed(m + n + k)
    x, w, b = torch.randn(m, k, generator=g), torch.randn(n, k, generator=g) / k ** .5, torch.randn(n, generator=g)
    r = torch.randn(m, n, generator=g)
    xd, wd, bd, rd = (t.to(DEV) for t in (x, w, b, r))
    for act, f in ((ops.ACT_NONE, lambda t: t), (ops.ACT_GELU, torch.nn.functional.gelu)):
        use_res = False   # the TMA kernel has no residual input (the layers never use one)
        y, pre = ops.linear_fwd(xd, wd, bd, residual=rd if use_res else None, act=act, want_preact=True)
        lin = bf(x) @ bf(w).T + b.double()
        assert_close(pre, lin, TOL, TOL, f"tc linear preact m={m}")
        assert_close(y, f(lin) + (r.double() if use_res else 0), TOL, TOL, f"tc linear act={act}")
    dy = torch.randn(m, n, generator=g)
    dx = ops.linear_bwd_data(dy.to(DEV), wd)
    assert_close(dx, bf(dy) @ bf(w), TOL, TOL, "tc dx")
    dx2 = ops.linear_bwd_data(dy.to(DEV), wd, dx=dx.clone(), accumulate=True)
    assert_close(dx2, 2 * (bf(dy) @ bf(w)), TOL, 2 * TOL, "tc dx accumulate")
    dw, db = torch.empty_like(wd), torch.empty_like(bd)
    ops.linear_bwd_weight(dy.to(DEV), xd, dw, db)
    assert_close(dw, bf(dy).T @ bf(x), TOL, TOL * max(1, m) ** .5, "tc dw")
    assert_close(db, dy.double().sum(0), 1e-5, 1e-4 * max(1, m) ** .5, "db")


def test_tc_weight_slices(_tc_mode):
    """packed in_proj: q/k/v slices addressed by row offset, as the encoder layers do."""
    bf = tf32 if _tc_mode == "tma_tf32" else globals()["bf"]
    g = torch.Generator().manual_seed(0)
    C, m = 128, 900
    x, w, b = torch.randn(m, C, generator=g), torch.randn(3 * C, C, generator=g) / C ** .5, torch.randn(3 * C, generator=g)
    for i in range(3):
        y = ops.linear_fwd(x.to(DEV), w.to(DEV), b.to(DEV), w_offset_rows=i * C, n=C)
        assert_close(y, bf(x) @ bf(w[i * C:(i + 1) * C]).T + b[i * C:(i + 1) * C].double(), 1e-3, 1e-3, f"slice {i}")


def _coords(seed, m, B, g):
    rng = np.random.default_rng(seed)
    cells = np.sort(rng.choice(B * g * g, size=min(m, B * g * g), replace=False))
    return torch.tensor(np.stack([cells // (g * g), (cells % (g * g)) // g, cells % g], 1), dtype=torch.int32)


@pytest.mark.parametrize("seed,m,B,g,cin,cout", [(0, 1500, 2, 96, 128, 128), (1, 700, 2, 47, 128, 256), (2, 3000, 1, 90, 256, 256)])
def test_tc_sparse_conv(seed, m, B, g, cin, cout, _tc_mode):
    """Forward and backward-data: cp.async-gathered TF32 tcgen05 GEMM in the default tensor-core mode (operands stay
    fp32 in shared memory, the tensor core drops the low 13 bits: tolerance = one TF32 ulp per product, as for the
    dense TMA GEMMs), thread-staged bf16 kernel otherwise (exact products of rounded operands: 1e-4).  The weight
    gradient takes the same two paths (gathered TN GEMM, split over the output rows)."""
    c = _coords(seed, m, B, g)
    m = c.shape[0]
    gen = torch.Generator().manual_seed(seed)
    x = torch.randn(m, cin, generator=gen)
    async_path = _tc_mode == "tma_tf32"
    rnd = tf32 if async_path else bf
    tol_f = (2e-3, 2e-3 * (9 * cin) ** .5 * 0.1) if async_path else (1e-4, 1e-4)
    for subm in (True, False):
        conv = restated.SparseConv(cin, cout, 3, 1 if subm else 2, 1, subm).double()
        w32 = conv.weight.detach().float()
        refs = {}
        for name, r in (("f", rnd), ("w", rnd)):   # references on the operands as the tensor core consumes them
            cv = restated.SparseConv(cin, cout, 3, 1 if subm else 2, 1, subm).double()
            with torch.no_grad():
                cv.weight.copy_(r(w32))
            xr = r(x).requires_grad_()
            refs[name] = (cv, xr, cv(restated.SparseTensor(xr, c, [g, g], B)))
        w = w32.to(DEV)
        if subm:
            table = ops.subm_table(c.to(DEV), B, g, g)
            table_t, flip, rows_out = table, True, m
        else:
            idx_out, n_out, table, table_t, _ = ops.strided_table(c.to(DEV), B, g, g)
            rows_out = int(n_out)
            table, flip = table[:rows_out], False
        y = ops.sparse_conv_fwd(x.to(DEV), table, w, rows_out)
        assert_close(y, refs["f"][2].features.detach(), *tol_f, f"tc sparse conv fwd subm={subm}")
        dy = torch.randn(rows_out, cout, generator=gen)
        refs["f"][2].features.backward(rnd(dy))
        refs["w"][2].features.backward(rnd(dy))
        dx = ops.sparse_conv_fwd(dy.to(DEV), table_t, ops.transpose_taps(w, flip), m)
        assert_close(dx, refs["f"][1].grad, *tol_f, f"tc sparse conv dx subm={subm}")
        dw = ops.sparse_conv_bwd_weight(dy.to(DEV), x.to(DEV), table, w.shape)
        tol_w = (2e-3, 2e-3 * m ** .5) if async_path else (1e-4, 1e-4 * m ** .5)
        assert_close(dw, refs["w"][0].weight.grad, *tol_w, f"tc sparse conv dw subm={subm}")
