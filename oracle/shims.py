"""TEST INFRASTRUCTURE ONLY -- leaf-library shims for the tier-1 oracle.

The reference hot path imports six libraries that are not installed in this
image (torch_scatter, spconv, pytorch3d, easydict, SharedArray) plus its own
compiled `sst_ops_cuda`.  `oracle/ref_loader.py` injects the stand-ins below
into `sys.modules` so that the reference's *own* Python files under
/root/reference run unmodified on CPU.  Nothing in the product package
(`t-mae_b200/`) may import this file; only tests, `__graft_entry__.smoke()` and
the `cpu_baseline` / `--impl reference` legs of `bench.py` do.

Every shim states the third-party semantics it restates:
parity for these three libraries is UNPINNED by the reference (it ships no
tests); the pins are the properties checked in tests/test_oracle.py.

Canonical choices:
  * slot assignment (`ingroup_inds`, `group_inner_inds`) = stable rank by
    original element index inside each group, i.e. what a serial execution of
    pcdet/ops/sst_ops/src/sst_ops_gpu.cu:14-39 yields;
  * strided sparse-conv output rows are in lexicographic (b, y, x) order.
"""
import types

import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------- easydict
class EasyDict(dict):
    """Attribute dict, recursive over dicts and over dicts inside lists
    (SST_BLOCK_LIST is a list of dicts read by attribute, SiamWCA_MAE.py:40-46)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            self[k] = v

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, EasyDict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._wrap(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._wrap(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


# --------------------------------------------------------------------------- torch_scatter
def scatter(src, index, dim=0, reduce="mean", dim_size=None):
    """torch_scatter.scatter(src, index, dim=0, reduce=...): segment reduce along dim 0.
    mean = segment sum / count (true division); used at temporal_dyn_vfe.py:85."""
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    out = src.new_zeros((n,) + tuple(src.shape[1:]))
    if reduce in ("sum", "add", "mean"):
        out.index_add_(0, index, src)
        if reduce == "mean":
            cnt = torch.zeros(n, dtype=src.dtype, device=src.device)
            cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
            out = out / cnt.clamp(min=1).view((n,) + (1,) * (src.dim() - 1))
        return out
    raise NotImplementedError(reduce)


def scatter_max(src, index, dim=0, dim_size=None):
    """torch_scatter.scatter_max(...) -> (out, argmax); the path uses [0] only
    (temporal_dyn_vfe.py:113).  Every segment is non-empty on this path."""
    assert dim == 0
    n = int(index.max()) + 1 if dim_size is None else dim_size
    idx = index.view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    out = src.new_full((n,) + tuple(src.shape[1:]), float("-inf"))
    out = out.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    # argmax: first index attaining the maximum
    hit = src == out[index]
    pos = torch.arange(src.shape[0], device=src.device).view((-1,) + (1,) * (src.dim() - 1)).expand_as(src)
    big = src.shape[0]
    cand = torch.where(hit, pos, torch.full_like(pos, big))
    arg = torch.full_like(out, big, dtype=torch.long).scatter_reduce(0, idx, cand, reduce="amin", include_self=True)
    return out, arg


def scatter_min(src, index, dim=0, dim_size=None):
    o, a = scatter_max(-src, index, dim, dim_size)
    return -o, a


# --------------------------------------------------------------------------- spconv (2-D only)
class SparseConvTensor:
    """The attribute surface spconv_utils.py / spt_backbone.py touch."""

    def __init__(self, features, indices, spatial_shape, batch_size, **kw):
        self.features = features
        self.indices = indices  # (M, 3) int [b, y, x]
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)

    def dense(self):
        Y, X = self.spatial_shape
        C = self.features.shape[1]
        out = self.features.new_zeros(self.batch_size, Y, X, C)
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2]] = self.features
        return out.permute(0, 3, 1, 2).contiguous()


class SparseModule(nn.Module):
    pass


class SparseConvolution(SparseModule):
    """Weight layout (Cout, kh, kw, Cin) -- spconv >= 2.2 native (KRSC)."""

    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None, subm=False):
        super().__init__()
        assert not bias
        self.in_channels, self.out_channels = cin, cout
        self.kernel_size, self.stride, self.padding, self.subm = k, stride, padding, subm
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))
        nn.init.kaiming_uniform_(self.weight, a=5 ** 0.5)

    def forward(self, x):
        k = self.kernel_size
        w = self.weight.permute(0, 3, 1, 2)  # (Cout, Cin, kh, kw): torch cross-correlation
        dense = x.dense()
        if self.subm:
            # SubMConv2d: output only at input sites = zero-filled dense conv sampled there
            out = F.conv2d(dense, w.to(dense.dtype), padding=k // 2)
            i = x.indices.long()
            feats = out[i[:, 0], :, i[:, 1], i[:, 2]]
            return SparseConvTensor(feats, x.indices, x.spatial_shape, x.batch_size)
        # SparseConv2d: active output = any active input in the receptive field
        out = F.conv2d(dense, w.to(dense.dtype), stride=self.stride, padding=self.padding)
        Y, X = x.spatial_shape
        occ = dense.new_zeros(x.batch_size, 1, Y, X)
        i = x.indices.long()
        occ[i[:, 0], 0, i[:, 1], i[:, 2]] = 1
        act = F.max_pool2d(occ, k, self.stride, self.padding)[:, 0] > 0
        idx = act.nonzero()  # lexicographic (b, y, x)
        feats = out[idx[:, 0], :, idx[:, 1], idx[:, 2]]
        return SparseConvTensor(feats, idx.int(), list(out.shape[-2:]), x.batch_size)


class SubMConv2d(SparseConvolution):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None, **kw):
        super().__init__(cin, cout, k, 1, k // 2, bias, indice_key, subm=True)


class SparseConv2d(SparseConvolution):
    def __init__(self, cin, cout, k, stride=1, padding=0, bias=False, indice_key=None, **kw):
        super().__init__(cin, cout, k, stride, padding, bias, indice_key, subm=False)


class SparseSequential(SparseModule):
    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def forward(self, x):
        for m in self._modules.values():
            if isinstance(m, SparseModule):
                x = m(x)
            else:
                x = x.replace_feature(m(x.features))
        return x


# --------------------------------------------------------------------------- pytorch3d.loss
def chamfer_distance(x, y, weights=None, batch_reduction="mean", point_reduction="mean", norm=2):
    """pytorch3d v0.7.1 chamfer_distance for equal-length clouds (SiamWCA_MAE.py:163):
    squared-L2 1-NN both ways, weights applied per cloud, point mean, batch mean divides by sum(w);
    sum(w)==0 -> 0."""
    assert batch_reduction == "mean" and point_reduction == "mean" and norm == 2
    N, P1, _ = x.shape
    P2 = y.shape[1]
    if weights is not None and weights.sum() == 0.0:
        z = (x.sum((1, 2)) * weights).sum() * 0.0
        return z, None
    d = ((x[:, :, None, :] - y[:, None, :, :]) ** 2).sum(-1)  # (N, P1, P2)
    cx = d.min(2).values
    cy = d.min(1).values
    if weights is not None:
        cx = cx * weights.view(N, 1)
        cy = cy * weights.view(N, 1)
    cx = cx.sum(1) / P1
    cy = cy.sum(1) / P2
    div = weights.sum() if weights is not None else max(N, 1)
    return cx.sum() / div + cy.sum() / div, None


# --------------------------------------------------------------------------- sst_ops_cuda
def _stable_rank(group):
    """rank of every element inside its group in original-index order."""
    n = group.shape[0]
    if n == 0:
        return group.clone()
    order = torch.sort(group, stable=True).indices
    g = group[order]
    pos = torch.arange(n, device=group.device)
    start = torch.ones(n, dtype=torch.bool, device=group.device)
    start[1:] = g[1:] != g[:-1]
    seg = torch.cummax(torch.where(start, pos, torch.zeros_like(pos)), 0).values
    rank = torch.empty_like(group)
    rank[order] = (pos - seg).to(group.dtype)
    return rank


def ingroup_inds_wrapper(group_inds, out_inds):
    """serial restatement of sst_ops_gpu.cu:14-20 (canonical arrival order = index order)."""
    out_inds.copy_(_stable_rank(group_inds))
    return 1


def group_inner_inds_wrapper(inverse_inds, group_inds):
    """serial restatement of sst_ops_gpu.cu:22-39: first K point indices per group, cyclic pad."""
    M, K = group_inds.shape
    rank = _stable_rank(inverse_inds)
    cnt = torch.bincount(inverse_inds, minlength=M)
    sel = rank < K
    pt = torch.arange(inverse_inds.shape[0], device=inverse_inds.device)
    group_inds[inverse_inds[sel], rank[sel]] = pt[sel]
    c = cnt.clamp(max=K)
    j = torch.arange(K, device=group_inds.device).view(1, K).expand(M, K)
    # ref pads i in [cnt, K) with slot i % cnt, where cnt is the *unclamped* count; for cnt >= K
    # the loop body never runs.
    src = torch.where(j < c.view(M, 1), j, j % cnt.clamp(min=1).view(M, 1))
    filled = torch.gather(group_inds, 1, src)
    group_inds.copy_(torch.where(cnt.view(M, 1) > 0, filled, group_inds))
    return 1


def make_modules():
    """name -> module objects to inject into sys.modules."""
    mods = {}
    m = types.ModuleType("easydict"); m.EasyDict = EasyDict; mods["easydict"] = m
    mods["SharedArray"] = types.ModuleType("SharedArray")
    m = types.ModuleType("torch_scatter")
    m.scatter, m.scatter_max, m.scatter_min = scatter, scatter_max, scatter_min
    mods["torch_scatter"] = m
    sp = types.ModuleType("spconv"); sp.__path__ = []
    spp = types.ModuleType("spconv.pytorch")
    conv = types.ModuleType("spconv.pytorch.conv"); conv.SparseConvolution = SparseConvolution
    for name in ("SparseConvTensor", "SparseModule", "SparseSequential", "SubMConv2d", "SparseConv2d"):
        setattr(spp, name, globals()[name])
    spp.conv = conv
    sp.pytorch = spp
    mods["spconv"], mods["spconv.pytorch"], mods["spconv.pytorch.conv"] = sp, spp, conv
    p3 = types.ModuleType("pytorch3d"); p3.__path__ = []
    p3l = types.ModuleType("pytorch3d.loss"); p3l.chamfer_distance = chamfer_distance
    p3.loss = p3l
    mods["pytorch3d"], mods["pytorch3d.loss"] = p3, p3l
    so = types.ModuleType("pcdet.ops.sst_ops.sst_ops_cuda")
    so.ingroup_inds_wrapper, so.group_inner_inds_wrapper = ingroup_inds_wrapper, group_inner_inds_wrapper
    mods["pcdet.ops.sst_ops.sst_ops_cuda"] = so
    return mods
