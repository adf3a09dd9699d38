"""Sparse 2-D tensors and the conv+BN+ReLU block of the path (spconv is not a dependency).

`SparseConvTensor` carries what pcdet reads from spconv's class of the same name (`features`, `indices`,
`spatial_shape`, `batch_size`, `replace_feature`, `dense`); pcdet/utils/spconv_utils.py:29-35, SiamWCA_MAE.py:187-235.
`ConvBNReLU` is `post_act_block` (spconv_utils.py:37-56): children named "0" (conv weight, spconv >= 2.2 layout
(Cout, kh, kw, Cin)), "1" (BatchNorm1d eps 1e-3 momentum 0.01), "2" (ReLU) so state_dict keys match.
"""
import math

import torch
import torch.nn as nn

from . import ops
from .vfe import bn_relu


class _DensifyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rows, indices, batch, Y, X, dtype):
        ctx.save_for_backward(indices)
        ctx.b16 = rows.dtype == torch.bfloat16
        if ctx.b16:   # bf16-storage mode: 16-bit rows into a 16-bit map, plain 16-byte copies
            if dtype != torch.bfloat16:
                raise RuntimeError("bf16 rows densify into a bf16 map (set decoder_autocast = torch.bfloat16)")
            return ops.densify_nhwc_b16(rows.contiguous(), indices, batch, Y, X)
        return ops.densify_nhwc(rows.contiguous(), indices, batch, Y, X, dtype)

    @staticmethod
    def backward(ctx, d):
        (indices,) = ctx.saved_tensors
        if ctx.b16:
            d = d.contiguous()
            return ops.gather_nhwc_b16(d if d.dtype == torch.bfloat16 else d.bfloat16(), indices), None, None, None, None, None
        return ops.gather_nhwc(d.contiguous(), indices), None, None, None, None, None


class _GatherNhwcFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dense, indices):
        ctx.save_for_backward(indices)
        ctx.shape, ctx.dtype = dense.shape, dense.dtype
        return ops.gather_nhwc(dense, indices)

    @staticmethod
    def backward(ctx, drows):
        (indices,) = ctx.saved_tensors
        B, Y, X, _ = ctx.shape
        return ops.densify_nhwc(drows.contiguous(), indices, B, Y, X, ctx.dtype), None


def gather_bev(spatial_features, indices):
    """spatial_features (B,C,Y,X) -> rows (M,C) at indices (M,3) [b,y,x]  (SiamWCA_MAE.py:311-312)."""
    nhwc = spatial_features.permute(0, 2, 3, 1).contiguous()  # free when the map is channels_last
    return _GatherNhwcFn.apply(nhwc, indices)


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features = features
        self.indices = indices  # (M,3) int32 [b, y, x], ascending lexicographic
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)

    def replace_feature(self, f):
        return SparseConvTensor(f, self.indices, self.spatial_shape, self.batch_size)

    def dense(self, dtype=torch.float32):
        """(B, C, Y, X), stored channels-last; dtype fp32 or bf16 (bf16 features need a bf16 map)."""
        Y, X = self.spatial_shape
        return _DensifyFn.apply(self.features, self.indices, self.batch_size, Y, X, dtype).permute(0, 3, 1, 2)


class _SparseConvFn(torch.autograd.Function):
    """Gather-GEMM sparse convolution; `table` (rows_out, 9) for forward, `table_t` (rows_in, 9) for backward-data
    (for submanifold convs table_t is table and the taps are flipped)."""

    @staticmethod
    def forward(ctx, x, w, table, table_t, flip, rows_out):
        ctx.save_for_backward(x, w, table, table_t)
        ctx.flip = flip
        return ops.sparse_conv_fwd(x, table, w, rows_out)

    @staticmethod
    def backward(ctx, dy):
        x, w, table, table_t = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.sparse_conv_fwd(dy, table_t, ops.transpose_taps(w, ctx.flip), x.shape[0])
        dw = ops.sparse_conv_bwd_weight(dy, x, table, w.shape)
        return dx, dw, None, None, None, None


class _SparseConvFnBF16(torch.autograd.Function):
    """_SparseConvFn in the bf16-storage mode: bf16 rows, bf16 copies of the fp32 master weight (ops.shadows), gathered
    tcgen05 kind::f16 GEMMs, fp32 weight gradient."""

    @staticmethod
    def forward(ctx, x, w, table, table_t, flip, rows_out):
        ctx.save_for_backward(x, w, table, table_t)
        ctx.flip = flip
        return ops.bf16_sparse_conv_fwd(x, table, ops.shadows.get(w), rows_out)

    @staticmethod
    def backward(ctx, dy):
        x, w, table, table_t = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.bf16_sparse_conv_fwd(dy, table_t, ops.transpose_taps_bf16(w, ctx.flip), x.shape[0])
        dw = ops.bf16_sparse_conv_bwd_weight(dy, x, table, w.shape)
        return dx, dw, None, None, None, None


class SparseConvWeight(nn.Module):
    def __init__(self, cin, cout, k=3):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))


class ConvBNReLU(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.add_module("0", SparseConvWeight(cin, cout))
        self.add_module("1", nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01))
        self.add_module("2", nn.ReLU())

    def forward(self, feats, table, table_t, flip, rows_out, bounds=None, order=None):
        fn = _SparseConvFnBF16 if feats.dtype == torch.bfloat16 else _SparseConvFn
        y = fn.apply(feats.contiguous(), self._modules["0"].weight, table, table_t, flip, rows_out)
        return bn_relu(y, self._modules["1"], relu=True, bounds=bounds, order=order)
