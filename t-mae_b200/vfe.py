"""Drop-in dynamic VFE modules (pcdet API): `DynVFE` and `TemporalDynVFE`.

Mirror of pcdet/models/backbones_3d/vfe/{dyn_vfe.py, temporal_dyn_vfe.py}: same constructor
signature, same `forward(batch_dict) -> batch_dict` keys, same parameter names
(`dvfe_mlps.0.{0,3}.weight`, `dvfe_mlps.0.{1,4}.*`), so pcdet's `vfe.__all__` registry and released
checkpoints work unchanged.  All arithmetic runs in libtmae_sm100.so (voxelize, point features, the two
Linear+BatchNorm+ReLU layers, per-voxel max); only `TYPE: mean` with one MLP group is on the T-MAE path
(tools/cfgs/once_models/t_mae_ssl.yaml:44-53) and anything else raises.
"""
import torch
import torch.nn as nn

from . import ops


class _LinearFn(torch.autograd.Function):
    """y = x @ w.T (no bias): network_utils.py:30."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return ops.linear_fwd(x, w)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.linear_bwd_data(dy, w) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w)
        ops.linear_bwd_weight(dy, x, dw)
        return dx, dw


class _BnReluFn(torch.autograd.Function):
    """BatchNorm1d over rows followed by ReLU (network_utils.py:31-33, spconv_utils.py:50-54)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, training, relu):
        if training:
            y, mean, rstd = ops.bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, relu)
        else:
            mean, rstd = running_mean, torch.rsqrt(running_var + eps)
            y = ops.bn_apply(x, mean, rstd, gamma, beta, relu)
        ctx.save_for_backward(x, beta, mean, rstd, gamma)   # the ReLU mask is recomputed from x: y is not kept
        ctx.flags = (relu, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, beta, mean, rstd, gamma = ctx.saved_tensors
        relu, training = ctx.flags
        dx, dg, db = ops.bn_bwd(dy.contiguous(), x, beta, mean, rstd, gamma, relu, training)
        return dx, dg, db, None, None, None, None, None, None


class _BnReluSegFn(torch.autograd.Function):
    """The same BatchNorm1d + ReLU applied to several row segments of one tensor, each with its OWN batch
    statistics -- what the reference computes when it calls the module once per frame (current and previous frame
    share the Siamese encoder weights, SiamWCA_MAE.py:265-284); `order` is the reference's call order, which fixes
    the running-statistics update sequence."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, training, relu, bounds, order):
        y = torch.empty_like(x)
        stats = [None] * (len(bounds) - 1)
        for i in order:
            a, b = bounds[i], bounds[i + 1]
            if b <= a:
                continue
            if training:
                _, mean, rstd = ops.bn_train_fwd(x[a:b], gamma, beta, running_mean, running_var, momentum, eps, relu, out=y[a:b])
            else:
                mean, rstd = running_mean, torch.rsqrt(running_var + eps)
                ops.bn_apply(x[a:b], mean, rstd, gamma, beta, relu, out=y[a:b])
            stats[i] = (mean, rstd)
        ctx.save_for_backward(x, beta, gamma, *[t for s in stats if s is not None for t in s])
        ctx.misc = (relu, training, bounds, [s is not None for s in stats])
        return y

    @staticmethod
    def backward(ctx, dy):
        x, beta, gamma, *flat = ctx.saved_tensors
        relu, training, bounds, present = ctx.misc
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = db = None
        k = 0
        for i, has in enumerate(present):
            if not has:
                continue
            mean, rstd = flat[2 * k], flat[2 * k + 1]
            k += 1
            a, b = bounds[i], bounds[i + 1]
            _, g, bb = ops.bn_bwd(dy[a:b], x[a:b], beta, mean, rstd, gamma, relu, training, out=dx[a:b])
            dg, db = (g, bb) if dg is None else (dg + g, db + bb)
        return dx, dg, db, None, None, None, None, None, None, None, None


class _BnReluSegFnBF16(torch.autograd.Function):
    """_BnReluSegFn on bf16 rows (the bf16-storage mode after the sparse convolutions): fp32 statistics and arithmetic,
    bf16 in and out (tmae_bn_bf16_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, momentum, eps, training, relu, bounds, order):
        y = torch.empty_like(x)
        stats = [None] * (len(bounds) - 1)
        for i in order:
            a, b = bounds[i], bounds[i + 1]
            if b <= a:
                continue
            if training:
                mean, rstd = ops.bn_bf16_fwd(x[a:b], gamma, beta, running_mean, running_var, momentum, eps, relu, True, y[a:b])
            else:
                mean, rstd = running_mean, torch.rsqrt(running_var + eps)
                ops.bn_bf16_fwd(x[a:b], gamma, beta, None, None, 0.0, eps, relu, False, y[a:b], mean, rstd)
            stats[i] = (mean, rstd)
        ctx.save_for_backward(x, beta, gamma, *[t for s in stats if s is not None for t in s])
        ctx.misc = (relu, training, bounds, [s is not None for s in stats])
        return y

    @staticmethod
    def backward(ctx, dy):
        x, beta, gamma, *flat = ctx.saved_tensors
        relu, training, bounds, present = ctx.misc
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dg = db = None
        k = 0
        for i, has in enumerate(present):
            if not has:
                continue
            mean, rstd = flat[2 * k], flat[2 * k + 1]
            k += 1
            a, b = bounds[i], bounds[i + 1]
            _, g, bb = ops.bn_bf16_bwd(dy[a:b], x[a:b], mean, rstd, gamma, beta, relu, training, out=dx[a:b])
            dg, db = (g, bb) if dg is None else (dg + g, db + bb)
        return dx, dg, db, None, None, None, None, None, None, None, None


def bn_relu(x, bn, relu=True, bounds=None, order=None):
    """Applies an nn.BatchNorm1d's parameters/buffers with the library kernels (train or eval).  `bounds` = row
    boundaries [0, m0, m0+m1, ...] of segments normalised independently, visited in `order`."""
    training = bn.training or bn.running_mean is None
    n_seg = 1 if bounds is None else sum(1 for i in range(len(bounds) - 1) if bounds[i + 1] > bounds[i])
    if training and bn.num_batches_tracked is not None:
        bn.num_batches_tracked += n_seg
    if x.dtype == torch.bfloat16:
        bounds = (0, x.shape[0]) if bounds is None else bounds
        order = list(range(len(bounds) - 1)) if order is None else order
        return _BnReluSegFnBF16.apply(x.contiguous(), bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, training, relu,
                                      tuple(int(b) for b in bounds), tuple(order))
    if bounds is None:
        return _BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, training, relu)
    order = list(range(len(bounds) - 1)) if order is None else order
    return _BnReluSegFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, training, relu,
                              tuple(int(b) for b in bounds), tuple(order))


class _SegMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, voxel_offset, pt_order, n_voxels):
        out, arg = ops.segment_max_fwd(x, voxel_offset, pt_order, n_voxels)
        ctx.save_for_backward(arg)
        ctx.n_points = x.shape[0]
        return out

    @staticmethod
    def backward(ctx, dout):
        (arg,) = ctx.saved_tensors
        return ops.segment_max_bwd(dout.contiguous(), arg, ctx.n_points), None, None, None


def _fc_bn_relu(cfg, c_in):
    layers = []
    for c in cfg:
        layers += [nn.Linear(c_in, c, bias=False), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
        c_in = c
    return nn.Sequential(*layers)


class TemporalDynVFE(nn.Module):
    """pcdet `TemporalDynVFE` (temporal_dyn_vfe.py:11-163) on the sm_100a kernels."""

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        get = model_cfg.get if hasattr(model_cfg, "get") else (lambda k, d=None: getattr(model_cfg, k, d))
        if get("TYPE", "mean") != "mean":
            raise NotImplementedError("only TYPE: mean is on the T-MAE path")
        mlps = get("MLPS", None)
        if mlps is None or len(mlps) != 1 or get("WITH_DISTANCE", False) or not get("USE_ABSLOTE_XYZ", True) \
                or not get("USE_CLUSTER_XYZ", True) or get("AGGREGATION_MLPS", None) is not None:
            raise NotImplementedError("VFE config outside the T-MAE path (t_mae_ssl.yaml:44-53)")
        nf = self._raw_features(num_point_features)
        self.point_stride = nf + 1
        self.dvfe_mlps = nn.ModuleList([_fc_bn_relu(list(mlps[0]), nf + 6)])
        self.finetuning = bool(get("FT", False))
        self.num_point_features = int(mlps[0][-1])
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.grid_size = [int(v) for v in grid_size]
        self.last = {}

    @staticmethod
    def _raw_features(num_point_features):
        return num_point_features - 1  # temporal_dyn_vfe.py:16 (group_id column)

    def get_output_feature_dim(self):
        return self.num_point_features

    # -- stage 1: integer work for both frames, one host read of the counts -------------------
    def _voxelize(self, points, batch_size):
        if points.dtype != torch.float32 or not points.is_cuda:
            raise RuntimeError("points must be a float32 CUDA tensor")
        if points.shape[1] != self.point_stride:
            raise RuntimeError(f"points must have {self.point_stride} columns [b,x,y,z,feat...]")
        return ops.voxelize(points.contiguous(), self.point_cloud_range, self.voxel_size, self.grid_size, batch_size)

    def _features(self, v, n_kept, n_vox):
        pts, coords, inv = v["points"][:n_kept], v["point_coords"][:n_kept], v["inverse"][:n_kept]
        mean = v["voxel_mean"][:n_vox]
        x = ops.vfe_point_features(pts, coords, inv, mean, self.point_cloud_range, self.voxel_size)
        seq = self.dvfe_mlps[0]
        for i in range(0, len(seq), 3):
            x = _LinearFn.apply(x, seq[i].weight)
            x = bn_relu(x, seq[i + 1], relu=True)
        return _SegMaxFn.apply(x, v["voxel_offset"], v["pt_order"], n_vox)

    def _run(self, frames, batch_size, side=None, ready=None):
        if side is None:
            vox = [self._voxelize(p, batch_size) for p in frames]
            counts = torch.stack([v["counts"] for v in vox]).cpu()  # the one host sync of the VFE
        else:
            # opt-in (batch_dict["side_stream"]): the inputs are ready on `side`; voxelise there and wait only for it
            main = torch.cuda.current_stream()
            if ready is not None:
                side.wait_event(ready)          # inputs produced on another stream (FrameAssembler, an H2D copy on main)
            for p in frames:
                p.record_stream(side)           # a no-op for tensors allocated on `side`
            with torch.cuda.stream(side):
                vox = [self._voxelize(p, batch_size) for p in frames]
                counts = torch.stack([v["counts"] for v in vox]).cpu()
            main.wait_stream(side)
            ops.record_all(vox, main)
        out = []
        for v, c in zip(vox, counts):
            n_kept, n_vox = int(c[0]), int(c[1])
            if n_vox == 0:
                raise RuntimeError("no point falls inside the point-cloud range")
            feats = self._features(v, n_kept, n_vox)
            starts = c[2:].tolist() + [n_vox]
            out.append(dict(points=v["points"][:n_kept], point_coords=v["point_coords"][:n_kept],
                            inverse=v["inverse"][:n_kept], voxel_coords=v["voxel_coords"][:n_vox], voxel_features=feats,
                            csr=(v["voxel_offset"], v["pt_order"], v["voxel_npts"][:n_vox]),
                            voxels_per_sample=[starts[i + 1] - starts[i] for i in range(batch_size)]))
        return out

    def _batch_size(self, batch_dict):
        return int(batch_dict["batch_size"])

    def forward(self, batch_dict, **kwargs):
        B = self._batch_size(batch_dict)
        cur, prev = self._run([batch_dict["points"], batch_dict["points_prev"]], B, batch_dict.get("side_stream"), batch_dict.get("inputs_ready_event"))
        for sfx, r in (("", cur), ("_prev", prev)):
            if self.finetuning:  # temporal_dyn_vfe.py:121-123,154-160
                batch_dict.pop("points" + sfx, None)
            else:
                batch_dict["points" + sfx] = r["points"]
                batch_dict["point_coords" + sfx] = r["point_coords"]
                batch_dict["point_inverse_indices" + sfx] = r["inverse"]
            batch_dict["voxel_coords" + sfx] = r["voxel_coords"]
            batch_dict["voxel_features" + sfx] = r["voxel_features"]
            # extra keys (ignored by pcdet): per-voxel point CSR and per-sample voxel counts, so the
            # backbone needs no .max().item() / .sum().item() round trips (sst_ops_utils.py:23, SiamWCA_MAE.py:172)
            batch_dict["voxel_point_csr" + sfx] = r["csr"]
            batch_dict["voxels_per_sample" + sfx] = r["voxels_per_sample"]
        return batch_dict


class DynVFE(TemporalDynVFE):
    """pcdet `DynVFE` (dyn_vfe.py:11-125): single frame; also writes `pillar_features`."""

    @staticmethod
    def _raw_features(num_point_features):
        return num_point_features  # dyn_vfe.py:22: no group_id column

    def forward(self, batch_dict, **kwargs):
        (r,) = self._run([batch_dict["points"]], self._batch_size(batch_dict), batch_dict.get("side_stream"), batch_dict.get("inputs_ready_event"))
        batch_dict.update(points=r["points"], point_coords=r["point_coords"], point_inverse_indices=r["inverse"],
                          voxel_coords=r["voxel_coords"], pillar_features=r["voxel_features"],
                          voxel_features=r["voxel_features"], voxel_point_csr=r["csr"],
                          voxels_per_sample=r["voxels_per_sample"])
        return batch_dict
