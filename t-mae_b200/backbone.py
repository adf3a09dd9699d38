"""Drop-in backbone_3d modules (pcdet API): `SiamWCA` (finetune encoder) and `SiamWCA_MAE` (pretraining).

Mirrors pcdet/models/backbones_3d/{spt_backbone.py, SiamWCA.py, SiamWCA_MAE.py} and
pcdet/models/model_utils/{sst_basic_block.py, wca_block.py, cosine_msa.py}: same constructor signature,
`forward(batch_dict)` keys, `get_loss()`, `num_point_features`, and the same parameter names
(`sst_blocks.{s}.encoder_blocks.{b}.encoder_list.{k}.win_attn.self_attn.in_proj_weight`, ...), so pcdet's
`backbones_3d.__all__` registry and released checkpoints work unchanged.

What differs is how it runs.  All coordinate-only work is done once per forward by `plan.build_plans`
(one host read), features stay in the flat voxel-major layout for the whole encoder (the padded
(windows, tokens, C) tensors of flat2window are never materialised), and every encoder layer is one
autograd node that launches the library kernels in sequence.  The dense decoder (ConvTranspose2d /
Conv2d + BatchNorm2d, SiamWCA_MAE.py:79-115) stays on cuDNN in channels-last layout.
"""
import os

import numpy as np
import torch
import torch.nn as nn

from . import ops
from .plan import build_plans, validate_levels
from .sparse import ConvBNReLU, SparseConvTensor, gather_bev
from .vfe import _LinearFn  # noqa: F401


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def pos_embed_table(C, temperature, win=(8, 8), normalize=False):
    """64-row table of the window position embedding, row = ly*8 + lx  (spt_backbone.py:186-222:
    the embedding depends only on the in-window cell, so it is built once with the reference's own op
    sequence and looked up in-kernel)."""
    wx, wy = win
    ly, lx = torch.meshgrid(torch.arange(wy), torch.arange(wx), indexing="ij")
    y, x = ly.reshape(-1) - wy / 2, lx.reshape(-1) - wx / 2
    if normalize:
        x, y = x / wx * 2 * 3.1415, y / wy * 2 * 3.1415
    L = C // 2
    inv_freq = torch.arange(L, dtype=torch.float32)
    inv_freq = temperature ** (2 * torch.div(inv_freq, 2, rounding_mode="floor") / L)
    ex, ey = x[:, None] / inv_freq[None, :], y[:, None] / inv_freq[None, :]
    ex = torch.stack([ex[:, ::2].sin(), ex[:, 1::2].cos()], dim=-1).flatten(1)
    ey = torch.stack([ey[:, ::2].sin(), ey[:, 1::2].cos()], dim=-1).flatten(1)
    return torch.cat([ex, ey], dim=-1).float().contiguous()


# ============================================================================ encoder layers
class CosineMHAParams(nn.Module):
    """Parameters of CosineMultiheadAttention (cosine_msa.py:441-458): packed in_proj (3C, C), out_proj, tau."""

    def __init__(self, C, H, tau_min=0.01):
        super().__init__()
        self.embed_dim, self.num_heads, self.tau_min = C, H, tau_min
        self.in_proj_weight = nn.Parameter(torch.empty(3 * C, C))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * C))
        self.out_proj = nn.Linear(C, C)
        self.tau = nn.Parameter(torch.ones(1, 1, 1))
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)


class _AttnHolder(nn.Module):
    def __init__(self, name, C, H, tau_min):
        super().__init__()
        setattr(self, name, CosineMHAParams(C, H, tau_min))


class _LayerFn(torch.autograd.Function):
    """One encoder layer = one library call each way (tmae_encoder_layer_fwd / _bwd): the SST self-attention layer
    (sst_basic_block.py:58-84) when x_kv is None, the WCA cross-attention layer (wca_block.py:70-103) otherwise."""

    @staticmethod
    def forward(ctx, x, x_kv, part, shift, lut, heads, tau_min, eps, *params):
        x = x.contiguous()
        if x_kv is not None:
            x_kv = x_kv.contiguous()
        T = ops.layer_tables(part, shift, x_kv is not None, x.shape[0], x_kv.shape[0] if x_kv is not None else 0)
        need_bwd = any(ctx.needs_input_grad)
        y, saved = ops.encoder_layer_fwd(x, x_kv, params, T, lut, tau_min, eps, heads, need_bwd)
        ctx.save_for_backward(x, x_kv, saved, lut, *params)
        ctx.misc = (T, part, heads, tau_min)  # part keeps the device tables alive
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x_kv, saved, lut, *params = ctx.saved_tensors
        T, _, heads, tau_min = ctx.misc
        dx, dkv, grads = ops.encoder_layer_bwd(dy.contiguous(), x, x_kv, params, T, lut, tau_min, heads, saved,
                                               x_kv is not None and ctx.needs_input_grad[1])
        return (dx, dkv, None, None, None, None, None, None, *grads)


class _LayerFnBF16(torch.autograd.Function):
    """_LayerFn in the bf16-storage mode (tmae_bf16_encoder_layer_fwd / _bwd): bf16 rows in and out, fp32 master parameters and
    parameter gradients, bf16 weight copies from ops.shadows."""

    @staticmethod
    def forward(ctx, x, x_kv, part, shift, lut, heads, tau_min, eps, *params):
        x = x.contiguous()
        if x_kv is not None:
            x_kv = x_kv.contiguous()
        T = ops.layer_tables(part, shift, x_kv is not None, x.shape[0], x_kv.shape[0] if x_kv is not None else 0)
        need_bwd = any(ctx.needs_input_grad)
        y, saved = ops.encoder_layer_fwd_bf16(x, x_kv, params, T, lut, tau_min, eps, heads, need_bwd)
        ctx.save_for_backward(x, x_kv, saved, lut, *params)
        ctx.misc = (T, part, heads, tau_min)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x_kv, saved, lut, *params = ctx.saved_tensors
        T, _, heads, tau_min = ctx.misc
        dx, dkv, grads = ops.encoder_layer_bwd_bf16(dy.contiguous(), x, x_kv, params, T, lut, tau_min, heads, saved,
                                                    x_kv is not None and ctx.needs_input_grad[1])
        return (dx, dkv, None, None, None, None, None, None, *grads)


class EncoderLayer(nn.Module):
    def __init__(self, C, H, FF, layer_cfg, cross):
        super().__init__()
        self.cross, self.nhead = cross, H
        if not _get(layer_cfg, "cosine", False):
            raise NotImplementedError("only cosine attention is on the T-MAE path (t_mae_ssl.yaml:86-89)")
        if _get(layer_cfg, "non_shared_tau", False):
            raise NotImplementedError("non_shared_tau is not used by the T-MAE configs")
        self.tau_min = float(_get(layer_cfg, "tau_min", 0.01))
        self.win_attn = _AttnHolder("cross_attn" if cross else "self_attn", C, H, self.tau_min)
        self.linear1, self.linear2 = nn.Linear(C, FF), nn.Linear(FF, C)
        self.norm1, self.norm2 = nn.LayerNorm(C), nn.LayerNorm(C)

    def _params(self):
        at = self.win_attn.cross_attn if self.cross else self.win_attn.self_attn
        return (at.in_proj_weight, at.in_proj_bias, at.out_proj.weight, at.out_proj.bias, at.tau, self.norm1.weight, self.norm1.bias,
                self.linear1.weight, self.linear1.bias, self.linear2.weight, self.linear2.bias, self.norm2.weight, self.norm2.bias)

    def forward_self(self, x, part, shift, lut):
        fn = _LayerFnBF16 if x.dtype == torch.bfloat16 else _LayerFn
        return fn.apply(x, None, part, shift, lut, self.nhead, self.tau_min, self.norm1.eps, *self._params())

    def forward_cross(self, x, xprev, tp, shift, lut):
        fn = _LayerFnBF16 if x.dtype == torch.bfloat16 else _LayerFn
        return fn.apply(x, xprev, tp, shift, lut, self.nhead, self.tau_min, self.norm1.eps, *self._params())


class ShiftBlock(nn.Module):
    """BasicShiftBlockV2 (sst_basic_block.py:87-114) / BasicShiftBlock_WCA (wca_block.py:106-145)."""

    def __init__(self, C, H, FF, layer_cfg, cross):
        super().__init__()
        self.encoder_list = nn.ModuleList([EncoderLayer(C, H, FF, layer_cfg, cross) for _ in range(2)])


class _EncBlockBase(nn.Module):
    def _common(self, cfg):
        enc, pre = cfg["ENCODER"], cfg["PREPROCESS"]
        if list(pre["WINDOW_SHAPE"]) != [8, 8, 1]:
            raise NotImplementedError("the kernels are specialised to 8x8x1 windows (t_mae_ssl.yaml:61)")
        if pre["SHUFFLE_VOXELS"]:
            raise NotImplementedError("SHUFFLE_VOXELS is False in both T-MAE configs (t_mae_ssl.yaml:74)")
        if enc["ACTIVATION"] != "gelu" or enc["DROPOUT"] != 0.0:
            raise NotImplementedError("encoder uses GELU and no dropout on the T-MAE path")
        validate_levels(pre)
        self.d_model = enc["D_MODEL"]
        self.register_buffer("pos_lut", pos_embed_table(self.d_model, pre["POS_TEMPERATURE"], normalize=pre["NORMALIZE_POS"]),
                             persistent=False)
        return enc


class SSTBlockV1(_EncBlockBase):
    """spt_backbone.py:267-353: [conv_down] -> 2 x (shift-0 layer, shift-1 layer) -> conv_out(x + enc(x))."""

    def __init__(self, cfg, cin, indice_key=None, **kw):
        super().__init__()
        enc = self._common(cfg)
        C = self.d_model
        self.conv_down = ConvBNReLU(cin, C) if enc["STRIDE"] > 1 else None
        self.encoder_blocks = nn.ModuleList(
            [ShiftBlock(C, enc["NHEAD"], enc["DIM_FEEDFORWARD"], enc["LAYER_CFG"], False) for _ in range(enc["NUM_BLOCKS"])])
        self.conv_out = ConvBNReLU(C, C)

    def forward(self, feats, stage, bounds=None, order=None):
        """`bounds` / `order`: row segments (frames of the Siamese-batched set) whose BatchNorm statistics stay separate."""
        if self.conv_down is not None:
            feats = self.conv_down(feats, stage.down, stage.down_t, False, stage.m, bounds, order)
        x = feats
        for blk in self.encoder_blocks:
            for s, layer in enumerate(blk.encoder_list):
                x = layer.forward_self(x, stage.part, s, self.pos_lut)
        return self.conv_out(feats + x, stage.subm, stage.subm, True, stage.m, bounds, order)


class WCABlock(_EncBlockBase):
    """SiamWCA.py:272-447: one BasicShiftBlock_WCA (NUM_BLOCKS forced to 1, :294-296) -> conv_out(x + enc(x))."""

    def __init__(self, cfg, cin, indice_key=None, **kw):
        super().__init__()
        enc = self._common(cfg)
        C = self.d_model
        n = 1 if enc["NUM_BLOCKS"] == 2 else enc["NUM_BLOCKS"]
        self.encoder_blocks = nn.ModuleList([ShiftBlock(C, enc["NHEAD"], enc["DIM_FEEDFORWARD"], enc["LAYER_CFG"], True) for _ in range(n)])
        self.conv_out = ConvBNReLU(C, C)

    def forward(self, feats, feats_prev, stage, tp):
        x = feats
        for s, layer in enumerate(self.encoder_blocks[0].encoder_list):
            x = layer.forward_cross(x, feats_prev, tp, s, self.pos_lut)
        return self.conv_out(feats + x, stage.subm, stage.subm, True, stage.m)


class _Bn2dReluCatFn(torch.autograd.Function):
    """BatchNorm2d + ReLU of several channels-last bf16 maps, each with its own parameters and batch statistics,
    written side by side into ONE (B, Y, X, sum C) buffer: the reference's `torch.cat([deblock_i(x_i)], dim=1)`
    (SiamWCA_MAE.py:246-249) without BatchNorm / ReLU / cat passes of their own.  Inputs are the raw
    ConvTranspose2d / Conv2d outputs (cuDNN).  args = n maps followed by (weight, bias) of every BatchNorm."""

    @staticmethod
    def forward(ctx, training, bns, n, *args):
        xs = args[:n]
        B, _, Y, X = xs[0].shape
        rows = B * Y * X
        cs = [x.shape[1] for x in xs]
        xr = [x.permute(0, 2, 3, 1).contiguous().view(rows, c) for x, c in zip(xs, cs)]  # free for channels-last maps
        out = torch.empty(rows, sum(cs), dtype=torch.bfloat16, device=xs[0].device)
        saved, off = [], 0
        for x, c, bn in zip(xr, cs, bns):
            if training:
                mean, rstd = ops.bn_bf16_fwd(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.momentum, bn.eps, True, True,
                                             out[:, off:off + c])
            else:
                mean, rstd = bn.running_mean, torch.rsqrt(bn.running_var + bn.eps)
                ops.bn_bf16_fwd(x, bn.weight, bn.bias, None, None, 0.0, bn.eps, True, False, out[:, off:off + c], mean, rstd)
            saved += [x, mean, rstd, bn.weight, bn.bias]
            off += c
        ctx.save_for_backward(*saved)
        ctx.misc = (training, cs, (B, Y, X))
        return out.view(B, Y, X, sum(cs)).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, d):
        training, cs, (B, Y, X) = ctx.misc
        rows = B * Y * X
        dr = d.permute(0, 2, 3, 1).contiguous().view(rows, sum(cs))
        if dr.dtype != torch.bfloat16:
            dr = dr.bfloat16()
        grads_x, grads_p, off = [], [], 0
        sv = ctx.saved_tensors
        for i, c in enumerate(cs):
            x, mean, rstd, gamma, beta = sv[5 * i:5 * i + 5]
            dx, dg, db = ops.bn_bf16_bwd(dr[:, off:off + c], x, mean, rstd, gamma, beta, True, training)
            grads_x.append(dx.view(B, Y, X, c).permute(0, 3, 1, 2))
            grads_p += [dg, db]
            off += c
        return (None, None, None, *grads_x, *grads_p)


def _bn2d_relu_cat(bns, xs):
    """Applies the nn.BatchNorm2d modules `bns` (+ReLU) to the bf16 maps `xs` and concatenates along channels."""
    training = bns[0].training
    if training:
        for bn in bns:
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
    return _Bn2dReluCatFn.apply(training, bns, len(xs), *xs, *[p for bn in bns for p in (bn.weight, bn.bias)])


def _deblocks(cfg):
    blocks, cin = nn.ModuleList(), 0
    for src in cfg["FEATURES_SOURCE"]:
        c = cfg["FUSE_LAYER"][src]
        blocks.append(nn.Sequential(
            nn.ConvTranspose2d(c["NUM_FILTER"], c["NUM_UPSAMPLE_FILTER"], c["UPSAMPLE_STRIDE"], stride=c["UPSAMPLE_STRIDE"], bias=False),
            nn.BatchNorm2d(c["NUM_UPSAMPLE_FILTER"], eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
        cin += c["NUM_UPSAMPLE_FILTER"]
    out = nn.Sequential(nn.Conv2d(cin, cin // len(blocks), 3, padding=1, bias=False),
                        nn.BatchNorm2d(cin // len(blocks), eps=1e-3, momentum=0.01), nn.ReLU(inplace=True))
    return blocks, out, cin // len(blocks)


class SiamWCA(nn.Module):
    """pcdet `SiamWCA` (SiamWCA.py:450-667), finetune-mode encoder: encode prev, encode cur, WCA x3, dense fuse."""
    _deblocks_name, _conv_out_name = "deblocks", "conv_out"

    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.grid_size = np.asarray(grid_size)
        self.voxel_size = [float(v) for v in voxel_size]
        self.point_cloud_range = [float(v) for v in point_cloud_range]
        self.sparse_shape = [int(grid_size[1]), int(grid_size[0])]
        asym = _get(model_cfg, "ASYMMETRIC", False)
        if asym and asym["ENABLED"]:
            raise NotImplementedError("ASYMMETRIC encoders are not enabled in either T-MAE config")
        self.block_cfgs = list(model_cfg["SST_BLOCK_LIST"])
        cin = input_channels
        self.sst_blocks = nn.ModuleList()
        for c in self.block_cfgs:
            self.sst_blocks.append(SSTBlockV1(c, cin, c["NAME"]))
            cin = c["ENCODER"]["D_MODEL"]
        self.wca_blocks = nn.ModuleList([WCABlock(c, c["ENCODER"]["D_MODEL"], c["NAME"]) for c in self.block_cfgs])
        de, out, self.num_point_features = _deblocks(model_cfg)
        setattr(self, self._deblocks_name, de)
        setattr(self, self._conv_out_name, out)
        self.decoder_autocast = None  # e.g. torch.bfloat16 for the cuDNN decoder in throughput runs
        self.fused_decoder_bn = os.environ.get("TMAE_FUSED_DECODER_BN", "1") != "0"
        self.debug_refs = False       # emit reference-format partition tables and check the status word
        self.siamese_batched = os.environ.get("TMAE_SIAMESE", "1") != "0"  # both frames through the shared SST blocks as one row set
        self.last_plan = None

    # ---- pieces -------------------------------------------------------------------------------
    @staticmethod
    def _indices(coords):
        # (bs_idx, y_idx, x_idx), SiamWCA.py:553-557; column views, no index tensor (a Python list index costs a
        # host-to-device copy and a sync per call)
        return torch.stack((coords[:, 0], coords[:, 2], coords[:, 3]), 1).int()

    def _bf16(self):
        """bf16-storage mode (ops.set_precision("bf16")): encoder rows travel as bf16 from the VFE output to the BEV map."""
        if ops.precision() != ops.PREC_BF16:
            return False
        if self.decoder_autocast != torch.bfloat16:
            raise RuntimeError("the bf16-storage mode needs the bf16 decoder (decoder_autocast = torch.bfloat16)")
        return True

    def _encode(self, feats, fp):
        hidden = []
        x = feats.to(torch.bfloat16) if self._bf16() else feats
        for blk, st in zip(self.sst_blocks, fp.stages):
            x = blk(x, st)
            hidden.append(x)
        return hidden

    def _encode_siamese(self, feats, feats_prev, fp_all, fp_cur):
        """Both frames through the shared-weight SST blocks as ONE row set [current rows ; previous rows] (samples
        B..2B-1 are the previous frame): every per-row op and every window (windows never span samples) is exactly
        what two separate passes compute; BatchNorm statistics stay per frame, visited previous-first like the
        reference (SiamWCA_MAE.py:265-284 / SiamWCA.py:630-640).  Halves the launches of the encoder."""
        hid, hid_prev = [], []
        x = torch.cat([feats, feats_prev], 0)
        if self._bf16():
            x = x.to(torch.bfloat16)
        for blk, st, st_cur in zip(self.sst_blocks, fp_all.stages, fp_cur.stages):
            x = blk(x, st, (0, st_cur.m, st.m), (1, 0))
            hid.append(x[:st_cur.m])
            hid_prev.append(x[st_cur.m:])
        return hid, hid_prev

    def _cross(self, hid, hid_prev, fp, tparts):
        return [blk(hid[i], hid_prev[i], fp.stages[i], tparts[i]) for i, blk in enumerate(self.wca_blocks)]

    def _sp(self, feats, st, B):
        return SparseConvTensor(feats, st.indices, [st.Y, st.X], B)

    def _dense(self, sps):
        de, out = getattr(self, self._deblocks_name), getattr(self, self._conv_out_name)
        if self.decoder_autocast is not None:
            # throughput mode: the BEV maps are written directly in bf16 channels-last, the cuDNN decoder runs under
            # autocast and `spatial_features` stays bf16 (the reference's AMP run returns fp16 here)
            maps = [sps[src].dense(self.decoder_autocast) for src in self.model_cfg["FEATURES_SOURCE"]]
            if self.decoder_autocast == torch.bfloat16 and self.fused_decoder_bn:
                # convolutions on cuDNN; BatchNorm2d + ReLU (+ the channel concat) on the library's bf16 row kernels
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    ups = [de[i][0](m) for i, m in enumerate(maps)]
                cat = _bn2d_relu_cat([de[i][1] for i in range(len(ups))], ups)
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    v = out[0](cat)
                return _bn2d_relu_cat([out[1]], [v])
            with torch.autocast("cuda", dtype=self.decoder_autocast):
                return out(torch.cat([de[i](m) for i, m in enumerate(maps)], 1))
        maps = [sps[src].dense() for src in self.model_cfg["FEATURES_SOURCE"]]
        return out(torch.cat([de[i](m) for i, m in enumerate(maps)], 1))

    def _strides(self, sps):
        return {k: 2 ** (i + 1) for i, k in enumerate(sps)}  # SiamWCA.py:581

    def _stride_out(self, strides):
        src = self.model_cfg["FEATURES_SOURCE"][0]
        return strides[src] // self.model_cfg["FUSE_LAYER"][src]["UPSAMPLE_STRIDE"]

    def _check_z(self, bd):
        # reference asserts voxel_coords[:, 1] == 0 (SiamWCA.py:620-621) with two host syncs per step; the grid
        # has a single z slab by construction, checked here once on the host
        if int(self.grid_size[2]) != 1:
            raise RuntimeError("the backbone needs a single z slab (pillars)")

    @staticmethod
    def _prepass(bd, fn):
        """Runs the coordinate-only part of the forward.  With batch_dict["side_stream"] (opt-in, see ops.side_stream)
        it runs on that stream -- its host reads of row counts then wait for a handful of tiny kernels, not for the
        main stream's backlog -- and the main stream is made to wait for it."""
        side = bd.get("side_stream")
        if side is None:
            return fn()
        main = torch.cuda.current_stream()
        with torch.cuda.stream(side):
            out = fn()
        main.wait_stream(side)
        ops.record_all(out, main)
        return out

    def _geometry(self, bd, coords, coords_prev):
        """Every coordinate-only table of the forward (plan.build_plans): no feature or weight is touched."""
        B = int(bd["batch_size"])
        idx_c, idx_p = self._indices(coords), self._indices(coords_prev)
        if self.siamese_batched:
            idx_p2 = idx_p.clone()
            idx_p2[:, 0] += B  # samples B..2B-1 of the Siamese-batched set are the previous frame
            idx_all = torch.cat([idx_c, idx_p2], 0)
            return build_plans([idx_c, idx_p, idx_all], B, self.sparse_shape, self.block_cfgs, temporal_pair=(0, 1),
                               want_ref=self.debug_refs, check=self.debug_refs, batches=[B, B, 2 * B],
                               need=[("subm", "part") if self.debug_refs else ("subm",), ("part",) if self.debug_refs else (), ("subm", "part")])
        return build_plans([idx_c, idx_p], B, self.sparse_shape, self.block_cfgs, temporal_pair=(0, 1), want_ref=self.debug_refs,
                           check=self.debug_refs)

    def _run(self, bd, feats, coords, feats_prev, coords_prev, geom=None):
        B = int(bd["batch_size"])
        plans, tparts = self._prepass(bd, lambda: self._geometry(bd, coords, coords_prev)) if geom is None else geom
        self.last_plan = (plans, tparts)
        if self._bf16():
            # bf16 copies of the tensor-core weight operands: all registered up front so that every copy whose fp32 master
            # changed (an optimizer step) is refreshed by ONE launch here
            ws = []
            for m in self.modules():
                if isinstance(m, EncoderLayer):
                    p = m._params()
                    ws += [p[0], p[2], p[7], p[9]]
                elif isinstance(m, ConvBNReLU):
                    ws.append(m._modules["0"].weight)
            # training: refresh unconditionally -- fused optimizers and `p.data` updates do not move Tensor._version (ops._weights_epoch)
            force = self.training and torch.is_grad_enabled()
            ops.shadows.register(ws)
            ops.shadows.refresh(force)
            ops.qkv_operands.refresh(force)
        if self.siamese_batched:
            hid, hid_prev = self._encode_siamese(feats, feats_prev, plans[2], plans[0])
        else:
            hid_prev = self._encode(feats_prev, plans[1])
            hid = self._encode(feats, plans[0])
        hid = self._cross(hid, hid_prev, plans[0], tparts)
        sps = {f"x_conv{i + 1}": self._sp(h, plans[0].stages[i], B) for i, h in enumerate(hid)}
        strides = self._strides(sps)
        sf = self._dense(sps)
        bd["multi_scale_3d_features"], bd["multi_scale_3d_strides"] = sps, strides
        bd["spatial_features"], bd["spatial_features_stride"] = sf, self._stride_out(strides)
        return sf

    def forward(self, batch_dict):
        self._check_z(batch_dict)
        self._run(batch_dict, batch_dict["voxel_features"], batch_dict["voxel_coords"], batch_dict["voxel_features_prev"],
                  batch_dict["voxel_coords_prev"])
        return batch_dict


class _GatherRowsFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, sel):
        ctx.save_for_backward(sel)
        ctx.rows = src.shape[0]
        out = torch.empty(sel.shape[0], src.shape[1], dtype=src.dtype, device=src.device)
        ops._call("gather_rows", ops._p(src.contiguous()), ops._p(sel), sel.shape[0], src.shape[1], ops._p(out), ops._stream())
        return out

    @staticmethod
    def backward(ctx, d):
        (sel,) = ctx.saved_tensors
        g = torch.zeros(ctx.rows, d.shape[1], dtype=d.dtype, device=d.device)
        ops._call("scatter_rows", ops._p(d.contiguous()), ops._p(sel), sel.shape[0], d.shape[1], ops._p(g), ops._stream())
        return g, None


class _LinearBiasFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        ctx.save_for_backward(x, w)
        return ops.linear_fwd(x.contiguous(), w, b)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dy = dy.contiguous()
        dw, db = torch.empty_like(w), torch.empty(w.shape[0], dtype=w.dtype, device=w.device)
        ops.linear_bwd_weight(dy, x.contiguous(), dw, db)
        return (ops.linear_bwd_data(dy, w) if ctx.needs_input_grad[0] else None), dw, db


class _ChamferFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, w, gtctx):
        loss, state = ops.chamfer_fwd(pred, None, w, gtctx)
        ctx.save_for_backward(pred, w, state)
        ctx.gtctx = gtctx
        return loss

    @staticmethod
    def backward(ctx, g):
        pred, w, state = ctx.saved_tensors
        return ops.chamfer_bwd(g.contiguous(), pred, None, w, state, ctx.gtctx), None, None


class SiamWCA_MAE(SiamWCA):
    """pcdet `SiamWCA_MAE` (SiamWCA_MAE.py): 75 % voxel masking, encode prev + visible cur, WCA, dense decode,
    16 predicted points per pillar, Chamfer loss against <= 64 grouped ground-truth points."""
    _deblocks_name, _conv_out_name = "decoder_deblocks", "decoder_conv_out"

    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kwargs):
        super().__init__(model_cfg, input_channels, grid_size, voxel_size, point_cloud_range)
        self.mask_cfg = model_cfg["MASK_CONFIG"]
        self.mask_ratio = float(self.mask_cfg["RATIO"])
        self.decoder_pred = nn.Linear(self.num_point_features, int(self.mask_cfg["NUM_PRD_POINTS"]) * 3, bias=True)
        self.forward_ret_dict = {}
        self.mask_generator = None  # optional torch.Generator on the device

    def _strides(self, sps):
        return {k: self.sparse_shape[0] // v.spatial_shape[0] for k, v in sps.items()}  # SiamWCA_MAE.py:214-216

    def mask_voxels(self, coords, voxels_per_sample):
        """SiamWCA_MAE.py:166-182 / common_utils.random_masking: per sample keep the int(L*(1-ratio)) voxels with
        the smallest uniform noise.  One batched draw + one sort, no per-sample .item(); returns mask (1 = removed)
        and the number of visible voxels (known on the host from the per-sample counts)."""
        M, dev = coords.shape[0], coords.device
        noise = torch.rand(M, device=dev, generator=self.mask_generator)
        sample = coords[:, 0].double()
        order = torch.argsort(sample + noise.double())  # samples are contiguous row blocks; noise orders inside
        starts = np.concatenate([[0], np.cumsum(voxels_per_sample)])
        keep_n = [int(L * (1 - self.mask_ratio)) for L in voxels_per_sample]
        pos = torch.arange(M, device=dev)
        lim = torch.repeat_interleave(torch.tensor([s + k for s, k in zip(starts[:-1], keep_n)], device=dev),
                                      torch.tensor(voxels_per_sample, device=dev), output_size=M)
        mask = torch.ones(M, device=dev)
        mask[order] = (pos >= lim).float()
        return mask, int(sum(keep_n))

    def forward(self, batch_dict):
        self._check_z(batch_dict)
        bd = batch_dict
        feats, coords = bd["voxel_features"], bd["voxel_coords"]

        def pre():  # coordinate-only: mask, visible set, geometry plans
            if "voxel_mae_mask_in" in bd:  # caller-supplied mask (tests, reproducible runs)
                mask = bd["voxel_mae_mask_in"].float()
                n_vis = int((mask == 0).sum())
            else:
                vps = bd.get("voxels_per_sample")
                if vps is None:
                    vps = torch.bincount(coords[:, 0], minlength=int(bd["batch_size"])).tolist()
                mask, n_vis = self.mask_voxels(coords, vps)
            vis = torch.nonzero_static(mask == 0, size=n_vis).view(-1).int()
            vis_coords = coords[vis.long()]
            return mask, vis, vis_coords, self._indices(coords), self._geometry(bd, vis_coords, bd["voxel_coords_prev"])

        mask, vis, vis_coords, all_idx, geom = self._prepass(bd, pre)
        bd["voxel_mae_mask"] = mask
        vis_feats = _GatherRowsFn.apply(feats, vis)
        sf = self._run(bd, vis_feats, vis_coords, bd["voxel_features_prev"], bd["voxel_coords_prev"], geom)
        vf = gather_bev(sf, all_idx)
        bd["voxel_features"], bd["voxel_coords"] = vf, coords
        bd["voxel_shuffle_inds"] = torch.arange(coords.shape[0], device=coords.device)
        pred = _LinearBiasFn.apply(vf, self.decoder_pred.weight, self.decoder_pred.bias).view(vf.shape[0], -1, 3)
        off, order, _ = bd["voxel_point_csr"]
        self._gtctx = (bd["points"], off, order, coords.contiguous(), self.point_cloud_range, self.voxel_size,
                       int(self.mask_cfg["NUM_GT_POINTS"]))
        self.forward_ret_dict = {"pred_points": pred, "mask": mask}
        return bd

    def gt_points(self):
        """(M, 64, 3) normalised ground-truth points (the reference's forward_ret_dict['gt_points'],
        SiamWCA_MAE.py:132-139); materialised only on request -- the loss kernel reads the CSR directly."""
        pk, off, order, vc, rng, vs, k = self._gtctx
        return ops.gt_group(pk, off, order, vc, rng, vs, vc.shape[0], k)

    def get_loss(self, tb_dict=None):
        tb_dict = {} if tb_dict is None else tb_dict
        r = self.forward_ret_dict
        loss = _ChamferFn.apply(r["pred_points"].float().contiguous(), r["mask"].contiguous(), self._gtctx)
        return loss, tb_dict
