"""TEST INFRASTRUCTURE ONLY -- tier-2 oracle: an in-repo pure-PyTorch restatement of the T-MAE
sparse-window voxel-encoder hot path.

It runs on CPU (and on CUDA with stock torch ops) where /root/reference does not exist, e.g. the
GPU box.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import it; the product package never does.

PARITY PIN: the reference ships no tests or golden vectors, so this restatement
is pinned to the reference ITSELF: tests/test_oracle.py runs the reference's own modules (tier 1,
oracle/ref_loader.py) and this file on the same seeded inputs and weights in this container, and
tests/golden/*.pt (made by tests/golden/make_golden.py from tier 1) pin it where the reference
tree is absent.  The three third-party leaves (torch_scatter, spconv, pytorch3d) are restated
from their published semantics; for those leaves parity is UNPINNED by the reference (no tests,
libraries not installed) -- see DESIGN.md.

Canonical choices (the reference is nondeterministic there): slots = stable rank
by element index inside the group; strided sparse-conv output rows in lexicographic (b,y,x)
order.

Module tree and parameter names equal the reference's, so state_dicts interchange.
Every function cites the reference lines it follows (paths relative to /root/reference).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ============================================================================ A1-A3  VFE
def in_range_coords(points, pc_range, voxel_size, grid_size):
    """pcdet/utils/common_utils.py:66-76.  fp32 (p - lo) / vs, truncation toward zero, keep iff
    0 <= c < grid on all axes.  Returns keep (N,) bool, coords (N,3) i64 [cx, cy, cz]."""
    lo = points.new_tensor(pc_range[:3])
    vs = points.new_tensor(voxel_size)
    g = torch.as_tensor(np.asarray(grid_size), device=points.device).to(torch.int64)
    c = ((points[:, 1:4] - lo) / vs).to(torch.int64)
    keep = ((c >= 0) & (c < g)).all(-1)
    return keep, c


def voxelize(points, pc_range, voxel_size, grid_size):
    """temporal_dyn_vfe.py:67-72,85.  Returns kept points, per-point coords [b,z,y,x], inverse,
    voxel coords (lexicographic order, = torch.unique(dim=0) order) and the per-voxel mean of
    points[:, 1:]."""
    keep, c = in_range_coords(points, pc_range, voxel_size, grid_size)
    pts, c = points[keep], c[keep]
    coords = torch.cat([pts[:, 0:1].long(), c.flip(-1)], -1)
    X, Y, Z = (int(v) for v in grid_size)
    key = ((coords[:, 0] * Z + coords[:, 1]) * Y + coords[:, 2]) * X + coords[:, 3]
    ukey, inv = torch.unique(key, sorted=True, return_inverse=True)
    vx = ukey % X
    vy = (ukey // X) % Y
    vz = (ukey // (X * Y)) % Z
    vb = ukey // (X * Y * Z)
    vcoords = torch.stack([vb, vz, vy, vx], -1)
    M = ukey.shape[0]
    s = pts.new_zeros(M, pts.shape[1] - 1).index_add_(0, inv, pts[:, 1:])
    n = pts.new_zeros(M).index_add_(0, inv, pts.new_ones(pts.shape[0]))
    return pts, coords, inv, vcoords, s / n[:, None]


def vfe_point_features(pts, coords, inv, mean, pc_range, voxel_size):
    """temporal_dyn_vfe.py:89-110 with USE_ABSLOTE_XYZ, USE_CLUSTER_XYZ, no distance:
    [f_center(3), x,y,z,feats, f_cluster(3)]."""
    vs = pts.new_tensor(voxel_size)
    lo = pts.new_tensor(pc_range[:3])
    f_cluster = pts[:, 1:4] - mean[inv, :3]
    f_center = torch.stack([
        pts[:, 1] - ((coords[:, 3] + 0.5) * vs[0] + lo[0]),
        pts[:, 2] - ((coords[:, 2] + 0.5) * vs[1] + lo[1]),
        pts[:, 3] - ((coords[:, 1] + 0.5) * vs[2] + lo[2])], -1)
    return torch.cat([f_center, pts[:, 1:], f_cluster], -1)


def segment_max(x, inv, M):
    idx = inv[:, None].expand_as(x)
    return x.new_full((M, x.shape[1]), float("-inf")).scatter_reduce(0, idx, x, "amax", include_self=True)


def _fc_bn_relu(cfg, c_in):
    """model_utils/network_utils.py:25-40 (BatchNorm1d with torch defaults)."""
    layers = []
    for c in cfg:
        layers += [nn.Linear(c_in, c, bias=False), nn.BatchNorm1d(c), nn.ReLU(inplace=True)]
        c_in = c
    return nn.Sequential(*layers)


class TemporalDynVFE(nn.Module):
    """vfe/temporal_dyn_vfe.py (TYPE mean, one MLP group)."""

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size, **kw):
        super().__init__()
        self.model_cfg = model_cfg
        nf = num_point_features - 1  # :16
        mlps = model_cfg["MLPS"]
        assert len(mlps) == 1 and model_cfg.get("TYPE", "mean") == "mean"
        self.dvfe_mlps = nn.ModuleList([_fc_bn_relu(mlps[0], nf + 6)])
        self.finetuning = model_cfg.get("FT", False)
        self.num_point_features = mlps[0][-1]
        self.voxel_size, self.point_cloud_range, self.grid_size = voxel_size, point_cloud_range, grid_size

    def get_output_feature_dim(self):
        return self.num_point_features

    def _forward(self, points):
        pts, coords, inv, vcoords, mean = voxelize(points, self.point_cloud_range, self.voxel_size, self.grid_size)
        x = vfe_point_features(pts, coords, inv, mean, self.point_cloud_range, self.voxel_size)
        x = self.dvfe_mlps[0](x)
        x = segment_max(x, inv, vcoords.shape[0])
        return pts, coords, inv, vcoords, x

    def forward(self, batch_dict, **kw):
        for sfx in ("", "_prev"):
            pts, coords, inv, vcoords, x = self._forward(batch_dict["points" + sfx])
            if self.finetuning:  # :121-123,154-160
                batch_dict.pop("points" + sfx)
            else:
                batch_dict["points" + sfx] = pts
                batch_dict["point_coords" + sfx] = coords
                batch_dict["point_inverse_indices" + sfx] = inv
            batch_dict["voxel_coords" + sfx] = vcoords
            batch_dict["voxel_features" + sfx] = x
        return batch_dict


class DynVFE(TemporalDynVFE):
    """vfe/dyn_vfe.py:52-125: one frame, also writes pillar_features."""

    def __init__(self, model_cfg, num_point_features, voxel_size, point_cloud_range, grid_size, **kw):
        super().__init__(model_cfg, num_point_features + 1, voxel_size, point_cloud_range, grid_size)

    def forward(self, batch_dict, **kw):
        pts, coords, inv, vcoords, x = self._forward(batch_dict["points"])
        batch_dict.update(points=pts, point_coords=coords, point_inverse_indices=inv,
                          voxel_coords=vcoords, pillar_features=x, voxel_features=x)
        return batch_dict


# ============================================================================ A4-A6  partition
def stable_rank(group):
    """serial semantics of pcdet/ops/sst_ops/src/sst_ops_gpu.cu:14-20 (index order)."""
    n = group.shape[0]
    if n == 0:
        return group.clone()
    order = torch.sort(group, stable=True).indices
    g = group[order]
    pos = torch.arange(n, device=group.device)
    first = torch.ones(n, dtype=torch.bool, device=group.device)
    first[1:] = g[1:] != g[:-1]
    seg = torch.cummax(torch.where(first, pos, torch.zeros_like(pos)), 0).values
    out = torch.empty_like(group)
    out[order] = (pos - seg).to(group.dtype)
    return out


def window_coords(coords, grid_xyz, win, shifted):
    """model_utils/sst_utils.py:6-58.  coords (M,4) [b,z,y,x] -> batch_win_inds (M,),
    coors_in_win (M,3) [z,y,x]."""
    wx, wy, wz = win
    gx, gy, gz = grid_xyz
    nx, ny, nz = (int(np.ceil(g / w) + 1) for g, w in ((gx, wx), (gy, wy), (gz, wz)))
    sx, sy, sz = (wx // 2, wy // 2, wz // 2) if shifted else (wx, wy, wz)
    if gz == wz:
        sz = 0
    x, y, z = coords[:, 3] + sx, coords[:, 2] + sy, coords[:, 1] + sz
    bwi = coords[:, 0] * (nx * ny * nz) + (x // wx) * (ny * nz) + (y // wy) * nz + (z // wz)
    return bwi, torch.stack([z % wz, y % wy, x % wx], -1)


def level_of_count(cnt, drop_info):
    """spt_backbone.py:56-60: level whose drop_range holds the count; later levels overwrite."""
    lvl = torch.full_like(cnt, -1)
    tgt = torch.zeros_like(cnt)
    for dl, info in drop_info.items():
        lo, hi = info["drop_range"]
        m = (cnt >= lo) & (cnt < hi)
        lvl[m] = dl
        tgt[m] = info["max_tokens"]
    return lvl, tgt


def drop_single(bwi, drop_info):
    """spt_backbone.py:47-71."""
    slot = stable_rank(bwi)
    cnt = torch.bincount(bwi)[bwi]
    lvl, tgt = level_of_count(cnt, drop_info)
    assert (tgt > 0).all() and (lvl >= 0).all()
    return slot < tgt, lvl


def drop_temporal(bwi, bwi_p, drop_info):
    """SiamWCA.py:65-140: level from max(count_cur, count_prev); keep needs the window non-empty
    in both frames."""
    n = int(max(bwi.max(), bwi_p.max())) + 1 if bwi.numel() and bwi_p.numel() else 1
    c, cp = torch.bincount(bwi, minlength=n), torch.bincount(bwi_p, minlength=n)
    empty = (c == 0) | (cp == 0)
    cm = torch.maximum(c, cp)
    out = []
    for b in (bwi, bwi_p):
        lvl, tgt = level_of_count(cm[b], drop_info)
        assert (tgt > 0).all() and (lvl >= 0).all()
        out.append(((stable_rank(b) < tgt) & ~empty[b], lvl))
    return out[0][0], out[0][1], out[1][0], out[1][1]


def flat2win_tables(bwi, lvl, drop_info):
    """sst_utils.py:61-115: per level, compact window ids in ascending order, slot by stable rank,
    flat2win = r * T + slot.  Returns {dl: (flat2win (n_l,), (where,))} + reference's extra keys."""
    out = {}
    for dl, info in drop_info.items():
        m = lvl == dl
        if not m.any():
            continue
        w = bwi[m]
        conti = torch.unique(w, sorted=True, return_inverse=True)[1]
        out[dl] = (conti * info["max_tokens"] + stable_rank(conti), torch.where(m))
    out["voxel_drop_level"] = lvl
    out["batching_info"] = drop_info
    return out


def flat2window(feat, tables):
    """sst_utils.py:118-160."""
    out = {}
    for dl, info in tables["batching_info"].items():
        if dl not in tables:
            continue
        inds, (pos,) = tables[dl]
        T = info["max_tokens"]
        R = int(inds.max()) // T + 1
        buf = feat.new_zeros(R * T, feat.shape[-1])
        buf[inds] = feat[pos].to(buf.dtype)   # no-op in fp32; under torch.autocast (mixed-precision yardstick) dtypes differ
        out[dl] = buf.view(R, T, -1)
    return out


def window2flat(feat3d, tables):
    """sst_utils.py:163-192."""
    n = sum(tables[dl][0].shape[0] for dl in tables if not isinstance(dl, str))
    any_ = next(iter(feat3d.values()))
    out = any_.new_zeros(n, any_.shape[-1])
    for dl, f in feat3d.items():
        inds, (pos,) = tables[dl]
        out[pos] = f.reshape(-1, f.shape[-1])[inds].to(out.dtype)
    return out


def pos_embed_flat(ciw, win, C, temperature, normalize=False):
    """spt_backbone.py:186-222 before the flat2window step: (M,3)[z,y,x] -> (M,C) fp32."""
    wx, wy = win[0], win[1]
    y, x = ciw[:, 1] - wy / 2, ciw[:, 2] - wx / 2
    if normalize:
        x, y = x / wx * 2 * 3.1415, y / wy * 2 * 3.1415
    L = C // 2
    j = torch.arange(L, dtype=torch.float32, device=ciw.device)
    inv_freq = temperature ** (2 * torch.div(j, 2, rounding_mode="floor") / L)
    ex, ey = x[:, None] / inv_freq[None], y[:, None] / inv_freq[None]
    ex = torch.stack([ex[:, ::2].sin(), ex[:, 1::2].cos()], -1).flatten(1)
    ey = torch.stack([ey[:, ::2].sin(), ey[:, 1::2].cos()], -1).flatten(1)
    return torch.cat([ex, ey], -1)


def key_padding_masks(tables):
    """spt_backbone.py:233-243: True = padded slot."""
    n = tables["voxel_drop_level"].shape[0]
    ones = torch.ones(n, 1, dtype=torch.bool, device=tables["voxel_drop_level"].device)
    return {dl: ~v.squeeze(2) for dl, v in flat2window(ones, tables).items()}


def _drop_info(pre_cfg):
    return {int(k): v for k, v in pre_cfg["DROP_INFO"]["train"].items()}  # spt_backbone.py:32-33


def sst_input(feat, coords, grid_xyz, pre_cfg):
    """SSTInputLayer.forward (spt_backbone.py:137-184), SHUFFLE_VOXELS False."""
    win, di = pre_cfg["WINDOW_SHAPE"], _drop_info(pre_cfg)
    info = {"voxel_features": feat, "voxel_coords": coords}
    bwi, ciw = zip(*(window_coords(coords, grid_xyz, win, s == 1) for s in range(2)))
    k0, l0 = drop_single(bwi[0], di)
    k1, l1 = drop_single(bwi[1][k0], di)
    keep = torch.arange(coords.shape[0], device=coords.device)[k0][k1]
    lv = [l0[k0][k1], l1[k1]]
    info["voxel_keep_inds"] = keep
    info["voxel_features"], info["voxel_coords"] = feat[keep], coords[keep]
    for s in range(2):
        b = bwi[s][keep]
        info[f"batch_win_inds_shift{s}"] = b
        info[f"voxel_drop_level_shift{s}"] = lv[s]
        info[f"coors_in_win_shift{s}"] = ciw[s][keep]
        t = flat2win_tables(b, lv[s], di)
        info[f"flat2win_inds_shift{s}"] = t
        pe = pos_embed_flat(ciw[s][keep], win, feat.shape[1], pre_cfg["POS_TEMPERATURE"], pre_cfg["NORMALIZE_POS"])
        info[f"pos_dict_shift{s}"] = flat2window(pe, t)
        info[f"key_mask_shift{s}"] = key_padding_masks(t)
    return info


def sst_input_temporal(feat, coords, feat_p, coords_p, grid_xyz, pre_cfg):
    """SSTInputLayer_Temporal.forward (SiamWCA.py:201-269)."""
    win, di = pre_cfg["WINDOW_SHAPE"], _drop_info(pre_cfg)
    infos = [{"voxel_features": feat, "voxel_coords": coords}, {"voxel_features": feat_p, "voxel_coords": coords_p}]
    for s in range(2):
        (b, ciw), (bp, ciwp) = (window_coords(c, grid_xyz, win, s == 1) for c in (coords, coords_p))
        k, l, kp, lp = drop_temporal(b, bp, di)
        for info, bb, cc, kk, ll in ((infos[0], b, ciw, k, l), (infos[1], bp, ciwp, kp, lp)):
            f = info["voxel_features"]
            info[f"voxel_keep_inds_shift{s}"] = torch.where(kk)[0]
            info[f"voxel_drop_level_shift{s}"] = ll[kk]
            info[f"batch_win_inds_shift{s}"] = bb[kk]
            info[f"coors_in_win_shift{s}"] = cc[kk]
            t = flat2win_tables(bb[kk], ll[kk], di)
            info[f"flat2win_inds_shift{s}"] = t
            pe = pos_embed_flat(cc[kk], win, f.shape[1], pre_cfg["POS_TEMPERATURE"], pre_cfg["NORMALIZE_POS"])
            info[f"pos_dict_shift{s}"] = flat2window(pe, t)
            info[f"key_mask_shift{s}"] = key_padding_masks(t)
    return infos


# ============================================================================ A7-A8  attention
class CosineMHA(nn.Module):
    """model_utils/cosine_msa.py:441-528 + :114-176,178-438 for the branch the path takes:
    packed in_proj split into three linears (q is k, k is not v; :44-62), per-head L2 normalise
    (eps 1e-12), logits / clamp(tau, tau_min), key-padding -inf, softmax, out_proj.  Inputs are
    (T, R, C) like nn.MultiheadAttention with batch_first=False."""

    def __init__(self, C, H, tau_min=0.01):
        super().__init__()
        self.embed_dim, self.num_heads, self.tau_min = C, H, tau_min
        self.in_proj_weight = nn.Parameter(torch.empty(3 * C, C))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * C))
        self.out_proj = nn.Linear(C, C)
        self.tau = nn.Parameter(torch.ones(1, 1, 1))
        nn.init.xavier_uniform_(self.in_proj_weight)
        nn.init.zeros_(self.out_proj.bias)

    def forward(self, q, k, v, key_padding_mask):
        Tq, R, C = q.shape
        Tk, H = k.shape[0], self.num_heads
        hd = C // H
        wq, wk, wv = self.in_proj_weight.chunk(3)
        bq, bk, bv = self.in_proj_bias.chunk(3)
        Q = F.linear(q, wq, bq).reshape(Tq, R * H, hd).transpose(0, 1)
        K = F.linear(k, wk, bk).reshape(Tk, R * H, hd).transpose(0, 1)
        V = F.linear(v, wv, bv).reshape(Tk, R * H, hd).transpose(0, 1)
        Q, K = F.normalize(Q, dim=2), F.normalize(K, dim=2)
        a = torch.bmm(Q, K.transpose(1, 2)) / self.tau.clamp(min=self.tau_min)
        mask = key_padding_mask.view(R, 1, 1, Tk).expand(-1, H, -1, -1).reshape(R * H, 1, Tk)
        a = a + torch.zeros_like(mask, dtype=a.dtype).masked_fill_(mask, float("-inf"))
        o = torch.bmm(a.softmax(-1), V)
        return self.out_proj(o.transpose(0, 1).reshape(Tq, R, C))


class _AttnHolder(nn.Module):
    def __init__(self, name, C, H, tau_min):
        super().__init__()
        setattr(self, name, CosineMHA(C, H, tau_min))


class EncoderLayer(nn.Module):
    """sst_basic_block.py:58-84 (self) / wca_block.py:70-103 (cross): post-LN, exact GELU."""

    def __init__(self, C, H, FF, layer_cfg, cross):
        super().__init__()
        self.cross = cross
        self.win_attn = _AttnHolder("cross_attn" if cross else "self_attn", C, H, layer_cfg.get("tau_min", 0.01))
        self.linear1, self.linear2 = nn.Linear(C, FF), nn.Linear(FF, C)
        self.norm1, self.norm2 = nn.LayerNorm(C), nn.LayerNorm(C)

    def _ffn(self, src):
        src = self.norm1(src)
        src = src + self.linear2(F.gelu(self.linear1(src)))
        return self.norm2(src)

    def forward_self(self, src, pos, tables, masks):
        f3 = flat2window(src, tables)
        out = {}
        for dl, f in f3.items():
            f = f.permute(1, 0, 2)
            qk = f + pos[dl].permute(1, 0, 2)
            out[dl] = self.win_attn.self_attn(qk, qk, f, masks[dl]).permute(1, 0, 2)
        return self._ffn(src + window2flat(out, tables))

    def forward_cross(self, src, pos, tables, keep, masks_p, src_p, pos_p, tables_p, keep_p):
        f3, f3p = flat2window(src[keep], tables), flat2window(src_p[keep_p], tables_p)
        out = {}
        for dl, f in f3.items():
            f, fp = f.permute(1, 0, 2), f3p[dl].permute(1, 0, 2)
            q = f + pos[dl].permute(1, 0, 2)
            k = fp + pos_p[dl].permute(1, 0, 2)
            out[dl] = self.win_attn.cross_attn(q, k, fp, masks_p[dl]).permute(1, 0, 2)
        src = src.clone()
        if out:
            src[keep] = (src[keep] + window2flat(out, tables)).to(src.dtype)
        return self._ffn(src)


class ShiftBlock(nn.Module):
    """BasicShiftBlockV2 (sst_basic_block.py:87-114) / BasicShiftBlock_WCA (wca_block.py:106-145)."""

    def __init__(self, C, H, FF, layer_cfg, cross):
        super().__init__()
        self.encoder_list = nn.ModuleList([EncoderLayer(C, H, FF, layer_cfg, cross) for _ in range(2)])


# ============================================================================ A9  sparse conv
class SparseTensor:
    def __init__(self, features, indices, spatial_shape, batch_size):
        self.features, self.indices = features, indices
        self.spatial_shape, self.batch_size = [int(s) for s in spatial_shape], int(batch_size)

    def replace_feature(self, f):
        return SparseTensor(f, self.indices, self.spatial_shape, self.batch_size)

    def dense(self):
        Y, X = self.spatial_shape
        out = self.features.new_zeros(self.batch_size, Y, X, self.features.shape[1])
        i = self.indices.long()
        out[i[:, 0], i[:, 1], i[:, 2]] = self.features.to(out.dtype)
        return out.permute(0, 3, 1, 2).contiguous()


class SparseConv(nn.Module):
    """spconv 2.x SubMConv2d / SparseConv2d semantics, weight (Cout,kh,kw,Cin)."""

    def __init__(self, cin, cout, k, stride, padding, subm):
        super().__init__()
        self.k, self.stride, self.padding, self.subm = k, stride, padding, subm
        self.weight = nn.Parameter(torch.empty(cout, k, k, cin))
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))

    def forward(self, x):
        w = self.weight.permute(0, 3, 1, 2)
        d = x.dense()
        i = x.indices.long()
        if self.subm:
            o = F.conv2d(d, w, padding=self.k // 2)
            return SparseTensor(o[i[:, 0], :, i[:, 1], i[:, 2]], x.indices, x.spatial_shape, x.batch_size)
        o = F.conv2d(d, w, stride=self.stride, padding=self.padding)
        occ = d.new_zeros(x.batch_size, 1, *x.spatial_shape)
        occ[i[:, 0], 0, i[:, 1], i[:, 2]] = 1
        idx = (F.max_pool2d(occ, self.k, self.stride, self.padding)[:, 0] > 0).nonzero()
        return SparseTensor(o[idx[:, 0], :, idx[:, 1], idx[:, 2]], idx.int(), o.shape[-2:], x.batch_size)


class ConvBNReLU(nn.Module):
    """post_act_block (utils/spconv_utils.py:37-56): children named 0 (conv), 1 (BN1d), 2 (ReLU)."""

    def __init__(self, cin, cout, k, stride=1, padding=0, subm=True):
        super().__init__()
        self.add_module("0", SparseConv(cin, cout, k, stride, padding, subm))
        self.add_module("1", nn.BatchNorm1d(cout, eps=1e-3, momentum=0.01))
        self.add_module("2", nn.ReLU())

    def forward(self, x):
        x = self._modules["0"](x)
        return x.replace_feature(self._modules["2"](self._modules["1"](x.features)))


# ============================================================================ blocks
def _coords4(sp):
    i = sp.indices.long()
    return torch.cat([i[:, :1], torch.zeros_like(i[:, :1]), i[:, 1:]], -1)


class SSTBlock(nn.Module):
    """SSTBlockV1 (spt_backbone.py:267-353)."""

    def __init__(self, cfg, cin):
        super().__init__()
        enc = cfg["ENCODER"]
        C = enc["D_MODEL"]
        self.pre_cfg = cfg["PREPROCESS"]
        self.conv_down = ConvBNReLU(cin, C, 3, enc["STRIDE"], 1, subm=False) if enc["STRIDE"] > 1 else None
        self.encoder_blocks = nn.ModuleList(
            [ShiftBlock(C, enc["NHEAD"], enc["DIM_FEEDFORWARD"], enc["LAYER_CFG"], False) for _ in range(enc["NUM_BLOCKS"])])
        self.conv_out = ConvBNReLU(C, C, 3)

    def forward(self, sp, trace=None):
        if self.conv_down is not None:
            sp = self.conv_down(sp)
        feat, coords = sp.features, _coords4(sp)
        grid = [sp.spatial_shape[1], sp.spatial_shape[0], 1]
        info = sst_input(feat, coords, grid, self.pre_cfg)
        if trace is not None:
            trace.append(info)
        x = info["voxel_features"]
        for blk in self.encoder_blocks:
            for s, layer in enumerate(blk.encoder_list):
                x = layer.forward_self(x, info[f"pos_dict_shift{s}"], info[f"flat2win_inds_shift{s}"],
                                       info[f"key_mask_shift{s}"])
        un = torch.zeros_like(feat)
        un[info["voxel_keep_inds"]] = x.to(un.dtype)
        return self.conv_out(sp.replace_feature(feat + un))


class WCABlock(nn.Module):
    """WCABlock (SiamWCA.py:272-447): one BasicShiftBlock_WCA regardless of NUM_BLOCKS (:294-296)."""

    def __init__(self, cfg, cin):
        super().__init__()
        enc = cfg["ENCODER"]
        C = enc["D_MODEL"]
        self.pre_cfg = cfg["PREPROCESS"]
        n = 1 if enc["NUM_BLOCKS"] == 2 else enc["NUM_BLOCKS"]
        self.encoder_blocks = nn.ModuleList(
            [ShiftBlock(C, enc["NHEAD"], enc["DIM_FEEDFORWARD"], enc["LAYER_CFG"], True) for _ in range(n)])
        self.conv_out = ConvBNReLU(C, C, 3)

    def forward(self, sp, sp_prev, trace=None):
        feat, coords = sp.features, _coords4(sp)
        grid = [sp.spatial_shape[1], sp.spatial_shape[0], 1]
        a, b = sst_input_temporal(feat, coords, sp_prev.features, _coords4(sp_prev), grid, self.pre_cfg)
        if trace is not None:
            trace.append((a, b))
        x = feat + 0
        for s, layer in enumerate(self.encoder_blocks[0].encoder_list):
            x = layer.forward_cross(x, a[f"pos_dict_shift{s}"], a[f"flat2win_inds_shift{s}"], a[f"voxel_keep_inds_shift{s}"],
                                    b[f"key_mask_shift{s}"], sp_prev.features, b[f"pos_dict_shift{s}"],
                                    b[f"flat2win_inds_shift{s}"], b[f"voxel_keep_inds_shift{s}"])
        return self.conv_out(sp.replace_feature(feat + x))


def _deblocks(cfg):
    blocks, cin = nn.ModuleList(), 0
    for src in cfg["FEATURES_SOURCE"]:
        c = cfg["FUSE_LAYER"][src]
        blocks.append(nn.Sequential(
            nn.ConvTranspose2d(c["NUM_FILTER"], c["NUM_UPSAMPLE_FILTER"], c["UPSAMPLE_STRIDE"],
                               stride=c["UPSAMPLE_STRIDE"], bias=False),
            nn.BatchNorm2d(c["NUM_UPSAMPLE_FILTER"], eps=1e-3, momentum=0.01), nn.ReLU(inplace=True)))
        cin += c["NUM_UPSAMPLE_FILTER"]
    out = nn.Sequential(nn.Conv2d(cin, cin // len(blocks), 3, padding=1, bias=False),
                        nn.BatchNorm2d(cin // len(blocks), eps=1e-3, momentum=0.01), nn.ReLU(inplace=True))
    return blocks, out, cin // len(blocks)


class SiamWCA(nn.Module):
    """backbones_3d/SiamWCA.py:450-667 (finetune-mode encoder; ASYMMETRIC off as in both YAMLs)."""
    _deblocks_name, _conv_out_name = "deblocks", "conv_out"

    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kw):
        super().__init__()
        self.model_cfg, self.grid_size = model_cfg, grid_size
        self.voxel_size, self.point_cloud_range = voxel_size, point_cloud_range
        self.sparse_shape = [int(grid_size[1]), int(grid_size[0])]
        cin = input_channels
        self.sst_blocks = nn.ModuleList()
        for c in model_cfg["SST_BLOCK_LIST"]:
            self.sst_blocks.append(SSTBlock(c, cin))
            cin = c["ENCODER"]["D_MODEL"]
        self.wca_blocks = nn.ModuleList([WCABlock(c, c["ENCODER"]["D_MODEL"]) for c in model_cfg["SST_BLOCK_LIST"]])
        de, out, self.num_point_features = _deblocks(model_cfg)
        setattr(self, self._deblocks_name, de)
        setattr(self, self._conv_out_name, out)
        self.trace = None

    def sparse_encode(self, feat, coords, batch_size):
        x = SparseTensor(feat, coords[:, [0, 2, 3]].contiguous().int(), self.sparse_shape, batch_size)
        hidden = []
        for blk in self.sst_blocks:
            x = blk(x, self.trace)
            hidden.append(x)
        return {f"x_conv{i + 1}": h for i, h in enumerate(hidden)}

    def _strides(self, feats):
        return {k: 2 ** (i + 1) for i, k in enumerate(feats)}  # SiamWCA.py:581

    def cross(self, feats, feats_prev):
        return {f"x_conv{i + 1}": blk(feats[f"x_conv{i + 1}"], feats_prev[f"x_conv{i + 1}"], self.trace)
                for i, blk in enumerate(self.wca_blocks)}

    def dense_conv(self, feats):
        de, out = getattr(self, self._deblocks_name), getattr(self, self._conv_out_name)
        maps = [de[i](feats[src].dense()) for i, src in enumerate(self.model_cfg["FEATURES_SOURCE"])]
        return out(torch.cat(maps, 1))

    def _stride_out(self, strides):
        src = self.model_cfg["FEATURES_SOURCE"][0]
        return strides[src] // self.model_cfg["FUSE_LAYER"][src]["UPSAMPLE_STRIDE"]

    def forward(self, bd):
        assert (bd["voxel_coords"][:, 1] == 0).all() and (bd["voxel_coords_prev"][:, 1] == 0).all()
        B = bd["batch_size"]
        prev = self.sparse_encode(bd["voxel_features_prev"], bd["voxel_coords_prev"], B)
        cur = self.sparse_encode(bd["voxel_features"], bd["voxel_coords"], B)
        cur = self.cross(cur, prev)
        strides = self._strides(cur)
        bd["multi_scale_3d_features"], bd["multi_scale_3d_strides"] = cur, strides
        bd["spatial_features"] = self.dense_conv(cur)
        bd["spatial_features_stride"] = self._stride_out(strides)
        return bd


def random_masking(L, ratio, device, generator=None):
    """common_utils.py:49-63 for N=1: keep the int(L*(1-ratio)) smallest of rand(L); 1 = removed."""
    keep = int(L * (1 - ratio))
    noise = torch.rand(1, L, device=device, generator=generator)
    ids = torch.argsort(noise, dim=1)[:, :keep]
    return torch.ones(1, L, device=device).scatter_(1, ids, 0)[0]


def group_inner_inds(inv, M, K):
    """sst_ops_gpu.cu:22-39 serial: first K point indices per voxel (index order), cyclic pad."""
    rank = stable_rank(inv)
    cnt = torch.bincount(inv, minlength=M)
    g = torch.full((M, K), -1, dtype=torch.long, device=inv.device)
    sel = rank < K
    g[inv[sel], rank[sel]] = torch.arange(inv.shape[0], device=inv.device)[sel]
    j = torch.arange(K, device=inv.device)[None].expand(M, K)
    src = torch.where(j < cnt.clamp(max=K)[:, None], j, j % cnt.clamp(min=1)[:, None])
    return torch.where(cnt[:, None] > 0, torch.gather(g, 1, src), g)


def chamfer(pred, gt, w):
    """pytorch3d v0.7.1 chamfer_distance(pred, gt, weights=w), defaults."""
    if w.sum() == 0:
        return (pred.sum((1, 2)) * w).sum() * 0.0
    d = ((pred[:, :, None] - gt[:, None]) ** 2).sum(-1)
    cx = (d.min(2).values * w[:, None]).sum(1) / pred.shape[1]
    cy = (d.min(1).values * w[:, None]).sum(1) / gt.shape[1]
    return cx.sum() / w.sum() + cy.sum() / w.sum()


class SiamWCA_MAE(SiamWCA):
    """backbones_3d/SiamWCA_MAE.py (pretraining): 75% voxel mask, encode prev + visible cur, WCA,
    dense decode, per-voxel point prediction, Chamfer loss."""
    _deblocks_name, _conv_out_name = "decoder_deblocks", "decoder_conv_out"

    def __init__(self, model_cfg, input_channels, grid_size, voxel_size, point_cloud_range, **kw):
        super().__init__(model_cfg, input_channels, grid_size, voxel_size, point_cloud_range)
        self.mask_cfg = model_cfg["MASK_CONFIG"]
        self.decoder_pred = nn.Linear(self.num_point_features, self.mask_cfg["NUM_PRD_POINTS"] * 3)
        self.forward_ret_dict = {}

    def _strides(self, feats):
        return {k: self.sparse_shape[0] // v.spatial_shape[0] for k, v in feats.items()}  # SiamWCA_MAE.py:214-216

    def mask_voxels(self, coords, B, generator=None):
        """:166-182; a caller-supplied mask (bd['voxel_mae_mask_in']) replaces the RNG draw."""
        return torch.cat([random_masking(int((coords[:, 0] == b).sum()), self.mask_cfg["RATIO"], coords.device, generator)
                          for b in range(B)])

    def forward(self, bd):
        assert (bd["voxel_coords"][:, 1] == 0).all() and (bd["voxel_coords_prev"][:, 1] == 0).all()
        B = bd["batch_size"]
        prev = self.sparse_encode(bd["voxel_features_prev"], bd["voxel_coords_prev"], B)
        feat, coords = bd["voxel_features"], bd["voxel_coords"]
        mask = bd["voxel_mae_mask_in"] if "voxel_mae_mask_in" in bd else self.mask_voxels(coords, B)
        bd["voxel_mae_mask"] = mask
        cur = self.sparse_encode(feat[mask == 0], coords[mask == 0], B)
        cur = self.cross(cur, prev)
        strides = self._strides(cur)
        sf = self.dense_conv(cur)
        bd["multi_scale_3d_features"], bd["multi_scale_3d_strides"] = cur, strides
        bd["spatial_features"], bd["spatial_features_stride"] = sf, self._stride_out(strides)
        vf = sf.permute(0, 2, 3, 1)[coords[:, 0], coords[:, 2], coords[:, 3]]  # :311-312
        bd.update(voxel_features=vf, voxel_coords=coords,
                  voxel_shuffle_inds=torch.arange(coords.shape[0], device=coords.device))
        # target_assigner :124-152
        pts, inv = bd["points"], bd["point_inverse_indices"]
        M = int(inv.max()) + 1
        gt = pts[:, 1:4][group_inner_inds(inv, M, self.mask_cfg["NUM_GT_POINTS"])]
        vs = torch.tensor(self.voxel_size[:3], device=pts.device).float()
        lo = torch.tensor(self.point_cloud_range[:3], device=pts.device).float()
        centers = (coords[:, 1:].flip(-1).float() + 0.5) * vs + lo  # common_utils.py:130-145
        self.forward_ret_dict = {"pred_points": self.decoder_pred(vf).view(vf.shape[0], -1, 3),
                                 "gt_points": gt - centers[:, None], "mask": mask}
        return bd

    def get_loss(self, tb_dict=None):
        r = self.forward_ret_dict
        return chamfer(r["pred_points"].to(r["gt_points"].dtype), r["gt_points"], r["mask"]), tb_dict or {}


# ============================================================================ configs
def _levels(spec):
    d = {str(i): {"max_tokens": t, "drop_range": [lo, hi]} for i, (t, lo, hi) in enumerate(spec)}
    return {"train": d, "test": {k: dict(v) for k, v in d.items()}}


def model_cfg(kind):
    """Plain-dict equivalent of cfg.MODEL in tools/cfgs/once_models/t_mae_ssl.yaml:44-176
    ('pretrain') and t_mae.yaml:58-195 ('finetune')."""
    if kind == "pretrain":
        lv = [(16, 0, 16), (32, 16, 32), (64, 32, 100000)]
    else:
        lv = [(8, 0, 8), (16, 8, 16), (32, 16, 32), (48, 32, 48), (64, 48, 100000)]

    def block(name, stride, C, FF):
        return {"NAME": name,
                "PREPROCESS": {"WINDOW_SHAPE": [8, 8, 1], "DROP_INFO": _levels(lv), "SHUFFLE_VOXELS": False,
                               "POS_TEMPERATURE": 1000, "NORMALIZE_POS": False},
                "ENCODER": {"NUM_BLOCKS": 2, "STRIDE": stride, "D_MODEL": C, "NHEAD": 8, "DIM_FEEDFORWARD": FF,
                            "DROPOUT": 0.0, "ACTIVATION": "gelu", "LAYER_CFG": {"cosine": True, "tau_min": 0.01}}}

    bb = {"NAME": "SiamWCA_MAE" if kind == "pretrain" else "SiamWCA",
          "SST_BLOCK_LIST": [block("sst_block_x1", 1, 128, 256), block("sst_block_x2", 2, 256, 512),
                             block("sst_block_x4", 2, 256, 512)],
          "FEATURES_SOURCE": ["x_conv1", "x_conv2", "x_conv3"],
          "FUSE_LAYER": {"x_conv1": {"UPSAMPLE_STRIDE": 1, "NUM_FILTER": 128, "NUM_UPSAMPLE_FILTER": 128},
                         "x_conv2": {"UPSAMPLE_STRIDE": 2, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128},
                         "x_conv3": {"UPSAMPLE_STRIDE": 4, "NUM_FILTER": 256, "NUM_UPSAMPLE_FILTER": 128}}}
    if kind == "pretrain":
        bb["MASK_CONFIG"] = {"RATIO": 0.75, "NUM_PRD_POINTS": 16, "NUM_GT_POINTS": 64}
    vfe = {"NAME": "TemporalDynVFE", "TYPE": "mean", "WITH_DISTANCE": False, "USE_ABSLOTE_XYZ": True,
           "USE_CLUSTER_XYZ": True, "MLPS": [[64, 128]], "FT": kind == "finetune"}
    return {"VFE": vfe, "BACKBONE_3D": bb}


BEV_CFG = {"NAME": "SSTBEVBackbone", "NUM_FILTER": 128, "CONV_SHORTCUT": [0, 1, 2],   # t_mae.yaml:197-206
           "CONV_KWARGS": [{"out_channels": 128, "kernel_size": 3, "dilation": d, "padding": d, "stride": 1} for d in (1, 1, 2, 1)]}


class SSTBEVBackbone(nn.Module):
    """Restatement of pcdet/models/backbones_2d/sst_bev_backbone.py:6-44: Conv2d(no bias) ->
    BatchNorm2d(eps 1e-3, momentum 0.01) -> ReLU per entry of CONV_KWARGS; the result is added to its input where the
    shapes agree and the layer index is in CONV_SHORTCUT."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        cin, blocks = model_cfg["NUM_FILTER"], []
        for kw in model_cfg["CONV_KWARGS"]:
            blocks.append(nn.Sequential(nn.Conv2d(cin, bias=False, **dict(kw)), nn.BatchNorm2d(kw["out_channels"], eps=1e-3, momentum=0.01),
                                        nn.ReLU()))
            cin = kw["out_channels"]
        self.conv_layer = nn.ModuleList(blocks)
        self.shortcut = set(model_cfg["CONV_SHORTCUT"])
        self.num_bev_features = cin

    def forward(self, data_dict):
        x = data_dict["spatial_features"]
        for i, blk in enumerate(self.conv_layer):
            y = blk(x)
            x = y + x if (i in self.shortcut and y.shape == x.shape) else y
        data_dict["spatial_features_2d"] = x
        return data_dict


def build(kind, grid_size, voxel_size, pc_range, num_point_features=5, seed=0):
    cfg = model_cfg(kind)
    torch.manual_seed(seed)
    vfe = TemporalDynVFE(cfg["VFE"], num_point_features, voxel_size, pc_range, grid_size)
    cls = SiamWCA_MAE if kind == "pretrain" else SiamWCA
    return vfe, cls(cfg["BACKBONE_3D"], vfe.get_output_feature_dim(), grid_size, voxel_size, pc_range)
