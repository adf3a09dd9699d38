"""TEST INFRASTRUCTURE ONLY -- oracle of row N4: CenterHead box decoding and rotated BEV NMS.

Restates (paths relative to /root/reference):
  pcdet/models/model_utils/centernet_utils.py:118-151  _gather_feat / _transpose_and_gather_feat / _topk
  pcdet/models/model_utils/centernet_utils.py:154-220  decode_bbox_from_heatmap
  pcdet/models/dense_heads/center_head.py:281-347       generate_predicted_boxes (NMS_TYPE: nms_gpu, t_mae.yaml:241-249)
  pcdet/models/model_utils/model_nms_utils.py:6-25      class_agnostic_nms
  pcdet/ops/iou3d_nms/iou3d_nms_utils.py:84-99          nms_gpu (sort, pre_maxsize, native call)
The rotated IoU and the greedy sweep are the plain-C restatement oracle/nms_ref.c; `ref_iou()` is the REFERENCE's own CPU
implementation (iou3d_cpu.cpp) compiled by oracle/build_ref.py into oracle/_ref/ (present wherever /root/reference was).
"""
import ctypes
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
_libs = {}


def _oracle():
    if "o" not in _libs:
        path = os.path.join(HERE, "_build", "libnms_oracle.so")
        if not os.path.exists(path):
            import importlib.util
            spec = importlib.util.spec_from_file_location("tmae_build_ref", os.path.join(HERE, "build_ref.py"))
            m = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(m)
            m.build()
        L = ctypes.CDLL(path)
        L.tmae_oracle_boxes_iou_bev.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        L.tmae_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p]
        L.tmae_oracle_nms.restype = ctypes.c_int64
        _libs["o"] = L
    return _libs["o"]


def ref_available():
    return os.path.exists(os.path.join(HERE, "_ref", "libiou3d_ref.so"))


def ref_iou(a, b):
    """(N,7), (M,7) float32 -> (N,M): the reference's own boxes_iou_bev_cpu."""
    if "r" not in _libs:
        L = ctypes.CDLL(os.path.join(HERE, "_ref", "libiou3d_ref.so"))
        L.ref_boxes_iou_bev.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
        _libs["r"] = L
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    _libs["r"].ref_boxes_iou_bev(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


def iou_bev(a, b):
    a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
    out = np.zeros((a.shape[0], b.shape[0]), np.float32)
    _oracle().tmae_oracle_boxes_iou_bev(a.ctypes.data, a.shape[0], b.ctypes.data, b.shape[0], out.ctypes.data)
    return out


def nms_gpu(boxes, scores, thresh, pre_maxsize=None):
    """iou3d_nms_utils.py:84-99 -> kept indices into `boxes` (torch int64)."""
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = np.ascontiguousarray(boxes[order].float().numpy()[:, :7])
    keep = np.zeros(b.shape[0], np.int64)
    n = _oracle().tmae_oracle_nms(b.ctypes.data, b.shape[0], ctypes.c_float(thresh), keep.ctypes.data)
    return order[torch.from_numpy(keep[:n])]


def class_agnostic_nms(box_scores, box_preds, nms_cfg):
    """model_nms_utils.py:6-25 with score_thresh=None (center_head.py:321-325)."""
    selected = torch.zeros(0, dtype=torch.long)
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(nms_cfg["NMS_PRE_MAXSIZE"], box_scores.shape[0]))
        keep = nms_gpu(box_preds[indices][:, :7], box_scores_nms, nms_cfg["NMS_THRESH"])
        selected = indices[keep[:nms_cfg["NMS_POST_MAXSIZE"]]]
    return selected, box_scores[selected]


def _gather(feat, ind):
    feat = feat.permute(0, 2, 3, 1).contiguous()
    feat = feat.view(feat.size(0), -1, feat.size(3))
    return feat.gather(1, ind.unsqueeze(2).expand(ind.size(0), ind.size(1), feat.size(2)))


def decode_bbox_from_heatmap(heatmap, rot_cos, rot_sin, center, center_z, dim, iou, point_cloud_range, voxel_size, feature_map_stride, K,
                             score_thresh, post_center_limit_range):
    """centernet_utils.py:154-220 (vel=None, circle_nms=False)."""
    B, ncls, H, W = heatmap.shape
    topk_scores, topk_inds = torch.topk(heatmap.flatten(2, 3), K)
    topk_inds = topk_inds % (H * W)
    topk_ys = torch.div(topk_inds, W, rounding_mode="floor").float()
    topk_xs = (topk_inds % W).int().float()
    scores, topk_ind = torch.topk(topk_scores.view(B, -1), K)
    class_ids = torch.div(topk_ind, K, rounding_mode="floor").int()
    inds = topk_inds.view(B, -1).gather(1, topk_ind)
    ys = topk_ys.view(B, -1).gather(1, topk_ind)
    xs = topk_xs.view(B, -1).gather(1, topk_ind)
    ious = _gather(iou, inds).view(B, K)
    center = _gather(center, inds).view(B, K, 2)
    rot_sin = _gather(rot_sin, inds).view(B, K, 1)
    rot_cos = _gather(rot_cos, inds).view(B, K, 1)
    center_z = _gather(center_z, inds).view(B, K, 1)
    dim = _gather(dim, inds).view(B, K, 3)
    angle = torch.atan2(rot_sin, rot_cos)
    xs = xs.view(B, K, 1) + center[:, :, 0:1]
    ys = ys.view(B, K, 1) + center[:, :, 1:2]
    xs = xs * feature_map_stride * voxel_size[0] + point_cloud_range[0]
    ys = ys * feature_map_stride * voxel_size[1] + point_cloud_range[1]
    boxes = torch.cat([xs, ys, center_z, dim, angle], dim=-1)
    lim = torch.as_tensor(post_center_limit_range, dtype=torch.float32)
    mask = (boxes[..., :3] >= lim[:3]).all(2) & (boxes[..., :3] <= lim[3:]).all(2)
    if score_thresh is not None:
        mask &= scores > score_thresh
    return [dict(pred_boxes=boxes[k, mask[k]], pred_scores=scores[k, mask[k]], pred_ious=ious[k, mask[k]], pred_labels=class_ids[k, mask[k]])
            for k in range(B)]


def generate_predicted_boxes(batch_size, pred_dicts, class_id_mapping_each_head, cfg, point_cloud_range, voxel_size, feature_map_stride):
    """center_head.py:281-347 for NMS_TYPE nms_gpu.  pred_dicts: per head {'hm','center','center_z','dim','rot'[,'iou']} raw head outputs."""
    pp = cfg["POST_PROCESSING"]
    ret = [dict(pred_boxes=[], pred_scores=[], pred_labels=[]) for _ in range(batch_size)]
    for idx, pd in enumerate(pred_dicts):
        hm = pd["hm"].sigmoid()
        iou = torch.clamp((pd["iou"] + 1) * 0.5, min=0, max=1) if "iou" in pd else torch.ones_like(hm[:, 0:1])
        finals = decode_bbox_from_heatmap(hm, pd["rot"][:, 0:1], pd["rot"][:, 1:2], pd["center"], pd["center_z"], pd["dim"].exp(), iou,
                                          point_cloud_range, voxel_size, feature_map_stride, pp["MAX_OBJ_PER_SAMPLE"], pp["SCORE_THRESH"],
                                          pp["POST_CENTER_LIMIT_RANGE"])
        for k, fd in enumerate(finals):
            fd["pred_labels"] = class_id_mapping_each_head[idx][fd["pred_labels"].long()]
            sel, sel_scores = class_agnostic_nms(fd["pred_scores"], fd["pred_boxes"], pp["NMS_CONFIG"])
            ret[k]["pred_boxes"].append(fd["pred_boxes"][sel])
            ret[k]["pred_scores"].append(sel_scores)
            ret[k]["pred_labels"].append(fd["pred_labels"][sel])
    for k in range(batch_size):
        ret[k]["pred_boxes"] = torch.cat(ret[k]["pred_boxes"], 0)
        ret[k]["pred_scores"] = torch.cat(ret[k]["pred_scores"], 0)
        ret[k]["pred_labels"] = torch.cat(ret[k]["pred_labels"], 0) + 1
    return ret


def random_boxes(seed, n, spread=20.0, clustered=True):
    """Seeded vehicle-like BEV boxes (N, 7); clustered so that many pairs overlap; includes exact duplicates, axis-aligned pairs sharing an
    edge, nested boxes and a degenerate zero-area box."""
    rng = np.random.default_rng(seed)
    ctr = rng.uniform(-spread, spread, (max(1, n // 6), 2))
    c = ctr[rng.integers(0, ctr.shape[0], n)] + rng.normal(0, 1.2, (n, 2)) if clustered else rng.uniform(-spread, spread, (n, 2))
    b = np.zeros((n, 7), np.float32)
    b[:, :2] = c
    b[:, 2] = rng.uniform(-2, 0, n)
    b[:, 3] = rng.uniform(3.5, 5.0, n)
    b[:, 4] = rng.uniform(1.5, 2.2, n)
    b[:, 5] = rng.uniform(1.4, 2.0, n)
    b[:, 6] = rng.uniform(-np.pi, np.pi, n)
    if n >= 8:
        b[1] = b[0]                                    # exact duplicate
        b[2] = b[0]; b[2, 6] += np.float32(np.pi)      # same box, heading flipped
        b[3, :] = [0, 0, 0, 4, 2, 1.5, 0]; b[4, :] = [4, 0, 0, 4, 2, 1.5, 0]    # axis-aligned, sharing an edge
        b[5, :] = [0, 0, 0, 2, 1, 1.5, 0.3]            # nested in b[3] (rotated)
        b[6, 3:5] = 0                                  # zero area
    return b
