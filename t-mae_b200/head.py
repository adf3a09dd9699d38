"""CenterHead post-processing: box decoding and rotated BEV NMS on the sm_100a kernels of csrc/nms.cu.

Mirrors pcdet/models/dense_heads/center_head.py:281-347 (`generate_predicted_boxes`, NMS_TYPE `nms_gpu` as in
tools/cfgs/once_models/t_mae.yaml:241-249), pcdet/models/model_utils/centernet_utils.py:154-220, model_nms_utils.py:6-25 and
pcdet/ops/iou3d_nms/iou3d_nms_utils.py (`boxes_iou_bev`, `nms_gpu`): same argument meaning, same outputs.  What differs is how it runs:
per head one top-K (torch), ONE decode kernel (gathers, exp, atan2, box assembly, range / score mask, stable compaction) and a batched
device-side NMS; the variable-length results of all heads and samples are cut with one host read at the end (the reference syncs per
head, per sample and inside nms_gpu).  The head's convolutions (shared_conv + SeparateHead) are dense cuDNN work and stay in pcdet.
"""
import ctypes

import torch

from . import ops
from ._lib import lib

F32, I32, I64 = torch.float32, torch.int32, torch.int64


def _f(v):
    return (ctypes.c_float * len(v))(*[float(x) for x in v])


def boxes_iou_bev(boxes_a, boxes_b):
    """iou3d_nms_utils.boxes_iou_bev: (N,7), (M,7) -> (N,M) rotated BEV IoU."""
    a, b = boxes_a[:, :7].contiguous().float(), boxes_b[:, :7].contiguous().float()
    out = torch.zeros(a.shape[0], b.shape[0], dtype=F32, device=a.device)
    ops._call("boxes_iou_bev", ops._p(a, F32), a.shape[0], ops._p(b, F32), b.shape[0], ops._p(out), ops._stream())
    return out


def nms_bev_batched(boxes, counts, thresh, post_max):
    """boxes (S, cap, 7), each sample's first counts[s] rows in descending score order -> keep (S, cap) i64, num_keep (S) i32 (device)."""
    L = lib()
    S, cap, _ = boxes.shape
    keep = torch.empty(S, cap, dtype=I64, device=boxes.device)
    num = torch.empty(S, dtype=I32, device=boxes.device)
    wsb = L.nms_bev_workspace_bytes(S, cap)
    ws = ops._ws(wsb, boxes.device)
    ops._call("nms_bev", ops._p(boxes, F32), ops._p(counts, I32), S, cap, float(thresh), int(post_max), ops._p(keep), ops._p(num), ops._p(ws), wsb,
              ops._stream())
    return keep, num


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    """iou3d_nms_utils.nms_gpu (:84-99): -> (kept indices into `boxes`, None)."""
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    b = boxes[order][:, :7].contiguous().float()
    n = b.shape[0]
    if n == 0:
        return order, None
    keep, num = nms_bev_batched(b.view(1, n, 7), torch.full((1,), n, dtype=I32, device=b.device), thresh, n)
    return order[keep[0, :int(num[0])]].contiguous(), None


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """model_nms_utils.class_agnostic_nms (:6-25)."""
    get = nms_config.get if hasattr(nms_config, "get") else (lambda k: getattr(nms_config, k))
    src = box_scores
    if score_thresh is not None:
        m = box_scores >= score_thresh
        box_scores, box_preds = box_scores[m], box_preds[m]
    selected = torch.zeros(0, dtype=I64, device=src.device)
    if box_scores.shape[0] > 0:
        s, idx = torch.topk(box_scores, k=min(get("NMS_PRE_MAXSIZE"), box_scores.shape[0]))
        keep, _ = nms_gpu(box_preds[idx][:, :7], s, get("NMS_THRESH"))
        selected = idx[keep[:get("NMS_POST_MAXSIZE")]]
    if score_thresh is not None:
        selected = m.nonzero().view(-1)[selected]
    return selected, src[selected]


def decode_head(pred_dict, class_map, point_cloud_range, voxel_size, feature_map_stride, K, score_thresh, limit_range):
    """One head: top-K of sigmoid(hm) over (class, y, x), then tmae_centerhead_decode.  -> boxes (B,K,7), scores, labels, ious (B,K), counts (B)
    i32: capacity-K device buffers whose first counts[b] rows are valid, in descending score order."""
    hm = pred_dict["hm"].float()
    B, ncls, H, W = hm.shape
    scores, inds = torch.topk(hm.sigmoid().flatten(1), K)          # == the reference's per-class top-K followed by a top-K over classes
    dev = hm.device
    boxes = torch.empty(B, K, 7, dtype=F32, device=dev)
    out_scores = torch.empty(B, K, dtype=F32, device=dev)
    labels = torch.empty(B, K, dtype=I64, device=dev)
    ious = torch.empty(B, K, dtype=F32, device=dev)
    counts = torch.empty(B, dtype=I32, device=dev)
    c = lambda k: pred_dict[k].float().contiguous()
    iou = c("iou") if "iou" in pred_dict else None
    cm = None if class_map is None else class_map.to(dev, I64).contiguous()
    ops._call("centerhead_decode", ops._p(scores.contiguous(), F32), ops._p(inds.contiguous(), I64), ops._p(c("center")), ops._p(c("center_z")),
              ops._p(c("dim")), ops._p(c("rot")), ops._p(iou), ops._p(cm), B, K, H, W, ncls, float(feature_map_stride), _f(voxel_size[:2]),
              _f(point_cloud_range[:2]), _f(limit_range), float(score_thresh if score_thresh is not None else -1e30), ops._p(boxes), ops._p(out_scores),
              ops._p(labels), ops._p(ious), ops._p(counts), ops._stream())
    return boxes, out_scores, labels, ious, counts


def generate_predicted_boxes(batch_size, pred_dicts, class_id_mapping_each_head, post_process_cfg, point_cloud_range, voxel_size, feature_map_stride):
    """CenterHead.generate_predicted_boxes (center_head.py:281-347) for NMS_TYPE nms_gpu: pred_dicts = the SeparateHead outputs per head
    ('hm' logits, 'center', 'center_z', 'dim' log sizes, 'rot' [cos, sin], optional 'iou').  -> [{pred_boxes, pred_scores, pred_labels}] per sample."""
    get = post_process_cfg.get if hasattr(post_process_cfg, "get") else (lambda k: getattr(post_process_cfg, k))
    nms = get("NMS_CONFIG")
    nget = nms.get if hasattr(nms, "get") else (lambda k: getattr(nms, k))
    if nget("NMS_TYPE") != "nms_gpu":
        raise NotImplementedError("only NMS_TYPE nms_gpu is on the T-MAE path (t_mae.yaml:246)")
    K = int(get("MAX_OBJ_PER_SAMPLE"))
    pre, post = int(nget("NMS_PRE_MAXSIZE")), int(nget("NMS_POST_MAXSIZE"))
    per_head = []
    for idx, pd in enumerate(pred_dicts):
        boxes, scores, labels, _, counts = decode_head(pd, class_id_mapping_each_head[idx], point_cloud_range, voxel_size, feature_map_stride, K,
                                                       get("SCORE_THRESH"), get("POST_CENTER_LIMIT_RANGE"))
        cnt = counts.clamp(max=pre) if pre < K else counts           # class_agnostic_nms: top NMS_PRE_MAXSIZE by score (rows are already sorted)
        keep, num = nms_bev_batched(boxes, cnt, float(nget("NMS_THRESH")), post)
        per_head.append((boxes, scores, labels, keep, num))
    nums = torch.stack([p[4] for p in per_head]).cpu()               # the one host read of the tail
    ret = []
    for k in range(batch_size):
        bs, ss, ls = [], [], []
        for h, (boxes, scores, labels, keep, _) in enumerate(per_head):
            sel = keep[k, :int(nums[h, k])]
            bs.append(boxes[k][sel]), ss.append(scores[k][sel]), ls.append(labels[k][sel])
        ret.append(dict(pred_boxes=torch.cat(bs, 0), pred_scores=torch.cat(ss, 0), pred_labels=torch.cat(ls, 0) + 1))
    return ret
