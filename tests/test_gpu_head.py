"""GPU parity of row N4 (CenterHead decode + rotated BEV NMS, csrc/nms.cu) against the oracle (oracle/head_ref.py + oracle/nms_ref.c, pinned
to the reference's own compiled iou3d_cpu.cpp by tests/test_oracle_nms.py) and the committed golden vectors made with that library."""
import os

import numpy as np
import pytest
import torch

from oracle import head_ref

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    import tmae_b200  # noqa: F401
    from tmae_b200 import head

DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nms.npz"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_iou_and_nms_match_reference_golden(seed):
    b, iou_ref, keep_ref = G[f"boxes{seed}"], G[f"iou{seed}"], G[f"keep{seed}"]
    bd = torch.from_numpy(b).to(DEV)
    iou = head.boxes_iou_bev(bd, bd).cpu().numpy()
    ok = np.isfinite(iou_ref)
    assert np.abs(iou[ok] - iou_ref[ok]).max() <= 2e-5     # sinf / cosf / atan2f differ from glibc's in the last bits; everything else is bit-equal
    keep, _ = head.nms_gpu(bd, torch.arange(b.shape[0], 0, -1, device=DEV).float(), 0.5)
    assert keep.cpu().tolist() == keep_ref.tolist()


@pytest.mark.parametrize("n", [1, 63, 64, 65, 500, 4096])
def test_nms_matches_oracle_sizes(n):
    b = head_ref.random_boxes(100 + n, n, spread=8.0 + n ** .5)
    scores = torch.rand(n, generator=torch.Generator().manual_seed(n))
    ref = head_ref.nms_gpu(torch.from_numpy(b), scores, 0.5)
    got, _ = head.nms_gpu(torch.from_numpy(b).to(DEV), scores.to(DEV), 0.5)
    iou = head_ref.iou_bev(b, b)
    near = np.abs(iou - 0.5) < 1e-4
    if not near.any():      # no pair sits on the threshold: the keep lists must be identical
        assert got.cpu().tolist() == ref.tolist()
    else:
        assert len(set(got.cpu().tolist()) ^ set(ref.tolist())) <= 2 * int(near.sum())
    sel_ref, _ = head_ref.class_agnostic_nms(scores, torch.from_numpy(b), dict(NMS_PRE_MAXSIZE=300, NMS_THRESH=0.5, NMS_POST_MAXSIZE=40))
    sel, sc = head.class_agnostic_nms(scores.to(DEV), torch.from_numpy(b).to(DEV), dict(NMS_PRE_MAXSIZE=300, NMS_THRESH=0.5, NMS_POST_MAXSIZE=40))
    if not near.any():
        assert sel.cpu().tolist() == sel_ref.tolist() and torch.equal(sc.cpu(), scores[sel_ref])


def _heads(seed, B, H, W):
    g = torch.Generator().manual_seed(seed)
    mk = lambda c, s=1.0, o=0.0: torch.randn(B, c, H, W, generator=g) * s + o
    return [dict(hm=mk(2, 1.5, -2), center=torch.rand(B, 2, H, W, generator=g), center_z=mk(1, 1, -1), dim=mk(3, 0.2, 1), rot=mk(2)),
            dict(hm=mk(1, 1.5, -2), center=torch.rand(B, 2, H, W, generator=g), center_z=mk(1, 1, -1), dim=mk(3, 0.2, 0.5), rot=mk(2), iou=mk(1)),
            dict(hm=mk(2, 1.5, -2), center=torch.rand(B, 2, H, W, generator=g), center_z=mk(1, 3, -1), dim=mk(3, 0.2, 0), rot=mk(2))]


@pytest.mark.parametrize("B,H,W,K", [(2, 24, 24, 100), (3, 59, 59, 500)])
def test_generate_predicted_boxes_matches_oracle(B, H, W, K):
    """The whole tail (center_head.py:281-347) on random head maps, 3 heads (one with an IoU branch), ONCE post-processing config."""
    heads = _heads(B * H, B, H, W)
    cfg = dict(SCORE_THRESH=0.1, POST_CENTER_LIMIT_RANGE=[-74.88, -74.88, -5.0, 74.88, 74.88, 3.0], MAX_OBJ_PER_SAMPLE=K,
               NMS_CONFIG=dict(NMS_TYPE="nms_gpu", NMS_THRESH=0.5, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=500))
    maps = [torch.tensor([0, 1]), torch.tensor([2]), torch.tensor([3, 4])]
    rng, vs, stride = [-74.88, -74.88, -5.0, 74.88, 74.88, 3.0], [0.32, 0.32, 8.0], 8
    ref = head_ref.generate_predicted_boxes(B, heads, maps, dict(POST_PROCESSING=cfg), rng, vs, stride)
    got = head.generate_predicted_boxes(B, [{k: v.to(DEV) for k, v in h.items()} for h in heads], maps, cfg, rng, vs, stride)
    # the decode alone (before NMS) for the first head: same candidates in the same order
    boxes, scores, labels, ious, counts = head.decode_head({k: v.to(DEV) for k, v in heads[1].items()}, maps[1], rng, vs, stride, K, 0.1, cfg["POST_CENTER_LIMIT_RANGE"])
    hm = heads[1]["hm"].sigmoid()
    iou_map = torch.clamp((heads[1]["iou"] + 1) * 0.5, 0, 1)
    dec = head_ref.decode_bbox_from_heatmap(hm, heads[1]["rot"][:, 0:1], heads[1]["rot"][:, 1:2], heads[1]["center"], heads[1]["center_z"],
                                            heads[1]["dim"].exp(), iou_map, rng, vs, stride, K, 0.1, cfg["POST_CENTER_LIMIT_RANGE"])
    for b in range(B):
        n = int(counts[b])
        assert n == dec[b]["pred_boxes"].shape[0]
        assert torch.allclose(boxes[b, :n].cpu(), dec[b]["pred_boxes"], rtol=1e-5, atol=1e-5)
        assert torch.allclose(scores[b, :n].cpu(), dec[b]["pred_scores"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(ious[b, :n].cpu(), dec[b]["pred_ious"], rtol=1e-6, atol=1e-7)
    for b in range(B):
        r, g_ = ref[b], got[b]
        assert g_["pred_boxes"].shape == r["pred_boxes"].shape, (g_["pred_boxes"].shape, r["pred_boxes"].shape)
        assert torch.equal(g_["pred_labels"].cpu(), r["pred_labels"])
        assert torch.allclose(g_["pred_scores"].cpu(), r["pred_scores"], rtol=1e-6, atol=1e-7)
        assert torch.allclose(g_["pred_boxes"].cpu(), r["pred_boxes"], rtol=1e-5, atol=1e-5)
