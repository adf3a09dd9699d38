"""CPU pins of the N4 oracle (oracle/nms_ref.c + oracle/head_ref.py): against golden vectors made with the reference's own compiled CPU
rotated-IoU (tests/golden/make_golden_nms.py), against that library live where it exists, and against hand-computable cases."""
import os

import numpy as np
import pytest
import torch

from oracle import head_ref

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "nms.npz"))


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_oracle_iou_and_keep_match_reference_golden(seed):
    b, iou_ref, keep_ref = G[f"boxes{seed}"], G[f"iou{seed}"], G[f"keep{seed}"]
    iou = head_ref.iou_bev(b, b)
    assert np.isfinite(iou[np.isfinite(iou_ref)]).all()
    ok = np.isfinite(iou_ref)
    assert np.abs(iou[ok] - iou_ref[ok]).max() <= 1e-6, "oracle IoU differs from the reference's iou3d_cpu.cpp"
    keep = head_ref.nms_gpu(torch.from_numpy(b), torch.arange(b.shape[0], 0, -1).float(), 0.5)
    assert keep.tolist() == keep_ref.tolist()


def test_oracle_equals_reference_library_live():
    if not head_ref.ref_available():
        pytest.skip("oracle/_ref/libiou3d_ref.so is built only where /root/reference exists")
    for seed in (10, 11):
        b = head_ref.random_boxes(seed, 150)
        a, r = head_ref.iou_bev(b, b), head_ref.ref_iou(b, b)
        ok = np.isfinite(r)
        assert np.abs(a[ok] - r[ok]).max() <= 1e-6


def test_known_answers():
    sq = np.array([[0, 0, 0, 2, 2, 1, 0]], np.float32)
    same = head_ref.iou_bev(sq, sq)[0, 0]
    assert abs(same - 1.0) < 1e-5
    half = np.array([[1, 0, 0, 2, 2, 1, 0]], np.float32)           # overlap 2 of union 6
    assert abs(head_ref.iou_bev(sq, half)[0, 0] - 1 / 3) < 1e-3    # the reference's 1e-2 corner margin is part of the semantics
    far = np.array([[10, 10, 0, 2, 2, 1, 0.7]], np.float32)
    assert head_ref.iou_bev(sq, far)[0, 0] == 0.0
    rot = np.array([[0, 0, 0, 2, 2, 1, np.pi / 4]], np.float32)    # square vs itself rotated 45 deg: octagon, area 8 (sqrt2 - 1)
    inter = 8 * (2 ** .5 - 1)
    assert abs(head_ref.iou_bev(sq, rot)[0, 0] - inter / (8 - inter)) < 2e-3


def test_decode_and_nms_pipeline_shapes():
    g = torch.Generator().manual_seed(0)
    B, H, W = 2, 24, 24
    heads = [dict(hm=torch.randn(B, 2, H, W, generator=g) - 1, center=torch.rand(B, 2, H, W, generator=g), center_z=torch.randn(B, 1, H, W, generator=g) - 1,
                  dim=torch.randn(B, 3, H, W, generator=g) * 0.2 + 1, rot=torch.randn(B, 2, H, W, generator=g)),
             dict(hm=torch.randn(B, 1, H, W, generator=g) - 1, center=torch.rand(B, 2, H, W, generator=g), center_z=torch.randn(B, 1, H, W, generator=g) - 1,
                  dim=torch.randn(B, 3, H, W, generator=g) * 0.2, rot=torch.randn(B, 2, H, W, generator=g), iou=torch.randn(B, 1, H, W, generator=g))]
    cfg = dict(POST_PROCESSING=dict(SCORE_THRESH=0.1, POST_CENTER_LIMIT_RANGE=[-74.88, -74.88, -5.0, 74.88, 74.88, 3.0], MAX_OBJ_PER_SAMPLE=100,
                                    NMS_CONFIG=dict(NMS_TYPE="nms_gpu", NMS_THRESH=0.5, NMS_PRE_MAXSIZE=4096, NMS_POST_MAXSIZE=50)))
    ret = head_ref.generate_predicted_boxes(B, heads, [torch.tensor([0, 1]), torch.tensor([2])], cfg, [-74.88, -74.88, -5, 74.88, 74.88, 3], [0.32, 0.32, 8], 8)
    for r in ret:
        n = r["pred_boxes"].shape[0]
        assert r["pred_boxes"].shape == (n, 7) and r["pred_scores"].shape == (n,) and r["pred_labels"].shape == (n,)
        assert n > 0 and set(r["pred_labels"].tolist()) <= {1, 2, 3} and (r["pred_scores"] > 0.1).all()
