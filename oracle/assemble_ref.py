"""TEST INFRASTRUCTURE ONLY -- CPU oracle of the batch-assembly step in front of the VFE.

Tier 2 (`assemble`): a numpy restatement of what the reference does per sample on its DataLoader workers, in the
reference's own order and dtypes:
    remove_ego_points            pcdet/datasets/once_temporal/once_eval/once_utils.py:43-45 (called with radius 2:
                                 once_temporal_dataset.py:167-168)
    convert_prv_frame_to_cur     once_utils.py:4-29 (float64; scipy Rotation.from_quat(...).as_matrix(), np.linalg.inv)
    mask_points_by_range         pcdet/utils/common_utils.py:124-127 (x, y only, closed interval; the range is a float32
                                 array: dataset.py:25) via data_processor.py:81-83
    collate_batch                pcdet/datasets/dataset.py:203-208 (np.pad with the sample index in column 0)
    load_data_to_gpu             pcdet/models/__init__.py:16-23 (`.float()`)
shuffle_points (data_processor.py:92-102) only permutes rows with the host RNG and is left out on both sides.

Tier 1 (`assemble_tier1`): the same pipeline calling the reference's OWN functions, loaded from /root/reference
(`once_utils.py` by file path -- it imports numpy and scipy only; `common_utils` through oracle/ref_loader.py).  Used by
tests/golden/make_golden_assemble.py to produce tests/golden/assemble.npz, which pins tier 2.
"""
import importlib.util
import os

import numpy as np
from scipy.spatial.transform import Rotation


def remove_ego_points(points, center_radius=1.0):
    return points[~((np.abs(points[:, 0]) < center_radius) & (np.abs(points[:, 1]) < center_radius))]


def convert_prv_frame_to_cur(pc_prv, pose_prv, pose_cur):
    xyz = pc_prv[:, :3]
    if np.any(pose_prv):                       # an all-zero pose (static vehicle) means: no transformation
        rot = Rotation.from_quat(pose_prv[:4]).as_matrix()
        xyz = np.dot(xyz, rot.T) + np.array(pose_prv[4:]).transpose()
    if np.any(pose_cur):
        m = np.zeros((4, 4))
        m[:3, :3] = Rotation.from_quat(pose_cur[:4]).as_matrix()
        m[:3, 3] = np.array(pose_cur[4:]).transpose()
        m[3, 3] = 1
        m = np.linalg.inv(m)
        xyz = np.dot(np.concatenate([xyz, np.ones((xyz.shape[0], 1))], axis=-1), m.T)[:, :3]
    return np.concatenate([xyz[:, :3], pc_prv[:, 3:]], axis=-1)


def mask_points_by_range(points, limit_range):
    return (points[:, 0] >= limit_range[0]) & (points[:, 0] <= limit_range[3]) \
        & (points[:, 1] >= limit_range[1]) & (points[:, 1] <= limit_range[4])


def _pipeline(samples, point_cloud_range, align_two_frames, ego_radius, f_ego, f_conv, f_mask):
    rng32 = np.asarray(point_cloud_range, np.float32)
    cur, prev = [], []
    for i, s in enumerate(samples):
        p, q = f_ego(s["points"], ego_radius), f_ego(s["points_prev"], ego_radius)
        if align_two_frames and s.get("frame_id", 0) != s.get("frame_id_prev", 1):
            q = f_conv(q, s["pose_prev"], s["pose"])
        for dst, a in ((cur, p), (prev, q)):
            a = a[f_mask(a, rng32)]
            dst.append(np.pad(a, ((0, 0), (1, 0)), mode="constant", constant_values=i))
    return np.concatenate(cur, 0).astype(np.float32), np.concatenate(prev, 0).astype(np.float32)


def assemble(samples, point_cloud_range, align_two_frames=True, ego_radius=2):
    """samples: list of dict(points, points_prev, pose, pose_prev[, frame_id, frame_id_prev]) -> (points, points_prev)
    collated float32 arrays (N', 1 + F)."""
    return _pipeline(samples, point_cloud_range, align_two_frames, ego_radius, remove_ego_points, convert_prv_frame_to_cur,
                     mask_points_by_range)


def tier1_available():
    from . import ref_loader
    return ref_loader.available()


def assemble_tier1(samples, point_cloud_range, align_two_frames=True, ego_radius=2):
    """Same pipeline through the reference's own functions (read in place, nothing copied)."""
    from . import ref_loader
    root = ref_loader.ref_root()
    path = os.path.join(root, "pcdet", "datasets", "once_temporal", "once_eval", "once_utils.py")
    spec = importlib.util.spec_from_file_location("_tmae_ref_once_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cu = ref_loader.load().common_utils
    return _pipeline(samples, point_cloud_range, align_two_frames, ego_radius, mod.remove_ego_points, mod.convert_prv_frame_to_cur,
                     cu.mask_points_by_range)
