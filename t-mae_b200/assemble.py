"""GPU batch assembly -- rows N2 / N3 of SURVEY.md section 8f, the step in front of the VFE.

The reference prepares `points` / `points_prev` per sample in numpy on DataLoader workers and then uploads them:

    remove_ego_points(points, 2)            pcdet/datasets/once_temporal/once_temporal_dataset.py:167-168
                                            (once_eval/once_utils.py:43-45)
    convert_prv_frame_to_cur(prev, ...)     once_temporal_dataset.py:169-173 (once_utils.py:4-29), only when
                                            ALIGN_TWO_FRAMES and frame_id != frame_id_prev
    mask_points_by_range                    processor/data_processor.py:81-83 (utils/common_utils.py:124-127)
    collate_batch                           datasets/dataset.py:203-208 (sample index prepended)
    load_data_to_gpu                        pcdet/models/__init__.py:16-23 (`.float().cuda()`, blocking, per key)

Here the raw float32 point arrays of all samples of a frame set go to the device in ONE pinned-memory copy and one library
call (`tmae_assemble_frames`, csrc/assemble.cu) does the rest; the 7-number poses become two 3x4 float64 affine maps per
sample on the host (`pose_affines`; quaternion -> matrix and the 4x4 inverse exactly as the reference computes them with
scipy / numpy).  The output rows keep the input order of the kept points, i.e. they equal the reference's collated tensor
before its `shuffle_points` permutation (data_processor.py:92-102; host RNG, not reproduced -- the VFE is invariant to
point order up to fp32 summation order).  Augmentations (`data_augmentor`) are out of scope.

`FrameAssembler.__call__(samples)` returns a `batch_dict` ready for `TemporalDynVFE.forward`.
"""
import numpy as np
import torch
from scipy.spatial.transform import Rotation

from . import ops


def pose_affines(pose_prev, pose_cur):
    """-> (xform (2, 3, 4) float64, flags (2,) uint8): prev -> global (`p @ R.T + t`) and global -> current (first three
    rows of inv([[R, t], [0, 1]])), each skipped (flag 0) when its pose vector is all zeros (once_utils.py:9,18: ONCE
    stores the zero vector for a static ego vehicle).  A zero-norm quaternion raises ValueError like the reference's
    scipy call does (caught and resampled at once_temporal_dataset.py:174-183)."""
    xform, flags = np.zeros((2, 3, 4), np.float64), np.zeros(2, np.uint8)
    pose_prev, pose_cur = np.asarray(pose_prev, np.float64), np.asarray(pose_cur, np.float64)
    if np.any(pose_prev):
        xform[0, :, :3] = Rotation.from_quat(pose_prev[:4]).as_matrix()
        xform[0, :, 3] = pose_prev[4:]
        flags[0] = 1
    if np.any(pose_cur):
        m = np.zeros((4, 4))
        m[:3, :3] = Rotation.from_quat(pose_cur[:4]).as_matrix()
        m[:3, 3] = pose_cur[4:]
        m[3, 3] = 1
        xform[1] = np.linalg.inv(m)[:3]
        flags[1] = 1
    return xform, flags


class _Staging:
    """Two pinned host buffers per role used alternately; a buffer is reused only after the copy that read it finished."""

    def __init__(self):
        self.bufs, self.events, self.turn = [None, None], [None, None], 0

    def get(self, nbytes):
        i = self.turn = self.turn ^ 1
        if self.events[i] is not None:
            self.events[i].synchronize()
        if self.bufs[i] is None or self.bufs[i].numel() < nbytes:
            self.bufs[i] = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, pin_memory=True)
        return self.bufs[i]

    def sent(self):
        ev = self.events[self.turn] or torch.cuda.Event()
        ev.record()
        self.events[self.turn] = ev


class FrameAssembler:
    """point_cloud_range as in the dataset config (x0, y0, z0, x1, y1, z1; stored as float32 like DatasetTemplate does,
    dataset.py:25); ego_radius = the reference's hard-coded 2 m; align_two_frames = DATA_CONFIG.ALIGN_TWO_FRAMES."""

    def __init__(self, point_cloud_range, ego_radius=2.0, align_two_frames=True, device="cuda"):
        r = np.asarray(point_cloud_range, np.float32)
        self.crop_xyxy = [float(r[0]), float(r[1]), float(r[3]), float(r[4])]
        self.ego_radius = float(ego_radius)
        self.align_two_frames = bool(align_two_frames)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FrameAssembler runs on a CUDA device (tmae_b200 has no CPU path)")
        self._stage = {}
        self.h2d_bytes = 0   # host->device bytes of the last assemble() call

    def assemble(self, frames, xforms=None, role="points", sync=True):
        """frames: list of (n_i, F) float32 arrays (one per sample, raw sensor order); xforms: None or a list of
        (xform, flags) per sample.  -> (N', 1 + F) float32 device tensor.  sync=True reads the kept count back (one 8-byte
        host read) and returns exactly the kept rows; sync=False returns all sum(n_i) rows, the tail holding the
        out-of-range sentinel rows the voxeliser drops, with no host read."""
        batch = len(frames)
        feats = int(frames[0].shape[1])
        sizes = [int(f.shape[0]) for f in frames]
        n = sum(sizes)
        # one pinned block: [raw points | sample offsets i64 | xform f64 | flags u8]
        raw_b = n * feats * 4
        off_at = (raw_b + 7) // 8 * 8
        xf_at = off_at + (batch + 1) * 8
        fl_at = xf_at + (batch * 24 * 8 if xforms is not None else 0)
        total = fl_at + (batch * 2 if xforms is not None else 0)
        st = self._stage.setdefault(role, _Staging())
        host = st.get(total)
        hv = host.numpy()
        row = 0
        raw_view = hv[:raw_b].view(np.float32).reshape(n, feats)
        for f, k in zip(frames, sizes):
            if f.dtype != np.float32 or f.ndim != 2 or f.shape[1] != feats:
                raise ValueError("frames must be (n, F) float32 arrays with one F")
            raw_view[row:row + k] = f
            row += k
        hv[off_at:xf_at].view(np.int64)[:] = np.concatenate([[0], np.cumsum(sizes)])
        if xforms is not None:
            hv[xf_at:fl_at].view(np.float64)[:] = np.stack([x for x, _ in xforms]).reshape(-1)
            hv[fl_at:total] = np.stack([fl for _, fl in xforms]).reshape(-1)
        with torch.cuda.device(self.device):
            dev = host[:total].to(self.device, non_blocking=True)
            st.sent()
            self.h2d_bytes = total
            raw = dev[:raw_b].view(torch.float32).view(n, feats)
            offs = dev[off_at:xf_at].view(torch.int64)
            xf = dev[xf_at:fl_at].view(torch.float64) if xforms is not None else None
            fl = dev[fl_at:total] if xforms is not None else None
            out, count = ops.assemble_frames(raw, offs, batch, xf, fl, self.ego_radius, self.crop_xyxy)
        if sync:
            return out[:int(count.item())]
        return out

    def __call__(self, samples, sync=True):
        """samples: list of dicts with `points`, `points_prev` (raw arrays as `get_lidar` returns them), `pose`,
        `pose_prev`, and optionally `frame_id`, `frame_id_prev` -> batch_dict(points, points_prev, batch_size)."""
        xforms = []
        for s in samples:
            same = "frame_id" in s and s["frame_id"] == s.get("frame_id_prev")
            if self.align_two_frames and not same:
                xforms.append(pose_affines(s["pose_prev"], s["pose"]))
            else:
                xforms.append((np.zeros((2, 3, 4)), np.zeros(2, np.uint8)))
        cur = self.assemble([s["points"] for s in samples], None, "points", sync)
        prev = self.assemble([s["points_prev"] for s in samples], xforms, "points_prev", sync)
        # the stream these tensors were produced on and an event recorded behind them: a consumer that runs its first kernels on ANOTHER
        # stream (the VFE / backbone pre-pass with batch_dict["side_stream"]) waits for the event and marks the tensors as used there
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        return dict(points=cur, points_prev=prev, batch_size=len(samples), inputs_ready_event=ev)
