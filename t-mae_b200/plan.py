"""Geometry plan: every coordinate-only table of one forward pass, built up front on the device with ONE
host read of the row counts (the reference does ~96 host round trips for the same information).

For each frame: the three stage site sets (stage 1 = the voxels, stages 2/3 = SparseConv2d k3 s2 p1 output sites,
spt_backbone.py:282-284), their submanifold / strided neighbour tables, and the two-shift window partition of
every stage (spt_backbone.py:137-184); for the frame pair: the temporal partition of every stage
(SiamWCA.py:201-269).
"""
import torch

from . import ops


class Stage:
    __slots__ = ("indices", "m", "Y", "X", "subm", "down", "down_t", "part")


class FramePlan:
    def __init__(self):
        self.stages = []


def _levels(pre_cfg):
    info = pre_cfg["DROP_INFO"]["train"]  # SSTInputLayer picks 'train' at construction (spt_backbone.py:32)
    keys = sorted(info, key=lambda k: int(k))
    return [(int(info[k]["max_tokens"]), int(info[k]["drop_range"][0]), int(info[k]["drop_range"][1])) for k in keys]


def check_status(part, what):
    s = int(part.status.item())
    if s:
        msgs = [m for b, m in ((1, "voxel coordinates are not in ascending (b,y,x) order"), (2, "coordinate outside the grid"),
                               (4, "a window's voxel count falls in no DROP_INFO level"),
                               (8, "a window exceeds its level's max_tokens (voxel dropping is not supported)")) if s & b]
        raise RuntimeError(f"window partition ({what}): " + "; ".join(msgs))


def build_plans(indices_list, batch, sparse_shape, block_cfgs, temporal_pair=None, want_ref=False, check=False, batches=None, need=None):
    """indices_list: per frame (M,3) i32 stage-1 sites.  temporal_pair = (i_cur, i_prev) adds the temporal
    partitions.  `batches[i]` overrides the sample count of frame set i (the Siamese-batched set holds 2B samples);
    `need[i]` is a subset of {"subm", "part"} (default both).  Returns (frame plans, temporal partitions per stage
    or None)."""
    Y, X = int(sparse_shape[0]), int(sparse_shape[1])
    n_stage = len(block_cfgs)
    batches = [batch] * len(indices_list) if batches is None else batches
    need = [("subm", "part")] * len(indices_list) if need is None else need
    plans, pending = [], []
    for idx, batch_i in zip(indices_list, batches):
        fp = FramePlan()
        st = Stage()
        st.indices, st.m, st.Y, st.X, st.down, st.down_t = idx, idx.shape[0], Y, X, None, None
        fp.stages.append(st)
        rows_dev, y, x = None, Y, X
        for s in range(1, n_stage):
            assert block_cfgs[s]["ENCODER"]["STRIDE"] == 2
            prev = fp.stages[-1]
            idx_out, n_out, table, table_t, (y, x) = ops.strided_table(prev.indices, batch_i, prev.Y, prev.X, rows_dev)
            st = Stage()
            st.indices, st.m, st.Y, st.X, st.down, st.down_t = idx_out, None, y, x, table, table_t
            fp.stages.append(st)
            pending.append(n_out)
            rows_dev = n_out
        plans.append(fp)
    if pending:
        counts = torch.cat(pending).cpu().tolist()  # the one host sync of the plan
        it = iter(counts)
        for fp in plans:
            for s in range(1, n_stage):
                st = fp.stages[s]
                st.m = int(next(it))
                st.indices = st.indices[:st.m]
                st.down = st.down[:st.m]
                if s + 1 < n_stage:
                    fp.stages[s + 1].down_t = fp.stages[s + 1].down_t[:st.m]
    for fp, batch_i, nd in zip(plans, batches, need):
        for s, st in enumerate(fp.stages):
            st.subm = ops.subm_table(st.indices, batch_i, st.Y, st.X) if "subm" in nd else None
            st.part = None
            if "part" in nd:
                st.part = ops.window_partition(st.indices, batch_i, st.X, st.Y, _levels(block_cfgs[s]["PREPROCESS"]), want_ref=want_ref)
                if check:
                    check_status(st.part, f"stage {s}")
    tparts = None
    if temporal_pair is not None:
        a, b = plans[temporal_pair[0]], plans[temporal_pair[1]]
        tparts = []
        for s in range(n_stage):
            tp = ops.window_partition(a.stages[s].indices, batch, a.stages[s].X, a.stages[s].Y,
                                      _levels(block_cfgs[s]["PREPROCESS"]), coords_b=b.stages[s].indices, want_ref=want_ref)
            tp.keep_a = (tp.win_a >= 0).to(torch.uint8)
            tp.keep_b = (tp.win_b >= 0).to(torch.uint8)
            if check:
                check_status(tp, f"temporal stage {s}")
            tparts.append(tp)
    return plans, tparts
