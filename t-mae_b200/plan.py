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


def validate_levels(pre_cfg, win_tokens=64):
    """Construction-time check of DROP_INFO (the reference asserts at run time that every voxel gets a level and a
    positive max_tokens, spt_backbone.py:62-63, and DROPS voxels beyond max_tokens, :65): the kernels do not drop, so
    the levels must cover every window population 1..win_tokens without gaps and each level must hold its largest
    window.  Both T-MAE configs satisfy this; anything else is rejected here instead of yielding silently
    wrong attention."""
    lv = _levels(pre_cfg)
    if not lv or len(lv) > 8:
        raise NotImplementedError("DROP_INFO needs 1..8 levels")
    covered = [False] * (win_tokens + 1)
    for tok, lo, hi in lv:
        if tok <= 0 or tok > win_tokens:
            raise NotImplementedError(f"DROP_INFO max_tokens {tok} outside 1..{win_tokens}")
        top = min(hi - 1, win_tokens)
        if top > tok:
            raise NotImplementedError(f"DROP_INFO level [{lo},{hi}) with max_tokens {tok} would drop voxels: not supported "
                                      "(both T-MAE configs never drop, t_mae_ssl.yaml:61-66)")
        for c in range(max(lo, 1), top + 1):
            if covered[c]:
                raise NotImplementedError(f"DROP_INFO levels overlap at {c} voxels per window")
            covered[c] = True
    missing = [c for c in range(1, win_tokens + 1) if not covered[c]]
    if missing:
        raise NotImplementedError(f"DROP_INFO has no level for windows holding {missing[0]} voxels")
    return lv


_STATUS_MSGS = ((1, "voxel coordinates are not in ascending (b,y,x) order"), (2, "coordinate outside the grid"),
                (4, "a window's voxel count falls in no DROP_INFO level"),
                (8, "a window exceeds its level's max_tokens (voxel dropping is not supported)"))


def _raise_status(s, what):
    raise RuntimeError(f"window partition ({what}): " + "; ".join(m for b, m in _STATUS_MSGS if s & b))


def check_status(part, what):
    s = int(part.status.item())
    if s:
        _raise_status(s, what)


class StatusLedger:
    """The partition kernels' status words (unsorted / out-of-grid coordinates, a window population in no level, a
    window larger than its level) of ONE forward, checked without a host sync of their own: all partitions of a plan
    write into one device buffer, which is copied to pinned host memory right behind them; the copy is inspected at the
    next build_plans call (or `flush()`), by when it has long completed.  A violation therefore raises one forward
    late, but always raises -- the reference asserts on each of these (spt_backbone.py:62-63,199-200, sst_utils.py:74-75)."""

    def __init__(self):
        self.pending = []  # [(pinned host tensor, event, names, device buffer)], oldest first
        self.free = []     # pinned buffers ready for reuse: a cudaHostAlloc per forward (the caching host allocator misses whenever the
                           # host runs ahead) costs milliseconds and can stall the launch queue

    def begin(self, n, device):
        self.flush()
        self.dev = torch.zeros(max(1, n), dtype=torch.int32, device=device)
        self.names, self.used = [], 0
        return self

    def slot(self, name):
        i = self.used
        self.used += 1
        self.names.append(name)
        return self.dev[i:i + 1]

    def commit(self):
        n = self.dev.shape[0]
        i = next((i for i, h in enumerate(self.free) if h.shape[0] >= n), None)
        host = torch.empty(max(n, 32), dtype=torch.int32, pin_memory=True) if i is None else self.free.pop(i)
        host[:n].copy_(self.dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.pending.append((host, ev, list(self.names), self.dev))

    def flush(self, wait=False):
        while self.pending:
            host, ev, names, _ = self.pending[0]
            if not wait and not ev.query():
                return  # still in flight (the host ran ahead): keep it for the next call
            ev.synchronize()
            self.pending.pop(0)
            vals = host[:len(names)].tolist()
            if len(self.free) < 8:
                self.free.append(host)
            for v, name in zip(vals, names):
                if v:
                    _raise_status(v, name + ", an earlier forward")


_ledgers = {}


def ledger(device):
    idx = torch.device(device).index or 0
    if idx not in _ledgers:
        _ledgers[idx] = StatusLedger()
    return _ledgers[idx]


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
    n_parts = sum(n_stage for nd in need if "part" in nd) + (n_stage if temporal_pair is not None else 0)
    led = ledger(indices_list[0].device).begin(n_parts, indices_list[0].device)
    for i, (fp, batch_i, nd) in enumerate(zip(plans, batches, need)):
        for s, st in enumerate(fp.stages):
            st.subm = ops.subm_table(st.indices, batch_i, st.Y, st.X) if "subm" in nd else None
            st.part = None
            if "part" in nd:
                st.part = ops.window_partition(st.indices, batch_i, st.X, st.Y, _levels(block_cfgs[s]["PREPROCESS"]), want_ref=want_ref,
                                               status=led.slot(f"frame set {i}, stage {s}"))
                if check:
                    check_status(st.part, f"stage {s}")
    tparts = None
    if temporal_pair is not None:
        a, b = plans[temporal_pair[0]], plans[temporal_pair[1]]
        tparts = []
        for s in range(n_stage):
            tp = ops.window_partition(a.stages[s].indices, batch, a.stages[s].X, a.stages[s].Y,
                                      _levels(block_cfgs[s]["PREPROCESS"]), coords_b=b.stages[s].indices, want_ref=want_ref,
                                      status=led.slot(f"temporal, stage {s}"))
            tp.keep_a = (tp.win_a >= 0).to(torch.uint8)
            tp.keep_b = (tp.win_b >= 0).to(torch.uint8)
            if check:
                check_status(tp, f"temporal stage {s}")
            tparts.append(tp)
    led.commit()
    return plans, tparts
