"""Thin tensor-level wrappers over the C ABI (include/tmae_sm100.h): pointer passing only.

Every function takes CUDA tensors, allocates the outputs/workspace with torch (device memory
plumbing) and enqueues the kernels on torch's current stream.  No arithmetic happens here and
there is no fallback path: a tensor that is not on a CUDA device raises.
"""
import ctypes
import weakref

import torch

from ._lib import lib

PREC_FP32, PREC_TF32, PREC_BF16 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_RELU = 0, 1, 2
ACT_GELU_DERIV = 3   # bf16 mode: y = GELU(.), the pre-activation output receives GELU'(.) (tmae_sm100.h TMAE_ACT_GELU_DERIV)
PRECISIONS = {"fp32": PREC_FP32, "tf32": PREC_TF32, "bf16": PREC_BF16}

_state = {"precision": PREC_FP32, "launches": 0}


def set_precision(p):
    """'fp32': parity mode (FFMA GEMMs, IEEE math, fp32 storage).
    'tf32': tcgen05 kind::tf32 GEMMs + TF32 mma attention, fp32 accumulate AND fp32 storage.
    'bf16': the throughput mode -- bf16 activation storage in the encoder, tcgen05 kind::f16 GEMMs, fp32 accumulate,
            fp32 statistics / master weights / weight gradients."""
    if PRECISIONS[p] == PREC_BF16 and not BF16_READY:
        raise NotImplementedError("the bf16-storage mode is not built yet")
    _state["precision"] = PRECISIONS[p]


BF16_READY = True
BENCH_PRECISION = "bf16"   # the mode bench.py and smoke() run by default


def precision():
    return _state["precision"]


def precision_name():
    return {v: k for k, v in PRECISIONS.items()}[_state["precision"]]


def _gemm_prec():
    """precision argument of the fp32-storage entry points (the bf16-storage layers have their own)."""
    return PREC_FP32 if _state["precision"] == PREC_FP32 else PREC_TF32


def set_option(name, value):
    """Library measurement switches (tmae_set_option): "wide_st", "attn_occ", "attn_occ_fwd", "bn_colsum_cap", "ln_bwd_cap"."""
    lib().set_option(name.encode(), int(value))


def dispatch_counts():
    """{'tma': n, 'simt_fp32': n, 'simt_in_tc_mode': n, 'thin_k': n}: which kernel served the GEMM-shaped calls so far
    (tmae_dispatch_counts).  'simt_in_tc_mode' must stay 0 on the hot path of a tensor-core mode."""
    out = (ctypes.c_int64 * 4)()
    lib().dispatch_counts(out, 4)
    return dict(zip(("tma", "simt_fp32", "simt_in_tc_mode", "thin_k"), [int(v) for v in out]))


def launch_count():
    """Kernel launches of this library so far, counted inside the library at its instrumented launch sites (a lower
    bound: memsets and a few small kernels are not counted)."""
    return int(lib().launch_count())


def _p(t, dtype=None):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("tmae_b200 ops need CUDA tensors (there is no CPU path)")
    if not t.is_contiguous():
        raise RuntimeError("tmae_b200 ops need contiguous tensors")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected {dtype}, got {t.dtype}")
    return t.data_ptr()


def _stream():
    # raw handle of torch's current stream (the Python Stream object costs ~15 us per call, ~160 calls per step)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


_side = {}


def side_stream(device=None):
    """The library's high-priority side stream of a device: the coordinate-only pre-pass of a step (voxelisation, masks,
    geometry plans and their host reads of row counts) runs there when the caller opts in with
    batch_dict["side_stream"], so those reads wait for a few tiny kernels instead of draining the main stream's
    backlog, and the host keeps running ahead of the GPU.
    The inputs (`points`, `points_prev`) must be complete ON the side stream: either they were produced there (an input pipeline that
    issues its H2D copies on it, as bench.py does), or batch_dict["inputs_ready_event"] holds an event recorded behind their producer
    (FrameAssembler returns one) -- the pre-pass waits for it and marks the tensors as used by the side stream."""
    idx = torch.cuda.current_device() if device is None else torch.device(device).index
    if idx not in _side:
        _side[idx] = torch.cuda.Stream(device=idx, priority=-1)
    return _side[idx]


def record_all(obj, stream, _seen=None):
    """record_stream(stream) on every CUDA tensor reachable from obj (dicts, sequences, objects with __slots__ or
    __dict__): tensors allocated on the side stream are consumed by kernels of the main stream.  No depth limit (a visited
    set guards cycles): a tensor that is missed returns to the side stream's pool the moment its last reference dies and
    is handed to the NEXT step's pre-pass while this step's backward still reads it."""
    if obj is None:
        return
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda:
            obj.record_stream(stream)
        return
    if isinstance(obj, (str, bytes, int, float, bool, torch.nn.Module, ctypes.Structure)):
        return
    if _seen is None:
        _seen = set()
    if id(obj) in _seen:
        return
    _seen.add(id(obj))
    if isinstance(obj, dict):
        for v in obj.values():
            record_all(v, stream, _seen)
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            record_all(v, stream, _seen)
    elif hasattr(obj, "__slots__"):
        for k in obj.__slots__:
            record_all(getattr(obj, k, None), stream, _seen)
    elif hasattr(obj, "__dict__"):
        for v in vars(obj).values():
            record_all(v, stream, _seen)


def _f3(v):
    return (ctypes.c_float * 3)(*[float(x) for x in v[:3]])


def _i32(v):
    return (ctypes.c_int32 * len(v))(*[int(x) for x in v])


TENSOR_BOUND = {"linear_fwd", "linear_bwd_data", "linear_bwd_weight", "sparse_conv_fwd", "sparse_conv_bwd_weight"}


def _call(name, *a, flops=0, nbytes=0):
    _state["launches"] += 1
    prof = _state.get("prof")
    if prof is None:
        getattr(lib(), name)(*a)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    getattr(lib(), name)(*a)
    e1.record()
    prof.append((name, e0, e1, flops, nbytes))


def profile_begin():
    """Time every ABI call with CUDA events on the launching stream (diagnostic pass, not the timed region)."""
    torch.cuda.synchronize()
    _state["prof"] = []


def profile_end():
    """-> {abi name: {ms, calls, flops, bytes}} (algorithmic work as stated in DESIGN.md)."""
    torch.cuda.synchronize()
    rec, _state["prof"] = _state.get("prof") or [], None
    out = {}
    for name, e0, e1, fl, nb in rec:
        r = out.setdefault(name, {"ms": 0.0, "calls": 0, "flops": 0, "bytes": 0})
        r["ms"] += e0.elapsed_time(e1)
        r["calls"] += 1
        r["flops"] += fl
        r["bytes"] += nb
    return out


def lib_profile(fn):
    """Runs fn() under the library's per-kernel CUDA-event profiler and returns the aggregated table."""
    L = lib()
    L.profile_begin()
    fn()
    buf = ctypes.create_string_buffer(1 << 16)
    L.profile_end(buf, len(buf))
    out = {}
    for line in buf.value.decode().splitlines():
        name, calls, ms, fl, by = line.split()
        out[name] = {"calls": int(calls), "ms": float(ms), "flops": float(fl), "bytes": float(by)}
    return out


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


F32, I32, I64, U8, BF16 = torch.float32, torch.int32, torch.int64, torch.uint8, torch.bfloat16


# ------------------------------------------------------------------------------------ voxelise
def voxelize(points, pc_range, voxel_size, grid_size, batch_size):
    """-> dict with capacity-sized device buffers and `counts` (2+B,) i64 on the device."""
    L = lib()
    n, stride = points.shape
    dev = points.device
    grid = _i32(grid_size)
    cells = batch_size * int(grid_size[0]) * int(grid_size[1]) * int(grid_size[2])
    mcap = max(1, min(n, cells))
    ncap = max(1, n)
    out = dict(
        points=torch.empty(ncap, stride, dtype=F32, device=dev),
        point_coords=torch.empty(ncap, 4, dtype=I64, device=dev),
        inverse=torch.empty(ncap, dtype=I64, device=dev),
        voxel_coords=torch.empty(mcap, 4, dtype=I64, device=dev),
        voxel_mean=torch.empty(mcap, stride - 1, dtype=F32, device=dev),
        voxel_npts=torch.empty(mcap, dtype=I32, device=dev),
        voxel_offset=torch.empty(mcap + 1, dtype=I32, device=dev),
        pt_order=torch.empty(ncap, dtype=I32, device=dev),
        counts=torch.empty(2 + batch_size, dtype=I64, device=dev),
    )
    wsb = L.voxelize_workspace_bytes(n, batch_size, grid)
    ws = _ws(wsb, dev)
    _call("voxelize", _p(points, F32), n, stride, _f3(pc_range), _f3(voxel_size), grid, batch_size, _p(out["points"]),
          _p(out["point_coords"]), _p(out["inverse"]), _p(out["voxel_coords"]), _p(out["voxel_mean"]), _p(out["voxel_npts"]),
          _p(out["voxel_offset"]), _p(out["pt_order"]), _p(out["counts"]), _p(ws), wsb, _stream())
    return out


def assemble_frames(raw, sample_offsets, batch, xform, xform_flags, ego_radius, crop_xyxy):
    """raw (n, F) f32 device rows [x,y,z,feat...] of `batch` samples back to back, sample_offsets (batch+1,) i64 device,
    xform (batch,2,3,4) f64 / xform_flags (batch,2) u8 device or None -> (out (n, 1+F) f32, count (1,) i64 device)."""
    L = lib()
    n, feats = raw.shape
    dev = raw.device
    out = torch.empty(n, feats + 1, dtype=F32, device=dev)
    count = torch.empty(1, dtype=I64, device=dev)
    wsb = L.assemble_frames_workspace_bytes(n)
    ws = _ws(wsb, dev)
    crop = (ctypes.c_float * 4)(*[float(v) for v in crop_xyxy])
    _call("assemble_frames", _p(raw, F32), _p(sample_offsets, I64), batch, feats, _p(xform, torch.float64), _p(xform_flags, U8),
          float(ego_radius), crop, _p(out), _p(count), n, _p(ws), wsb, _stream())
    return out, count


def vfe_point_features(points_kept, point_coords, inverse, voxel_mean, pc_range, voxel_size):
    n, stride = points_kept.shape
    x = torch.empty(n, stride - 1 + 6, dtype=F32, device=points_kept.device)
    _call("vfe_point_features", _p(points_kept, F32), n, stride, _p(point_coords, I64), _p(inverse, I64), _p(voxel_mean, F32),
          _f3(pc_range), _f3(voxel_size), _p(x), _stream())
    return x


# ------------------------------------------------------------------------------------ partition
class Partition:
    """Device tables of one window partition (both shifts); see tmae_window_partition."""
    __slots__ = ("m_a", "m_b", "wcap", "n_levels", "tokens", "win_a", "slot_a", "posidx_a", "tok_a", "cnt_a", "win_b", "slot_b",
                 "posidx_b", "tok_b", "cnt_b", "win_level", "n_win", "level_base", "status", "ref_a", "ref_b", "temporal",
                 "keep_a", "keep_b", "onehot_a", "onehot_b", "tcache")


def window_partition(coords_a, batch, grid_x, grid_y, levels, coords_b=None, want_ref=False, status=None):
    """coords_* (m,3) i32 ascending [b,y,x]; levels = [(max_tokens, lo, hi), ...]."""
    L = lib()
    dev = coords_a.device
    P = Partition()
    P.temporal = coords_b is not None
    P.m_a, P.m_b = coords_a.shape[0], (coords_b.shape[0] if P.temporal else 0)
    P.n_levels, P.tokens = len(levels), [int(l[0]) for l in levels]
    wcap = P.wcap = int(L.partition_window_capacity(batch, grid_x, grid_y))

    def per_voxel(m):
        m = max(1, m)
        return (torch.empty(2, m, dtype=I32, device=dev), torch.empty(2, m, dtype=I32, device=dev),
                torch.empty(2, m, dtype=U8, device=dev))

    P.win_a, P.slot_a, P.posidx_a = per_voxel(P.m_a)
    P.tok_a = torch.empty(2, wcap * 64, dtype=I32, device=dev)
    P.cnt_a = torch.empty(2, wcap, dtype=I32, device=dev)
    if P.temporal:
        P.win_b, P.slot_b, P.posidx_b = per_voxel(P.m_b)
        P.tok_b = torch.empty(2, wcap * 64, dtype=I32, device=dev)
        P.cnt_b = torch.empty(2, wcap, dtype=I32, device=dev)
    else:
        P.win_b = P.slot_b = P.posidx_b = P.tok_b = P.cnt_b = None
    P.win_level = torch.empty(2, wcap, dtype=I32, device=dev)
    P.n_win = torch.empty(2, dtype=I32, device=dev)
    P.level_base = torch.empty(2, P.n_levels + 1, dtype=I32, device=dev)
    P.status = torch.empty(1, dtype=I32, device=dev) if status is None else status   # (1,) i32; may be a slice of a shared buffer
    P.ref_a = P.ref_b = None
    if want_ref:
        P.ref_a = [torch.empty(2, max(1, P.m_a), dtype=I64, device=dev) for _ in range(3)]
        if P.temporal:
            P.ref_b = [torch.empty(2, max(1, P.m_b), dtype=I64, device=dev) for _ in range(3)]
    ra = [_p(t) for t in P.ref_a] if P.ref_a else [None] * 3
    rb = [_p(t) for t in P.ref_b] if P.ref_b else [None] * 3
    wsb = L.window_partition_workspace_bytes(batch, grid_x, grid_y, P.n_levels)
    ws = _ws(wsb, dev)
    _call("window_partition", _p(coords_a, I32), P.m_a, _p(coords_b, I32) if P.temporal else None, P.m_b, batch, grid_x, grid_y,
          P.n_levels, _i32([l[1] for l in levels]), _i32([l[2] for l in levels]), _i32(P.tokens),
          _p(P.win_a), _p(P.slot_a), _p(P.posidx_a), _p(P.tok_a), _p(P.cnt_a),
          _p(P.win_b), _p(P.slot_b), _p(P.posidx_b), _p(P.tok_b), _p(P.cnt_b),
          _p(P.win_level), _p(P.n_win), _p(P.level_base), P.status.data_ptr(), *ra, *rb, _p(ws), wsb, _stream())
    P.keep_a = P.keep_b = None
    # (2, m, 64) one-hot form of posidx: second A operand of the packed q/k/v projection in tensor-core mode
    # (2, m, 64) one-hot form of posidx: second operand of the packed q/k/v projection (tf32 mode, fp32) / of its weight-gradient
    # GEMM (bf16 mode, bf16)
    oh = onehot64 if _state["precision"] == PREC_TF32 else (onehot64_bf16 if _state["precision"] == PREC_BF16 else None)
    P.onehot_a = oh(P.posidx_a) if (oh and P.m_a > 0) else None
    P.onehot_b = oh(P.posidx_b) if (oh and P.temporal and P.m_b > 0) else None
    return P


def onehot64(idx):
    """idx (..., m) u8 -> (..., m, 64) fp32 one-hot."""
    out = torch.empty(*idx.shape, 64, dtype=F32, device=idx.device)
    _call("onehot64", _p(idx, U8), _p(out), idx.numel(), _stream())
    return out


# ------------------------------------------------------------------------------------ GEMMs
def linear_fwd(x, w, bias=None, residual=None, act=ACT_NONE, want_preact=False, w_offset_rows=0, n=None):
    """y = act(x @ w[w_offset_rows : w_offset_rows + n].T + bias) + residual."""
    m, k = x.shape
    n = w.shape[0] if n is None else n
    y = torch.empty(m, n, dtype=F32, device=x.device)
    pre = torch.empty(m, n, dtype=F32, device=x.device) if want_preact else None
    wp = _p(w, F32) + w_offset_rows * k * 4
    bp = None if bias is None else _p(bias, F32) + w_offset_rows * 4
    _call("linear_fwd", _p(x, F32), wp, bp, _p(residual), _p(y), _p(pre), m, n, k, act, _gemm_prec(), _stream(),
          flops=2 * m * n * k, nbytes=4 * (m * k + n * k + m * n))
    return (y, pre) if want_preact else y


def linear_fwd_lut(x, w, lut, rowidx):
    """y = x @ w.T + lut[rowidx]  (lut (64, n) from pos_table, rowidx (m,) u8)."""
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=F32, device=x.device)
    _call("linear_fwd_lut", _p(x, F32), _p(w, F32), _p(lut, F32), _p(rowidx, U8), _p(y), m, n, k, _gemm_prec(), _stream(),
          flops=2 * m * n * k, nbytes=4 * (m * k + n * k + m * n))
    return y


def pos_table(pos_lut, w, bias, n_pos):
    """(64, n) table  pos_lut @ w[:n_pos].T (zero for the other columns) + bias, and its transpose (n, 64)."""
    n, c = w.shape
    t = torch.empty(64, n, dtype=F32, device=w.device)
    tt = torch.empty(n, 64, dtype=F32, device=w.device)
    _call("pos_table", _p(pos_lut, F32), _p(w, F32), _p(bias, F32), _p(t), _p(tt), n, n_pos, c, _stream())
    return t, tt


def linear_fwd_dual(x, w, x2, w2):
    """y = x @ w.T + x2 @ w2.T in one GEMM (tensor-core mode) or two accumulating SIMT GEMMs (fp32)."""
    m, k = x.shape
    n, k2 = w.shape[0], x2.shape[1]
    y = torch.empty(m, n, dtype=F32, device=x.device)
    _call("linear_fwd_dual", _p(x, F32), _p(w, F32), _p(x2, F32), _p(w2, F32), _p(y), m, n, k, k2, _gemm_prec(), _stream(),
          flops=2 * m * n * (k + k2), nbytes=4 * (m * (k + k2) + n * (k + k2) + m * n))
    return y


def onehot64_bf16(idx):
    """idx (..., m) u8 -> (..., m, 64) bf16 one-hot."""
    out = torch.empty(*idx.shape, 64, dtype=BF16, device=idx.device)
    _call("onehot64_bf16", _p(idx, U8), out.data_ptr(), idx.numel(), _stream())
    return out


def binned_colsum(dy, rowidx):
    t = torch.empty(64, dy.shape[1], dtype=F32, device=dy.device)
    _call("binned_colsum", _p(dy, F32), _p(rowidx, U8), _p(t), dy.shape[0], dy.shape[1], _stream())
    return t


def pos_table_bwd(dtable, pos_lut, dw, n_pos, transposed=False):
    """dbias = dtable.sum(0); dw[:n_pos] += dtable[:, :n_pos].T @ pos_lut (in place).  transposed: dtable is (n, 64)."""
    n = dtable.shape[0] if transposed else dtable.shape[1]
    db = torch.empty(n, dtype=F32, device=dtable.device)
    _call("pos_table_bwd", _p(dtable, F32), int(transposed), _p(pos_lut, F32), _p(dw, F32), _p(db), n, n_pos, pos_lut.shape[1], _stream())
    return db


def linear_bwd_data(dy, w, dx=None, accumulate=False, w_offset_rows=0):
    m, n = dy.shape
    k = w.shape[1]
    if dx is None:
        dx = torch.empty(m, k, dtype=F32, device=dy.device)
        accumulate = False
    _call("linear_bwd_data", _p(dy, F32), _p(w, F32) + w_offset_rows * k * 4, _p(dx), m, n, k, int(accumulate), _gemm_prec(), _stream(),
          flops=2 * m * n * k, nbytes=4 * (m * k + n * k + m * n))
    return dx


def linear_bwd_weight(dy, x, dw, dbias=None, w_offset_rows=0):
    """dw[w_offset_rows : +n] = dy.T @ x (overwrites that slice); dbias slice likewise."""
    m, n = dy.shape
    k = x.shape[1]
    dbp = None if dbias is None else _p(dbias, F32) + w_offset_rows * 4
    _call("linear_bwd_weight", _p(dy, F32), _p(x, F32), _p(dw, F32) + w_offset_rows * k * 4, dbp, m, n, k, _gemm_prec(), _stream(),
          flops=2 * m * n * k, nbytes=4 * (m * k + n * k + m * n))


def gelu_bwd(dy, preact):
    dx = torch.empty_like(dy)
    _call("gelu_bwd", _p(dy, F32), _p(preact, F32), _p(dx), dy.numel(), _stream())
    return dx


# ------------------------------------------------------------------------------------ sparse conv
def subm_table(indices, batch, Y, X, rows_dev=None):
    L = lib()
    m = indices.shape[0]
    table = torch.empty(max(1, m), 9, dtype=I32, device=indices.device)
    wsb = L.subm_table_workspace_bytes(batch, Y, X)
    ws = _ws(wsb, indices.device)
    _call("subm_table", _p(indices, I32), m, _p(rows_dev), batch, Y, X, _p(table), _p(ws), wsb, _stream())
    return table


def strided_table(indices, batch, Y, X, rows_dev=None):
    """-> indices_out (cap,3), n_out (device i32[1]), table (cap,9), table_t (m,9), (Yo, Xo)."""
    L = lib()
    m = indices.shape[0]
    Yo, Xo = (Y + 2 - 3) // 2 + 1, (X + 2 - 3) // 2 + 1
    cap = max(1, min(4 * m, batch * Yo * Xo))
    dev = indices.device
    idx_out = torch.empty(cap, 3, dtype=I32, device=dev)
    n_out = torch.empty(1, dtype=I32, device=dev)
    table = torch.empty(cap, 9, dtype=I32, device=dev)
    table_t = torch.empty(max(1, m), 9, dtype=I32, device=dev)
    wsb = L.strided_table_workspace_bytes(batch, Y, X)
    ws = _ws(wsb, dev)
    _call("strided_table", _p(indices, I32), m, _p(rows_dev), batch, Y, X, _p(idx_out), cap, _p(n_out), _p(table), _p(table_t),
          _p(ws), wsb, _stream())
    return idx_out, n_out, table, table_t, (Yo, Xo)


def sparse_conv_fwd(x, table, w, rows_out):
    cout, taps, cin = w.shape[0], w.shape[1] * w.shape[2] if w.dim() == 4 else w.shape[1], w.shape[-1]
    y = torch.empty(rows_out, cout, dtype=F32, device=x.device)
    _call("sparse_conv_fwd", _p(x, F32), _p(table, I32), _p(w, F32), _p(y), rows_out, taps, cin, cout, 0, _gemm_prec(), _stream(),
          flops=2 * rows_out * taps * cin * cout, nbytes=4 * (x.numel() + w.numel() + rows_out * cout) + 4 * rows_out * taps)
    return y


def sparse_conv_bwd_weight(dy, x, table, w_shape):
    cout, taps, cin = w_shape[0], w_shape[1] * w_shape[2], w_shape[3]
    dw = torch.empty(w_shape, dtype=F32, device=x.device)
    _call("sparse_conv_bwd_weight", _p(dy, F32), _p(x, F32), _p(table, I32), _p(dw), dy.shape[0], taps, cin, cout, _gemm_prec(), _stream(),
          flops=2 * dy.shape[0] * taps * cin * cout, nbytes=4 * (x.numel() + dy.numel() + dw.numel()))
    return dw


def transpose_taps(w, flip):
    cout, taps, cin = w.shape[0], w.shape[1] * w.shape[2], w.shape[3]
    wt = torch.empty(cin, taps, cout, dtype=F32, device=w.device)
    _call("transpose_taps", _p(w, F32), _p(wt), cout, taps, cin, int(flip), _stream())
    return wt


# ------------------------------------------------------------------------------------ row ops
def add_pos(x, posidx, lut):
    y = torch.empty_like(x)
    _call("add_pos", _p(x, F32), _p(posidx, U8), _p(lut, F32), _p(y), x.shape[0], x.shape[1], _stream())
    return y


def add_layernorm_fwd(x, res, rowmask, gamma, beta, eps):
    rows, c = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(max(1, rows), dtype=F32, device=x.device)
    rstd = torch.empty(max(1, rows), dtype=F32, device=x.device)
    _call("add_layernorm_fwd", _p(x, F32), _p(res), _p(rowmask), _p(gamma, F32), _p(beta, F32), _p(y), _p(mean), _p(rstd), rows, c,
          float(eps), _stream())
    return y, mean, rstd


def add_layernorm_bwd(dy, x, res, rowmask, gamma, mean, rstd, dgamma, dbeta, want_dres=False, dcolsum=None):
    """dcolsum (c,) optional: receives the column sums of dres (when want_dres) or of dv -- the bias gradient of the linear
    layer that produced `res`."""
    rows, c = x.shape
    dv = torch.empty_like(x)
    dres = torch.empty_like(x) if want_dres else None
    _call("add_layernorm_bwd_colsum", _p(dy, F32), _p(x, F32), _p(res), _p(rowmask), _p(gamma, F32), _p(mean), _p(rstd), _p(dv), _p(dres),
          _p(dgamma, F32), _p(dbeta, F32), _p(dcolsum, F32), rows, c, _stream())
    return dv, dres


def bn_train_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, relu, out=None):
    L = lib()
    rows, c = x.shape
    y = torch.empty_like(x) if out is None else out
    mean = torch.empty(c, dtype=F32, device=x.device)
    rstd = torch.empty(c, dtype=F32, device=x.device)
    wsb = L.bn_workspace_bytes(c)
    ws = _ws(wsb, x.device)
    _call("bn_train_fwd", _p(x, F32), _p(gamma, F32), _p(beta, F32), _p(running_mean), _p(running_var), float(momentum), float(eps),
          _p(y), _p(mean), _p(rstd), rows, c, int(relu), _p(ws), wsb, _stream())
    return y, mean, rstd


def bn_apply(x, mean, rstd, gamma, beta, relu, out=None):
    y = torch.empty_like(x) if out is None else out
    _call("bn_apply", _p(x, F32), _p(mean, F32), _p(rstd, F32), _p(gamma, F32), _p(beta, F32), _p(y), x.shape[0], x.shape[1], int(relu),
          _stream())
    return y


def bn_bwd(dy, x, beta, mean, rstd, gamma, relu, training, out=None):
    L = lib()
    rows, c = x.shape
    dx = torch.empty_like(x) if out is None else out
    dgamma = torch.empty(c, dtype=F32, device=x.device)
    dbeta = torch.empty(c, dtype=F32, device=x.device)
    wsb = L.bn_workspace_bytes(c)
    ws = _ws(wsb, x.device)
    _call("bn_bwd", _p(dy, F32), _p(x, F32), None, _p(mean), _p(rstd), _p(gamma, F32), _p(beta, F32), _p(dx), _p(dgamma), _p(dbeta), rows, c,
          int(relu), int(training), _p(ws), wsb, _stream())
    return dx, dgamma, dbeta


def bn_bf16_fwd(x, gamma, beta, running_mean, running_var, momentum, eps, relu, training, out, mean=None, rstd=None):
    """x (rows, C) bf16 contiguous; out (rows, C) bf16 view with any row pitch (a column slice of the concat buffer)."""
    L = lib()
    rows, c = x.shape
    if not (x.is_cuda and x.dtype == BF16 and out.dtype == BF16 and x.stride(1) == 1 and out.stride(1) == 1):
        raise RuntimeError("bn_bf16_fwd needs bf16 CUDA tensors with unit column stride")
    if training:
        mean = torch.empty(c, dtype=F32, device=x.device)
        rstd = torch.empty(c, dtype=F32, device=x.device)
    wsb = L.bn_workspace_bytes(c)
    ws = _ws(wsb, x.device)
    _call("bn_bf16_fwd", x.data_ptr(), x.stride(0), _p(gamma, F32), _p(beta, F32), _p(running_mean), _p(running_var), float(momentum), float(eps),
          out.data_ptr(), out.stride(0), _p(mean), _p(rstd), rows, c, int(relu), int(training), _p(ws), wsb, _stream())
    return mean, rstd


def bn_bf16_bwd(dy, x, mean, rstd, gamma, beta, relu, training, out=None):
    """dy (rows, C) bf16 view with any row pitch; x (rows, C) bf16 contiguous -> dx bf16 contiguous, dgamma, dbeta."""
    L = lib()
    rows, c = x.shape
    if not (dy.dtype == BF16 and x.dtype == BF16 and dy.stride(1) == 1 and x.stride(1) == 1):
        raise RuntimeError("bn_bf16_bwd needs bf16 tensors with unit column stride")
    dx = torch.empty(rows, c, dtype=BF16, device=x.device) if out is None else out
    dgamma = torch.empty(c, dtype=F32, device=x.device)
    dbeta = torch.empty(c, dtype=F32, device=x.device)
    wsb = L.bn_workspace_bytes(c)
    ws = _ws(wsb, x.device)
    _call("bn_bf16_bwd", dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), _p(mean, F32), _p(rstd, F32), _p(gamma, F32), _p(beta, F32),
          dx.data_ptr(), dx.stride(0), _p(dgamma), _p(dbeta), rows, c, int(relu), int(training), _p(ws), wsb, _stream())
    return dx, dgamma, dbeta


def segment_max_fwd(x, voxel_offset, pt_order, n_voxels):
    c = x.shape[1]
    out = torch.empty(n_voxels, c, dtype=F32, device=x.device)
    arg = torch.empty(n_voxels, c, dtype=I32, device=x.device)
    _call("segment_max_fwd", _p(x, F32), _p(voxel_offset, I32), _p(pt_order, I32), n_voxels, c, _p(out), _p(arg), _stream())
    return out, arg


def segment_max_bwd(dout, arg, n_points):
    n_voxels, c = dout.shape
    dx = torch.empty(n_points, c, dtype=F32, device=dout.device)
    _call("segment_max_bwd", _p(dout, F32), _p(arg, I32), n_voxels, c, _p(dx), n_points, _stream())
    return dx


def densify_nhwc(rows, indices, batch, Y, X, dtype=F32):
    """(B,Y,X,C) map of dtype fp32 or bf16 from fp32 rows."""
    m, c = rows.shape
    dense = torch.empty(batch, Y, X, c, dtype=dtype, device=rows.device)
    name = "densify_nhwc" if dtype == F32 else "densify_nhwc_bf16"
    _call(name, _p(rows, F32), _p(indices, I32), m, c, batch, Y, X, _p(dense), 1, _stream())
    return dense


def gather_nhwc(dense, indices):
    """fp32 rows from a (B,Y,X,C) fp32 or bf16 map."""
    _, Y, X, c = dense.shape
    m = indices.shape[0]
    rows = torch.empty(m, c, dtype=F32, device=dense.device)
    name = "gather_nhwc" if dense.dtype == F32 else "gather_nhwc_bf16"
    _call(name, _p(dense), _p(indices, I32), m, c, Y, X, _p(rows), _stream())
    return rows


# ------------------------------------------------------------------------------------ attention
def small_end(part, shift, cap=16):
    """Device pointer holder: number of leading (level-sorted) windows whose level holds <= cap tokens."""
    l = next((i for i, t in enumerate(part.tokens) if t > cap), part.n_levels)
    return part.level_base[shift, l:l + 1]


def mid_end(part, shift):
    return small_end(part, shift, 32)


def _pv(t):
    """pointer of a (rows, C) fp32 view with unit column stride (a column block of a packed projection is allowed)."""
    if not (t.is_cuda and t.dtype == F32 and t.stride(1) == 1 and t.stride(0) % 4 == 0):
        raise RuntimeError("expected an fp32 CUDA matrix with unit column stride and a row pitch that is a multiple of 4")
    return t.data_ptr()


def window_attention_fwd(q, k, v, qtok, qcnt, ktok, kcnt, n_win, small, mid, max_windows, tau, tau_min, heads, zero_out):
    mq, c = q.shape
    o = torch.zeros(mq, c, dtype=F32, device=q.device) if zero_out else torch.empty(mq, c, dtype=F32, device=q.device)
    lse = torch.empty(max(1, mq), heads, dtype=F32, device=q.device)
    _call("window_attention_fwd", _pv(q), _pv(k), _pv(v), _p(o), _p(lse), _p(qtok), _p(qcnt), _p(ktok), _p(kcnt), _p(n_win),
          _p(small), _p(mid), max_windows, _p(tau, F32), float(tau_min), c, heads, q.stride(0), k.stride(0), v.stride(0), mq, k.shape[0],
          _gemm_prec(), _stream(), nbytes=4 * (q.numel() * 2 + k.numel() * 2))
    return o, lse


def window_attention_bwd(dout, q, k, v, o, lse, qtok, qcnt, ktok, kcnt, n_win, small, mid, max_windows, tau, tau_min, heads, dtau, zero):
    alloc = torch.zeros if zero else torch.empty
    packed = q.stride(0) == 3 * q.shape[1] and k.data_ptr() == q.data_ptr() + 4 * q.shape[1]   # gradients mirror the packed layout
    if packed:
        buf = alloc(q.shape[0], 3 * q.shape[1], dtype=F32, device=q.device)
        dq, dk, dv = buf[:, :q.shape[1]], buf[:, q.shape[1]:2 * q.shape[1]], buf[:, 2 * q.shape[1]:]
    else:
        dq, dk, dv = (alloc(t.shape[0], t.shape[1], dtype=F32, device=t.device) for t in (q, k, v))
    dsum = torch.empty_like(lse)
    _call("window_attention_bwd", _p(dout, F32), _pv(q), _pv(k), _pv(v), _p(o, F32), _p(lse, F32), _p(dsum), _pv(dq), _pv(dk),
          _pv(dv), _p(dtau, F32), _p(qtok), _p(qcnt), _p(ktok), _p(kcnt), _p(n_win), _p(small), _p(mid), max_windows, _p(tau, F32), float(tau_min), q.shape[1],
          heads, q.stride(0), k.stride(0), v.stride(0), q.shape[0], k.shape[0], _gemm_prec(), _stream(), nbytes=4 * (q.numel() * 4 + k.numel() * 4))
    return dq, dk, dv


# ------------------------------------------------------------------------------------ the reference's native op, 1:1
def get_inner_win_inds(group_inds):
    """pcdet/ops/sst_ops/sst_ops_utils.py:5-12 with the canonical (stable, deterministic) slot order."""
    L = lib()
    g = group_inds.contiguous()
    out = torch.zeros_like(g) - 1
    wsb = L.sst_ops_workspace_bytes(g.numel())
    ws = _ws(wsb, g.device)
    _call("ingroup_inds", _p(g, I64), _p(out, I64), g.numel(), _p(ws), wsb, _stream())
    return out


def group_inner_inds(points, inverse_inds, K, n_groups=None):
    """pcdet/ops/sst_ops/sst_ops_utils.py:15-27.  n_groups (= number of voxels) skips the reference's .max().item() host sync."""
    L = lib()
    inv = inverse_inds.contiguous()
    m = int(inv.max().item()) + 1 if n_groups is None else int(n_groups)
    group_inds = torch.full((m, K), -1, dtype=I64, device=points.device)
    wsb = L.sst_ops_workspace_bytes(inv.numel())
    ws = _ws(wsb, inv.device)
    _call("group_inner_inds", _p(inv, I64), inv.numel(), _p(group_inds, I64), m, K, _p(ws), wsb, _stream())
    return points[group_inds]


# ------------------------------------------------------------------------------------ loss
def gt_group(points_kept, voxel_offset, pt_order, voxel_coords, pc_range, voxel_size, n_voxels, k, want_inds=False):
    dev = points_kept.device
    gt = torch.empty(n_voxels, k, 3, dtype=F32, device=dev)
    inds = torch.empty(n_voxels, k, dtype=I64, device=dev) if want_inds else None
    _call("gt_group", _p(points_kept, F32), points_kept.shape[1], _p(voxel_offset, I32), _p(pt_order, I32), _p(voxel_coords, I64),
          _f3(pc_range), _f3(voxel_size), n_voxels, k, _p(gt), _p(inds), _stream())
    return (gt, inds) if want_inds else gt


def chamfer_fwd(pred, gt, w, gtctx=None):
    """gt (M,P2,3) or None with gtctx = (points_kept, voxel_offset, pt_order, voxel_coords, pc_range, voxel_size, P2)."""
    m, p1, _ = pred.shape
    dev = pred.device
    loss = torch.empty((), dtype=F32, device=dev)
    state = torch.empty(4, dtype=torch.float64, device=dev)
    if gt is not None:
        p2 = gt.shape[1]
        ctx = (None, 0, None, None, None, None, None)
    else:
        pk, off, order, vc, rng, vs, p2 = gtctx
        ctx = (_p(pk, F32), pk.shape[1], _p(off, I32), _p(order, I32), _p(vc, I64), _f3(rng), _f3(vs))
    _call("chamfer_fwd", _p(pred, F32), _p(gt), _p(w, F32), m, p1, p2, *ctx, _p(loss), _p(state), _stream())
    return loss, state


def chamfer_bwd(grad_loss, pred, gt, w, state, gtctx=None):
    m, p1, _ = pred.shape
    dpred = torch.empty_like(pred)
    if gt is not None:
        p2 = gt.shape[1]
        ctx = (None, 0, None, None, None, None, None)
    else:
        pk, off, order, vc, rng, vs, p2 = gtctx
        ctx = (_p(pk, F32), pk.shape[1], _p(off, I32), _p(order, I32), _p(vc, I64), _f3(rng), _f3(vs))
    _call("chamfer_bwd", _p(grad_loss, F32), _p(pred, F32), _p(gt), _p(w, F32), m, p1, p2, *ctx, _p(state), _p(dpred), _stream())
    return dpred


# ------------------------------------------------------------------------------------ whole encoder layer
class LayerParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("in_w", "in_b", "out_w", "out_b", "tau", "ln1_g", "ln1_b", "w1", "b1", "w2", "b2",
                                               "ln2_g", "ln2_b")]


class LayerTables(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("posidx_q", "posidx_kv", "qtok", "qcnt", "ktok", "kcnt", "n_win", "small_end", "mid_end",
                                               "rowmask")] + [("max_windows", ctypes.c_int64), ("onehot_q", ctypes.c_void_p),
                                                              ("onehot_kv", ctypes.c_void_p)]


def layer_tables(part, shift, cross, m_q, m_kv):
    """tmae_layer_tables for one shift of a partition (self: frame a on both sides; cross: a = current, b = previous).  Built once
    per (partition, shift): the 2-4 layers that share a partition reuse the struct."""
    cache = getattr(part, "tcache", None)
    if cache is None:
        cache = part.tcache = {}
    key = (shift, cross, m_q, m_kv)
    if key in cache:
        return cache[key]
    T = cache[key] = LayerTables()
    T.posidx_q = _p(part.posidx_a[shift])
    T.qtok, T.qcnt = _p(part.tok_a[shift]), _p(part.cnt_a[shift])
    if cross:
        T.posidx_kv = _p(part.posidx_b[shift])
        T.ktok, T.kcnt = _p(part.tok_b[shift]), _p(part.cnt_b[shift])
        T.rowmask = _p(part.keep_a[shift])
        T.max_windows = min(part.wcap, m_q, m_kv)
    else:
        T.posidx_kv = None
        T.ktok, T.kcnt = T.qtok, T.qcnt
        T.rowmask = None
        T.max_windows = min(part.wcap, m_q)
    oa, ob = getattr(part, "onehot_a", None), getattr(part, "onehot_b", None)
    T.onehot_q = _p(oa[shift]) if oa is not None else None
    T.onehot_kv = _p(ob[shift]) if (cross and ob is not None) else None
    T.n_win = _p(part.n_win[shift:shift + 1])
    T.small_end = _p(small_end(part, shift))
    T.mid_end = _p(mid_end(part, shift))
    return T


def _layer_params(tensors):
    P = LayerParams()
    for (name, _), t in zip(LayerParams._fields_, tensors):
        setattr(P, name, _p(t, F32))
    return P


_param_structs = {}


def _layer_params_cached(params):
    """LayerParams of a layer's 13 parameter tensors, rebuilt only when their storage moves (module.to(...), load_state_dict with
    assign): 13 pointer checks per call otherwise."""
    key = id(params[0])
    hit = _param_structs.get(key)
    ptrs = tuple(t.data_ptr() for t in params)
    if hit is None or hit[0] != ptrs:
        hit = _param_structs[key] = (ptrs, _layer_params(params))
    return hit[1]


def encoder_layer_fwd(x, x_kv, params, T, lut, tau_min, eps, heads, need_backward=True):
    """params: the 13 tensors in tmae_layer_params order.  -> (y, saved buffer)."""
    L = lib()
    m_q, c = x.shape
    m_kv = x_kv.shape[0] if x_kv is not None else m_q
    ff = params[7].shape[0]
    cross = int(x_kv is not None)
    nb = L.encoder_layer_saved_bytes(m_q, m_kv, c, ff, heads, cross)
    saved = _ws(nb, x.device)
    y = torch.empty_like(x)
    P = _layer_params(params)
    _call("encoder_layer_fwd", _p(x, F32), _p(x_kv), ctypes.byref(P), ctypes.byref(T), _p(lut, F32), float(tau_min), float(eps), m_q, m_kv,
          c, ff, heads, _gemm_prec(), int(need_backward), _p(y), _p(saved), nb, _stream())
    return y, saved


def encoder_layer_bwd(dy, x, x_kv, params, T, lut, tau_min, heads, saved, want_dkv):
    """-> (dx, dx_kv or None, [13 parameter gradients])."""
    L = lib()
    m_q, c = x.shape
    m_kv = x_kv.shape[0] if x_kv is not None else m_q
    ff = params[7].shape[0]
    cross = int(x_kv is not None)
    sizes = [t.numel() for t in params]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + (n + 63) // 64 * 64)
    gbuf = torch.empty(offs[-1], dtype=F32, device=x.device)
    grads = [gbuf[o:o + n].view(t.shape) for o, n, t in zip(offs, sizes, params)]
    G = _layer_params(grads)
    P = _layer_params(params)
    nb = L.encoder_layer_scratch_bytes(m_q, m_kv, c, ff, heads, cross)
    scratch = _ws(nb, x.device)
    dx = torch.empty_like(x)
    dkv = torch.empty_like(x_kv) if (cross and want_dkv) else None
    _call("encoder_layer_bwd", _p(dy, F32), _p(x, F32), _p(x_kv), ctypes.byref(P), ctypes.byref(T), _p(lut, F32), float(tau_min), m_q, m_kv, c, ff, heads,
          _gemm_prec(), _p(saved), saved.numel(), _p(dx), _p(dkv), ctypes.byref(G), _p(scratch), nb, _stream())
    return dx, dkv, grads


# ==================================================================================== bf16-storage mode
def _pb(t):
    """pointer of a contiguous bf16 CUDA tensor (32-byte aligned: the TMEM epilogues store 256-bit words)."""
    if t is None:
        return None
    if not (t.is_cuda and t.dtype == BF16 and t.is_contiguous()):
        raise RuntimeError("expected a contiguous bf16 CUDA tensor")
    if t.data_ptr() % 32:
        raise RuntimeError("bf16 activations must be 32-byte aligned")
    return t.data_ptr()


def cast_bf16(x):
    """fp32 -> bf16 copy (round to nearest even)."""
    y = torch.empty(x.shape, dtype=BF16, device=x.device)
    _call("cast_f32_bf16", _p(x, F32), y.data_ptr(), x.numel(), _stream())
    return y


def cast_f32(x):
    y = torch.empty(x.shape, dtype=F32, device=x.device)
    _call("cast_bf16_f32", _p(x, BF16), y.data_ptr(), x.numel(), _stream())
    return y


# Fused / foreach optimizers update parameters WITHOUT moving `Tensor._version` (measured: torch.optim.AdamW(fused=True) leaves it
# unchanged), and so does any update through `p.data`.  Derived weight operands (bf16 shadows, [W | table] projections) therefore also
# key on an "epoch" that every torch optimizer step advances (global post-step hook), and training-mode forwards refresh unconditionally.
_weights_epoch = [0]


def _on_optimizer_step(*_args, **_kw):
    _weights_epoch[0] += 1


def weights_changed():
    """Tell the library that parameters were modified by means it cannot see (`p.data` arithmetic, a custom optimizer)."""
    _weights_epoch[0] += 1


try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook
    _reg_post_hook(_on_optimizer_step)
except ImportError:   # very old torch: training-mode forwards still refresh unconditionally
    pass


class WeightShadows:
    """bf16 copies of fp32 master weights (the tensor-core operands of the bf16 mode), refreshed when the master changes
    (`Tensor._version`, the optimizer epoch above, or `force`): ALL stale copies of a device in one launch through a device-resident
    segment table that is built once (the parameters and their shadows never move).  Parameters are held by weak reference: the copies
    of a model that was dropped go with it."""

    def __init__(self):
        self.items = {}      # id(param) -> [weakref(param), shadow, version, data_ptr]
        self.tables = {}     # device -> (tuple of item ids, (n, 3) int64 device table: src ptr, dst ptr, numel)
        self.epoch = -1      # _weights_epoch at the last full refresh
        self._keep = []

    def _fresh(self, p):
        it = self.items.get(id(p))
        if it is None or it[0]() is not p or it[3] != p.data_ptr():   # new parameter, or its storage moved (module.to(...), p.data = ...)
            it = self.items[id(p)] = [weakref.ref(p), torch.empty(p.shape, dtype=BF16, device=p.device), -1, p.data_ptr()]
            self.tables.pop(p.device, None)
        return it

    def get(self, p):
        it = self._fresh(p)
        if it[2] != p._version or self.epoch != _weights_epoch[0]:
            self.refresh()
        return it[1]

    def register(self, params):
        for p in params:
            self._fresh(p)

    def refresh(self, force=False):
        """force (or an optimizer step since the last refresh): every copy is stale, whatever the version counters say."""
        live = []
        for key, it in list(self.items.items()):
            p = it[0]()
            if p is None:                                   # the model is gone
                del self.items[key]
                self.tables.pop(it[1].device, None)
                continue
            if it[3] != p.data_ptr():                       # storage moved: re-create
                it = self._fresh(p)
            live.append((key, it, p))
        force = force or self.epoch != _weights_epoch[0]
        self.epoch = _weights_epoch[0]
        self._keep = []
        by_dev = {}
        for key, it, p in live:
            by_dev.setdefault(p.device, []).append((key, it, p))
        for dev, group in by_dev.items():
            stale = group if force else [g for g in group if g[1][2] != g[2]._version]
            if not stale:
                continue
            keys = tuple(k for k, _, _ in stale)
            cached = self.tables.get(dev) if len(stale) == len(group) else None
            if cached is not None and cached[0] == keys:
                table = cached[1]
            else:
                table = torch.tensor([[p.data_ptr(), it[1].data_ptr(), p.numel()] for _, it, p in stale], dtype=I64, device=dev)
                if len(stale) == len(group):
                    self.tables[dev] = (keys, table)
            with torch.cuda.device(dev):
                _call("cast_f32_bf16_multi", table.data_ptr(), len(stale), _stream())
            for _, it, p in stale:
                it[2] = p._version
            self._keep.append(table)   # alive until the next refresh (the launch reads it asynchronously)


shadows = WeightShadows()


def bf16_linear_fwd(x, w, bias=None, act=ACT_NONE, want_preact=False, out=None, accumulate=False):
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=BF16, device=x.device) if out is None else out
    pre = torch.empty(m, n, dtype=BF16, device=x.device) if want_preact else None
    _call("bf16_linear_fwd", _pb(x), _pb(w), _p(bias, F32), _pb(y), _pb(pre), m, n, k, act, int(accumulate), _stream())
    return (y, pre) if want_preact else y


def bf16_qkv_fwd(x, w, table, posidx, norm_cols, hd):
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=BF16, device=x.device)
    inv = torch.empty(m, norm_cols // hd, dtype=F32, device=x.device)
    _call("bf16_qkv_fwd", _pb(x), _pb(w), _p(table, F32), _p(posidx, U8), _pb(y), _p(inv), m, n, k, norm_cols, hd, _stream())
    return y, inv


def bf16_qkv_wcat(pos_lut, w, bias, n_pos):
    """(n, c + 64) bf16 [w | table^T] from the fp32 masters: the B operand of bf16_qkv_fwd_onehot."""
    n, c = w.shape
    wcat = torch.empty(n, c + 64, dtype=BF16, device=w.device)
    _call("bf16_qkv_wcat", _p(pos_lut, F32), _p(w, F32), _p(bias, F32), _pb(wcat), n, n_pos, c, _stream())
    return wcat


def bf16_qkv_fwd_onehot(x, onehot, wcat, norm_cols, hd):
    """bf16_qkv_fwd with the position term inside the MMA: y = [x | onehot] wcat^T, then the per-head normalisation."""
    m, k = x.shape
    n = wcat.shape[0]
    y = torch.empty(m, n, dtype=BF16, device=x.device)
    inv = torch.empty(m, norm_cols // hd, dtype=F32, device=x.device)
    _call("bf16_qkv_fwd_onehot", _pb(x), _pb(onehot), _pb(wcat), _pb(y), _p(inv), m, n, k, norm_cols, hd, _stream())
    return y, inv


def bf16_linear_ln_fwd(a, w, bias, res, rowmask, gamma, beta, eps, want_v=True):
    m, k = a.shape
    n = w.shape[0]
    y = torch.empty(m, n, dtype=BF16, device=a.device)
    v = torch.empty(m, n, dtype=BF16, device=a.device) if want_v else None
    mean = torch.empty(max(1, m), dtype=F32, device=a.device)
    rstd = torch.empty(max(1, m), dtype=F32, device=a.device)
    _call("bf16_linear_ln_fwd", _pb(a), _pb(w), _p(bias, F32), _pb(res), _p(rowmask, U8), _p(gamma, F32), _p(beta, F32), float(eps), _pb(v), _pb(y),
          _p(mean), _p(rstd), m, n, k, _stream())
    return y, v, mean, rstd


def bf16_linear_bwd_data(dy, w, gelu_pre=None, dx=None, accumulate=False, pre_is_derivative=False):
    """dx = dy w [* gelu'(gelu_pre)] [+= dx]; pre_is_derivative: gelu_pre already holds GELU' (an ACT_GELU_DERIV forward wrote it)."""
    m, n = dy.shape
    k = w.shape[1]
    if dx is None:
        dx = torch.empty(m, k, dtype=BF16, device=dy.device)
        accumulate = False
    _call("bf16_linear_bwd_data", _pb(dy), _pb(w), _pb(gelu_pre), _pb(dx), m, n, k, int(accumulate) | (2 if pre_is_derivative else 0), _stream())
    return dx


def bf16_linear_bwd_weight(dy, x, onehot=None):
    """-> dw (n, k) fp32 [, dtab_t (n, 64) fp32 = dy^T onehot when onehot (m, 64) bf16 is given]."""
    m, n = dy.shape
    k = x.shape[1]
    dw = torch.empty(n, k, dtype=F32, device=dy.device)
    dt = torch.empty(n, 64, dtype=F32, device=dy.device) if onehot is not None else None
    _call("bf16_linear_bwd_weight", _pb(dy), _pb(x), _p(dw), _pb(onehot), _p(dt), m, n, k, _stream())
    return dw if onehot is None else (dw, dt)


def bf16_layernorm_bwd(dy, v, rowmask, gamma, mean, rstd, want_dres=False, want_colsum=False):
    rows, c = v.shape
    dv = torch.empty_like(v)
    dres = torch.empty_like(v) if want_dres else None
    dg = torch.empty(c, dtype=F32, device=v.device)
    db = torch.empty(c, dtype=F32, device=v.device)
    dc = torch.empty(c, dtype=F32, device=v.device) if want_colsum else None
    _call("bf16_layernorm_bwd", _pb(dy), _pb(v), _p(rowmask, U8), _p(gamma, F32), _p(mean, F32), _p(rstd, F32), _pb(dv), _pb(dres), _p(dg), _p(db),
          _p(dc), rows, c, _stream())
    return dv, dres, dg, db, dc


def bf16_colsum(x):
    out = torch.empty(x.shape[1], dtype=F32, device=x.device)
    _call("bf16_colsum", _pb(x), _p(out), x.shape[0], x.shape[1], _stream())
    return out


def bf16_binned_colsum(dy, rowidx):
    t = torch.empty(64, dy.shape[1], dtype=F32, device=dy.device)
    _call("bf16_binned_colsum", _pb(dy), _p(rowidx, U8), _p(t), dy.shape[0], dy.shape[1], _stream())
    return t


def bf16_sparse_conv_fwd(x, table, w, rows_out):
    """x (rows_in, cin) bf16, w (cout, taps, cin) or (cout, kh, kw, cin) bf16 -> (rows_out, cout) bf16."""
    cout, cin = w.shape[0], w.shape[-1]
    taps = w.numel() // (cout * cin)
    y = torch.empty(rows_out, cout, dtype=BF16, device=x.device)
    _call("bf16_sparse_conv_fwd", _pb(x), _p(table, I32), _pb(w), _pb(y), rows_out, taps, cin, cout, _stream())
    return y


def bf16_sparse_conv_bwd_weight(dy, x, table, w_shape):
    cout, taps, cin = w_shape[0], w_shape[1] * w_shape[2], w_shape[3]
    dw = torch.empty(w_shape, dtype=F32, device=x.device)
    _call("bf16_sparse_conv_bwd_weight", _pb(dy), _pb(x), _p(table, I32), _p(dw), dy.shape[0], taps, cin, cout, _stream())
    return dw


def transpose_taps_bf16(w, flip):
    cout, taps, cin = w.shape[0], w.shape[1] * w.shape[2], w.shape[3]
    wt = torch.empty(cin, taps, cout, dtype=BF16, device=w.device)
    _call("transpose_taps_bf16", _p(w, F32), wt.data_ptr(), cout, taps, cin, int(flip), _stream())
    return wt


def densify_nhwc_b16(rows, indices, batch, Y, X):
    m, c = rows.shape
    dense = torch.empty(batch, Y, X, c, dtype=BF16, device=rows.device)
    _call("densify_nhwc_b16", _pb(rows), _p(indices, I32), m, c, batch, Y, X, dense.data_ptr(), 1, _stream())
    return dense


def gather_nhwc_b16(dense, indices):
    _, Y, X, c = dense.shape
    m = indices.shape[0]
    rows = torch.empty(m, c, dtype=BF16, device=dense.device)
    _call("gather_nhwc_b16", _pb(dense), _p(indices, I32), m, c, Y, X, _pb(rows), _stream())
    return rows


class BF16Weights(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("in_w", "out_w", "w1", "w2", "in_wcat")]


class QkvOperands:
    """Per attention layer the (3C, C + 64) bf16 operand [in_proj_weight | table^T] of the packed projection (tmae_bf16_qkv_wcat): it
    depends on the weights only, so it is kept across calls and ALL stale ones of a device are rebuilt by one launch when the masters
    changed (as WeightShadows: version counters, optimizer epoch, `force`); an inference loop never rebuilds it.  Weak references."""

    def __init__(self):
        self.items = {}      # id(in_w) -> [weakref(in_w), weakref(in_b), weakref(lut), wcat, versions, pointers]
        self.tables = {}     # device -> (item ids, record table)
        self.epoch = -1
        self._keep = []

    def get(self, w, b, lut):
        it = self.items.get(id(w))
        ptrs = (w.data_ptr(), b.data_ptr(), lut.data_ptr())
        if it is None or it[0]() is not w or it[5] != ptrs:
            n, c = w.shape
            it = self.items[id(w)] = [weakref.ref(w), weakref.ref(b), weakref.ref(lut), torch.empty(n, c + 64, dtype=BF16, device=w.device), None, ptrs]
            self.tables.pop(w.device, None)
        if it[4] != (w._version, b._version) or self.epoch != _weights_epoch[0]:
            self.refresh()
        return it[3]

    def refresh(self, force=False):
        live = []
        for key, it in list(self.items.items()):
            w, b, lut = it[0](), it[1](), it[2]()
            if w is None or b is None or lut is None or it[5] != (w.data_ptr(), b.data_ptr(), lut.data_ptr()):
                del self.items[key]                         # the layer is gone or was moved: get() re-creates the entry on its next call
                self.tables.pop(it[3].device, None)
                continue
            live.append((key, it, w, b, lut))
        force = force or self.epoch != _weights_epoch[0]
        self.epoch = _weights_epoch[0]
        self._keep = []
        by_dev = {}
        for rec in live:
            by_dev.setdefault(rec[2].device, []).append(rec)
        for dev, group in by_dev.items():
            stale = group if force else [g for g in group if g[1][4] != (g[2]._version, g[3]._version)]
            if not stale:
                continue
            keys = tuple(g[0] for g in stale)
            cached = self.tables.get(dev) if len(stale) == len(group) else None
            if cached is not None and cached[0] == keys:
                table = cached[1]
            else:
                table = torch.tensor([[lut.data_ptr(), w.data_ptr(), b.data_ptr(), it[3].data_ptr(), w.shape[0], 2 * w.shape[1], w.shape[1]]
                                      for _, it, w, b, lut in stale], dtype=I64, device=dev)
                if len(stale) == len(group):
                    self.tables[dev] = (keys, table)
            with torch.cuda.device(dev):
                _call("bf16_qkv_wcat_multi", table.data_ptr(), len(stale), max(w.shape[0] for _, _, w, _, _ in stale), _stream())
            for _, it, w, b, _ in stale:
                it[4] = (w._version, b._version)
            self._keep.append(table)   # alive until the next refresh (the launch reads it asynchronously)


qkv_operands = QkvOperands()


def attention_tc_available():
    return bool(lib().cdll.tmae_bf16_attention_tc_available())


def set_attention_impl(impl):
    """bf16 layers: 1 = tcgen05 window-attention kernel, 0 = cast bridge to the fp32-I/O mma.sync kernels (checker)."""
    lib().bf16_set_attention_impl(int(impl))
    _state["attn_impl"] = int(impl)


def encoder_layer_fwd_bf16(x, x_kv, params, T, lut, tau_min, eps, heads, need_backward=True):
    """bf16-storage form of encoder_layer_fwd: x, x_kv bf16; params = the 13 fp32 master tensors."""
    L = lib()
    m_q, c = x.shape
    m_kv = x_kv.shape[0] if x_kv is not None else m_q
    ff = params[7].shape[0]
    cross = int(x_kv is not None)
    nb = L.bf16_encoder_layer_saved_bytes(m_q, m_kv, c, ff, heads, cross)
    saved = _ws(nb, x.device)
    y = torch.empty_like(x)
    P = _layer_params_cached(params)
    W = BF16Weights()
    W.in_w, W.out_w, W.w1, W.w2 = (shadows.get(params[i]).data_ptr() for i in (0, 2, 7, 9))
    W.in_wcat = qkv_operands.get(params[0], params[1], lut).data_ptr() if (c % 64 == 0 and T.onehot_q) else None
    _call("bf16_encoder_layer_fwd", _pb(x), _pb(x_kv), ctypes.byref(P), ctypes.byref(W), ctypes.byref(T), _p(lut, F32), float(tau_min), float(eps),
          m_q, m_kv, c, ff, heads, int(need_backward), _pb(y), _p(saved), nb, _stream())
    return y, saved


def encoder_layer_bwd_bf16(dy, x, x_kv, params, T, lut, tau_min, heads, saved, want_dkv):
    L = lib()
    m_q, c = x.shape
    m_kv = x_kv.shape[0] if x_kv is not None else m_q
    ff = params[7].shape[0]
    cross = int(x_kv is not None)
    sizes = [t.numel() for t in params]
    offs = [0]
    for n in sizes:
        offs.append(offs[-1] + (n + 63) // 64 * 64)
    gbuf = torch.empty(offs[-1], dtype=F32, device=x.device)
    grads = [gbuf[o:o + n].view(t.shape) for o, n, t in zip(offs, sizes, params)]
    G = LayerParams()
    base = gbuf.data_ptr()
    for (name, _), o in zip(LayerParams._fields_, offs):
        setattr(G, name, base + 4 * o)
    P = _layer_params_cached(params)
    W = BF16Weights()
    W.in_w, W.out_w, W.w1, W.w2 = (shadows.get(params[i]).data_ptr() for i in (0, 2, 7, 9))
    nb = L.bf16_encoder_layer_scratch_bytes(m_q, m_kv, c, ff, heads, cross)
    scratch = _ws(nb, x.device)
    dx = torch.empty_like(x)
    dkv = torch.empty_like(x_kv) if (cross and want_dkv) else None
    _call("bf16_encoder_layer_bwd", _pb(dy), _pb(x), _pb(x_kv), ctypes.byref(P), ctypes.byref(W), ctypes.byref(T), _p(lut, F32), float(tau_min),
          m_q, m_kv, c, ff, heads, _p(saved), saved.numel(), _pb(dx), _pb(dkv), ctypes.byref(G), gbuf.data_ptr(), gbuf.numel() * 4, _p(scratch), nb,
          _stream())
    return dx, dkv, grads


def _pvb(t):
    """pointer of a (rows, C) bf16 view with unit column stride (a column block of a packed projection is allowed)."""
    if not (t.is_cuda and t.dtype == BF16 and t.stride(1) == 1 and t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0):
        raise RuntimeError("expected a bf16 CUDA matrix with unit column stride, a row pitch that is a multiple of 8 and 16-byte alignment")
    return t.data_ptr()


def bf16_window_attention_fwd(q, k, v, qtok, qcnt, ktok, kcnt, n_win, small, mid, max_windows, tau, tau_min, heads, zero_out):
    """tcgen05 window attention: q, k unit vectors per head (bf16), -> o (bf16), lse (fp32)."""
    mq, c = q.shape
    o = (torch.zeros if zero_out else torch.empty)(mq, c, dtype=BF16, device=q.device)
    lse = torch.zeros(max(1, mq), heads, dtype=F32, device=q.device)
    _call("bf16_window_attention_fwd", _pvb(q), _pvb(k), _pvb(v), _pb(o), _p(lse), _p(qtok), _p(qcnt), _p(ktok), _p(kcnt), _p(n_win), _p(small), _p(mid),
          max_windows, _p(tau, F32), float(tau_min), c, heads, q.stride(0), k.stride(0), v.stride(0), mq, k.shape[0], _stream())
    return o, lse


def bf16_window_attention_bwd(dout, q, k, v, o, lse, inv_q, inv_k, qtok, qcnt, ktok, kcnt, n_win, small, mid, max_windows, tau, tau_min, heads, dtau, zero):
    """-> dq, dk, dv (bf16, same packing as q / k / v): gradients wrt the UN-normalised projections (see tmae_bf16_window_attention_bwd)."""
    alloc = torch.zeros if zero else torch.empty
    C = q.shape[1]
    packed = q.stride(0) == 3 * C and k.data_ptr() == q.data_ptr() + 2 * C
    if packed:
        buf = alloc(q.shape[0], 3 * C, dtype=BF16, device=q.device)
        dq, dk, dv = buf[:, :C], buf[:, C:2 * C], buf[:, 2 * C:]
    elif k.stride(0) == 2 * C and v.data_ptr() == k.data_ptr() + 2 * C:
        dq = alloc(q.shape[0], C, dtype=BF16, device=q.device)
        buf = alloc(k.shape[0], 2 * C, dtype=BF16, device=q.device)
        dk, dv = buf[:, :C], buf[:, C:]
    else:
        dq, dk, dv = (alloc(t.shape[0], C, dtype=BF16, device=t.device) for t in (q, k, v))
    _call("bf16_window_attention_bwd", _pb(dout), _pvb(q), _pvb(k), _pvb(v), _pb(o), _p(lse, F32), inv_q.data_ptr(), inv_q.stride(0), inv_k.data_ptr(), inv_k.stride(0),
          _pvb(dq), _pvb(dk), _pvb(dv), _p(dtau, F32), _p(qtok), _p(qcnt), _p(ktok), _p(kcnt), _p(n_win), _p(small), _p(mid), max_windows, _p(tau, F32),
          float(tau_min), C, heads, q.stride(0), k.stride(0), v.stride(0), q.shape[0], k.shape[0], _stream())
    return dq, dk, dv
