"""Micro-benchmarks of single library kernels at the bench's shapes (CUDA-event timing, L2 flushed between
iterations).  Used to pick the ncu target and to compare kernel variants:  python tools/kernel_bench.py [gemm|attn|all]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa: E402,F401
from tmae_b200 import ops  # noqa: E402

DEV = "cuda"
FLUSH = None


def timeit(fn, iters=10, flush=True):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        if flush:
            FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def timeit_queue(fns, reps=3):
    """GPU-only time per call: the GPU is parked on a spin kernel while the host enqueues every call, so host launch
    overhead is not in the interval; `fns` should cycle over distinct buffers whose total size exceeds L2."""
    for f in fns:
        f()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        torch.cuda._sleep(40_000_000)  # ~20 ms
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for f in fns:
            f()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / len(fns))
    return best


def gemm(prec):
    ops.set_precision(prec)
    for key, val in os.environ.items():   # TMAE_OPT_<name>=<int> -> tmae_set_option (A/B runs)
        if key.startswith("TMAE_OPT_"):
            ops.set_option(key[9:].lower(), int(val))
    print(f"--- linear ({prec}) : m n k | fwd us TFLOP/s GB/s | bwd_data us GB/s | bwd_weight us GB/s")
    for m, n, k in [(56000, 128, 128), (56000, 384, 128), (56000, 256, 128), (56000, 128, 256), (50000, 256, 256), (50000, 768, 256), (50000, 512, 256),
                    (50000, 256, 512), (14000, 128, 128), (14000, 256, 256), (240000, 128, 64)]:
        w, b = torch.randn(n, k, device=DEV), torch.randn(n, device=DEV)
        sets = [(torch.randn(m, k, device=DEV), torch.randn(m, n, device=DEV)) for _ in range(4)]
        dw, db = torch.empty_like(w), torch.empty_like(b)
        t = timeit_queue([lambda s=s: ops.linear_fwd(s[0], w, b) for s in sets] * 2)
        t2 = timeit_queue([lambda s=s: ops.linear_bwd_data(s[1], w) for s in sets] * 2)
        t3 = timeit_queue([lambda s=s: ops.linear_bwd_weight(s[1], s[0], dw, db) for s in sets] * 2)
        fl, by = 2 * m * n * k, 4 * (m * k + m * n + n * k)
        print(f"{m:7d} {n:4d} {k:4d} | {t * 1e3:8.1f} {fl / t / 1e9:7.1f} {by / t / 1e6:7.0f} | {t2 * 1e3:8.1f} {by / t2 / 1e6:7.0f} | {t3 * 1e3:8.1f} {by / t3 / 1e6:7.0f}")


def attn(tc=1):
    import numpy as np
    ops.set_option("attn_tc", tc)
    for key, val in os.environ.items():   # TMAE_OPT_<name>=<int> -> tmae_set_option (A/B runs)
        if key.startswith("TMAE_OPT_"):
            ops.set_option(key[9:].lower(), int(val))
    print(f"--- window attention (attn_tc={tc}) : M C | fwd us (GB/s of 4*M*C*4) | bwd us (GB/s of 8*M*C*4)")
    for M, C, g, B in [(56000, 128, 468, 4), (14000, 128, 468, 4), (50000, 256, 234, 4), (28000, 256, 117, 4)]:
        rng = np.random.default_rng(0)
        # clustered occupancy like a lidar BEV: sample cells with a radial density
        cells = np.unique(np.clip((rng.normal(0, g / 5, (M * 3, 2)) + g / 2).astype(np.int64), 0, g - 1) @ np.array([g, 1]))
        per = min(M // B, cells.shape[0])
        cells = np.sort(rng.choice(cells, per, replace=False))
        c = np.concatenate([np.stack([np.full(per, b), cells // g, cells % g], 1) for b in range(B)])
        coords = torch.tensor(c, dtype=torch.int32, device=DEV)
        P = ops.window_partition(coords, B, g, g, [(16, 0, 16), (32, 16, 32), (64, 32, 100000)])
        m = coords.shape[0]
        tau = torch.ones(1, device=DEV)
        H = 8
        args = (P.tok_a[0], P.cnt_a[0], P.tok_a[0], P.cnt_a[0], P.n_win[0:1], ops.small_end(P, 0), ops.mid_end(P, 0), min(P.wcap, m), tau, 0.01, H)
        sets = []
        for _ in range(4):
            q, k, v = (torch.randn(m, C, device=DEV) for _ in range(3))
            o, lse = ops.window_attention_fwd(q, k, v, *args, False)
            sets.append((q, k, v, o, lse, torch.randn_like(o)))
        dtau = torch.zeros(1, device=DEV)
        t = timeit_queue([lambda s=s: ops.window_attention_fwd(s[0], s[1], s[2], *args, False) for s in sets] * 2)
        t2 = timeit_queue([lambda s=s: ops.window_attention_bwd(s[5], s[0], s[1], s[2], s[3], s[4], *args, dtau, False) for s in sets] * 2)
        nw = int(P.n_win[0])
        lb = P.level_base[0].tolist()
        cnt = P.cnt_a[0][:nw].long()
        pairs = int((cnt * cnt).sum())
        print(f"{m:7d} {C:4d} windows {nw} levels {lb} pairs {pairs} | {t * 1e3:8.1f} ({4 * 4 * m * C / t / 1e6:.0f} GB/s) | {t2 * 1e3:8.1f} ({8 * 4 * m * C / t2 / 1e6:.0f} GB/s)")
        tab = ops.lib_profile(lambda: [ops.window_attention_fwd(s[0], s[1], s[2], *args, False) for s in sets] +
                              [ops.window_attention_bwd(s[5], s[0], s[1], s[2], s[3], s[4], *args, dtau, False) for s in sets])
        print("     " + "  ".join(f"{k} {v['ms'] / v['calls'] * 1e3:.1f}us" for k, v in tab.items()))
    ops.set_option("attn_tc", 0)


def _lidar_partition(M, g, B, seed=0):
    import numpy as np
    rng = np.random.default_rng(seed)
    cells = np.unique(np.clip((rng.normal(0, g / 5, (M * 3, 2)) + g / 2).astype(np.int64), 0, g - 1) @ np.array([g, 1]))
    per = min(M // B, cells.shape[0])
    cells = np.sort(rng.choice(cells, per, replace=False))
    c = np.concatenate([np.stack([np.full(per, b), cells // g, cells % g], 1) for b in range(B)])
    coords = torch.tensor(c, dtype=torch.int32, device=DEV)
    return coords, ops.window_partition(coords, B, g, g, [(16, 0, 16), (32, 16, 32), (64, 32, 100000)])


def attn_bf16(shapes=None):
    """tcgen05 window attention (bf16 storage) at the bench's stage shapes: forward / backward us and GB/s of the algorithmic bytes."""
    BF = torch.bfloat16
    print("--- tcgen05 window attention (bf16) : M C | fwd us (GB/s of 4*M*C*2) | bwd us (GB/s of 7*M*C*2)")
    for M, C, g, B in shapes or [(68000, 128, 468, 8), (14000, 128, 468, 4), (64000, 256, 234, 8), (35000, 256, 117, 8)]:
        coords, P = _lidar_partition(M, g, B)
        m = coords.shape[0]
        tau = torch.ones(1, device=DEV)
        H = 8
        args = (P.tok_a[0], P.cnt_a[0], P.tok_a[0], P.cnt_a[0], P.n_win[0:1], ops.small_end(P, 0), ops.mid_end(P, 0), min(P.wcap, m), tau, 0.01, H)
        sets = []
        for _ in range(4):
            qkv = torch.randn(m, 3 * C, device=DEV).to(BF)
            q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
            o, lse = ops.bf16_window_attention_fwd(q, k, v, *args, False)
            inv = torch.ones(m, 2 * H, device=DEV)
            sets.append((q, k, v, o, lse, torch.randn(m, C, device=DEV).to(BF), inv))
        dtau = torch.zeros(1, device=DEV)
        t = timeit_queue([lambda s=s: ops.bf16_window_attention_fwd(s[0], s[1], s[2], *args, False) for s in sets] * 2)
        t2 = timeit_queue([lambda s=s: ops.bf16_window_attention_bwd(s[5], s[0], s[1], s[2], s[3], s[4], s[6][:, :H], s[6][:, H:], *args, dtau, False) for s in sets] * 2)
        nw = int(P.n_win[0])
        lb = P.level_base[0].tolist()
        print(f"{m:7d} {C:4d} windows {nw} levels {lb} | {t * 1e3:8.1f} ({4 * 2 * m * C / t / 1e6:.0f} GB/s) | {t2 * 1e3:8.1f} ({7 * 2 * m * C / t2 / 1e6:.0f} GB/s)")


def gemm_bf16():
    BF = torch.bfloat16
    print("--- bf16 GEMMs : m n k | fwd us GB/s | ln-fused us GB/s | bwd_data us GB/s | bwd_weight us GB/s")
    for m, n, k in [(68000, 128, 128), (68000, 384, 128), (68000, 256, 128), (68000, 128, 256), (64000, 256, 256), (64000, 768, 256), (64000, 512, 256),
                    (64000, 256, 512), (35000, 256, 256), (14000, 128, 128)]:
        w, b = torch.randn(n, k, device=DEV).to(BF), torch.randn(n, device=DEV)
        sets = [(torch.randn(m, k, device=DEV).to(BF), torch.randn(m, n, device=DEV).to(BF)) for _ in range(4)]
        t = timeit_queue([lambda s=s: ops.bf16_linear_fwd(s[0], w, b) for s in sets] * 2)
        t2 = timeit_queue([lambda s=s: ops.bf16_linear_bwd_data(s[1], w) for s in sets] * 2)
        t3 = timeit_queue([lambda s=s: ops.bf16_linear_bwd_weight(s[1], s[0]) for s in sets] * 2)
        by = 2 * (m * k + m * n + n * k)
        ln = ""
        if n in (128, 256):
            g_, be = torch.ones(n, device=DEV), torch.zeros(n, device=DEV)
            t4 = timeit_queue([lambda s=s: ops.bf16_linear_ln_fwd(s[0], w, b, s[1], None, g_, be, 1e-5) for s in sets] * 2)
            ln = f"{t4 * 1e3:8.1f} {(by + 2 * 2 * m * n) / t4 / 1e6:7.0f}"
        print(f"{m:7d} {n:4d} {k:4d} | {t * 1e3:8.1f} {by / t / 1e6:7.0f} | {ln:>16s} | {t2 * 1e3:8.1f} {by / t2 / 1e6:7.0f} | {t3 * 1e3:8.1f} {(by + 2 * n * k) / t3 / 1e6:7.0f}")


def bn():
    """BatchNorm1d + ReLU over sparse-conv rows (fp32) and over a decoder map (bf16): forward (train) and backward."""
    for key, val in os.environ.items():
        if key.startswith("TMAE_OPT_"):
            ops.set_option(key[9:].lower(), int(val))
    for rows, C in [(70000, 128), (60000, 256), (240000, 64)]:
        g, b = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
        rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
        sets = [(torch.randn(rows, C, device=DEV), torch.randn(rows, C, device=DEV)) for _ in range(4)]
        t = timeit_queue([lambda s=s: ops.bn_train_fwd(s[0], g, b, rm, rv, 0.01, 1e-3, True) for s in sets] * 2)
        _, mean, rstd = ops.bn_train_fwd(sets[0][0], g, b, rm, rv, 0.01, 1e-3, True)
        t2 = timeit_queue([lambda s=s: ops.bn_bwd(s[1], s[0], b, mean, rstd, g, True, True) for s in sets] * 2)
        by = rows * C * 4
        print(f"bn fp32 {rows:7d} x {C:3d} | fwd {t * 1e3:7.1f} us ({3 * by / t / 1e6:.0f} GB/s) | bwd {t2 * 1e3:7.1f} us ({5 * by / t2 / 1e6:.0f} GB/s)")


def bev(batch=8):
    """SSTBEVBackbone forward (eval) and forward+backward (train) on a finetune-sized bf16 channels-last map."""
    from tmae_b200 import config
    m = tmae_b200.SSTBEVBackbone(config.model_cfg("finetune")["BACKBONE_2D"]).to(DEV)
    torch.backends.cudnn.benchmark = True
    xs = [(torch.randn(batch, 128, 468, 468, device=DEV).clamp_min(0) * (torch.rand(batch, 1, 468, 468, device=DEV) < 0.3)).to(torch.bfloat16)
          .contiguous(memory_format=torch.channels_last) for _ in range(2)]
    nbytes = xs[0].numel() * 2
    for train in (False, True):
        m.train(train)

        def run(x):
            if not train:
                with torch.no_grad():
                    return m(dict(spatial_features=x))["spatial_features_2d"]
            x = x.detach().requires_grad_()
            y = m(dict(spatial_features=x))["spatial_features_2d"]
            y.backward(y.detach())
            return y
        t = timeit_queue([lambda x=x: run(x) for x in xs] * 2)
        tab = ops.lib_profile(lambda: [run(x) for x in xs])
        ours = sum(v["ms"] for v in tab.values()) / len(xs)
        print(f"SSTBEVBackbone batch {batch} 468x468x128 bf16 {'fwd+bwd (train)' if train else 'fwd (eval)'}: {t:.2f} ms per call "
              f"({batch / t * 1e3:.0f} scans/s); library BatchNorm kernels {ours:.2f} ms of it; one map = {nbytes / 1e6:.0f} MB")
        print("     " + "  ".join(f"{k} {v['ms'] / v['calls'] * 1e3:.0f}us x{v['calls'] // len(xs)} ({v['bytes'] / v['ms'] / 1e6:.0f} GB/s)" for k, v in tab.items() if v["ms"] > 0))


def small():
    """The kernels in front of and behind the encoder (voxelise, window partition, Chamfer) at the bench's size -- where a frame set is
    5-10 MB and the multi-kernel calls are launch-latency-bound -- and at sizes where the bytes dominate: GB/s of the ALGORITHMIC bytes
     from the library's own CUDA-event profiler, against the measured 6.45 TB/s copy peak."""
    import numpy as np
    from tmae_b200 import synth
    shape = synth.SHAPES["once"]
    grid = synth.grid_size(shape).tolist()
    print("voxelize (points x 5 floats -> kept points, coords, inverse, CSR, voxel mean):")
    for B, npts in ((4, 60000), (16, 120000), (64, 120000), (64, 500000)):
        rng = np.random.default_rng(B)
        n = B * npts
        pts = np.empty((n, 5), np.float32)
        pts[:, 0] = np.repeat(np.arange(B), npts)
        r = np.abs(rng.normal(0, 30, n)) + 2
        th = rng.uniform(0, 2 * np.pi, n)
        pts[:, 1], pts[:, 2] = r * np.cos(th), r * np.sin(th)
        pts[:, 3], pts[:, 4] = rng.normal(-1, 0.5, n), rng.uniform(0, 1, n)
        p = torch.from_numpy(pts).to(DEV)
        ops.voxelize(p, shape["range"], shape["voxel"], grid, B)
        torch.cuda.synchronize()
        tab = ops.lib_profile(lambda: [ops.voxelize(p, shape["range"], shape["voxel"], grid, B) for _ in range(5)])
        v = tab["voxelize"]
        print(f"   batch {B:3d} x {npts:7d} points ({n * 20 / 1e6:6.1f} MB in): {v['ms'] / v['calls'] * 1e3:8.1f} us  {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s")
    print("window_partition (both shifts, 3 levels):")
    for M, g, B in ((70000, 468, 8), (560000, 468, 64), (2000000, 936, 64)):
        coords, P = _lidar_partition(M, g, B)
        lv = [(16, 0, 16), (32, 16, 32), (64, 32, 100000)]
        torch.cuda.synchronize()
        tab = ops.lib_profile(lambda: [ops.window_partition(coords, B, g, g, lv) for _ in range(5)])
        v = tab["window_partition"]
        print(f"   {coords.shape[0]:8d} voxels, grid {g}, batch {B:3d}: {v['ms'] / v['calls'] * 1e3:8.1f} us  {v['bytes'] / v['ms'] / 1e6:7.0f} GB/s")
    print("chamfer (16 predictions vs <= 64 ground-truth points per pillar, dense (M, 64, 3) ground truth):")
    for M in (55000, 500000, 4000000):
        g = torch.Generator(device=DEV).manual_seed(M)
        pred = torch.randn(M, 16, 3, device=DEV, generator=g)
        gt = torch.randn(M, 64, 3, device=DEV, generator=g)
        gt[:, 40:] = float("nan") if False else gt[:, 40:]
        w = (torch.rand(M, device=DEV, generator=g) < 0.75).float()
        ops.chamfer_fwd(pred, gt, w)
        torch.cuda.synchronize()
        tab = ops.lib_profile(lambda: [ops.chamfer_fwd(pred, gt, w) for _ in range(5)])
        v = tab["chamfer_fwd"]
        nb = M * (16 * 12 + 4) + 0.75 * M * 64 * 12
        print(f"   {M:8d} pillars: {v['ms'] / v['calls'] * 1e3:8.1f} us  {nb / (v['ms'] / v['calls']) / 1e6:7.0f} GB/s (bytes actually needed: masked pillars skip their ground truth)")


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "small":
        small()
    if what in ("attn_bf16", "bf16"):
        attn_bf16()
    if what == "attn_bf16_one":
        attn_bf16([(68000, 128, 468, 8)])
    if what in ("gemm_bf16", "bf16"):
        gemm_bf16()
    if what in ("gemm", "all"):
        gemm("bf16")
    if what in ("gemm32",):
        gemm("fp32")
    if what in ("attn", "all"):
        attn()
    if what == "bev":
        bev()
    if what == "bn":
        bn()
