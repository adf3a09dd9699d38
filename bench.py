#!/usr/bin/env python
"""Benchmark of the T-MAE sparse-window voxel-encoder hot path on B200 (driver contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pretrain|finetune|waymo] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic ONCE-shaped scan pairs:
  pretrain (default, BASELINE.json configs[1]): TemporalDynVFE -> SiamWCA_MAE.forward -> Chamfer loss -> backward
            -> AdamW step, batch 4 scan pairs per GPU, 75 % voxel mask;
  finetune (configs[3]): TemporalDynVFE -> SiamWCA.forward, batch 8, 120k-point scans, no gradient;
  waymo    (configs[4]): the pretraining step on Waymo-shaped scans (180k points, 5 point features, 472x472 grid), batch 4.
The default run also carries short `extra_workloads` lines for finetune and waymo, the same-GPU stock-PyTorch comparator
(`gpu_torch_baseline`) and the host-core baseline (`cpu_baseline`).
Prints ONE JSON line.  `value` = scan pairs per second with the inputs resident in HBM; `e2e` = the same through
the public module API from pinned HOST buffers (H2D copies and the loss read-back inside the timed region).
`--impl reference` times the reference's algorithm on the host cores (the tier-2 oracle port; /root/reference does
not exist on the GPU box) on a bounded sample of the same workload.
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

if "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is ONE process that should use every host thread it can
    # (measured here: 34 s instead of 12.5 s per step with the variable left at 1, torch.set_num_threads notwithstanding)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        if os.environ.get(_v) == "1":
            del os.environ[_v]

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "pretrain": dict(kind="pretrain", shape="once", npf=5, batch=4, n_points=60000, train=True,
                     name="T-MAE pretraining forward+backward+AdamW, synthetic ONCE scan pairs (60k pts/frame, 0.32 m pillars, "
                          "468x468 grid), 75% voxel mask, Chamfer loss, batch 4 scan pairs per GPU"),
    "finetune": dict(kind="finetune", shape="once", npf=5, batch=8, n_points=120000, train=False,
                     name="finetune-mode encoder forward (no masking, temporal cross-attention), synthetic ONCE scans "
                          "(120k pts/frame, 0.32 m pillars), batch 8 scan pairs per GPU"),
    "waymo": dict(kind="pretrain", shape="waymo", npf=6, batch=4, n_points=180000, train=True,
                  name="T-MAE pretraining forward+backward+AdamW, synthetic Waymo-shaped scan pairs (180k pts/frame, 64 beams, 5 point features, "
                       "0.32 m pillars, 472x472 grid), 75% voxel mask, Chamfer loss, batch 4 scan pairs per GPU"),
}


def dist_env():
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    return rank, world, int(os.environ.get("LOCAL_RANK", 0))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured (MEASURED_PEAKS.json, sustained bf16)")
    return dict(hbm=6650.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region through NVML, inline at step boundaries (`sample`; a looping
    nvidia-smi process was measured to slow the timed region by ~15 %); falls back to a nvidia-smi child when NVML is absent."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period=float(os.environ.get("TMAE_CLOCK_PERIOD", "0.25"))):
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons, self.stop, self.t, self.proc, self.lines = [], 0, set(), False, None, None, []
        self.nv = self.h = None

    def sample(self):
        """One NVML sample, taken INLINE by the timing loop at a few step boundaries (two light queries, tens of microseconds).  A
        sampling thread costs more than it looks: every wake-up has to take the GIL from the thread that enqueues the step (5 ms
        switch interval), seen as 5-7 ms steps once or twice per timed region and, in lockstep, on every rank."""
        if self.nv is None:
            return
        nv, h = self.nv, self.h
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        try:
            self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            for name, bit in bits.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.nv, self.h = nv, h
            for _ in range(3):      # the FIRST call of each query can take ~75 ms (measured: one 94 ms step at the first in-loop sample)
                self.sample()
            self.sm, self.reasons = [], set()   # warm-up samples are not part of the record (the GPU is idle here)
            return self
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "500",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
        if self.t:
            self.t.join(timeout=2)

    def summary(self):
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            self.sm.append(int(f[0]))
            self.mx = max(self.mx, int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx or None, "reasons": sorted(self.reasons),
                "samples": len(self.sm)}


def shard_seeds(w, n_batches, rank):
    """Scan-pair sharding: rank r owns scan indices (seeds) [r*n_batches*B, (r+1)*n_batches*B) -- disjoint across ranks."""
    B = w["batch"]
    return [[1000 + (rank * n_batches + i) * B + j for j in range(B)] for i in range(n_batches)]


def aggregate_value(batch, steps, world, total_ms):
    """Whole-job scan pairs per second: all ranks' units over the slowest rank's time."""
    return batch * steps * world / (total_ms / 1e3)


def scan_cost(pair, shape):
    """Cost proxy of a scan pair: occupied pillars of both frames (the encoder's row count)."""
    lo, vs = np.asarray(shape["range"][:2]), np.asarray(shape["voxel"][:2])
    n = 0
    for pts in pair:
        c = np.floor((pts[:, :2] - lo) / vs).astype(np.int64)
        n += np.unique(c[:, 0] * 100000 + c[:, 1]).shape[0]
    return n


def make_batches(w, n_batches, rank, balanced=True):
    """The rank's own scan pairs (shard_seeds: disjoint across ranks) as `n_batches` collated batches in pinned host memory.
    balanced: the rank forms its batches from its scans SORTED by cost (occupied pillars), lightest batch first.  Data-parallel training
    steps run in lockstep (the gradient exchange), so a step costs the slowest rank's batch; with every rank's i-th batch made of its
    i-th cost quantile the ranks' steps line up (a length-grouped sampler, rank-local: no scan changes owner, nothing is exchanged).  The
    mean cost per rank is unchanged -- the single-GPU number does not move."""
    from tmae_b200 import synth
    shape = synth.SHAPES[w.get("shape", "once")]
    seeds = [sd for b in shard_seeds(w, n_batches, rank) for sd in b]
    pairs = [synth.scan_pair(sd, w["n_points"], w.get("shape", "once")) for sd in seeds]
    if balanced:
        order = np.argsort([scan_cost(p, shape) for p in pairs], kind="stable")
        pairs = [pairs[i] for i in order]
    out = []
    B = w["batch"]
    for i in range(n_batches):
        grp = pairs[i * B:(i + 1) * B]
        pts, ptsp = synth.collate([p[0] for p in grp]), synth.collate([p[1] for p in grp])
        a, b = torch.from_numpy(pts), torch.from_numpy(ptsp)
        out.append((a.pin_memory(), b.pin_memory()) if torch.cuda.is_available() else (a, b))   # (no GPU: the host-logic tests)
    return out


# ------------------------------------------------------------------------------------------- our arm
class Ctx:
    """Per-process state shared by the workloads of one bench.py run."""
    pass


def setup(args):
    import tmae_b200  # noqa: F401
    from tmae_b200 import ops
    c = Ctx()
    c.rank, c.world, c.local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(c.local)
    c.dev = torch.device("cuda", c.local)
    if c.world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=c.dev, timeout=datetime.timedelta(seconds=180))
    # One slab for the caching allocator: the step's tensor sizes change with every batch (row counts), and blocks recorded
    # on two streams return to the pool late, so a pool made of many cudaMalloc'd segments sized to past requests keeps
    # hitting a size nothing fits -- a cudaMalloc in the middle of a step synchronises the device (one 70-100 ms step in
    # about a quarter of the runs).  A single large cached segment is split and re-merged on demand instead.
    c.slab_gib = min(64, int(torch.cuda.mem_get_info(c.dev)[0] * 0.45) >> 30)
    if c.slab_gib > 0:
        del_me = torch.empty(c.slab_gib << 30, dtype=torch.uint8, device=c.dev)
        del del_me
    # ... and one for the side stream's pool (the caching allocator keeps a pool per stream: the coordinate-only pre-pass allocates its
    # tables there, with sizes that change every batch -- measured: one 40-90 ms cudaMalloc step in ~1 of 8 without it)
    if args.side_stream and c.slab_gib > 8:
        with torch.cuda.stream(ops.side_stream(c.dev)):
            del_me = torch.empty(4 << 30, dtype=torch.uint8, device=c.dev)
            del del_me
    args.precision = args.precision or ops.BENCH_PRECISION
    ops.set_precision(args.precision)
    torch.backends.cudnn.benchmark = True
    return c


def measure(w, args, c, steps, n_batches, full):
    """One workload: build the modules, warm up, time `steps` steps from HBM-resident inputs and again from pinned host
    buffers.  full: also the per-kernel event table / roofline and the workload's row counts (rank 0)."""
    import tmae_b200
    from tmae_b200 import dist as tdist, ops, synth
    rank, world, dev = c.rank, c.world, c.dev
    shape = synth.SHAPES[w["shape"]]
    grid = synth.grid_size(shape).tolist()
    torch.manual_seed(0)
    vfe, bb = tmae_b200.build_model(w["kind"], grid, shape["voxel"], shape["range"], num_point_features=w["npf"])
    bb.decoder_autocast = torch.bfloat16 if args.decoder == "bf16" else None

    class Step(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.vfe, self.backbone_3d = vfe, bb

        def forward(self, pts, ptsp, side=None):
            bd = dict(points=pts, points_prev=ptsp, batch_size=w["batch"])
            if side is not None:
                bd["side_stream"] = side  # the inputs were produced on this stream (input pipeline): see ops.side_stream
            bd = self.vfe(bd)
            bd = self.backbone_3d(bd)
            if w["train"]:
                return self.backbone_3d.get_loss()[0]
            return bd["spatial_features"]

    model = Step().to(dev)
    model.train(w["train"])
    net = model
    opt = None
    if w["train"]:
        if world > 1 and args.ddp:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[c.local], find_unused_parameters=False)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    params = [p for p in model.parameters() if p.requires_grad]
    # gradient exchange: persistent flat buffer, buckets all-reduced on a communication stream (tmae_b200/dist.py)
    flat = None
    bucket_opts = None
    if w["train"] and world > 1 and not args.ddp:
        flat = tdist.OverlappedGradients(params, world).attach() if args.grad_sync == "overlap" else tdist.FlatGradients(params, world, n_buckets=args.grad_buckets)
        if args.grad_sync == "flat" and not args.single_optimizer:
            # one fused AdamW per bucket (same hyper-parameters: the update of every parameter is what the single optimizer computes):
            # bucket b is updated while bucket b+1 is still on the wire
            bucket_opts = [torch.optim.AdamW(ps, lr=1e-4, weight_decay=0.01, fused=True) for ps in flat.bucket_params()]
    if world > 1:  # same initial weights on every rank (DDP broadcasts them; the flat all-reduce path does it here)
        for p in list(model.parameters()) + list(model.buffers()):
            torch.distributed.broadcast(p.data, 0)
    bb.mask_generator = torch.Generator(device=dev).manual_seed(2000 + rank)

    host = make_batches(w, n_batches, rank + int(os.environ.get("TMAE_BENCH_RANK_OFFSET", "0")), balanced=not args.unbalanced)   # debugging aid: another rank's scans on one GPU
    resident = [(a.to(dev), b.to(dev)) for a, b in host]
    h2d = sum(t.numel() * 4 for t in host[0])
    side = ops.side_stream(dev) if args.side_stream else None

    def step(pts, ptsp, module=None):
        m = net if module is None else module
        if flat is not None and args.grad_sync == "overlap":
            flat.enabled = module is None   # the rank-local profiling pass must not enter collectives
        if w["train"]:
            dbg = os.environ.get("TMAE_SYNC_DEBUG")

            def chk(tag):
                if dbg:
                    torch.cuda.synchronize()
                    print(f"[dbg rank {rank}] {tag} ok", file=sys.stderr, flush=True)
            loss = m(pts, ptsp, side)
            chk("forward")
            loss.backward()
            chk("backward")
            if flat is not None and module is None and bucket_opts is not None:
                flat.reduce(step_fns=[o.step for o in bucket_opts])   # exchange and per-bucket updates pipelined
                chk("finish+step")
                for o in bucket_opts:
                    o.zero_grad(set_to_none=True)
                return loss
            if flat is not None and module is None:
                # overlap: buckets whose gradients were complete have been in flight since, wait for the communication stream;
                # flat: copy into the persistent flat buffer and all-reduce its buckets now
                (flat.finish if args.grad_sync == "overlap" else flat.reduce)()
                chk("finish")
            opt.step()
            opt.zero_grad(set_to_none=True)
            chk("step")
            return loss
        with torch.no_grad():
            return m(pts, ptsp, side)

    def step_local(pts, ptsp):
        """Same step on the un-wrapped module: no collective, so rank 0 may run it alone (profiling pass)."""
        return step(pts, ptsp, model)

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    host_ms = []
    host_out = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(max(steps, 1))]

    def timed(from_host, k, sampler=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        d2h = 0
        gc.collect()
        gc.disable()   # a generation-2 collection of the Python heap is a 50-250 ms host pause (seen as one slow step in ~half the runs)
        barrier()
        c0 = ops.launch_count()
        host_t0 = time.perf_counter()
        ev[0].record()
        marks = {k // 4, k // 2, (3 * k) // 4} if sampler is not None else ()
        for i in range(k):
            if i in marks:
                sampler.sample()
            if i >= 2:
                ev[i - 1].synchronize()  # bounded run-ahead: the host stays at most two steps in front of the GPU, so blocks
                                         # recorded on two streams return to the allocator pool before it has to grow
            if from_host:
                a, b = host[i % len(host)]
                if side is not None:  # input pipeline: the H2D copies of this step's scans are issued on the side stream
                    with torch.cuda.stream(side):
                        a, b = a.to(dev, non_blocking=True), b.to(dev, non_blocking=True)
                    a.record_stream(torch.cuda.current_stream()), b.record_stream(torch.cuda.current_stream())
                    out = step(a, b)
                else:
                    out = step(a.to(dev, non_blocking=True), b.to(dev, non_blocking=True))
                # device->host read of this step's result (loss / mean feature) into pinned memory, enqueued in-stream every
                # step and complete before the closing synchronize of the timed region; the host does not stall on it
                # (a training loop logs the loss the same way)
                host_out[i % len(host_out)].copy_((out if w["train"] else out.float().mean()).detach().reshape(1), non_blocking=True)
                d2h = 4
            else:
                out = step(*resident[i % len(resident)])
            ev[i + 1].record()
        host_ms.append((time.perf_counter() - host_t0) * 1e3 / k)  # host time to ENQUEUE the steps (includes the bounded run-ahead waits)
        barrier()
        gc.enable()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(k)]
        ms = torch.cuda.memory_stats(dev)
        print(f"[bench] {w['kind']}/{w['shape']} {'e2e' if from_host else 'resident'} per-step ms: " + " ".join(f"{p:.1f}" for p in per) +
              f" | cudaMalloc calls so far {ms.get('num_device_alloc', 0)}, frees {ms.get('num_device_free', 0)}, retries {ms.get('num_alloc_retries', 0)}, "
              f"reserved {ms.get('reserved_bytes.all.current', 0) / 2**30:.1f} GiB", file=sys.stderr)
        total = ev[0].elapsed_time(ev[-1])
        return total, per, ops.launch_count() - c0, d2h

    # every distinct batch several times: its row counts are new sizes for the caching allocator, and with the host running a
    # step ahead of the GPU (no sync in the loop) blocks recorded on two streams return to the pool late, so the pool keeps
    # growing (cudaMalloc stalls of 50-250 ms) for ~14 steps before it is stationary
    # (and a one-off 100-300 ms driver-side stall was observed at the ~23rd step of a process in a third of the runs,
    # with or without the clock sampler: the warm-up runs past it)
    warmup = args.warmup if args.profile_run else max(args.warmup, 28, 2 * len(resident))
    for i in range(warmup):
        step(*resident[i % len(resident)])
    clocks = None
    if full and args.clock_sampler:
        with ClockSampler(c.local) as cs:
            total, per, launches, _ = timed(False, steps, cs)
        clocks = cs.summary()
    else:
        total, per, launches, _ = timed(False, steps)
    timed(True, len(host))   # untimed pass over the host-input path: every distinct batch once (new allocator sizes)
    e_total, e_per, _, d2h = timed(True, steps)
    assert all(torch.isfinite(t).all() for t in host_out), "non-finite step result"

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    def minr(x):
        return -maxr(-x)

    rank_ms = total / steps
    spread = (minr(rank_ms), maxr(rank_ms))
    total, e_total = maxr(total), maxr(e_total)
    out = dict(value=aggregate_value(w["batch"], steps, world, total), e_value=aggregate_value(w["batch"], steps, world, e_total),
               ms_per_step=total / steps, e_ms_per_step=e_total / steps, per=per, launches=launches, host_ms=host_ms[0], h2d=h2d, d2h=d2h,
               warmup=warmup, clocks=clocks, rank_ms_min_max=spread, roof=None, table=None, step_roof=None, counts=None, attention_block=None,
               dispatch=ops.dispatch_counts())
    if full and rank == 0:
        out["roof"], out["table"], out["step_roof"] = roofline(ops, step_local, resident, args, total / steps)
        try:
            out["counts"] = workload_counts(bb, resident)
        except Exception as e:  # diagnostics only: never lose the measurement over them
            out["counts"] = {"error": repr(e)}
        try:
            out["attention_block"] = attention_block(bb, out["table"], w["train"], peaks()["tensor"])
        except Exception as e:
            out["attention_block"] = {"error": repr(e)}
    if flat is not None and args.grad_sync == "overlap":
        flat.detach()
    del model, net, opt, params, flat, resident, host, vfe, bb
    gc.collect()
    return out


def run_ours(args):
    c = setup(args)
    rank, world = c.rank, c.world
    w = WORKLOADS[args.workload]
    m = measure(w, args, c, args.steps, args.batches, True)
    extras = {}
    if args.extra and args.workload == "pretrain":
        # the other single-GPU configurations of BASELINE.json, a few steps each, in the same process: configs[3] (finetune
        # forward, 120k-point scans, batch 8) and configs[4] (Waymo-shaped pretraining step)
        for name in ("finetune", "waymo"):
            try:
                e = measure(WORKLOADS[name], args, c, max(3, min(args.steps, 6)), 2, False)
                per = sorted(e["per"])
                extras[name] = {"workload": WORKLOADS[name]["name"], "value": round(e["value"], 3), "unit": "scans/s", "ms_per_step": round(e["ms_per_step"], 3),
                                "e2e_value": round(e["e_value"], 3), "steps": len(per), "p50_ms_per_scan": round(per[len(per) // 2] / WORKLOADS[name]["batch"], 3),
                                "gpu_launches": e["launches"]}
            except Exception as ex:   # an extra line never costs the headline
                extras[name] = {"error": repr(ex)[:300]}
    gpu_torch, cpu = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_torch = gpu_torch_baseline(w, c.dev)
        cpu = cpu_baseline(w, args)
    if rank == 0:
        per = sorted(m["per"])
        line = {
            "metric": "encoder scans/sec (scan pairs through vfe -> backbone_3d -> loss)", "value": round(m["value"], 3), "unit": "scans/s",
            "n_gpus": world, "steps": args.steps, "warmup": m["warmup"], "ms_per_step": round(m["ms_per_step"], 3),
            "p50_ms_per_scan": round(per[len(per) // 2] / w["batch"], 3), "p95_ms_per_scan": round(per[min(len(per) - 1, int(0.95 * len(per)))] / w["batch"], 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"fp32": "f32", "tf32": "tf32 (tensor-core operands; f32 accumulate and storage)",
                      "bf16": "bf16 (activation storage and tensor-core operands; f32 accumulate, statistics, master weights)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": w["name"], "precision": f"encoder kernels {args.precision}; cuDNN decoder {args.decoder}",
                       "parallelism": f"dp{world} (scan-pair sharding" + ((", DistributedDataParallel NCCL gradient all-reduce)" if args.ddp else ", bucketed NCCL gradient all-reduce over one persistent flat buffer)") if w["train"] else ", no collective)"),
                       "l2": f"inputs cycle over {args.batches} distinct batches; per-step activation working set >> 126 MB L2",
                       "batching": "seed order" if args.unbalanced else "each rank sorts its own scan pairs by occupied pillars and batches neighbours (rank-local length-grouped sampler: lockstep steps line up across ranks)",
                       "allocator": f"torch caching allocator over one pre-reserved {c.slab_gib} GiB slab"},
            "clocks": m["clocks"],
            "e2e": {"value": round(m["e_value"], 3), "unit": "scans/s", "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"],
                    "ms_per_step": round(m["e_ms_per_step"], 3)},
            "gpu_launches": m["launches"], "gpu_launches_note": "kernel-family launches of libtmae_sm100.so inside the timed region (counted in the library; torch / cuDNN / NCCL kernels not included)",
            "gemm_dispatch": m["dispatch"], "host_enqueue_ms_per_step": round(m["host_ms"], 2),
            "rank_ms_per_step_min_max": [round(v, 3) for v in m["rank_ms_min_max"]],
            "roofline": m["roof"], "step_roofline": m["step_roof"], "cpu_baseline": cpu, "gpu_torch_baseline": gpu_torch, "extra_workloads": extras,
            "kernel_time_table": m["table"], "kernel_time_table_note": "per-kernel CUDA events inside the library over 2 extra steps with the weight-gradient GEMMs on the caller's stream (in the timed steps they run on the library's auxiliary stream and overlap the data-gradient kernels)",
            "counts": m["counts"], "attention_block": m["attention_block"],
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


def gpu_torch_baseline(w, dev, steps=2):
    """The reference's algorithm in STOCK PyTorch on the same B200 (tier-2 oracle on CUDA: torch / cuBLAS / cuDNN kernels,
    TF32 allowed, no AMP): the honest same-GPU comparator of SURVEY 8d.  Same workload (batch, points, fwd+bwd+AdamW);
    sparse convs dense-emulated (spconv is not installed here), so this is a reported baseline, not a target."""
    from oracle import restated
    from tmae_b200 import synth
    shape = synth.SHAPES[w["shape"]]
    grid = synth.grid_size(shape).tolist()
    try:
        tf = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = True
        torch.manual_seed(0)
        vfe, bb = restated.build(w["kind"], grid, shape["voxel"], shape["range"], num_point_features=w["npf"])
        vfe.to(dev).train(w["train"]), bb.to(dev).train(w["train"])
        opt = torch.optim.AdamW(list(vfe.parameters()) + list(bb.parameters()), lr=1e-4, weight_decay=0.01, fused=True) if w["train"] else None
        pts, ptsp = synth.batch(1000, w["batch"], w["n_points"], w["shape"])
        pts, ptsp = torch.from_numpy(pts).to(dev), torch.from_numpy(ptsp).to(dev)

        def step():
            if w["train"]:
                bb(vfe(dict(points=pts, points_prev=ptsp, batch_size=w["batch"])))
                bb.get_loss()[0].backward()
                opt.step()
                opt.zero_grad(set_to_none=True)
            else:
                with torch.no_grad():
                    bb(vfe(dict(points=pts, points_prev=ptsp, batch_size=w["batch"])))
        step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        return {"value": round(w["batch"] / ms * 1e3, 3), "unit": "scans/s", "ms_per_step": round(ms, 1), "steps": steps,
                "what": "tier-2 oracle (oracle/restated.py) on cuda:0 with stock torch ops, fp32 storage, TF32 matmul/conv allowed, same batch and step; "
                        "sparse convs dense-emulated; its ~100 host round trips per forward are part of the reference's algorithm (SURVEY F8)"}
    except Exception as e:
        return {"error": repr(e)[:300]}
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf
        torch.cuda.empty_cache()


def workload_counts(bb, resident):
    """Row counts of the last step that ran (log N, M per stage and the window counts with every run);
    read from the backbone's geometry plan after the timed region."""
    plans, _ = bb.last_plan
    names = ["current (visible)" if bb.__class__.__name__ == "SiamWCA_MAE" else "current", "previous", "both frames (Siamese-batched set)"]
    out = {"points_per_frame_set_first_batch": [int(resident[0][0].shape[0]), int(resident[0][1].shape[0])],
           "pillars_per_stage": {n: [int(st.m) for st in fp.stages] for n, fp in zip(names, plans)}}
    win = {}
    for s, st in enumerate(plans[-1].stages):   # the set that runs through the SST blocks carries the partition tables
        if st.part is not None:
            lb = st.part.level_base.cpu().tolist()   # (2 shifts, levels + 1): windows in front of each level, last = all
            win[f"stage{s + 1}"] = {"windows_shift0": int(lb[0][-1]), "windows_shift1": int(lb[1][-1]), "windows_by_level_shift0":
                                    [int(lb[0][i + 1] - lb[0][i]) for i in range(len(lb[0]) - 1)], "level_max_tokens": list(st.part.tokens)}
    out["windows"] = win
    return out


def attention_block(bb, table, train, tensor_peak_tflops):
    """Useful and executed tensor-core FLOPs of the window-attention core in one step, from the partition tables of the last plan
    (token counts per window, both shifts) and the module structure, next to the time of its kernel families.
    Per layer and window with q queries and k keys: S = QK^T and PV are 2 * 2 * q * k * C flops forward, five such products
    backward.  The tcgen05 kernels run windows of 17..64 tokens as 128 x 128 tiles (4 / 2 windows per tile): executed = 128 * 128 per
    tile and product whatever the occupancy; windows of <= 16 tokens run on warp kernels (SIMT, useful flops only)."""
    plans, tparts = bb.last_plan
    useful_tc = useful_simt = tiles_flops = 0.0

    def part_sums(part, shift, cross):
        lb = part.level_base[shift].cpu().tolist()
        n = int(lb[-1])
        cq = part.cnt_a[shift][:n].double().cpu()
        ck = part.cnt_b[shift][:n].double().cpu() if cross else cq
        prod = cq * ck
        out = []
        for li, tok in enumerate(part.tokens):
            a, b = int(lb[li]), int(lb[li + 1])
            out.append((tok, b - a, float(prod[a:b].sum())))
        return out

    passes = 14.0 if train else 4.0          # (2 products forward + 5 backward) x 2 flops per multiply-add
    tile_passes = 14.0 if train else 4.0
    stages = plans[-1].stages
    for si, blk in enumerate(bb.sst_blocks):
        C = blk.d_model
        for sb in blk.encoder_blocks:
            for shift, _ in enumerate(sb.encoder_list):
                for tok, nwin, pq in part_sums(stages[si].part, shift, False):
                    if tok <= 16:
                        useful_simt += passes * pq * C
                    else:
                        useful_tc += passes * pq * C
                        tiles_flops += tile_passes * (-(-nwin // (128 // tok))) * 128 * 128 * C
    if tparts is not None:
        for si, blk in enumerate(getattr(bb, "wca_blocks", [])):
            C = blk.d_model
            for shift, _ in enumerate(blk.encoder_blocks[0].encoder_list):
                for tok, nwin, pq in part_sums(tparts[si], shift, True):
                    if tok <= 16:
                        useful_simt += passes * pq * C
                    else:
                        useful_tc += passes * pq * C
                        tiles_flops += tile_passes * (-(-nwin // (128 // tok))) * 128 * 128 * C
    fam = {k: v["ms_per_step"] for k, v in (table or {}).items() if k.startswith("attn_")}
    ms_tc = sum(v for k, v in fam.items() if k.startswith("attn_tc"))
    ms_all = sum(fam.values())
    out = {"useful_gflop_per_step": round((useful_tc + useful_simt) / 1e9, 2), "useful_gflop_tcgen05_kernels": round(useful_tc / 1e9, 2),
           "executed_tile_gflop_tcgen05_kernels": round(tiles_flops / 1e9, 2), "useful_gflop_warp_kernels": round(useful_simt / 1e9, 2),
           "ms_per_step": round(ms_all, 3), "ms_per_step_tcgen05_kernels": round(ms_tc, 3), "kernel_families_ms": fam}
    if ms_tc > 0:
        out["tcgen05_kernels_useful_tflops"] = round(useful_tc / (ms_tc * 1e-3) / 1e12, 2)
        out["tcgen05_kernels_executed_tflops"] = round(tiles_flops / (ms_tc * 1e-3) / 1e12, 2)
        out["tcgen05_kernels_executed_frac_of_tensor_peak"] = round(tiles_flops / (ms_tc * 1e-3) / 1e12 / tensor_peak_tflops, 4)
    out["note"] = ("windows hold <= 64 tokens and heads are 16 / 32 wide: the op is HBM- and latency-bound (17-30 useful flop per byte against a ridge "
                   "of 220), see DESIGN.md section 6; ncu sm__pipe_tensor_cycles_active of the tcgen05 kernels: 5.1-5.3 % (profiles/r02_ncu_attn_in_bench.txt)")
    return out


def roofline(ops, step, resident, args, step_ms):
    """Per-kernel CUDA-event timing inside the library (tmae_profile_begin/_end, events recorded on the launching
    stream around every kernel family) over extra steps after the timed region; the dominant kernel family is
    reported against its roofline with ALGORITHMIC flops / bytes, and the whole step against
    the HBM time of the sum of its kernels' algorithmic bytes."""
    pk = peaks()
    n = 2

    def run():
        for i in range(n):
            step(*resident[i % len(resident)])
    # one stream for this pass: kernels that overlap (the weight-gradient GEMMs on the library's auxiliary stream) share the GPU and
    # stretch each other's event-timed duration -- the table is about the kernels, the overlap is in `ms_per_step`
    ops.set_option("wgrad_stream", 0)
    try:
        table = ops.lib_profile(run)
    finally:
        ops.set_option("wgrad_stream", int(os.environ.get("TMAE_OPT_WGRAD_STREAM", "1")))
    if not table:
        return None, None, None
    tot = sum(r["ms"] for r in table.values())
    rows = sorted(table.items(), key=lambda kv: -kv[1]["ms"])
    name, r = rows[0]
    ridge = pk["tensor"] * 1e12 / (pk["hbm"] * 1e9)  # flop per byte
    ai = r["flops"] / r["bytes"] if r["bytes"] else float("inf")
    tf = r["flops"] / (r["ms"] / 1e3) / 1e12
    gb = r["bytes"] / (r["ms"] / 1e3) / 1e9
    if r["flops"] > 0 and ai >= ridge:
        roof = {"kernel": name, "bound": "tensor", "achieved": round(tf, 3), "peak": pk["tensor"], "unit": "TFLOP/s", "frac": round(tf / pk["tensor"], 5)}
    else:
        roof = {"kernel": name, "bound": "hbm", "achieved": round(gb, 1), "peak": pk["hbm"], "unit": "GB/s", "frac": round(gb / pk["hbm"], 5)}
    traffic = traffic_req = None
    for tp in ("r02_traffic.json", "r01_traffic.json"):
        tp = os.path.join(ROOT, "profiles", tp)
        if os.path.exists(tp):  # dram__bytes_read + dram__bytes_write per launch of this kernel family from the committed ncu --set full capture
            rec = json.load(open(tp)).get(name, {})
            traffic = rec.get("traffic_bytes_per_launch")
            traffic_req = rec.get("l2_requested_bytes_per_launch")
            if traffic is not None:
                break
    roof.update(traffic=traffic, traffic_note=(f"ncu --set full, launches of this family captured inside the same program; those launches asked the L2 for {traffic_req} bytes each "
                                               "(a family's launches differ in size: compare traffic with THIS figure, not with the family average below)") if traffic_req else None,
                algorithmic_bytes_per_launch=round(r["bytes"] / r["calls"]) if r["calls"] else None, arithmetic_intensity_flop_per_byte=round(ai, 1) if r["bytes"] else None, ridge_flop_per_byte=round(ridge, 1),
                tensor_tflops=round(tf, 2), peak_source=pk["src"], avg_launch_us=round(r["ms"] * 1e3 / r["calls"], 2),
                launches_per_step=r["calls"] // n, share_of_library_time=round(r["ms"] / tot, 4))
    short = {k: {"ms_per_step": round(v["ms"] / n, 3), "launches_per_step": v["calls"] // n,
                 "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 2) if v["flops"] else None,
                 "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["bytes"] else None} for k, v in rows[:20]}
    step_bytes = sum(v["bytes"] for v in table.values()) / n
    step_roof = {"algorithmic_bytes_per_step": round(step_bytes), "ms_at_hbm_peak": round(step_bytes / (pk["hbm"] * 1e9) * 1e3, 3), "ms_per_step": round(step_ms, 3),
                 "frac": round(step_bytes / (pk["hbm"] * 1e9) * 1e3 / step_ms, 4), "library_kernel_ms_per_step": round(tot / n, 3),
                 "note": "sum over the library's kernel families of their algorithmic bytes (cuDNN decoder, torch glue and attention kernels without a byte model excluded) / measured HBM copy peak, against the measured step"}
    return roof, short, step_roof


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(w):
    """The reference's algorithm on the host (tier-2 oracle port): one B=1 scan pair of the same workload."""
    from oracle import restated
    from tmae_b200 import synth
    shape = synth.SHAPES[w["shape"]]
    grid = synth.grid_size(shape).tolist()
    torch.manual_seed(0)
    vfe, bb = restated.build(w["kind"], grid, shape["voxel"], shape["range"], num_point_features=w["npf"])
    vfe.train(w["train"]), bb.train(w["train"])
    pts, ptsp = synth.batch(1000, 1, w["n_points"], w["shape"])
    pts, ptsp = torch.from_numpy(pts), torch.from_numpy(ptsp)

    def step():
        if w["train"]:
            bd = bb(vfe(dict(points=pts, points_prev=ptsp, batch_size=1)))
            loss = bb.get_loss()[0]
            loss.backward()
            for p in list(vfe.parameters()) + list(bb.parameters()):
                p.grad = None
        else:
            with torch.no_grad():
                bb(vfe(dict(points=pts, points_prev=ptsp, batch_size=1)))
    return step


def cpu_baseline(w, args, steps=2):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(w)
    step()
    t = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t) / steps
    return {"value": round(1.0 / dt, 4), "unit": "scans/s", "cores": cores, "kind": "port",
            "sample": f"{steps} steps of ONE scan pair (batch 1) of the same workload, fp32, torch CPU ops, {cores} threads; "
                      "sparse convs dense-emulated"}


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(w)
    for _ in range(args.warmup):
        step()
    t = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t
    v = round(args.steps / dt, 4)
    sample = f"each step = ONE scan pair (batch 1) of the workload, fp32 torch CPU ops, {cores} threads; sparse convs dense-emulated"
    emit({
        "impl": "reference", "metric": "encoder scans/sec (scan pairs through vfe -> backbone_3d -> loss)", "value": v, "unit": "scans/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt / args.steps * 1e3, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "sample": sample},
        "cpu_baseline": {"value": v, "unit": "scans/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


_OUT = None


def emit(obj):
    """The ONE JSON line goes to the real stdout; everything libraries print (e.g. NCCL's version banner) was
    redirected to stderr in main()."""
    print(json.dumps(obj), file=_OUT or sys.stdout, flush=True)


def main():
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pretrain", choices=list(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32", "bf16"],
                    help="bf16 = bf16 activation storage + tcgen05 kind::f16 GEMMs (fp32 accumulate); tf32 = fp32 storage + tcgen05 kind::tf32 "
                         "GEMMs + TF32 mma attention; fp32 = FFMA parity mode.  Default: tmae_b200.ops.BENCH_PRECISION")
    ap.add_argument("--decoder", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--batches", type=int, default=16,
                    help="distinct scan-pair batches per rank (the timed steps cycle over them): enough that every rank's mean step cost is the population's, "
                         "not the luck of four scenes (measured: 21.1 .. 22.0 ms between ranks with 4)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-core and the stock-PyTorch-on-GPU baselines")
    ap.add_argument("--no-extra", dest="extra", action="store_false", help="skip the short finetune / Waymo-shaped lines (extra_workloads)")
    ap.add_argument("--grad-buckets", type=int, default=4, help="--grad-sync flat: slices of the flat gradient buffer all-reduced separately")
    ap.add_argument("--single-optimizer", action="store_true", help="N > 1, --grad-sync flat: one optimizer step behind the whole exchange instead of one per bucket")
    ap.add_argument("--unbalanced", action="store_true", help="batches in seed order instead of the rank-local cost-sorted order (make_batches)")
    ap.add_argument("--profile-run", action="store_true", help="for runs under ncu: warm up exactly --warmup steps (no allocator-stationarity minimum)")
    ap.add_argument("--no-clock-sampler", dest="clock_sampler", action="store_false")
    ap.add_argument("--grad-sync", default="flat", choices=["flat", "overlap"],
                    help="gradient exchange of the multi-GPU pretraining step: 'flat' = after backward, the persistent flat buffer is all-reduced in buckets "
                         "on a communication stream; 'overlap' = buckets are all-reduced from post-accumulate hooks while backward runs")
    ap.add_argument("--ddp", action="store_true", help="wrap the step in torch DistributedDataParallel instead of the flat gradient all-reduce")
    ap.add_argument("--no-side-stream", dest="side_stream", action="store_false",
                    help="run the coordinate-only pre-pass (voxelise, mask, plans) on the main stream instead of the library's side stream")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
