#!/usr/bin/env python
"""Benchmark of the T-MAE sparse-window voxel-encoder hot path on B200 (driver contract: see DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload pretrain|finetune] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one pass of the hot path over one batch of synthetic ONCE-shaped scan pairs:
  pretrain (default, BASELINE.json configs[1]): TemporalDynVFE -> SiamWCA_MAE.forward -> Chamfer loss -> backward
            -> AdamW step, batch 4 scan pairs per GPU, 75 % voxel mask;
  finetune (configs[3]): TemporalDynVFE -> SiamWCA.forward, batch 8, 120k-point scans, no gradient.
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

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "pretrain": dict(kind="pretrain", batch=4, n_points=60000, train=True,
                     name="T-MAE pretraining forward+backward+AdamW, synthetic ONCE scan pairs (60k pts/frame, 0.32 m pillars, "
                          "468x468 grid), 75% voxel mask, Chamfer loss, batch 4 scan pairs per GPU"),
    "finetune": dict(kind="finetune", batch=8, n_points=120000, train=False,
                     name="finetune-mode encoder forward (no masking, temporal cross-attention), synthetic ONCE scans "
                          "(120k pts/frame, 0.32 m pillars), batch 8 scan pairs per GPU"),
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
    """SM clock and throttle reasons sampled DURING the timed region through NVML in a background thread (a looping
    nvidia-smi process was measured to slow the timed region by ~15 %); falls back to nvidia-smi when NVML is absent."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index, period=float(os.environ.get("TMAE_CLOCK_PERIOD", "0.25"))):
        self.index, self.period = index, period
        self.sm, self.mx, self.reasons, self.stop, self.t, self.proc, self.lines = [], 0, set(), False, None, None, []

    def _nvml_loop(self, nv, h):
        bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else 0x8,
                "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}
        while not self.stop:
            try:
                self.sm.append(int(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[self.index]) if vis and vis.split(",")[self.index].isdigit() else self.index
            h = nv.nvmlDeviceGetHandleByIndex(idx)
            self.mx = int(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.t = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.t.start()
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


def make_batches(w, n_batches, rank):
    from tmae_b200 import synth
    out = []
    for seeds in shard_seeds(w, n_batches, rank):
        pts, ptsp = synth.batch(seeds[0], w["batch"], w["n_points"])
        out.append((torch.from_numpy(pts).pin_memory(), torch.from_numpy(ptsp).pin_memory()))
    return out


# ------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import tmae_b200
    from tmae_b200 import dist as tdist, ops, synth
    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    w = WORKLOADS[args.workload]
    shape = synth.ONCE
    grid = synth.grid_size(shape).tolist()
    # One slab for the caching allocator: the step's tensor sizes change with every batch (row counts), and blocks recorded
    # on two streams return to the pool late, so a pool made of many cudaMalloc'd segments sized to past requests keeps
    # hitting a size nothing fits -- a cudaMalloc in the middle of a step synchronises the device (one 70-100 ms step in
    # about a quarter of the runs).  A single large cached segment is split and re-merged on demand instead.
    slab_gib = min(64, int(torch.cuda.mem_get_info(dev)[0] * 0.45) >> 30)
    if slab_gib > 0:
        del_me = torch.empty(slab_gib << 30, dtype=torch.uint8, device=dev)
        del del_me
    torch.manual_seed(0)
    vfe, bb = tmae_b200.build_model(w["kind"], grid, shape["voxel"], shape["range"])
    args.precision = args.precision or ops.BENCH_PRECISION
    ops.set_precision(args.precision)
    bb.decoder_autocast = torch.bfloat16 if args.decoder == "bf16" else None
    torch.backends.cudnn.benchmark = True

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
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local], find_unused_parameters=False)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0.01, fused=True)
    params = [p for p in model.parameters() if p.requires_grad]
    if world > 1:  # same initial weights on every rank (DDP broadcasts them; the flat all-reduce path does it here)
        for p in list(model.parameters()) + list(model.buffers()):
            torch.distributed.broadcast(p.data, 0)
    bb.mask_generator = torch.Generator(device=dev).manual_seed(2000 + rank)

    host = make_batches(w, args.batches, rank)
    resident = [(a.to(dev), b.to(dev)) for a, b in host]
    h2d = sum(t.numel() * 4 for t in host[0])

    side = ops.side_stream(dev) if args.side_stream else None

    def step(pts, ptsp, module=None):
        m = net if module is None else module
        if w["train"]:
            loss = m(pts, ptsp, side)
            loss.backward()
            if world > 1 and not args.ddp and module is None:
                tdist.allreduce_gradients(params, world)  # ONE flat NCCL all-reduce of the 47 MB of gradients (tmae_b200/dist.py)
            opt.step()
            opt.zero_grad(set_to_none=True)
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
    host_out = [torch.zeros(1, dtype=torch.float32).pin_memory() for _ in range(max(args.steps, 1))]

    def timed(from_host):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        d2h = 0
        gc.collect()
        gc.disable()   # a generation-2 collection of the Python heap is a 50-250 ms host pause (seen as one slow step in ~half the runs)
        barrier()
        c0 = ops.launch_count()
        host_t0 = time.perf_counter()
        ev[0].record()
        for i in range(args.steps):
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
        host_ms.append((time.perf_counter() - host_t0) * 1e3 / args.steps)  # host time to ENQUEUE the steps (no sync when resident)
        barrier()
        gc.enable()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
        ms = torch.cuda.memory_stats(dev)
        print(f"[bench] {'e2e' if from_host else 'resident'} per-step ms: " + " ".join(f"{p:.1f}" for p in per) +
              f" | cudaMalloc calls so far {ms.get('num_device_alloc', 0)}, frees {ms.get('num_device_free', 0)}, retries {ms.get('num_alloc_retries', 0)}, "
              f"reserved {ms.get('reserved_bytes.all.current', 0) / 2**30:.1f} GiB", file=sys.stderr)
        total = ev[0].elapsed_time(ev[-1])
        return total, per, ops.launch_count() - c0, d2h

    # every distinct batch four times: its row counts are new sizes for the caching allocator, and with the host running a
    # step ahead of the GPU (no sync in the loop) blocks recorded on two streams return to the pool late, so the pool keeps
    # growing (cudaMalloc stalls of 50-250 ms) for ~14 steps before it is stationary
    # (and a one-off 100-300 ms driver-side stall was observed at the ~23rd step of a process in a third of the runs,
    # with or without the clock sampler: the warm-up runs past it)
    args.warmup = max(args.warmup, 7 * len(resident))
    for i in range(args.warmup):
        step(*resident[i % len(resident)])
    if args.clock_sampler:
        with ClockSampler(local) as cs:
            total, per, launches, _ = timed(False)
        clocks = cs.summary()
    else:
        total, per, launches, _ = timed(False)
        clocks = None
    steps_saved, args.steps = args.steps, len(host)   # untimed pass over the host-input path: every distinct batch once (new allocator sizes)
    timed(True)
    args.steps = steps_saved
    e_total, e_per, _, d2h = timed(True)
    assert all(torch.isfinite(t).all() for t in host_out), "non-finite step result"

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t)

    total, e_total = maxr(total), maxr(e_total)
    value = aggregate_value(w["batch"], args.steps, world, total)
    e_value = aggregate_value(w["batch"], args.steps, world, e_total)

    roof, prof_table = None, None
    if rank == 0:
        roof, prof_table = roofline(ops, step_local, resident, args)
    counts = None
    if rank == 0:
        try:
            counts = workload_counts(bb, resident)
        except Exception as e:  # diagnostics only: never lose the measurement over them
            counts = {"error": repr(e)}
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(w, args)

    if rank == 0:
        line = {
            "metric": "encoder scans/sec (scan pairs through vfe -> backbone_3d -> loss)", "value": round(value, 3), "unit": "scans/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(total / args.steps, 3),
            "p50_ms_per_scan": round(float(np.median(per)) / w["batch"], 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"fp32": "f32", "tf32": "tf32 (tensor-core operands; f32 accumulate and storage)",
                                     "bf16": "bf16 (activation storage and tensor-core operands; f32 accumulate, statistics, master weights)"}[args.precision], "data": "synthetic",
            "config": {"workload": w["name"], "precision": f"encoder kernels {args.precision}; cuDNN decoder {args.decoder}",
                       "parallelism": f"dp{world} (scan-pair sharding" + ((", DistributedDataParallel NCCL gradient all-reduce)" if args.ddp else ", one flat NCCL gradient all-reduce per step)") if w["train"] else ", no collective)"),
                       "l2": f"inputs cycle over {args.batches} distinct batches; per-step activation working set >> 126 MB L2",
                       "allocator": f"torch caching allocator over one pre-reserved {slab_gib} GiB slab"},
            "clocks": clocks,
            "e2e": {"value": round(e_value, 3), "unit": "scans/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": round(e_total / args.steps, 3)},
            "gpu_launches": launches, "host_enqueue_ms_per_step": round(host_ms[0], 2),
            "roofline": roof, "cpu_baseline": cpu, "kernel_time_table": prof_table, "counts": counts,
        }
        emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()


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


def roofline(ops, step, resident, args):
    """Per-kernel CUDA-event timing inside the library (tmae_profile_begin/_end, events recorded on the launching
    stream around every kernel family) over extra steps after the timed region; the dominant kernel family is
    reported against its roofline with ALGORITHMIC flops / bytes."""
    pk = peaks()
    n = 2

    def run():
        for i in range(n):
            step(*resident[i % len(resident)])
    table = ops.lib_profile(run)
    if not table:
        return None, None
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
    traffic = None
    tp = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tp):  # dram__bytes_read + dram__bytes_write per launch of this kernel family from the committed ncu --set full capture
        traffic = json.load(open(tp)).get(name, {}).get("traffic_bytes_per_launch")
    roof.update(traffic=traffic, algorithmic_bytes_per_launch=round(r["bytes"] / r["calls"]) if r["calls"] else None, arithmetic_intensity_flop_per_byte=round(ai, 1) if r["bytes"] else None, ridge_flop_per_byte=round(ridge, 1),
                tensor_tflops=round(tf, 2), peak_source=pk["src"], avg_launch_us=round(r["ms"] * 1e3 / r["calls"], 2),
                launches_per_step=r["calls"] // n, share_of_library_time=round(r["ms"] / tot, 4))
    short = {k: {"ms_per_step": round(v["ms"] / n, 3), "launches_per_step": v["calls"] // n,
                 "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 2) if v["flops"] else None,
                 "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["bytes"] else None} for k, v in rows[:14]}
    return roof, short


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_step_fn(w):
    """The reference's algorithm on the host (tier-2 oracle port): one B=1 scan pair of the same workload."""
    from oracle import restated
    from tmae_b200 import synth
    shape = synth.ONCE
    grid = synth.grid_size(shape).tolist()
    torch.manual_seed(0)
    vfe, bb = restated.build(w["kind"], grid, shape["voxel"], shape["range"])
    vfe.train(w["train"]), bb.train(w["train"])
    pts, ptsp = synth.batch(1000, 1, w["n_points"])
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
    ap.add_argument("--batches", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-sampler", dest="clock_sampler", action="store_false")
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
