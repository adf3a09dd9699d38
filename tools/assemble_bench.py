"""Measures the batch-assembly step at the bench's shape: 4 ONCE scan pairs of 60k raw points.

    python tools/assemble_bench.py [--batch 4] [--points 60000] [--kind once] > gpurun_out/assemble.json

Reports (a) the library kernels alone on device-resident raw points (CUDA events, L2 flushed between iterations, GPU
parked while the host enqueues) against the measured HBM copy peak -- algorithmic bytes = 4 F read + 4 (1 + F) written
per point; (b) the public call `FrameAssembler(samples)` from host numpy arrays (pinned staging copy + H2D + kernels +
the 8-byte count read-back) in scan pairs per second; (c) the oracle's numpy path on the host cores for the same samples.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa: E402,F401
from tmae_b200 import ops, synth  # noqa: E402
from tmae_b200.assemble import FrameAssembler, pose_affines  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--points", type=int, default=60000)
    ap.add_argument("--kind", default="once")
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    S = synth.SHAPES[a.kind]
    samples = [synth.raw_scan_pair(900 + i, a.points, a.kind) for i in range(a.batch)]
    asm = FrameAssembler(S["range"])
    dev = "cuda"
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]

    # (a) kernels alone, previous frame set (both affine maps on)
    raw = torch.from_numpy(np.concatenate([s["points_prev"] for s in samples])).to(dev)
    n, F = raw.shape
    offs = torch.from_numpy(np.concatenate([[0], np.cumsum([s["points_prev"].shape[0] for s in samples])]).astype(np.int64)).to(dev)
    xf = [pose_affines(s["pose_prev"], s["pose"]) for s in samples]
    xform = torch.from_numpy(np.stack([x for x, _ in xf])).to(dev)
    flags = torch.from_numpy(np.stack([f for _, f in xf])).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def run():
        return ops.assemble_frames(raw, offs, a.batch, xform, flags, 2.0, asm.crop_xyxy)
    for _ in range(3):
        run()
    ts = []
    for _ in range(a.iters):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    k_ms = ts[len(ts) // 2]
    out, count = run()
    kept = int(count.item())
    alg = 4.0 * (n * F + kept * (1 + F))          # rows read + kept rows written (sentinel rows not counted)

    # (b) public call from host arrays
    for _ in range(3):
        asm(samples)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.iters):
        bd = asm(samples)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) / a.iters * 1e3
    t0 = time.perf_counter()
    for _ in range(a.iters):
        bd = asm(samples, sync=False)
    torch.cuda.synchronize()
    nosync_ms = (time.perf_counter() - t0) / a.iters * 1e3

    # (c) oracle on the host
    from oracle import assemble_ref
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        p, q = assemble_ref.assemble(samples, S["range"])
    cpu_ms = (time.perf_counter() - t0) / reps * 1e3
    same = bool(np.array_equal(bd["points_prev"][:q.shape[0]].cpu().numpy(), q) and np.array_equal(asm(samples)["points"].cpu().numpy(), p))

    print(json.dumps({
        "workload": f"{a.batch} raw {a.kind} scan pairs x {a.points} points/frame, F={F}",
        "kernel": {"name": "assemble_frames (asm_count_kernel + asm_write_kernel), previous frame set", "ms": round(k_ms, 4),
                   "points": n, "kept": kept, "algorithmic_bytes": alg, "achieved_gbs": round(alg / k_ms / 1e6, 1),
                   "peak_gbs": peak, "frac": round(alg / k_ms / 1e6 / peak, 4),
                   "note": "two launches of ~%.1f MB: launch-latency bound at this size" % (alg / 1e6)},
        "public_call": {"ms_per_batch": round(e2e_ms, 3), "ms_per_batch_no_host_read": round(nosync_ms, 3),
                        "scan_pairs_per_s": round(a.batch / e2e_ms * 1e3, 1), "h2d_bytes_per_batch": 2 * asm.h2d_bytes},
        "cpu_oracle": {"ms_per_batch": round(cpu_ms, 2), "scan_pairs_per_s": round(a.batch / cpu_ms * 1e3, 1), "kind": "port (numpy, 1 thread + BLAS)"},
        "bit_exact_vs_oracle": same}))


if __name__ == "__main__":
    main()
