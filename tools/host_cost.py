"""Host cost of one pretraining step: the same step as the bench on TINY scans (2k points per frame), so that the ~970
kernel launches take almost no GPU time and the step time is what the host needs to enqueue them (Python, autograd, ctypes,
tensor-map encoding, cudaLaunch).  That is the floor of the step time until the launch sequence is captured in CUDA graphs.

    python tools/host_cost.py [n_points]
"""
import gc
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tmae_b200  # noqa: E402
from tmae_b200 import ops, synth  # noqa: E402

n_points = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
w = dict(bench.WORKLOADS["pretrain"], n_points=n_points)
dev = torch.device("cuda", 0)
grid = synth.grid_size(synth.ONCE).tolist()
torch.manual_seed(0)
vfe, bb = tmae_b200.build_model("pretrain", grid, synth.ONCE["voxel"], synth.ONCE["range"])
ops.set_precision(ops.BENCH_PRECISION)
bb.decoder_autocast = torch.bfloat16
torch.backends.cudnn.benchmark = True
vfe.to(dev), bb.to(dev)
opt = torch.optim.AdamW(list(vfe.parameters()) + list(bb.parameters()), lr=1e-4, weight_decay=0.01, fused=True)
batches = [(a.to(dev), b.to(dev)) for a, b in bench.make_batches(w, 2, 0)]
side = ops.side_stream(dev)


def step(p, pp):
    bd = bb(vfe(dict(points=p, points_prev=pp, batch_size=w["batch"], side_stream=side)))
    bb.get_loss()[0].backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(10):
    step(*batches[i % 2])
torch.cuda.synchronize()
gc.collect()
gc.disable()
c0 = ops.launch_count()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
steps = 20
t0 = time.perf_counter()
e0.record()
for i in range(steps):
    step(*batches[i % 2])
host = (time.perf_counter() - t0) / steps * 1e3
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / steps * 1e3
print(f"{n_points} points per frame: host enqueue {host:.2f} ms per step, wall {wall:.2f} ms per step, GPU span {e0.elapsed_time(e1) / steps:.2f} ms per step, "
      f"{(ops.launch_count() - c0) // steps} library calls per step")
