"""cProfile of the host side of the pretraining step (the bench is host-bound when enqueue time ~ GPU time)."""
import cProfile
import os
import pstats
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tmae_b200  # noqa: E402
from tmae_b200 import ops, synth  # noqa: E402

w = bench.WORKLOADS["pretrain"]
dev = torch.device("cuda", 0)
shape = synth.ONCE
grid = synth.grid_size(shape).tolist()
torch.manual_seed(0)
vfe, bb = tmae_b200.build_model(w["kind"], grid, shape["voxel"], shape["range"])
ops.set_precision(ops.BENCH_PRECISION)
bb.decoder_autocast = torch.bfloat16
torch.backends.cudnn.benchmark = True
vfe.to(dev), bb.to(dev)
params = list(vfe.parameters()) + list(bb.parameters())
opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=0.01, fused=True)
batches = [(a.to(dev), b.to(dev)) for a, b in bench.make_batches(w, 2, 0)]


side = ops.side_stream(dev) if "--no-side" not in sys.argv else None


def step(p, pp):
    bd = dict(points=p, points_prev=pp, batch_size=w["batch"])
    if side is not None:
        bd["side_stream"] = side
    bd = vfe(bd)
    bd = bb(bd)
    loss = bb.get_loss()[0]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(6):
    step(*batches[i % 2])
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(6):
    step(*batches[i % 2])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(60)
st.sort_stats("tottime").print_stats(30)
