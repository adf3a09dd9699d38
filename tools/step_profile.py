"""torch.profiler summary of one pretraining step (all CUDA kernels, ours and torch/cuDNN's), to see what is left
outside the library:  python tools/step_profile.py [bf16|fp32]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import tmae_b200  # noqa: E402
from tmae_b200 import ops, synth  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
w = bench.WORKLOADS["pretrain"]
dev = torch.device("cuda", 0)
grid = synth.grid_size(synth.ONCE).tolist()
torch.manual_seed(0)
vfe, bb = tmae_b200.build_model("pretrain", grid, synth.ONCE["voxel"], synth.ONCE["range"])
vfe.to(dev), bb.to(dev)
ops.set_precision(prec)
bb.decoder_autocast = torch.bfloat16
torch.backends.cudnn.benchmark = True
opt = torch.optim.AdamW(list(vfe.parameters()) + list(bb.parameters()), lr=1e-4, fused=True)
host = bench.make_batches(w, 2, 0)
res = [(a.to(dev), b.to(dev)) for a, b in host]


def step(i):
    bd = vfe(dict(points=res[i % 2][0], points_prev=res[i % 2][1], batch_size=w["batch"]))
    bd = bb(bd)
    loss = bb.get_loss()[0]
    loss.backward()
    opt.step()
    opt.zero_grad(set_to_none=True)


for i in range(3):
    step(i)
torch.cuda.synchronize()
import time
t = time.perf_counter()
for i in range(5):
    step(i)
t_issue = time.perf_counter() - t
torch.cuda.synchronize()
t_all = time.perf_counter() - t
print(f"5 steps: host issue {t_issue / 5 * 1e3:.1f} ms/step, wall {t_all / 5 * 1e3:.1f} ms/step")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(2):
        step(i)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
ka = prof.key_averages()
kern = [(e.key, e.device_time_total / 2e3, e.count // 2) for e in ka if e.device_type == DeviceType.CUDA]
kern.sort(key=lambda r: -r[1])
tot = sum(r[1] for r in kern)
ours = sum(r[1] for r in kern if "tmae::" in r[0])
print(f"GPU kernel time {tot:.2f} ms/step ({sum(r[2] for r in kern)} launches), of which library kernels {ours:.2f} ms; wall {t_all / 5 * 1e3:.1f} ms/step")
print("top CUDA kernels by device time (ms per step, launches per step):")
for k, ms, n in kern[:40]:
    print(f"{ms:9.3f} {n:6d}  {k[:120]}")
ops_ = [(e.key, e.self_cpu_time_total / 2e3, e.count // 2) for e in ka if e.device_type == DeviceType.CPU]
ops_.sort(key=lambda r: -r[1])
print("top host-side ops by self CPU time (ms per step):")
for k, ms, n in ops_[:15]:
    print(f"{ms:9.3f} {n:6d}  {k[:100]}")
