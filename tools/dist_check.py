"""2-rank NCCL check of tmae_b200.dist on real GPUs: OverlappedGradients / FlatGradients against a plain all-reduce of cloned gradients.
    torchrun --nproc-per-node 2 tools/dist_check.py [overlap|flat]"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tmae_b200  # noqa: E402,F401
from tmae_b200 import dist as tdist  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "overlap"
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(0)
layers = []
for i in range(24):
    layers += [torch.nn.Linear(256, 256), torch.nn.LayerNorm(256), torch.nn.GELU()]
net = torch.nn.Sequential(*layers).to(dev)
params = list(net.parameters())
og = tdist.OverlappedGradients(params, world).attach() if mode == "overlap" else tdist.FlatGradients(params, world)
worst = 0.0
for it in range(30):
    x = torch.randn(4096 + 64 * ((it + rank) % 5), 256, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank))
    net(x).square().mean().backward()
    ref = [p.grad.clone() for p in params]
    (og.finish if mode == "overlap" else og.reduce)()
    for r in ref:
        dist.all_reduce(r)
        r.div_(world)
    worst = max(worst, max((p.grad - r).abs().max().item() / (r.abs().max().item() + 1e-12) for p, r in zip(params, ref)))
    for p in params:
        p.grad = None
torch.cuda.synchronize()
print(f"rank {rank}: mode {mode}, 30 steps, worst relative difference {worst:.2e}", flush=True)
assert worst < 1e-5
dist.destroy_process_group()
