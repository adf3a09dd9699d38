"""Prints the largest parameter-gradient deviations of the product (fp32 parity mode) from the tier-2 oracle on the
golden pretraining case, relative to each tensor's largest entry:  python tools/grad_noise.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import golden_inputs, load_golden, run_tier2  # noqa: E402
from oracle import cases  # noqa: E402
import tmae_b200  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
g = load_golden("pretrain")
pts, ptsp = golden_inputs(g)
B, ms = g["meta"]["batch"], g["meta"]["mask_seed"]
S = cases.SMALL
vfe, bb = tmae_b200.build_model("pretrain", S["grid"], S["voxel"], S["range"])
cases.fill_params(vfe), cases.fill_params(bb)
vfe.cuda(), bb.cuda()
bd = vfe(dict(points=torch.from_numpy(pts).cuda(), points_prev=torch.from_numpy(ptsp).cuda(), batch_size=B))
bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"].cpu(), B, 0.75, ms).cuda()
bd = bb(bd)
bb.forward_ret_dict["pred_points"].retain_grad()
loss, _ = bb.get_loss()
loss.backward()
ovfe, obb, _, obd = run_tier2("pretrain", pts, ptsp, B, ms)
obb.forward_ret_dict["pred_points"].retain_grad()
oloss, _ = obb.get_loss()
oloss.backward()
pp, op_ = bb.forward_ret_dict["pred_points"], obb.forward_ret_dict["pred_points"]
dpe = (pp.grad.cpu() - op_.grad).abs().flatten(1).max(1).values
print("pred max err %.2e ; dpred: scale %.2e, pillars with err > 1e-3 of scale: %d of %d, max err %.2e" % (
    (pp.detach().cpu() - op_.detach()).abs().max().item(), op_.grad.abs().max().item(),
    int((dpe > 1e-3 * op_.grad.abs().max()).sum()), dpe.numel(), dpe.max().item()))
print("loss", loss.item(), oloss.item())
worst = []
op = dict(obb.named_parameters())
for k, p in bb.named_parameters():
    ref = op[k].grad
    worst.append(((p.grad.cpu() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12), k, tuple(p.shape)))
worst.sort(reverse=True)
for w in worst[:6]:
    print("%.2e %s %s" % w)
n = len(worst)
print("percentiles: median %.2e  p75 %.2e  p90 %.2e  (n=%d)" % (worst[n // 2][0], worst[n // 4][0], worst[n // 10][0], n))
for st in bb.last_plan[0][-1].stages:
    print("stage rows", st.m)
