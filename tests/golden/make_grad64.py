"""Float64 ground truth for the parameter gradients of the golden pretraining case.

The fp32 tier-2 oracle deviates from exact arithmetic by up to 6e-3 of a tensor's scale on this case (median 2e-4:
18 encoder layers + 13 BatchNorms amplify rounding), which is the same size as any correct fp32 implementation's
distance from it.  Gradient parity is therefore pinned to the SAME oracle evaluated in float64 (weights filled in
fp32 by `cases.fill_params`, then cast; voxel coordinates verified identical to the fp32 run).  Stored: for every
parameter its max |grad| and a strided sample (every STRIDE-th element) in float32.

    cd tests && python golden/make_grad64.py      (CPU, ~1 min; writes golden/small_pretrain_grad64.pt)
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from common import golden_inputs, load_golden  # noqa: E402
from oracle import cases, restated  # noqa: E402

STRIDE = 53


def main():
    g = load_golden("pretrain")
    pts, ptsp = golden_inputs(g)
    B, ms = g["meta"]["batch"], g["meta"]["mask_seed"]
    S = cases.SMALL
    vfe, bb = restated.build("pretrain", S["grid"], S["voxel"], S["range"])
    cases.fill_params(vfe), cases.fill_params(bb)
    vfe.double(), bb.double()
    bd = vfe(dict(points=torch.from_numpy(pts).double(), points_prev=torch.from_numpy(ptsp).double(), batch_size=B))
    assert torch.equal(bd["voxel_coords"], g["voxel_coords"].long()), "float64 voxelisation moved a point"
    bd["voxel_mae_mask_in"] = cases.fixed_mask(bd["voxel_coords"], B, 0.75, ms).double()
    bd = bb(bd)
    loss, _ = bb.get_loss()
    loss.backward()
    out = {"loss": loss.item(), "stride": STRIDE, "grads": {}}
    for pre, m in (("vfe.", vfe), ("backbone_3d.", bb)):
        for k, p in m.named_parameters():
            out["grads"][pre + k] = (p.grad.abs().max().item(), p.grad.flatten()[::STRIDE].float().clone())
    torch.save(out, os.path.join(HERE, "small_pretrain_grad64.pt"))
    print("loss", out["loss"], "tensors", len(out["grads"]))


if __name__ == "__main__":
    main()
