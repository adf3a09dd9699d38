"""Max abs error of the attention core (fwd + bwd) against the float64 oracle, small and large windows, self and cross."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_gpu_ops as T  # noqa: E402
from tmae_b200 import ops  # noqa: E402
from oracle import restated  # noqa: E402

DEV = "cuda"


def case(C, cross, tau_v, na=2500, nb=2200):
    H, B, g = 8, 2, 64
    _, _, levels = T._levels("pretrain")
    ca = T._coords(0, na if not cross else max(na // 4, 50), B, g)
    cb = T._coords(5, nb, B, g) if cross else None
    P = ops.window_partition(ca.to(DEV), B, g, g, levels, coords_b=cb.to(DEV) if cross else None)
    gen = torch.Generator().manual_seed(C)
    ma, mb = ca.shape[0], (cb.shape[0] if cross else ca.shape[0])
    q, k, v = torch.randn(ma, C, generator=gen), torch.randn(mb, C, generator=gen), torch.randn(mb, C, generator=gen)
    tau = torch.tensor([[[tau_v]]])
    shift = 1
    nw = int(P.n_win[shift])
    qt, qc = P.tok_a[shift], P.cnt_a[shift]
    kt, kc = (P.tok_b[shift], P.cnt_b[shift]) if cross else (qt, qc)
    args = (qt, qc, kt, kc, P.n_win[shift:shift + 1], ops.small_end(P, shift), ops.mid_end(P, shift), min(P.wcap, ma), tau.to(DEV), 0.01, H)
    o, lse = ops.window_attention_fwd(q.to(DEV), k.to(DEV), v.to(DEV), *args, zero_out=cross)
    mha = restated.CosineMHA(C, H).double()
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.eye(C).repeat(3, 1)), mha.in_proj_bias.zero_()
        mha.out_proj.weight.copy_(torch.eye(C)), mha.out_proj.bias.zero_(), mha.tau.copy_(tau)
    wa, sa = P.win_a[shift, :ma].cpu().long(), P.slot_a[shift, :ma].cpu().long()
    wb, sb = (P.win_b[shift, :mb].cpu().long(), P.slot_b[shift, :mb].cpu().long()) if cross else (wa, sa)
    ka, kb = wa >= 0, wb >= 0
    qd, kd, vd = q.double().requires_grad_(), k.double().requires_grad_(), v.double().requires_grad_()
    ref = T._attn_ref(mha, qd[ka], kd[kb], vd[kb], wa[ka], wb[kb], sa[ka], sb[kb], nw)
    do = torch.randn(ma, C, generator=gen)
    ref.backward(do.double()[ka])
    dtau = torch.zeros(1, 1, 1, device=DEV)
    dq, dk, dv = ops.window_attention_bwd(do.to(DEV), q.to(DEV), k.to(DEV), v.to(DEV), o, lse, *args, dtau, zero=cross)
    cnt = qc[:nw].cpu()
    small = cnt[wa.clamp(min=0)] <= 16
    e = lambda a, b, m: ((a.cpu().double() - b)[m].abs().max().item() if m.any() else 0.0)
    print(f"C={C} cross={cross} tau={tau_v} na={na} small windows {int((cnt <= 16).sum())}/{nw}: o {e(o, torch.zeros_like(qd).index_put((torch.where(ka)[0],), ref.detach()), ka):.2e} | "
          f"dq small {e(dq, qd.grad, ka & small):.2e} large {e(dq, qd.grad, ka & ~small):.2e} | dk {e(dk, kd.grad, kb):.2e} dv {e(dv, vd.grad, kb):.2e} | "
          f"dtau rel {abs(dtau.item() - mha.tau.grad.item()) / abs(mha.tau.grad.item()):.2e}")


for C in (128, 256):
    for cross in (False, True):
        for na, nb in ((300, 300), (1200, 1000), (2500, 2200)):
            case(C, cross, 0.37, na, nb)
