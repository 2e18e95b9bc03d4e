#!/usr/bin/env python
"""Golden fixture of the Stage-2 training-step tail from LIVE torch + the reference's EMA (build container only):

    python tests/golden/make_golden_optim.py     ->  tests/golden/optim.npz

The loss expression of train_interp_levels.py:1144-1156 (both weight branches) with autograd's d loss / d delta_hat, then
three steps of clip_grad_norm_(1.0) -> torch.optim.AdamW(lr 2e-4, wd 1e-2).step() -> EMA(0.999).update() (src/utils/ema.py)
on a small set of parameter tensors with recorded gradients.
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("IDB200_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
from src.utils.ema import EMA  # noqa: E402

torch.set_num_threads(4)
g = torch.Generator().manual_seed(23)
out = {}
# ---- loss + gradient seed ----
B, T, D = 7, 64, 2
for name, anchor_conf in (("conf", True), ("mask", False)):
    delta_hat = torch.randn((B, T, D), generator=g, requires_grad=True)
    target = torch.randn((B, T, D), generator=g) * 0.3
    weight = torch.rand((B, T), generator=g) if anchor_conf else (torch.rand((B, T), generator=g) < 0.3)
    grad_accum = 2 if anchor_conf else 1
    diff = ((delta_hat - target) ** 2).sum(dim=-1)
    if anchor_conf:
        w = float(1.0) + (float(0.1) - float(1.0)) * weight
    else:
        w = torch.where(weight, torch.tensor(0.1), torch.tensor(1.0))
    loss = (diff * w).sum() / (w.sum() * D + 1e-8)
    loss = loss / grad_accum
    loss.backward()
    out.update({f"loss_{name}_delta_hat": delta_hat.detach(), f"loss_{name}_target": target, f"loss_{name}_weight": weight,
                f"loss_{name}_value": loss.detach(), f"loss_{name}_grad": delta_hat.grad, f"loss_{name}_grad_accum": torch.tensor(grad_accum)})
# ---- clip + AdamW + EMA, three steps ----
shapes = [(37, 5), (64,), (3, 3, 3, 2), (1,), (129, 7)]
params = [torch.nn.Parameter(torch.randn(s, generator=g) * 0.5) for s in shapes]
opt = torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-2)
ema = EMA(params, decay=0.999)
for i, p in enumerate(params):
    out[f"opt_p0_{i}"] = p.detach().clone()
for step in range(1, 4):
    for i, p in enumerate(params):
        p.grad = torch.randn(p.shape, generator=g) * (3.0 if step == 2 else 0.02)     # step 2 is clipped, 1 and 3 are not
        out[f"opt_g{step}_{i}"] = p.grad.clone()
    total = torch.nn.utils.clip_grad_norm_(params, 1.0)
    out[f"opt_norm{step}"] = total.detach().clone()
    opt.step()
    opt.zero_grad(set_to_none=True)
    ema.update(params)
    for i, p in enumerate(params):
        out[f"opt_p{step}_{i}"] = p.detach().clone()
        out[f"opt_ema{step}_{i}"] = ema.shadow[i].clone()
        st = opt.state[p]
        out[f"opt_m{step}_{i}"] = st["exp_avg"].clone()
        out[f"opt_v{step}_{i}"] = st["exp_avg_sq"].clone()
np.savez_compressed(os.path.join(HERE, "optim.npz"), **{k: v.detach().cpu().numpy() for k, v in out.items()})
print("wrote optim.npz", len(out), "arrays")
