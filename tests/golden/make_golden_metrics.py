#!/usr/bin/env python
"""Golden fixture of the batched trajectory metrics from the LIVE reference (build container only):

    python tests/golden/make_golden_metrics.py     ->  tests/golden/metrics.npz

Outputs of ``src.eval.metrics.compute_metrics_batch`` (metrics.py:68-128) on seeded inputs: in-maze random walks,
out-of-bounds excursions, exact half-cell positions (round-half-even), broadcast occ / goal, T = 2 (< 3: smoothness 0).
"""
import os
import sys

import numpy as np
import torch

REF = os.environ.get("IDB200_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
from src.eval.metrics import compute_metrics_batch  # noqa: E402

torch.set_num_threads(4)
g = torch.Generator().manual_seed(17)
out = {}
for name, (B, T, D, H, W) in {"a": (33, 64, 2, 21, 21), "b": (7, 256, 2, 9, 12), "c": (5, 2, 2, 21, 21), "d": (12, 64, 4, 21, 21)}.items():
    occ = (torch.rand((B, H, W), generator=g) < 0.25).float()
    start = torch.rand((B, 1, D), generator=g)
    traj = start + 0.03 * torch.randn((B, T, D), generator=g).cumsum(dim=1)
    traj[0, T // 2:, 0] += 0.8                                   # leaves the unit square
    if T > 4:
        traj[1, :4, 0] = torch.tensor([0.5 / (W - 1), 1.5 / (W - 1), 2.5 / (W - 1), 0.0])   # exact .5 cells: round-half-even
        traj[1, :4, 1] = torch.tensor([0.5 / (H - 1), 1.5 / (H - 1), 2.5 / (H - 1), 1.0])
    goal = traj[:, -1].clone() + 0.05 * torch.randn((B, D), generator=g)
    goal[2] = traj[2, -1]                                        # distance exactly 0 -> success
    gt = traj + 0.01 * torch.randn((B, T, D), generator=g)
    res = compute_metrics_batch(occ, traj, goal, gt)
    out.update({f"{name}_occ": occ, f"{name}_traj": traj, f"{name}_goal": goal, f"{name}_gt": gt})
    out.update({f"{name}_out_{k}": v for k, v in res.items()})
    if name == "a":                                              # broadcast occ [H,W] and goal [D], no gt
        res1 = compute_metrics_batch(occ[0], traj, goal[0])
        out.update({f"a1_out_{k}": v for k, v in res1.items()})
np.savez_compressed(os.path.join(HERE, "metrics.npz"), **{k: v.detach().cpu().numpy() for k, v in out.items()})
print("wrote metrics.npz", len(out), "arrays")
