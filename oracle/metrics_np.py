"""CPU restatement (numpy, fp32) of the reference's batched trajectory metrics -- TEST INFRASTRUCTURE ONLY.

Follows ``src/eval/metrics.py``: ``_pos_to_cell`` :13-24, ``compute_metrics_batch`` :68-128 of
EquilibriaW/Interpolated_Diffusion.  Pinned against ``tests/golden/metrics.npz`` (outputs of the live reference,
``tests/golden/make_golden_metrics.py``).  Only tests / smoke / the bench CPU leg may import this module.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np

F32 = np.float32


def pos_to_cell(pos: np.ndarray, h: int, w: int):
    """metrics.py:13-24 (torch.round = round-half-to-even = np.rint)."""
    x, y = pos[..., 0], pos[..., 1]
    oob = (x < 0) | (x > 1) | (y < 0) | (y > 1)
    j = np.rint(x * F32(max(w - 1, 1))).astype(np.int64)
    i = np.rint(y * F32(max(h - 1, 1))).astype(np.int64)
    return np.clip(i, 0, h - 1), np.clip(j, 0, w - 1), oob


def compute_metrics_batch(occ: np.ndarray, traj: np.ndarray, goal: np.ndarray, gt: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """metrics.py:68-128.  occ [B|1,H,W] or [H,W]; traj [B,T,D] or [T,D]; goal [B|1,D] or [D]; gt like traj."""
    traj = np.asarray(traj, dtype=F32)
    occ = np.asarray(occ, dtype=F32)
    goal = np.asarray(goal, dtype=F32)
    if traj.ndim == 2:
        traj = traj[None]
    if occ.ndim == 2:
        occ = occ[None]
    if goal.ndim == 1:
        goal = goal[None]
    B, T = traj.shape[:2]
    if occ.shape[0] != B:
        if occ.shape[0] != 1:
            raise ValueError("occ batch size does not match traj batch size")
        occ = np.broadcast_to(occ, (B,) + occ.shape[1:])
    if goal.shape[0] != B:
        if goal.shape[0] != 1:
            raise ValueError("goal batch size does not match traj batch size")
        goal = np.broadcast_to(goal, (B, goal.shape[1]))
    h, w = occ.shape[-2:]
    i, j, oob = pos_to_cell(traj, h, w)
    coll = (occ[np.arange(B)[:, None], i, j] > 0.5) | oob
    out = {"collision_rate": coll.astype(F32).mean(axis=1, dtype=F32)}
    goal_dist = np.sqrt(((traj[:, -1] - goal) ** 2).sum(axis=-1, dtype=F32))
    out["goal_dist"] = goal_dist
    out["success"] = (goal_dist < F32(1.0 / float(w))).astype(F32)
    out["path_length"] = np.sqrt(((traj[:, 1:] - traj[:, :-1]) ** 2).sum(axis=-1, dtype=F32)).sum(axis=1, dtype=F32)
    if T < 3:
        out["smoothness"] = np.zeros_like(goal_dist)
    else:
        acc = traj[:, 2:] - F32(2) * traj[:, 1:-1] + traj[:, :-2]
        out["smoothness"] = np.sqrt((acc ** 2).sum(axis=-1, dtype=F32)).mean(axis=1, dtype=F32)
    if gt is not None:
        gt = np.asarray(gt, dtype=F32)
        if gt.ndim == 2:
            gt = gt[None]
        if gt.shape[0] != B:
            if gt.shape[0] != 1:
                raise ValueError("gt batch size does not match traj batch size")
            gt = np.broadcast_to(gt, traj.shape)
        out["mse_to_gt"] = ((traj - gt) ** 2).mean(axis=(1, 2), dtype=F32)
    return out
