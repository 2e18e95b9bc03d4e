"""Mirror of the reference's ``src/eval/metrics.py`` batched metrics (``compute_metrics_batch`` :68-128,
``compute_metrics`` :131-139) on the sm_100a kernel ``idb200_traj_metrics``: one launch for the whole batch instead of
the per-sample host loop of ``sample_generate.py:1323-1398``.  Same names, argument order, broadcasting rules,
``ValueError`` messages and result keys; CUDA tensors only (no CPU fallback)."""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from .. import _lib as L


def _as_cuda_f32(x, device) -> torch.Tensor:
    t = x if torch.is_tensor(x) else torch.as_tensor(x)
    return L.f32c(t.to(device))


@torch.no_grad()
def compute_metrics_batch(occ, traj, goal, gt: Optional[object] = None) -> Dict[str, torch.Tensor]:
    if not torch.is_tensor(traj):
        traj = torch.as_tensor(traj)
    dev = L.require_cuda(traj)
    traj_t = L.f32c(traj)
    occ_t = _as_cuda_f32(occ, dev)
    goal_t = _as_cuda_f32(goal, dev)
    if traj_t.dim() == 2:
        traj_t = traj_t.unsqueeze(0)
    if occ_t.dim() == 2:
        occ_t = occ_t.unsqueeze(0)
    if goal_t.dim() == 1:
        goal_t = goal_t.unsqueeze(0)
    B, T, D = traj_t.shape
    if occ_t.shape[0] != B and occ_t.shape[0] != 1:
        raise ValueError("occ batch size does not match traj batch size")
    if goal_t.shape[0] != B and goal_t.shape[0] != 1:
        raise ValueError("goal batch size does not match traj batch size")
    gt_t = None
    if gt is not None:
        gt_t = _as_cuda_f32(gt, dev)
        if gt_t.dim() == 2:
            gt_t = gt_t.unsqueeze(0)
        if gt_t.shape[0] != B and gt_t.shape[0] != 1:
            raise ValueError("gt batch size does not match traj batch size")
    H, W = occ_t.shape[-2:]
    n_out = 6 if gt_t is not None else 5
    out = torch.empty((B, n_out), device=dev, dtype=torch.float32)
    L.call("idb200_traj_metrics", occ_t.data_ptr(), 0 if occ_t.shape[0] == 1 and B != 1 else H * W, traj_t.data_ptr(),
           goal_t.data_ptr(), 0 if goal_t.shape[0] == 1 and B != 1 else goal_t.shape[1], L.ptr(gt_t),
           0 if gt_t is None or (gt_t.shape[0] == 1 and B != 1) else T * D, B, T, D, H, W, float(np.float32(1.0 / float(W))),
           out.data_ptr(), n_out, L.stream(dev))
    res = {"collision_rate": out[:, 0], "goal_dist": out[:, 1], "success": out[:, 2], "path_length": out[:, 3], "smoothness": out[:, 4]}
    if gt_t is not None:
        res["mse_to_gt"] = out[:, 5]
    return res


def compute_metrics(occ, traj, goal, gt: Optional[object] = None) -> Dict[str, float]:
    batch = compute_metrics_batch(occ, traj, goal, gt)
    return {k: float(v[0].item()) for k, v in batch.items()}
