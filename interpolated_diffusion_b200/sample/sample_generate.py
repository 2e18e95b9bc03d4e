"""Mirror of the hot-path functions of the reference's ``src/sample/sample_generate.py`` (lines 260-404
and the per-batch body 944-1285) on libidb200 kernels.  CLI, checkpoint reconciliation, plotting and
per-sample metric loops of the reference file are out of scope."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .. import _lib as L
from ..diffusion.ddpm import _timesteps, ddim_step_scalar

ANNEAL = {"none": 0, "linear": 1, "cosine": 2}


def _build_known_mask_values(idx: torch.Tensor, cond: dict, D: int, T: int, clamp_endpoints: bool = True, *,
                             logit_space: bool = False, logit_eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """sample_generate.py:260-280 -> (known_mask bool [B,K,D], known_values fp32 [B,K,D]).
    ``logit_space=True`` also applies ``logit_pos`` (sample_generate.py:1021-1022) in the same launch."""
    dev = L.require_cuda(idx)
    B, K = idx.shape
    sg = None
    if clamp_endpoints:
        if "start_goal" not in cond:
            raise ValueError("clamp_endpoints=True but start_goal missing from cond")
        sg = L.f32c(cond["start_goal"])
        L.require_cuda(sg)
    known_mask = torch.empty((B, K, D), device=dev, dtype=torch.bool)
    known_values = torch.empty((B, K, D), device=dev, dtype=torch.float32)
    L.call("idb200_known_mask_values", L.ptr(L.i64c(idx)), L.ptr(sg), B, K, D, T, int(bool(clamp_endpoints)),
           int(bool(logit_space)), float(logit_eps), L.ptr(known_mask), L.ptr(known_values), L.stream(dev))
    return known_mask, known_values


def _kp_feat_from_idx(idx: torch.Tensor, T: int, kp_feat_dim: int, left_diff: Optional[torch.Tensor] = None,
                      right_diff: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sample_generate.py:283-307: [left_gap, right_gap, t_norm(, left_diff, right_diff)] per keypoint."""
    B, K = idx.shape
    feat = torch.zeros((B, K, kp_feat_dim), device=idx.device, dtype=torch.float32)
    if kp_feat_dim <= 0:
        return feat
    denom = float(max(1, T - 1))
    if K > 1:
        gaps = (idx[:, 1:] - idx[:, :-1]).float() / denom
        feat[:, 1:, 0] = gaps
        feat[:, :-1, 1] = gaps
    if kp_feat_dim >= 3:
        feat[:, :, 2] = idx.float() / denom
    if kp_feat_dim >= 5 and left_diff is not None and right_diff is not None:
        feat[:, :, 3] = left_diff
        feat[:, :, 4] = right_diff
    return feat


def anchor_conf_mask_in(mask_s: torch.Tensor, student_mask: Optional[torch.Tensor], mask_prev: Optional[torch.Tensor],
                        s, levels: int, anneal_mode: str, conf_teacher: float, conf_student: float,
                        conf_endpoints: float, conf_missing: float, clamp_endpoints: bool, *, want_conf: bool = True,
                        channels: int = 0):
    """``idb200_anchor_conf``: conf [B,T] (+ optional anneal) and, if ``channels`` in {2,3}, the stacked
    ``mask_in`` of sample_generate.py:1163-1178 / :1255-1256 in one launch.  ``s`` is an int or int64 [B]."""
    dev = L.require_cuda(mask_s, student_mask, mask_prev)
    B, T = mask_s.shape
    ms = L.u8c(mask_s)
    st = L.u8c(student_mask) if student_mask is not None else None
    mp = L.u8c(mask_prev) if mask_prev is not None else None
    conf = torch.empty((B, T), device=dev, dtype=torch.float32) if want_conf else None
    mask_in = torch.empty((B, T, channels), device=dev, dtype=torch.float32) if channels else None
    s_row, s_scalar = (L.i64c(s), 0) if isinstance(s, torch.Tensor) else (None, int(s))
    L.call("idb200_anchor_conf", L.ptr(ms), L.ptr(st), L.ptr(mp), L.ptr(s_row), s_scalar, int(levels),
           ANNEAL.get(anneal_mode, 0) if levels > 0 else 0, float(conf_teacher), float(conf_student),
           float(conf_endpoints), float(conf_missing), int(bool(clamp_endpoints)), B, T, max(channels, 2), L.ptr(conf),
           L.ptr(mask_in), L.stream(dev))
    return conf, mask_in


def _build_anchor_conf(mask_s: torch.Tensor, student_mask: Optional[torch.Tensor], use_student: bool,
                       conf_teacher: float, conf_student: float, conf_endpoints: float, conf_missing: float,
                       clamp_endpoints: bool) -> torch.Tensor:
    """sample_generate.py:319-336"""
    st = student_mask if (student_mask is not None and use_student) else None
    conf, _ = anchor_conf_mask_in(mask_s, st, None, 0, 0, "none", conf_teacher, conf_student, conf_endpoints,
                                  conf_missing, clamp_endpoints)
    return conf


def _soft_clamp_lambda(s: int, levels: int, schedule: str, max_val: float) -> float:
    """sample_generate.py:339-347"""
    if levels <= 0:
        return float(max_val)
    frac = float(s) / float(levels)
    if schedule == "linear":
        return float(max_val) * frac
    if schedule == "cosine":
        return float(max_val) * 0.5 * (1.0 + math.cos(math.pi * (1.0 - frac)))
    return float(max_val)


def _anneal_conf(conf: torch.Tensor, s: int, levels: int, mode: str) -> torch.Tensor:
    """sample_generate.py:350-360 (a scalar-lambda axpy on a [B,T] tensor; the fused form is
    ``anchor_conf_mask_in``)."""
    if conf is None or mode == "none" or levels <= 0:
        return conf
    frac = float(s) / float(levels)
    if mode == "linear":
        lam = 1.0 - frac
    elif mode == "cosine":
        lam = 0.5 * (1.0 + math.cos(math.pi * frac))
    else:
        lam = 0.0
    return conf + (1.0 - conf) * float(lam)


def _compute_sigma_for_level(K_s: int, K_min: int, sigma_max: float, sigma_min: float, sigma_pow: float) -> float:
    """train_interp_levels.py:386-401 (imported by the sampler for s2 noise)."""
    if sigma_max <= 0.0:
        return 0.0
    K_s = max(1, int(K_s))
    K_min = max(1, int(K_min))
    ratio = float(K_min) / float(K_s)
    sigma = float(sigma_max) * (ratio ** float(sigma_pow))
    sigma = min(float(sigma_max), sigma)
    return max(float(sigma_min), sigma)


def _sample_keypoints_ddim(model, schedule, idx: torch.Tensor, known_mask: torch.Tensor, known_values: torch.Tensor,
                           cond: dict, steps: int, T: int, schedule_name: str = "linear",
                           return_intermediates: bool = False, pos_clip: bool = False, pos_clip_min: float = 0.0,
                           pos_clip_max: float = 1.0, *, z_T: Optional[torch.Tensor] = None):
    """sample_generate.py:363-404.  The DDIM update and the known-value ``torch.where`` (:397-399) are one
    launch per step with the step's two alpha-bar entries as kernel arguments (no per-row table gather,
    no host sync).  ``z_T=`` injects the initial noise (the reference draws it from the global RNG, :389)."""
    dev = L.require_cuda(idx, known_mask, known_values)
    B, K = idx.shape
    D = known_values.shape[-1]
    alpha_bar = schedule["alpha_bar"].detach().to("cpu", torch.float32)
    n_train = alpha_bar.shape[0]
    times = _timesteps(n_train, steps, schedule=schedule_name).tolist()
    z = torch.randn((B, K, D), device=dev) if z_T is None else L.f32c(z_T).clone()
    z = torch.where(known_mask, known_values, z)
    if pos_clip:
        z[..., :2] = z[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
    intermediates = [z.detach().clone()] if return_intermediates else None
    for i in range(len(times) - 1):
        t = torch.full((B,), int(times[i]), device=dev, dtype=torch.long)
        eps = model(z, t, idx, known_mask, cond, T)
        z = ddim_step_scalar(z, eps, float(alpha_bar[times[i]]), float(alpha_bar[times[i + 1]]), known_mask=known_mask,
                             known_values=known_values, pos_clip=pos_clip, pos_clip_min=pos_clip_min,
                             pos_clip_max=pos_clip_max)
        if return_intermediates:
            intermediates.append(z.detach().clone())
    if return_intermediates:
        return z, intermediates
    return z
