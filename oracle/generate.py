"""CPU restatement of the generation hot loop of ``src/sample/sample_generate.py``
(lines 363-404 and 944-1285; the exact recipe is SURVEY.md section 3.5).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

``z_T`` (the ``torch.randn`` at sample_generate.py:389) and, for ``adj`` mode, the nested
masks are inputs so that both sides of a parity test see identical randomness.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch

from . import denoiser_torch as dn
from . import diffusion_np as df
from . import keyframes_np as kf
from . import sampling_np as sp

F32 = np.float32


def _t(a) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a))


def sample_keypoints_ddim(sd_kp, n_heads: int, schedule: Dict[str, np.ndarray], idx: np.ndarray,
                          known_mask: np.ndarray, known_values: np.ndarray, cond: Dict[str, torch.Tensor],
                          steps: int, T: int, z_T: np.ndarray, schedule_name: str = "linear",
                          return_intermediates: bool = False, pos_clip: bool = False,
                          pos_clip_min: float = 0.0, pos_clip_max: float = 1.0,
                          teacher_forced: Optional[List[np.ndarray]] = None):
    """sample_generate.py:363-404.  ``teacher_forced[i]`` (optional) replaces z before eval i:
    the per-step parity protocol of SURVEY 7.3-1 (DDIM step 1 amplifies by ~3243)."""
    B, K = idx.shape
    n_train = schedule["alpha_bar"].shape[0]
    times = df.timesteps(n_train, steps, schedule=schedule_name)

    def clip(z):
        if pos_clip:
            z[..., :2] = np.clip(z[..., :2], F32(pos_clip_min), F32(pos_clip_max))
        return z

    z = np.where(known_mask, known_values, z_T.astype(F32)).astype(F32)
    z = clip(z)
    inter = [z.copy()] if return_intermediates else None
    cond_vec = dn.cond_encoder(sd_kp, cond)             # loop-invariant (reference recomputes it, :396)
    eps_list = []
    for i in range(len(times) - 1):
        if teacher_forced is not None:
            z = teacher_forced[i].astype(F32).copy()
        t = np.full((B,), int(times[i]), dtype=np.int64)
        t_prev = np.full((B,), int(times[i + 1]), dtype=np.int64)
        eps = dn.keypoint_denoiser(sd_kp, n_heads, _t(z), _t(t), _t(idx), _t(known_mask), cond, T,
                                   cond_vec=cond_vec).numpy()
        eps_list.append(eps)
        z = df.ddim_step(z, eps, t, t_prev, schedule, eta=0.0)
        z = np.where(known_mask, known_values, z).astype(F32)
        z = clip(z)
        if return_intermediates:
            inter.append(z.copy())
    if return_intermediates:
        return z, inter, eps_list
    return z


def generate(sd_kp, sd_interp, n_heads: int, cond: Dict[str, torch.Tensor], z_T: np.ndarray, *, T: int,
             K_min: int, levels: int, D: int = 2, ddim_steps: int = 20, ddim_schedule: str = "quadratic",
             n_train: int = 1000, beta_schedule: str = "cosine", stage2_mode: str = "x0",
             clamp_policy: str = "endpoints", clamp_dims: str = "pos", logit_space: bool = True,
             logit_eps: float = 1e-5, anchor_conf: bool = True, conf_teacher: float = 0.95,
             conf_student: float = 0.5, conf_endpoints: float = 1.0, conf_missing: float = 0.0,
             anneal_mode: str = "linear", soft_anchor_clamp: bool = True, soft_clamp_schedule: str = "linear",
             soft_clamp_max: float = 1.0, recompute_vel: bool = True, clamp_endpoints: bool = True,
             idx: Optional[np.ndarray] = None, masks_levels: Optional[np.ndarray] = None,
             causal: bool = False) -> Dict[str, np.ndarray]:
    """SURVEY.md 3.5 steps 1-10 (sample_generate.py:974-1285), kp_index_mode=uniform unless idx given."""
    B = z_T.shape[0]
    schedule = df.make_alpha_bars(df.make_beta_schedule(beta_schedule, n_train))
    if idx is None:
        idx, masks = kf.sample_fixed_k_indices_uniform_batch(B, T, K_min)
    else:
        masks = kf._mask_from_idx(idx, T)
    sg = cond["start_goal"].numpy()
    known_mask, known_values = sp.build_known_mask_values(idx, sg, D, T, clamp_endpoints)
    if logit_space:
        known_values = sp.logit_pos(known_values, eps=logit_eps)
    z = sample_keypoints_ddim(sd_kp, n_heads, schedule, idx, known_mask, known_values, cond, ddim_steps, T,
                              z_T, schedule_name=ddim_schedule)
    z_pred = sp.sigmoid_pos(z) if logit_space else z
    x_pred = kf.interpolate_from_indices(idx, z_pred, T, recompute_velocity=recompute_vel)
    conf_pred = None
    if anchor_conf:
        conf_pred = sp.build_anchor_conf(masks, masks, True, conf_teacher, conf_student, conf_endpoints,
                                         conf_missing, clamp_endpoints)
    out = {"idx": idx, "masks": masks, "z": z, "z_pred": z_pred, "x_pred": x_pred}
    cond_vec2 = dn.cond_encoder(sd_interp, cond)

    def run_interp(x, s_int, mask_in):
        s = torch.full((B,), s_int, dtype=torch.long)
        return dn.interp_level_denoiser(sd_interp, n_heads, _t(x), s, _t(mask_in), cond, causal=causal,
                                        cond_vec=cond_vec2).numpy()

    def policy_mask(m):
        if clamp_policy == "all_anchors":
            return m
        if clamp_policy == "endpoints":
            c = np.zeros_like(m)
            c[:, 0] = True
            c[:, -1] = True
            return c
        return None

    if stage2_mode == "adj":                                   # sample_generate.py:1160-1204
        x_curr = x_pred
        for s in range(levels, 0, -1):
            m_s, m_prev = masks_levels[:, s], masks_levels[:, s - 1]
            if anchor_conf:
                conf_s = sp.build_anchor_conf(m_s, None, False, conf_teacher, conf_student, conf_endpoints,
                                              conf_missing, clamp_endpoints)
                conf_s = sp.anneal_conf(conf_s, s, levels, anneal_mode)
                mask_in = np.stack([m_s.astype(F32), m_prev.astype(F32), conf_s], axis=-1)
            else:
                conf_s = None
                mask_in = np.stack([m_s, m_prev], axis=-1)
            x_curr = (x_curr + run_interp(x_curr, s, mask_in)).astype(F32)
            if soft_anchor_clamp and conf_s is not None:
                lam = sp.soft_clamp_lambda(s, levels, soft_clamp_schedule, soft_clamp_max)
                x_curr = sp.apply_soft_clamp(x_curr, x_pred, conf_s, lam, clamp_dims)
            cm = policy_mask(m_s)
            if cm is not None:
                x_curr = sp.apply_clamp(x_curr, x_pred, cm, clamp_dims)
        out["x_hat"] = x_curr
        return out

    # "x0" one-step jump, sample_generate.py:1252-1285
    if anchor_conf and conf_pred is not None:
        conf_s = sp.anneal_conf(conf_pred, levels, levels, anneal_mode)
        mask_in = np.stack([masks.astype(F32), conf_s], axis=-1)
    else:
        mask_in = masks
    delta = run_interp(x_pred, levels, mask_in)
    x_hat = (x_pred + delta).astype(F32)
    out["delta"] = delta
    if soft_anchor_clamp and conf_pred is not None:
        lam = sp.soft_clamp_lambda(levels, levels, soft_clamp_schedule, soft_clamp_max)
        x_hat = sp.apply_soft_clamp(x_hat, x_pred, conf_pred, lam, clamp_dims)
    cm = policy_mask(masks)
    if cm is not None:
        x_hat = sp.apply_clamp(x_hat, x_pred, cm, clamp_dims)
    out["x_hat"] = x_hat
    return out
