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


def generate_causal_chunked(sd_kp, sd_interp, n_heads: int, cond: Dict[str, torch.Tensor], *, T: int, chunk: int, K_min: int,
                            levels: int, idx_chunks: List[np.ndarray], z_T_chunks: List[np.ndarray], D: int = 2,
                            ddim_steps: int = 20, n_train: int = 1000, beta_schedule: str = "cosine", logit_space: bool = False,
                            logit_eps: float = 1e-5, recompute_vel: bool = True, clamp_endpoints: bool = True,
                            clamp_policy: str = "endpoints", clamp_dims: str = "pos") -> np.ndarray:
    """The chunk loop of ``src/sample/sample_generate_causal.py:485-583`` (long-horizon causal generation), restated for a
    batch: the reference runs it sample by sample (B = 1) but every quantity that shapes the loop (cur, end, local_T, full_len)
    depends only on T and chunk, so the samples advance in lockstep with identical per-sample arithmetic.  The random anchors
    of each chunk (``sample_fixed_k_indices_batch``, :511) and the DDIM noise (:194) are inputs.  PARITY: pinned against a run of the
    live script's own ``main()`` (``tests/golden/make_golden_r2.py::gold_causal_chunks``: tiny checkpoints, the particle-maze
    samples handed over as a prepared dataset.npz, plotting stubbed, every draw recorded) by
    ``tests/test_oracle_golden_r2.py::test_causal_chunk_loop_against_live_main``."""
    B = cond["start_goal"].shape[0]
    schedule = df.make_alpha_bars(df.make_beta_schedule(beta_schedule, n_train))
    sg = cond["start_goal"].numpy().astype(F32)
    start, goal = sg[:, :2], sg[:, 2:]
    x_gen = np.zeros((B, T, D), F32)
    x_gen[:, 0, :2] = start                                                    # :493-496
    cur, c = 1, 0
    while cur < T:                                                             # :504
        end = min(T - 1, cur + chunk - 1)
        L = end - cur + 1
        remaining = T - cur
        left = x_gen[:, cur - 1, :2].copy()
        if end == T - 1:
            right = goal.copy()
        else:                                                                  # _heuristic_right, :86-88
            frac = F32(min(1.0, float(L) / max(1, remaining)))
            right = (left + frac * (goal - left)).astype(F32)
        local_T = L + 1
        idx = idx_chunks[c].astype(np.int64)
        mask_local = kf._mask_from_idx(idx, local_T)
        K = idx.shape[1]
        known_mask = np.zeros((B, K, D), bool)
        known_values = np.zeros((B, K, D), F32)
        if clamp_endpoints:                                                    # :516-524
            first, last = (idx == 0)[..., None], (idx == local_T - 1)[..., None]
            known_mask[:, :, :2] = first | last
            known_values[:, :, :2] = np.where(first, left[:, None, :], known_values[:, :, :2])
            known_values[:, :, :2] = np.where(last, right[:, None, :], known_values[:, :, :2])
        if logit_space:
            known_values = sp.logit_pos(known_values, eps=logit_eps)
        cond_chunk = dict(cond)
        cond_chunk["start_goal"] = _t(np.concatenate([left, right], axis=1).astype(F32))          # :528-529
        z_hat = sample_keypoints_ddim(sd_kp, n_heads, schedule, idx, known_mask, known_values, cond_chunk, ddim_steps, local_T,
                                      z_T_chunks[c], schedule_name="linear")
        if logit_space:
            z_hat = sp.sigmoid_pos(z_hat)
        x_s = kf.interpolate_from_indices(idx, z_hat, local_T, recompute_velocity=recompute_vel)   # :555
        full_len = end + 1                                                      # :558-566
        x_full = np.zeros((B, full_len, D), F32)
        mask_full = np.zeros((B, full_len), bool)
        if cur > 1:
            x_full[:, :cur - 1] = x_gen[:, :cur - 1]
            mask_full[:, :cur - 1] = True
        x_full[:, cur - 1:full_len] = x_s
        mask_full[:, cur - 1:full_len] = mask_local
        delta = dn.interp_level_denoiser(sd_interp, n_heads, _t(x_full), torch.full((B,), levels), _t(mask_full), cond_chunk,
                                         causal=True).numpy()
        x_hat = (x_full + delta).astype(F32)
        if clamp_policy == "all_anchors":
            clamp_mask = mask_full
        elif clamp_policy == "endpoints":
            clamp_mask = np.zeros_like(mask_full)
            clamp_mask[:, cur - 1] = True
            clamp_mask[:, full_len - 1] = True
        else:
            clamp_mask = None
        if clamp_mask is not None:
            x_hat = sp.apply_clamp(x_hat, x_full, clamp_mask, clamp_dims)
        x_gen[:, cur:end + 1, :2] = x_hat[:, cur:end + 1, :2]                   # :583-585
        if D > 2 and recompute_vel:
            x_gen[:, cur:end + 1, 2:] = x_hat[:, cur:end + 1, 2:]
        cur = end + 1
        c += 1
    if D > 2 and recompute_vel:                                                 # :632-638
        pos = x_gen[:, :, :2]
        v = np.zeros_like(pos)
        v[:, :-1] = (pos[:, 1:] - pos[:, :-1]) / F32(1.0 / float(T))
        x_gen = np.concatenate([pos, v], axis=-1).astype(F32)
    return x_gen


def causal_chunk_plan(T: int, chunk: int, K_min: int):
    """(cur, end, local_T, K) of every chunk of the loop above."""
    plan, cur = [], 1
    while cur < T:
        end = min(T - 1, cur + chunk - 1)
        local_T = end - cur + 2
        plan.append((cur, end, local_T, min(K_min, local_T)))
        cur = end + 1
    return plan
