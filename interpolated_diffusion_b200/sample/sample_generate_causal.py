"""Batched long-horizon causal generation: the chunk loop of the reference's ``src/sample/sample_generate_causal.py:485-583``.

The reference generates one trajectory at a time (B = 1) and, for every chunk, re-runs Stage 1 on the chunk's keypoints and the
causal Stage-2 denoiser over the whole prefix + chunk.  Everything that shapes that loop (``cur``, ``end``, ``local_T``,
``full_len``) depends only on ``T`` and ``chunk``, so a batch of trajectories advances in lockstep and every step is one of the
batched kernels of this package.  The prefix is recomputed each chunk exactly as the reference does: the conditioning changes
from chunk to chunk (``start_goal`` becomes [left, right] of the chunk, :528), so FiLM changes and a KV cache of the prefix would
NOT reproduce the reference.

RNG: the reference draws the chunk's anchors with ``sample_fixed_k_indices_batch(1, ...)`` on its generator sample by sample; here
one batched draw per chunk (same distribution, different stream order); ``idx_chunks`` / ``z_T_chunks`` inject them for parity."""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from .. import _lib as L
from ..corruptions import keyframes as kf
from ..diffusion.schedules import make_alpha_bars, make_beta_schedule
from ..utils.clamp import apply_clamp
from ..utils.normalize import logit_pos, sigmoid_pos
from .sample_generate import _build_known_mask_values, _sample_keypoints_ddim


def _heuristic_right(left: torch.Tensor, goal: torch.Tensor, L_: int, remaining: int) -> torch.Tensor:
    """sample_generate_causal.py:86-88"""
    frac = min(1.0, float(L_) / max(1, remaining))
    return left + frac * (goal - left)


def _maze_embedding(model, cond: Dict[str, torch.Tensor]) -> Optional[torch.Tensor]:
    """maze.fc(mean-pooled conv stack(occ[, sdf])) of the model's conditioning encoder (encoders.py:56-65), or None when the
    encoder is not the built-in MazeConditionEncoder."""
    enc = getattr(model, "cond_enc", None)
    if enc is None or not hasattr(enc, "maze"):
        return None
    x = cond["occ"]
    if enc.use_sdf:
        if cond.get("sdf") is None:
            raise ValueError("use_sdf is True but sdf missing from cond")
        x = torch.cat([x, cond["sdf"]], dim=1)
    enc.maze.precision = getattr(model, "precision", "bf16")
    return enc.maze(x)


def _chunk_cond_vec(model, maze_emb: Optional[torch.Tensor], cond_chunk: Dict[str, torch.Tensor]) -> Optional[torch.Tensor]:
    """cond_vec of a chunk = hoisted maze embedding + start/goal MLP of the chunk's [left, right] (encoders.py:66-70)."""
    if maze_emb is None:
        return None
    enc = model.cond_enc
    emb = maze_emb.clone()
    if enc.use_start_goal:
        enc.sg(cond_chunk["start_goal"], out=emb)
    return emb


def chunk_plan(T: int, chunk: int, K_min: int):
    """(cur, end, local_T, K) per chunk (:504-513)."""
    plan, cur = [], 1
    while cur < T:
        end = min(T - 1, cur + chunk - 1)
        local_T = end - cur + 2
        plan.append((cur, end, local_T, min(K_min, local_T)))
        cur = end + 1
    return plan


@torch.no_grad()
def generate_causal_chunked(kp_model, interp_model, cond: Dict[str, torch.Tensor], *, T: int, chunk: int = 16, K_min: int = 8,
                            levels: int = 3, data_dim: int = 2, ddim_steps: int = 20, n_train: int = 1000,
                            beta_schedule: str = "cosine", logit_space: bool = False, logit_eps: float = 1e-5,
                            recompute_vel: bool = True, clamp_endpoints: bool = True, clamp_policy: str = "endpoints",
                            clamp_dims: str = "pos", generator: Optional[torch.Generator] = None,
                            idx_chunks: Optional[List[torch.Tensor]] = None, z_T_chunks: Optional[List[torch.Tensor]] = None,
                            schedule: Optional[Dict[str, torch.Tensor]] = None) -> torch.Tensor:
    """-> x_gen fp32 [B, T, data_dim].  Defaults are the reference CLI's (:27-79)."""
    if clamp_policy not in ("none", "endpoints", "all_anchors"):
        raise ValueError(f"Unknown clamp_policy: {clamp_policy}")
    sg = L.f32c(cond["start_goal"])
    dev = L.require_cuda(sg, cond["occ"])
    B, D = sg.shape[0], data_dim
    if schedule is None:
        schedule = {k: v.to(dev) for k, v in make_alpha_bars(make_beta_schedule(beta_schedule, n_train)).items()}
    start, goal = sg[:, :2], sg[:, 2:]
    x_gen = torch.zeros((B, T, D), device=dev, dtype=torch.float32)
    x_gen[:, 0, :2] = start
    s_level = torch.full((B,), levels, device=dev, dtype=torch.long)
    # the maze part of the conditioning does not change from chunk to chunk (only start_goal does): encode it once per model
    maze_emb = [_maze_embedding(m, cond) for m in (kp_model, interp_model)]
    for c, (cur, end, local_T, K) in enumerate(chunk_plan(T, chunk, K_min)):
        L_ = end - cur + 1
        left = x_gen[:, cur - 1, :2].contiguous()
        right = goal if end == T - 1 else _heuristic_right(left, goal, L_, T - cur)
        if idx_chunks is not None:
            idx_local = L.i64c(idx_chunks[c])
            mask_local = torch.zeros((B, local_T), device=dev, dtype=torch.bool).scatter_(1, idx_local, True)
        else:
            idx_local, mask_local = kf.sample_fixed_k_indices_batch(B, local_T, K, generator=generator, device=dev, ensure_endpoints=True)
        cond_chunk = dict(cond)
        cond_chunk["start_goal"] = torch.cat([left, right], dim=1).contiguous()
        # endpoint tokens known = [left, right] of the chunk (:514-526): the known-mask kernel with the chunk's start_goal
        known_mask, known_values = _build_known_mask_values(idx_local, cond_chunk, D, local_T, clamp_endpoints)
        if logit_space:
            known_values = logit_pos(known_values, eps=logit_eps)
        cv_kp, cv_il = (_chunk_cond_vec(m, e, cond_chunk) for m, e in zip((kp_model, interp_model), maze_emb))
        z_hat = _sample_keypoints_ddim(kp_model, schedule, idx_local, known_mask, known_values, cond_chunk, ddim_steps, local_T,
                                       z_T=None if z_T_chunks is None else z_T_chunks[c], cond_vec=cv_kp)
        if logit_space:
            z_hat = sigmoid_pos(z_hat)
        x_s = kf.interpolate_from_indices(idx_local, z_hat, local_T, recompute_velocity=recompute_vel)
        full_len = end + 1
        x_full = torch.zeros((B, full_len, D), device=dev, dtype=torch.float32)
        mask_full = torch.zeros((B, full_len), device=dev, dtype=torch.bool)
        if cur > 1:
            x_full[:, :cur - 1] = x_gen[:, :cur - 1]
            mask_full[:, :cur - 1] = True
        x_full[:, cur - 1:full_len] = x_s
        mask_full[:, cur - 1:full_len] = mask_local
        delta_hat = interp_model(x_full, s_level, mask_full, cond_chunk, cond_vec=cv_il)
        x_hat = x_full + delta_hat
        if clamp_policy == "all_anchors":
            clamp_mask = mask_full
        elif clamp_policy == "endpoints":
            clamp_mask = torch.zeros_like(mask_full)
            clamp_mask[:, cur - 1] = True
            clamp_mask[:, full_len - 1] = True
        else:
            clamp_mask = None
        if clamp_mask is not None:
            x_hat = apply_clamp(x_hat, x_full, clamp_mask, clamp_dims)
        x_gen[:, cur:end + 1, :2] = x_hat[:, cur:end + 1, :2]
        if D > 2 and recompute_vel:
            x_gen[:, cur:end + 1, 2:] = x_hat[:, cur:end + 1, 2:]
    if D > 2 and recompute_vel:                                   # :632-638
        pos = x_gen[:, :, :2]
        v = torch.zeros_like(pos)
        v[:, :-1] = (pos[:, 1:] - pos[:, :-1]) / (1.0 / float(T))
        x_gen = torch.cat([pos, v], dim=-1)
    return x_gen
