"""numpy restatement of the elementwise helpers around the two denoisers:
``src/utils/clamp.py``, ``src/utils/normalize.py``, the conf / known-mask helpers of
``src/sample/sample_generate.py`` and the corruption builders of
``src/train/train_interp_levels.py``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
Random draws are replayed from a :class:`NoiseTape` (the reference draws them from a
``torch.Generator`` in a fixed order; the golden generator records that order).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .keyframes_np import interpolate_from_indices, segment_lookup, build_nested_masks_batch

F32 = np.float32


class NoiseTape:
    """Replays recorded generator draws in order; shape-checked."""

    def __init__(self, draws: Sequence[np.ndarray]):
        self.draws = list(draws)
        self.pos = 0

    def take(self, shape: Tuple[int, ...]) -> np.ndarray:
        if self.pos >= len(self.draws):
            raise RuntimeError("noise tape exhausted")
        a = np.asarray(self.draws[self.pos])
        self.pos += 1
        if tuple(a.shape) != tuple(shape):
            raise RuntimeError(f"noise tape shape mismatch: recorded {a.shape}, wanted {shape}")
        return a


# --------------------------------------------------------------------------- #
# utils/clamp.py, utils/normalize.py
# --------------------------------------------------------------------------- #
def apply_clamp(x_hat: np.ndarray, x_ref: np.ndarray, clamp_mask: Optional[np.ndarray], clamp_dims: str) -> np.ndarray:
    """clamp.py:4-10 (in-place on x_hat for 'pos', like the reference)."""
    if clamp_mask is None:
        return x_hat
    m = clamp_mask[..., None]
    if clamp_dims == "pos":
        x_hat[:, :, :2] = np.where(m, x_ref[:, :, :2], x_hat[:, :, :2])
        return x_hat
    return np.where(m, x_ref, x_hat)


def apply_soft_clamp(x_hat: np.ndarray, x_ref: np.ndarray, conf: Optional[np.ndarray], lam: float,
                     clamp_dims: str) -> np.ndarray:
    """clamp.py:13-32"""
    if conf is None or lam <= 0.0:
        return x_hat
    w = conf[..., None] if conf.ndim == 2 else conf
    w = (w.astype(F32) * F32(float(lam))).astype(F32)
    if clamp_dims == "pos":
        x_hat[:, :, :2] = x_hat[:, :, :2] + w * (x_ref[:, :, :2] - x_hat[:, :, :2])
        return x_hat
    return (x_hat + w * (x_ref - x_hat)).astype(F32)


def logit_pos(x: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """normalize.py:4-11"""
    if x.shape[-1] < 2:
        return x
    out = x.astype(F32).copy()
    pos = np.clip(out[..., :2], F32(eps), F32(1.0 - eps))
    out[..., :2] = np.log(pos / (F32(1.0) - pos))
    return out


def sigmoid_pos(x: np.ndarray) -> np.ndarray:
    """normalize.py:14-20"""
    if x.shape[-1] < 2:
        return x
    out = x.astype(F32).copy()
    out[..., :2] = F32(1.0) / (F32(1.0) + np.exp(-out[..., :2]))
    return out


# --------------------------------------------------------------------------- #
# sample_generate.py helpers
# --------------------------------------------------------------------------- #
def build_known_mask_values(idx: np.ndarray, start_goal: Optional[np.ndarray], D: int, T: int,
                            clamp_endpoints: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """sample_generate.py:260-280"""
    B, K = idx.shape
    known_mask = np.zeros((B, K, D), dtype=bool)
    known_values = np.zeros((B, K, D), dtype=F32)
    if clamp_endpoints:
        if start_goal is None:
            raise ValueError("clamp_endpoints=True but start_goal missing from cond")
        if D < 2:
            return known_mask, known_values
        start = start_goal[:, None, :2]
        goal = start_goal[:, None, 2:]
        m_start = (idx == 0)[..., None]
        m_goal = (idx == T - 1)[..., None]
        known_mask[:, :, :2] = m_start | m_goal
        known_values[:, :, :2] = np.where(m_start, start, known_values[:, :, :2])
        known_values[:, :, :2] = np.where(m_goal, goal, known_values[:, :, :2])
    return known_mask, known_values


def build_anchor_conf(mask_s: np.ndarray, student_mask: Optional[np.ndarray], use_student: bool,
                      conf_teacher: float, conf_student: float, conf_endpoints: float,
                      conf_missing: float, clamp_endpoints: bool) -> np.ndarray:
    """sample_generate.py:319-336 (train twin: train_interp_levels.py:546-562 with use_student
    == (student_mask is not None))."""
    conf = np.full(mask_s.shape, F32(conf_missing), dtype=F32)
    conf = np.where(mask_s, F32(conf_teacher), conf)
    if student_mask is not None and use_student:
        conf = np.where(student_mask & mask_s, F32(conf_student), conf)
    conf = conf.astype(F32)
    if clamp_endpoints:
        conf[:, 0] = F32(conf_endpoints)
        conf[:, -1] = F32(conf_endpoints)
    return conf


def soft_clamp_lambda(s: int, levels: int, schedule: str, max_val: float) -> float:
    """sample_generate.py:339-347"""
    if levels <= 0:
        return float(max_val)
    frac = float(s) / float(levels)
    if schedule == "linear":
        return float(max_val) * frac
    if schedule == "cosine":
        return float(max_val) * 0.5 * (1.0 + math.cos(math.pi * (1.0 - frac)))
    return float(max_val)


def anneal_conf(conf: Optional[np.ndarray], s, levels: int, mode: str) -> Optional[np.ndarray]:
    """sample_generate.py:350-360 (scalar s) and train_interp_levels.py:565-576 (per-row s)."""
    if conf is None or mode == "none" or levels <= 0:
        return conf
    if np.isscalar(s):
        frac = float(s) / float(levels)
        if mode == "linear":
            lam = 1.0 - frac
        elif mode == "cosine":
            lam = 0.5 * (1.0 + math.cos(math.pi * frac))
        else:
            lam = 0.0
        return (conf + (F32(1.0) - conf) * F32(float(lam))).astype(F32)
    frac = np.asarray(s).astype(F32) / F32(float(levels))
    if mode == "linear":
        lam = F32(1.0) - frac
    elif mode == "cosine":
        lam = F32(0.5) * (F32(1.0) + np.cos(F32(math.pi) * frac))
    else:
        lam = np.zeros_like(frac)
    lam = lam.astype(F32).reshape(-1, 1)
    return (conf + (F32(1.0) - conf) * lam).astype(F32)


# --------------------------------------------------------------------------- #
# train_interp_levels.py corruption builders
# --------------------------------------------------------------------------- #
def compute_sigma_for_level(K_s: int, K_min: int, sigma_max: float, sigma_min: float, sigma_pow: float) -> float:
    """train_interp_levels.py:386-401"""
    if sigma_max <= 0.0:
        return 0.0
    K_s = max(1, int(K_s))
    K_min = max(1, int(K_min))
    ratio = float(K_min) / float(K_s)
    sigma = float(sigma_max) * (ratio ** float(sigma_pow))
    sigma = min(float(sigma_max), sigma)
    return max(float(sigma_min), sigma)


def compute_jitter_for_level(K_s: int, K_min: int, jitter_max: int, jitter_pow: float) -> int:
    """train_interp_levels.py:433-441"""
    if jitter_max <= 0:
        return 0
    K_s = max(1, int(K_s))
    K_min = max(1, int(K_min))
    ratio = float(K_min) / float(K_s)
    jitter = int(round(float(jitter_max) * (ratio ** float(jitter_pow))))
    return max(0, min(int(jitter_max), jitter))


def distance_alpha(idx: np.ndarray, T: int) -> np.ndarray:
    """train_interp_levels.py:444-455 -> [B,T,1] tent weight."""
    _, left_idx, right_idx = segment_lookup(idx, T)
    t_grid = np.arange(T, dtype=np.int64)[None, :]
    gap = np.maximum(right_idx - left_idx, 1)
    dist = np.minimum(t_grid - left_idx, right_idx - t_grid)
    alpha = np.clip((F32(2.0) * dist.astype(F32)) / gap.astype(F32), F32(0.0), F32(1.0))
    return alpha.astype(F32)[..., None]


def corrupt_from_anchors(source: np.ndarray, idx: np.ndarray, T: int, tape: NoiseTape, sigma: float,
                         anchor_sigma: float, index_jitter: int, index_jitter_prob: float, mode: str,
                         clamp_endpoints: bool, recompute_velocity: bool) -> np.ndarray:
    """train_interp_levels.py:458-510.  Draw order on the tape: [randint jitter, rand use]
    (only if jitter is on), randn(B,K,2) (if anchor_sigma>0), randn(B,T,2) (if sigma>0)."""
    B, _, D = source.shape
    K = idx.shape[1]
    idx_j = idx
    if index_jitter > 0 and index_jitter_prob > 0.0:
        jit = tape.take((B, K)).astype(np.int64) - int(index_jitter)
        use = tape.take((B, K)).astype(F32) < F32(float(index_jitter_prob))
        if clamp_endpoints:
            use = use & ~(idx == 0) & ~(idx == (T - 1))
        idx_j = np.clip(np.where(use, idx + jit, idx), 0, T - 1)
    rows = np.arange(B)[:, None]
    vals = source.astype(F32)[rows, idx_j].copy()
    if anchor_sigma > 0.0:
        noise_vals = np.zeros_like(vals)
        noise_vals[:, :, :2] = tape.take((B, K, 2)).astype(F32) * F32(float(anchor_sigma))
        if clamp_endpoints:
            mask_end = (idx == 0) | (idx == T - 1)
            noise_vals[mask_end] = 0.0
        vals[:, :, :2] = vals[:, :, :2] + noise_vals[:, :, :2]
    x = interpolate_from_indices(idx, vals, T, recompute_velocity=False)
    if sigma > 0.0:
        alpha = distance_alpha(idx, T) if mode == "dist" else F32(1.0)
        noise = np.zeros_like(x)
        noise[:, :, :2] = tape.take((B, T, 2)).astype(F32) * F32(float(sigma))
        x[:, :, :2] = x[:, :, :2] + noise[:, :, :2] * alpha
    if recompute_velocity and D == 4:
        pos = x[:, :, :2]
        v = np.zeros_like(pos)
        dt = F32(1.0 / float(T))
        v[:, :-1] = (pos[:, 1:] - pos[:, :-1]) / dt
        v[:, -1] = 0.0
        x = np.concatenate([pos, v], axis=-1)
    return x.astype(F32)


def build_interp_adjacent_batch(x0: np.ndarray, K_min: int, levels: int, masks_levels: np.ndarray,
                                idx_levels: List[np.ndarray], s_idx: np.ndarray, tape: Optional[NoiseTape] = None,
                                recompute_velocity: bool = False, x0_override: Optional[np.ndarray] = None,
                                corrupt_mode: str = "none", corrupt_sigma_max: float = 0.0,
                                corrupt_sigma_min: float = 0.0, corrupt_sigma_pow: float = 1.0,
                                corrupt_anchor_frac: float = 0.0, corrupt_index_jitter_max: int = 0,
                                corrupt_index_jitter_prob: float = 0.0, corrupt_index_jitter_pow: float = 1.0,
                                clamp_endpoints: bool = True, pos_clip: bool = False,
                                pos_clip_min: float = 0.0, pos_clip_max: float = 1.0):
    """train_interp_levels.py:294-383 with masks / idx / s_idx given (they are drawn earlier
    in the reference step, :896-906 and :1036)."""
    B, T, D = x0.shape
    x_s = np.zeros_like(x0, dtype=F32)
    x_prev = np.zeros_like(x0, dtype=F32)
    mask_s = np.zeros((B, T), dtype=bool)
    mask_prev = np.zeros((B, T), dtype=bool)
    source = (x0_override if x0_override is not None else x0).astype(F32)
    for s in range(1, levels + 1):
        sel = s_idx == s
        if not np.any(sel):
            continue
        idx = idx_levels[s][sel]
        idx_prev = idx_levels[s - 1][sel]
        src = source[sel]
        rows = np.arange(src.shape[0])[:, None]
        if corrupt_mode != "none":
            K_s, K_prev = idx.shape[1], idx_prev.shape[1]
            sig_s = compute_sigma_for_level(K_s, K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
            sig_p = compute_sigma_for_level(K_prev, K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
            jit_s = compute_jitter_for_level(K_s, K_min, corrupt_index_jitter_max, corrupt_index_jitter_pow)
            jit_p = compute_jitter_for_level(K_prev, K_min, corrupt_index_jitter_max, corrupt_index_jitter_pow)
            xs = corrupt_from_anchors(src, idx, T, tape, sig_s, sig_s * float(corrupt_anchor_frac), jit_s,
                                      corrupt_index_jitter_prob, corrupt_mode, clamp_endpoints, recompute_velocity)
            xp = corrupt_from_anchors(src, idx_prev, T, tape, sig_p, sig_p * float(corrupt_anchor_frac), jit_p,
                                      corrupt_index_jitter_prob, corrupt_mode, clamp_endpoints, recompute_velocity)
        else:
            xs = interpolate_from_indices(idx, src[rows, idx], T, recompute_velocity)
            xp = interpolate_from_indices(idx_prev, src[rows, idx_prev], T, recompute_velocity)
        if pos_clip:
            xs[:, :, :2] = np.clip(xs[:, :, :2], F32(pos_clip_min), F32(pos_clip_max))
            xp[:, :, :2] = np.clip(xp[:, :, :2], F32(pos_clip_min), F32(pos_clip_max))
        x_s[sel] = xs
        x_prev[sel] = xp
        mask_s[sel] = masks_levels[sel, s]
        mask_prev[sel] = masks_levels[sel, s - 1]
    return x_s, x_prev, mask_s, mask_prev


def build_interp_level_batch(x0: np.ndarray, K_min: int, levels: int, masks_levels: np.ndarray,
                             idx_levels: List[np.ndarray], s_idx: np.ndarray, tape: Optional[NoiseTape] = None,
                             recompute_velocity: bool = False, corrupt_mode: str = "none",
                             corrupt_sigma_max: float = 0.0, corrupt_sigma_min: float = 0.0,
                             corrupt_sigma_pow: float = 1.0, corrupt_anchor_frac: float = 0.0,
                             clamp_endpoints: bool = True):
    """train_interp_levels.py:227-291 (x0-target variant), masks / idx / s_idx given."""
    B, T, D = x0.shape
    x_s = np.zeros_like(x0, dtype=F32)
    mask_s = np.zeros((B, T), dtype=bool)
    source = x0.astype(F32)
    for s in range(1, levels + 1):
        sel = s_idx == s
        if not np.any(sel):
            continue
        idx = idx_levels[s][sel]
        src = source[sel]
        rows = np.arange(src.shape[0])[:, None]
        if corrupt_mode != "none":
            sig = compute_sigma_for_level(idx.shape[1], K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
            xs = corrupt_from_anchors(src, idx, T, tape, sig, sig * float(corrupt_anchor_frac), 0, 0.0,
                                      corrupt_mode, clamp_endpoints, recompute_velocity)
        else:
            xs = interpolate_from_indices(idx, src[rows, idx], T, recompute_velocity)
        x_s[sel] = xs
        mask_s[sel] = masks_levels[sel, s]
    return x_s, mask_s


def stage2_loss(delta_hat: np.ndarray, target: np.ndarray, weight_mask: np.ndarray,
                w_anchor: float = 0.1, w_missing: float = 1.0) -> float:
    """train_interp_levels.py:1144-1154 (anchor_conf branch): weighted MSE."""
    diff = ((delta_hat.astype(np.float64) - target.astype(np.float64)) ** 2).sum(axis=-1)
    w = float(w_missing) + (float(w_anchor) - float(w_missing)) * weight_mask.astype(np.float64)
    return float((diff * w).sum() / (w.sum() * delta_hat.shape[-1] + 1e-8))
