"""numpy restatement of ``src/corruptions/keyframes.py`` (reference).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.

Random draws are *inputs* here (``scores``, ``u``, ``perms``): the reference
draws them with ``torch.rand(..., generator=gen)`` and everything after the
draw is deterministic integer / fp32 arithmetic, restated below op by op (each
numpy fp32 op is individually rounded, like the reference's eager torch ops).
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence, Tuple

import numpy as np

F32 = np.float32


# --------------------------------------------------------------------------- #
# K schedule  (keyframes.py:135-169)
# --------------------------------------------------------------------------- #
def compute_k_schedule(T: int, K_min: int, levels: int, schedule: str = "doubling",
                       geom_gamma: Optional[float] = None) -> List[int]:
    """keyframes.py:135-169.  Python ``round`` (banker's) is part of the contract."""
    K_min = min(K_min, T)
    K_list = [0 for _ in range(levels + 1)]
    K_list[levels] = K_min
    if levels <= 0:
        return K_list
    if schedule == "doubling":
        for s in range(levels, 0, -1):
            K_list[s - 1] = min(T, max(K_list[s] + 1, 2 * K_list[s]))
        return K_list
    if schedule == "linear":
        for s in range(levels - 1, -1, -1):
            frac = float(levels - s) / float(levels)
            target = int(round(K_min + frac * (T - K_min)))
            K_list[s] = min(T, max(K_list[s + 1] + 1, target))
        return K_list
    if schedule == "geom":
        if geom_gamma is None:
            geom_gamma = (float(T) / float(K_min)) ** (1.0 / float(levels)) if K_min > 0 else 1.0
        for s in range(levels - 1, -1, -1):
            target = int(round(K_min * (geom_gamma ** float(levels - s))))
            K_list[s] = min(T, max(K_list[s + 1] + 1, target))
        return K_list
    raise ValueError(f"Unknown k schedule: {schedule}")


def _check_tk(T: int, K: int, ensure_endpoints: bool) -> None:
    # keyframes.py:51-59 (same messages)
    if T <= 0:
        raise ValueError("T must be positive")
    if K <= 0:
        raise ValueError("K must be positive")
    if ensure_endpoints:
        if T < 2:
            raise ValueError("T must be >= 2 when ensure_endpoints is True")
        if K < 2:
            raise ValueError("K must be >= 2 when ensure_endpoints is True")


def stable_rank(scores: np.ndarray) -> np.ndarray:
    """rank[b, j] = position of element j in the stable ascending argsort of row b.

    Equivalent closed form used by the CUDA kernel:
    rank(j) = #{u : s_u < s_j  or  (s_u == s_j and u < j)}.
    """
    perm = np.argsort(scores, axis=1, kind="stable")
    rank = np.empty_like(perm)
    B, n = scores.shape
    rank[np.arange(B)[:, None], perm] = np.arange(n)[None, :]
    return rank


def _mask_from_idx(idx: np.ndarray, T: int) -> np.ndarray:
    B = idx.shape[0]
    mask = np.zeros((B, T), dtype=bool)
    mask[np.arange(B)[:, None], idx] = True
    return mask


# --------------------------------------------------------------------------- #
# fixed-K index sampling  (keyframes.py:42-81, 84-132)
# --------------------------------------------------------------------------- #
def sample_fixed_k_indices_batch(scores: Optional[np.ndarray], B: int, T: int, K: int,
                                 ensure_endpoints: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """keyframes.py:42-81 with ``scores`` = the ``torch.rand((B, T-2))`` draw at :61
    (``(B, T)`` at :74 when ``ensure_endpoints`` is False)."""
    _check_tk(T, K, ensure_endpoints)
    K = min(K, T)
    if ensure_endpoints and T > 2 and K > 2:
        perm = np.argsort(scores, axis=1, kind="stable")
        chosen = perm[:, : K - 2] + 1
        idx = np.concatenate([np.zeros((B, 1), np.int64), chosen.astype(np.int64),
                              np.full((B, 1), T - 1, np.int64)], axis=1)
    elif ensure_endpoints:
        idx = np.concatenate([np.zeros((B, 1), np.int64), np.full((B, 1), T - 1, np.int64)], axis=1)
    else:
        perm = np.argsort(scores, axis=1, kind="stable")
        idx = perm[:, :K].astype(np.int64)
    idx = np.sort(idx, axis=1)
    return idx, _mask_from_idx(idx, T)


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    """fp32 ``torch.linspace``.  ATen's CPU kernel evaluates the symmetric two-sided formula
    (``start + step*i`` below the midpoint, ``end - step*(steps-1-i)`` above) in SIMD chunks,
    so its last bit depends on the host's vector width; ``_timesteps``' ``.long()`` truncation
    (ddpm.py:85-91) and ``round`` (keyframes.py:118) are sensitive to that bit.  The primitive
    itself is therefore the definition and the oracle calls it rather than imitating it."""
    import torch
    return torch.linspace(float(start), float(end), int(steps), dtype=torch.float32).numpy().copy()


def sample_fixed_k_indices_uniform_batch(B: int, T: int, K: int, ensure_endpoints: bool = True,
                                         jitter: float = 0.0,
                                         u: Optional[np.ndarray] = None) -> Tuple[np.ndarray, np.ndarray]:
    """keyframes.py:84-132; ``u`` = the ``torch.rand((B, K))`` draw at :109 (jitter only)."""
    _check_tk(T, K, ensure_endpoints)
    K = min(K, T)
    base = _linspace_f32(0.0, float(T - 1), K)
    if jitter and K > 2 and T > 2:
        spacing = float(T - 1) / float(K - 1)
        max_jitter = spacing * float(jitter) * 0.5
        # (rand - 0.5) * 2.0 * max_jitter : three fp32 tensor-scalar ops (:109)
        noise = ((u.astype(F32) - F32(0.5)) * F32(2.0)) * F32(max_jitter)
        noise[:, 0] = 0.0
        noise[:, -1] = 0.0
        pos = base[None, :] + noise
    else:
        pos = np.broadcast_to(base[None, :], (B, K))
    idx = np.rint(pos).astype(np.int64)          # torch.round = half-to-even
    idx = np.clip(idx, 0, T - 1)
    if ensure_endpoints and K >= 2:
        idx[:, 0] = 0
        idx[:, -1] = T - 1
    for k in range(1, K):                        # :122-125 strictly-increasing fix-ups
        idx[:, k] = np.maximum(idx[:, k], idx[:, k - 1] + 1)
    for k in range(K - 2, -1, -1):
        idx[:, k] = np.minimum(idx[:, k], idx[:, k + 1] - 1)
    idx = np.clip(idx, 0, T - 1)
    if ensure_endpoints and K >= 2:
        idx[:, 0] = 0
        idx[:, -1] = T - 1
    return idx, _mask_from_idx(idx, T)


# --------------------------------------------------------------------------- #
# nested masks  (keyframes.py:172-209, 212-257, 260-294, 297-345)
# --------------------------------------------------------------------------- #
def build_nested_masks_batch(scores: np.ndarray, T: int, K_min: int, levels: int,
                             k_schedule: str = "doubling",
                             k_geom_gamma: Optional[float] = None) -> Tuple[np.ndarray, List[np.ndarray]]:
    """keyframes.py:172-209 with ``scores`` = the ``torch.rand((B, T-2))`` draw at :188."""
    if levels < 1:
        raise ValueError("levels must be >= 1")
    K_list = compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    B = scores.shape[0]
    perm = np.argsort(scores, axis=1, kind="stable")
    masks_levels = np.zeros((B, levels + 1, T), dtype=bool)
    idx_levels: List[np.ndarray] = []
    for s in range(levels + 1):
        K_s = K_list[s]
        if K_s <= 2 or T <= 2:
            idx = np.concatenate([np.zeros((B, 1), np.int64), np.full((B, 1), T - 1, np.int64)], axis=1)
        else:
            interior = perm[:, : K_s - 2].astype(np.int64) + 1
            idx = np.concatenate([np.zeros((B, 1), np.int64), interior,
                                  np.full((B, 1), T - 1, np.int64)], axis=1)
        idx = np.sort(idx, axis=1)
        idx_levels.append(idx)
        masks_levels[np.arange(B)[:, None], s, idx] = True
    return masks_levels, idx_levels


def build_nested_masks_from_base(idx_base: np.ndarray, T: int, levels: int, perms: Iterable[np.ndarray],
                                 k_schedule: str = "doubling",
                                 k_geom_gamma: Optional[float] = None) -> Tuple[np.ndarray, List[np.ndarray]]:
    """keyframes.py:212-257.  ``perms`` yields, in call order, the ``torch.randperm``
    results the reference draws at :250 (one per sample per level that needs anchors)."""
    if levels < 1:
        raise ValueError("levels must be >= 1")
    if idx_base.ndim != 2:
        raise ValueError("idx_base must be [B, K]")
    B, K_base = idx_base.shape
    K_list = compute_k_schedule(T, K_base, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    masks_levels = np.zeros((B, levels + 1, T), dtype=bool)
    idx_levels: List[Optional[np.ndarray]] = [None] * (levels + 1)
    idx_s = np.sort(idx_base.astype(np.int64), axis=1)
    idx_levels[levels] = idx_s
    masks_levels[np.arange(B)[:, None], levels, idx_s] = True
    perms = iter(perms)
    for s in range(levels - 1, -1, -1):
        K_s = K_list[s]
        prev_mask = masks_levels[:, s + 1]
        need = K_s - prev_mask.sum(axis=1)
        rows = []
        for b in range(B):
            k_need = int(need[b])
            if k_need <= 0:
                rows.append(idx_levels[s + 1][b])
                continue
            available = np.nonzero(~prev_mask[b])[0]
            if available.size == 0:
                rows.append(idx_levels[s + 1][b])
                continue
            perm = np.asarray(next(perms))
            chosen = available[perm[:k_need]]
            rows.append(np.sort(np.concatenate([idx_levels[s + 1][b], chosen])))
        idx_t = np.stack(rows, axis=0).astype(np.int64)
        idx_levels[s] = idx_t
        masks_levels[np.arange(B)[:, None], s, idx_t] = True
    return masks_levels, idx_levels


def build_nested_masks_from_logits(logits: np.ndarray, K_min: int, levels: int,
                                   k_schedule: str = "doubling",
                                   k_geom_gamma: Optional[float] = None) -> Tuple[np.ndarray, List[np.ndarray]]:
    """keyframes.py:260-294: interior ranked by descending logit, prefix per level."""
    if logits.ndim != 2:
        raise ValueError("logits must be [B, T]")
    if levels < 1:
        raise ValueError("levels must be >= 1")
    B, T = logits.shape
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    K_list = compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    if K_list[levels] < 2:
        raise ValueError("K_min must be >= 2 to include endpoints")
    interior = logits[:, 1:-1]
    # descending, ties keep the lower index first (stable sort with a "greater" comparator)
    order_interior = np.argsort(-interior.astype(np.float64), axis=1, kind="stable") + 1
    endpoints = np.zeros((B, 2), np.int64)
    endpoints[:, 1] = T - 1
    order = np.concatenate([endpoints, order_interior.astype(np.int64)], axis=1)
    masks_levels = np.zeros((B, levels + 1, T), dtype=bool)
    idx_levels: List[Optional[np.ndarray]] = [None] * (levels + 1)
    for s in range(levels + 1):
        idx_s = np.sort(order[:, : K_list[s]], axis=1)
        idx_levels[s] = idx_s
        masks_levels[np.arange(B)[:, None], s, idx_s] = True
    return masks_levels, idx_levels


def build_nested_masks_from_level_logits(logits_levels: np.ndarray, K_min: int, levels: int,
                                         k_schedule: str = "doubling",
                                         k_geom_gamma: Optional[float] = None) -> Tuple[np.ndarray, List[np.ndarray]]:
    """keyframes.py:297-345: coarse-to-fine masked top-k per level."""
    if logits_levels.ndim != 3:
        raise ValueError("logits_levels must be [B, L, T]")
    B, L, T = logits_levels.shape
    if levels < 1:
        raise ValueError("levels must be >= 1")
    if L != levels + 1:
        raise ValueError(f"logits_levels second dim must be levels+1 ({levels+1}), got {L}")
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    K_list = compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    masks_levels = np.zeros((B, levels + 1, T), dtype=bool)
    selected = np.zeros((B, T), dtype=bool)
    selected[:, 0] = True
    selected[:, -1] = True
    for s in range(levels, -1, -1):
        need = K_list[s] - selected.sum(axis=1)
        if np.any(need < 0):
            raise ValueError("K_schedule produced decreasing K values; ensure nestedness.")
        if np.any(need > 0):
            scores = logits_levels[:, s, :].astype(F32).copy()
            scores[selected] = F32(-1e9)
            for b in range(B):
                k_need = int(need[b])
                if k_need <= 0:
                    continue
                top = np.argsort(-scores[b].astype(np.float64), kind="stable")[:k_need]
                selected[b, top] = True
        masks_levels[:, s] = selected
    idx_levels = []
    for s in range(levels + 1):
        idx_levels.append(np.stack([np.nonzero(masks_levels[b, s])[0] for b in range(B)]).astype(np.int64))
    return masks_levels, idx_levels


# --------------------------------------------------------------------------- #
# interpolation  (keyframes.py:348-380, 413-455)
# --------------------------------------------------------------------------- #
def segment_lookup(idx: np.ndarray, T: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """keyframes.py:361-366: seg = searchsorted(idx, t, right=True) - 1 clamped to [0, K-2]."""
    B, K = idx.shape
    t_grid = np.arange(T, dtype=np.int64)
    seg = (idx[:, None, :] <= t_grid[None, :, None]).sum(axis=2) - 1      # idx rows are sorted
    seg = np.clip(seg, 0, K - 2)
    rows = np.arange(B)[:, None]
    return seg, idx[rows, seg], idx[rows, seg + 1]


def interpolate_from_indices(idx: np.ndarray, vals: np.ndarray, T: int,
                             recompute_velocity: bool = False) -> np.ndarray:
    """keyframes.py:348-380 (the vectorised function is the contract)."""
    if idx.ndim != 2:
        raise ValueError("idx must be [B, K]")
    if vals.ndim != 3:
        raise ValueError("vals must be [B, K, D]")
    B, K = idx.shape
    D = vals.shape[2]
    vals = vals.astype(F32, copy=False)
    seg, left_idx, right_idx = segment_lookup(idx, T)
    rows = np.arange(B)[:, None]
    left_val = vals[rows, seg]                    # [B,T,D]
    right_val = vals[rows, seg + 1]
    denom = np.maximum(right_idx - left_idx, 1).astype(F32)[..., None]
    t_grid = np.arange(T, dtype=np.int64)[None, :]
    w = (t_grid - left_idx).astype(F32)[..., None] / denom
    y = left_val + w * (right_val - left_val)
    for k in range(K):                            # scatter_ :372 (sequential: last duplicate wins)
        y[np.arange(B), idx[:, k]] = vals[:, k]
    if recompute_velocity and D == 4:
        pos = y[:, :, :2]
        v = np.zeros_like(pos)
        dt = F32(1.0 / float(T))
        v[:, :-1] = (pos[:, 1:] - pos[:, :-1]) / dt
        v[:, -1] = 0.0
        y = np.concatenate([pos, v], axis=-1)
    return y.astype(F32)


def interpolate_from_mask(x: np.ndarray, mask: np.ndarray, recompute_velocity: bool = False) -> np.ndarray:
    """keyframes.py:413-455 restated through idx (thin wrapper).  The legacy
    loop uses linspace weights and differs from the vectorised form by <= 1 ulp; rows
    with fewer than two anchors are returned unchanged (:439-440)."""
    single = x.ndim == 2
    if single:
        x = x[None]
        mask = mask[None]
    if x.ndim != 3:
        raise ValueError("x must have shape [T, D] or [B, T, D]")
    B, T, D = x.shape
    if mask.ndim == 1:
        mask = np.broadcast_to(mask[None], (B, T))
    y = np.empty_like(x, dtype=F32)
    for b in range(B):
        kf = np.nonzero(mask[b])[0]
        if kf.size < 2:
            yb = x[b].astype(F32).copy()
            if recompute_velocity and D == 4:
                yb = interpolate_from_indices(np.arange(T)[None], yb[None], T, True)[0]
            y[b] = yb
            continue
        # outside [kf[0], kf[-1]] the legacy loop keeps x; the idx form would extrapolate,
        # so add those positions as anchors (they reproduce x exactly).
        keep = np.concatenate([np.arange(0, kf[0]), kf, np.arange(kf[-1] + 1, T)]).astype(np.int64)
        y[b] = interpolate_from_indices(keep[None], x[b][keep][None], T, recompute_velocity)[0]
    return y[0] if single else y
