"""Drop-in mirror of the reference's ``src/corruptions/keyframes.py`` (same names, positional
order, defaults, dtypes, return structure and error messages) executing on libidb200's sm_100a
kernels.  CUDA only: ``device=None`` means the current CUDA device (the reference defaults to CPU).

New keyword-only arguments (defaults reproduce the reference): ``scores=`` / ``u=`` inject the random
draw the reference makes with ``torch.rand(..., generator=generator, device=device)``.

Tie-breaking contract: anchors are ranked by the STABLE ascending order of the
scores (lower index first).  ``torch.argsort`` itself is not stable for rows longer than 16 on CPU
and differs again on CUDA, so on the ~1e-4 of rows that contain an exact fp32 tie the reference is
implementation-defined; everywhere else results are bit-identical.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Tuple

import torch

from .. import _lib as L

F_DESCENDING = 1
F_RECOMPUTE_VELOCITY = 2
F_NO_ENDPOINTS = 4

__all__ = [
    "sample_fixed_k_mask", "sample_fixed_k_indices_batch", "sample_fixed_k_indices_uniform_batch",
    "_compute_k_schedule", "build_nested_masks_batch", "build_nested_masks_from_base",
    "build_nested_masks_from_logits", "build_nested_masks_from_level_logits", "interpolate_from_indices",
    "build_nested_masks", "interpolate_from_mask", "interpolate_keyframes", "nested_masks_interp",
]


def _check_tk(T: int, K: int, ensure_endpoints: bool) -> None:
    # keyframes.py:14-22 / :51-59 / :94-102 (same messages)
    if T <= 0:
        raise ValueError("T must be positive")
    if K <= 0:
        raise ValueError("K must be positive")
    if ensure_endpoints:
        if T < 2:
            raise ValueError("T must be >= 2 when ensure_endpoints is True")
        if K < 2:
            raise ValueError("K must be >= 2 when ensure_endpoints is True")


def _compute_k_schedule(T: int, K_min: int, levels: int, schedule: str = "doubling",
                        geom_gamma: float = None) -> List[int]:
    """keyframes.py:135-169 (host-side integers; python ``round`` is banker's rounding)."""
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


def _idx_widths(K_list, T: int, no_endpoints: bool) -> List[int]:
    if no_endpoints:
        return [min(k, T) for k in K_list]
    return [2 if (k <= 2 or T <= 2) else min(k, T) for k in K_list]


def nested_masks_interp(scores: torch.Tensor, T: int, K_list, *, x0: Optional[torch.Tensor] = None,
                        levels_out: Optional[Tuple[int, int]] = None, want_masks: bool = True,
                        want_idx: bool = True, flags: int = 0, score_offset: int = 0):
    """One launch of K1 (``idb200_nested_masks_interp``): masks for every level in ``K_list`` and,
    if ``x0`` is given, ``Interp(x0 | M_s)`` for ``levels_out = (s_lo, s_hi)``.

    Returns ``(masks bool [B,L,T] | None, [idx_s int64 [B,W_s]] | None, x_levels fp32 [n,B,T,D] | None)``.
    """
    dev = L.require_cuda(scores, x0)
    B = scores.shape[0]
    n_levels = len(K_list)
    no_end = bool(flags & F_NO_ENDPOINTS)
    scores = L.f32c(scores) if score_offset == 0 else scores
    score_stride = scores.stride(0) if scores.dim() == 2 and scores.shape[1] > 0 else max(T - 2, 1)
    masks = torch.empty((B, n_levels, T), dtype=torch.bool, device=dev) if want_masks else None
    widths = _idx_widths(K_list, T, no_end)
    idx_buf = torch.empty((B * sum(widths),), dtype=torch.int64, device=dev) if want_idx else None
    x_levels = None
    s_lo = s_hi = 0
    D = 0
    if x0 is not None:
        x0 = L.f32c(x0)
        D = x0.shape[-1]
        s_lo, s_hi = levels_out if levels_out is not None else (0, n_levels - 1)
        x_levels = torch.empty((s_hi - s_lo + 1, B, T, D), dtype=torch.float32, device=dev)
    karr = (ctypes.c_int * n_levels)(*[int(k) for k in K_list])
    sptr = scores.data_ptr() + 4 * score_offset if scores.numel() > 0 else None
    L.call("idb200_nested_masks_interp", L.ptr(x0), sptr, score_stride, B, T, D, n_levels, karr, L.ptr(masks),
           L.ptr(idx_buf), L.ptr(x_levels), B * T * D, s_lo, s_hi, flags, L.stream(dev))
    idx_levels = None
    if want_idx:
        idx_levels, off = [], 0
        for w in widths:
            idx_levels.append(idx_buf[B * off: B * (off + w)].view(B, w))
            off += w
    return masks, idx_levels, x_levels


def sample_fixed_k_mask(T: int, K: int, generator: torch.Generator = None, device: torch.device = None,
                        ensure_endpoints: bool = True) -> torch.Tensor:
    """keyframes.py:6-39 (single-sample legacy helper; same ``torch.randperm`` draw as the reference)."""
    device = L.resolve_device(device)
    _check_tk(T, K, ensure_endpoints)
    K = min(K, T)
    mask = torch.zeros(T, dtype=torch.bool, device=device)
    if ensure_endpoints:
        mask[0] = True
        mask[T - 1] = True
        remaining = K - 2
        if remaining > 0 and T > 2:
            perm = torch.randperm(T - 2, generator=generator, device=device)
            mask[perm[:remaining] + 1] = True
    else:
        perm = torch.randperm(T, generator=generator, device=device)
        mask[perm[:K]] = True
    return mask


def sample_fixed_k_indices_batch(B: int, T: int, K: int, generator: torch.Generator = None,
                                 device: torch.device = None, ensure_endpoints: bool = True, *,
                                 scores: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """keyframes.py:42-81 -> (idx int64 [B,K], mask bool [B,T])."""
    device = L.resolve_device(device if scores is None else scores.device)
    _check_tk(T, K, ensure_endpoints)
    K = min(K, T)
    if ensure_endpoints:
        if T > 2 and K > 2:
            if scores is None:
                scores = torch.rand((B, T - 2), generator=generator, device=device)
        else:
            scores = torch.zeros((B, max(T - 2, 0)), device=device)
        masks, idxs, _ = nested_masks_interp(scores, T, [K])
    else:
        if scores is None:
            scores = torch.rand((B, T), generator=generator, device=device)
        masks, idxs, _ = nested_masks_interp(scores, T, [K], flags=F_NO_ENDPOINTS)
    return idxs[0], masks[:, 0]


def sample_fixed_k_indices_uniform_batch(B: int, T: int, K: int, generator: torch.Generator = None,
                                         device: torch.device = None, ensure_endpoints: bool = True,
                                         jitter: float = 0.0, *, u: Optional[torch.Tensor] = None
                                         ) -> Tuple[torch.Tensor, torch.Tensor]:
    """keyframes.py:84-132.  Without jitter the row is batch-invariant: it is computed once on the host
    with the reference's own op sequence (linspace -> round half-even -> monotone fix-ups) and
    broadcast; with jitter the same sequence runs per row on the device."""
    device = L.resolve_device(device)
    _check_tk(T, K, ensure_endpoints)
    K = min(K, T)
    if K > T:
        raise ValueError("K must be <= T for uniform spacing")
    use_jitter = bool(jitter) and K > 2 and T > 2
    work_dev = device if use_jitter else torch.device("cpu")
    base = torch.linspace(0, T - 1, K, device=work_dev)
    if use_jitter:
        spacing = float(T - 1) / float(K - 1)
        max_jitter = spacing * float(jitter) * 0.5
        if u is None:
            u = torch.rand((B, K), generator=generator, device=device)
        noise = (u - 0.5) * 2.0 * max_jitter
        noise[:, 0] = 0.0
        noise[:, -1] = 0.0
        pos = base.unsqueeze(0) + noise
    else:
        pos = base.unsqueeze(0)
    idx = torch.round(pos).long().clamp(0, T - 1)
    if ensure_endpoints and K >= 2:
        idx[:, 0] = 0
        idx[:, -1] = T - 1
    for k in range(1, K):
        idx[:, k] = torch.maximum(idx[:, k], idx[:, k - 1] + 1)
    for k in range(K - 2, -1, -1):
        idx[:, k] = torch.minimum(idx[:, k], idx[:, k + 1] - 1)
    idx = idx.clamp(0, T - 1)
    if ensure_endpoints and K >= 2:
        idx[:, 0] = 0
        idx[:, -1] = T - 1
    if not use_jitter:
        idx = idx.to(device).expand(B, K).contiguous()
    mask = torch.zeros((B, T), dtype=torch.bool, device=device)
    mask.scatter_(1, idx, True)
    return idx, mask


def build_nested_masks_batch(B: int, T: int, K_min: int, levels: int, generator: torch.Generator = None,
                             device: torch.device = None, k_schedule: str = "doubling",
                             k_geom_gamma: float = None, *, scores: Optional[torch.Tensor] = None
                             ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """keyframes.py:172-209 -> (masks bool [B,S+1,T], [idx_s int64 [B,K_s]])."""
    if levels < 1:
        raise ValueError("levels must be >= 1")
    device = L.resolve_device(device if scores is None else scores.device)
    K_list = _compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    if scores is None:
        scores = torch.rand((B, T - 2), generator=generator, device=device)
    masks, idxs, _ = nested_masks_interp(scores, T, K_list)
    return masks, idxs


def build_nested_masks_from_base(idx_base: torch.Tensor, T: int, levels: int, generator: torch.Generator = None,
                                 device: torch.device = None, k_schedule: str = "doubling",
                                 k_geom_gamma: float = None, *, scores: Optional[torch.Tensor] = None
                                 ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """keyframes.py:212-257, re-formulated for a parallel machine: the reference draws one
    ``randperm`` per sample per level in a python loop (O(B*S) host syncs), which no batched kernel can
    replay bit-for-bit.  Here one ``torch.rand((B, T))`` ranks the non-base positions; base anchors are
    forced to the front, so level S == idx_base, every finer level is a superset with exactly K_s
    anchors, and the added anchors are a uniformly random nested choice -- the same distribution."""
    if levels < 1:
        raise ValueError("levels must be >= 1")
    if idx_base.dim() != 2:
        raise ValueError("idx_base must be [B, K]")
    L.require_cuda(idx_base)
    device = idx_base.device
    B, K_base = idx_base.shape
    K_list = _compute_k_schedule(T, K_base, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    if scores is None:
        scores = torch.rand((B, T), generator=generator, device=device)
    scores = L.f32c(scores).clone()
    # base anchors first, in index order (rank ties resolve to the lower index)
    scores.scatter_(1, L.i64c(idx_base), -1.0)
    masks, idxs, _ = nested_masks_interp(scores, T, K_list, flags=F_NO_ENDPOINTS)
    return masks, idxs


def build_nested_masks_from_logits(logits: torch.Tensor, K_min: int, levels: int, k_schedule: str = "doubling",
                                   k_geom_gamma: float = None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """keyframes.py:260-294: interior ranked by descending logit; endpoints always kept."""
    if logits.dim() != 2:
        raise ValueError("logits must be [B, T]")
    if levels < 1:
        raise ValueError("levels must be >= 1")
    B, T = logits.shape
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    K_list = _compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    if K_list[levels] < 2:
        raise ValueError("K_min must be >= 2 to include endpoints")
    logits = L.f32c(logits)
    masks, idxs, _ = nested_masks_interp(logits, T, K_list, flags=F_DESCENDING, score_offset=1)
    return masks, idxs


def build_nested_masks_from_level_logits(logits_levels: torch.Tensor, K_min: int, levels: int,
                                         k_schedule: str = "doubling", k_geom_gamma: float = None
                                         ) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    """keyframes.py:297-345: coarse-to-fine, level s adds the top (K_s - |selected|) unselected
    positions by ``logits_levels[:, s]``.  One K1 launch per level: already-selected positions are
    pushed to the front (score +inf-like, in index order) exactly as the reference pushes them to
    the back of its top-k with -1e9."""
    if logits_levels.dim() != 3:
        raise ValueError("logits_levels must be [B, L, T]")
    B, Lv, T = logits_levels.shape
    if levels < 1:
        raise ValueError("levels must be >= 1")
    if Lv != levels + 1:
        raise ValueError(f"logits_levels second dim must be levels+1 ({levels+1}), got {Lv}")
    if T < 2:
        raise ValueError("T must be >= 2 when using endpoints")
    L.require_cuda(logits_levels)
    device = logits_levels.device
    K_list = _compute_k_schedule(T, K_min, levels, schedule=k_schedule, geom_gamma=k_geom_gamma)
    for s in range(levels):
        if K_list[s] < K_list[s + 1]:
            raise ValueError("K_schedule produced decreasing K values; ensure nestedness.")
    masks_levels = torch.zeros((B, levels + 1, T), dtype=torch.bool, device=device)
    idx_levels: List[torch.Tensor] = [None for _ in range(levels + 1)]
    selected = torch.zeros((B, T), dtype=torch.bool, device=device)
    selected[:, 0] = True
    selected[:, -1] = True
    big = torch.finfo(torch.float32).max
    for s in range(levels, -1, -1):
        sc = L.f32c(logits_levels[:, s, :]).clone()
        sc.masked_fill_(selected, big)               # selected first, ties by index
        m, ix, _ = nested_masks_interp(sc, T, [max(K_list[s], 2)], flags=F_DESCENDING | F_NO_ENDPOINTS)
        selected = m[:, 0]
        masks_levels[:, s] = selected
        idx_levels[s] = ix[0]
    return masks_levels, idx_levels


def interpolate_from_indices(idx: torch.Tensor, vals: torch.Tensor, T: int,
                             recompute_velocity: bool = False) -> torch.Tensor:
    """keyframes.py:348-380 -> fp32 [B,T,D] (bit-identical arithmetic; see include/idb200.h)."""
    if idx.dim() != 2:
        raise ValueError("idx must be [B, K]")
    if vals.dim() != 3:
        raise ValueError("vals must be [B, K, D]")
    dev = L.require_cuda(idx, vals)
    B, K = idx.shape
    D = vals.shape[2]
    idx = L.i64c(idx)
    vals = L.f32c(vals)
    y = torch.empty((B, T, D), dtype=torch.float32, device=dev)
    L.call("idb200_interpolate_from_indices", L.ptr(idx), L.ptr(vals), B, K, T, D, int(bool(recompute_velocity)),
           L.ptr(y), L.stream(dev))
    return y


def build_nested_masks(T: int, K_min: int, levels: int, generator: torch.Generator = None,
                       device: torch.device = None) -> List[torch.Tensor]:
    """keyframes.py:383-410 (single-sample legacy list of masks) through the batched kernel."""
    if levels < 1:
        raise ValueError("levels must be >= 1")
    masks, _ = build_nested_masks_batch(1, T, K_min, levels, generator=generator, device=device)
    return [masks[0, s] for s in range(levels + 1)]


def interpolate_from_mask(x: torch.Tensor, mask: torch.Tensor, recompute_velocity: bool = False) -> torch.Tensor:
    """keyframes.py:413-455 as a thin wrapper over the idx kernel.  Rows keep ``x`` outside
    [first anchor, last anchor] and unchanged when they hold fewer than two anchors, like the legacy loop."""
    single = x.dim() == 2
    if single:
        x = x.unsqueeze(0)
        mask = mask.unsqueeze(0)
    if x.dim() != 3:
        raise ValueError("x must have shape [T, D] or [B, T, D]")
    dev = L.require_cuda(x, mask)
    B, T, D = x.shape
    if mask.dim() == 1:
        mask = mask.unsqueeze(0).expand(B, T)
    mask = mask.bool()
    t = torch.arange(T, device=dev).unsqueeze(0)
    cnt = mask.sum(dim=1, keepdim=True)
    first = torch.where(mask, t, T).min(dim=1, keepdim=True).values
    last = torch.where(mask, t, -1).max(dim=1, keepdim=True).values
    keep = mask | (t < first) | (t > last) | (cnt < 2)
    # ragged anchor sets -> fixed width T: sorted anchor positions, padded with the last anchor (duplicates
    # are legal for the kernel: denominators clamp to 1 and anchors are copied exactly)
    key = torch.where(keep, t, T + t)
    order = torch.sort(key, dim=1).values
    last_keep = torch.where(keep, t, -1).max(dim=1, keepdim=True).values
    idx = torch.where(order < T, order, last_keep)
    vals = x.float().gather(1, idx.unsqueeze(-1).expand(B, T, D))
    y = interpolate_from_indices(idx, vals, T, recompute_velocity=recompute_velocity)
    return y[0] if single else y


def interpolate_keyframes(x: torch.Tensor, mask: torch.Tensor, recompute_velocity: bool = False) -> torch.Tensor:
    """keyframes.py:496-498"""
    return interpolate_from_mask(x, mask, recompute_velocity=recompute_velocity)
