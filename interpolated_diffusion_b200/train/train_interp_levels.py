"""Mirror of the corruption builders of the reference's ``src/train/train_interp_levels.py`` (lines 227-596)
on libidb200 kernels.  The trainer CLI / checkpoint / selector / bootstrap plumbing is out of scope
(SURVEY.md section 2, row 9).

Generator draw order is the reference's: every ``torch.randn`` / ``torch.rand`` /
``torch.randint`` below is the same call with the same shape in the same place, so with the same generator
state the noise is identical and the fused kernel (``idb200_corrupt_from_anchors``) reproduces the
reference bit for bit."""
from __future__ import annotations

from typing import List, Optional

import torch

from .. import _lib as L
from ..corruptions.keyframes import build_nested_masks_batch, interpolate_from_indices
from ..sample.sample_generate import _compute_sigma_for_level, anchor_conf_mask_in


def _compute_jitter_for_level(K_s: int, K_min: int, jitter_max: int, jitter_pow: float) -> int:
    """train_interp_levels.py:433-441"""
    if jitter_max <= 0:
        return 0
    K_s = max(1, int(K_s))
    K_min = max(1, int(K_min))
    ratio = float(K_min) / float(K_s)
    jitter = int(round(float(jitter_max) * (ratio ** float(jitter_pow))))
    return max(0, min(int(jitter_max), jitter))


def _corrupt_from_anchors(source: torch.Tensor, idx: torch.Tensor, T: int, generator: torch.Generator, sigma: float,
                          anchor_sigma: float, index_jitter: int, index_jitter_prob: float, mode: str,
                          clamp_endpoints: bool, recompute_velocity: bool, *, out: Optional[torch.Tensor] = None,
                          row_index: Optional[torch.Tensor] = None) -> torch.Tensor:
    """train_interp_levels.py:458-510 in one launch.  ``row_index`` (new, optional): rows of ``idx`` / noise
    map to rows ``row_index`` of ``source`` / ``out`` -- the boolean-mask gather / scatter of :328-382 without
    materialising ``source[sel]``."""
    dev = L.require_cuda(source, idx)
    n, K = idx.shape
    D = source.shape[-1]
    src = L.f32c(source)
    idx = L.i64c(idx)
    idx_j = None
    if index_jitter > 0 and index_jitter_prob > 0.0:           # :471-483
        jitter = torch.randint(0, 2 * index_jitter + 1, (n, K), generator=generator, device=dev) - int(index_jitter)
        use = torch.rand((n, K), generator=generator, device=dev) < float(index_jitter_prob)
        if clamp_endpoints:
            use = use & ~(idx == 0) & ~(idx == (T - 1))
        idx_j = torch.where(use, idx + jitter, idx).clamp(0, T - 1)
    anchor_noise = None
    if anchor_sigma > 0.0:                                     # :486-490
        anchor_noise = torch.randn((n, K, 2), generator=generator, device=dev, dtype=src.dtype)
    path_noise = None
    if sigma > 0.0:                                            # :496-501
        path_noise = torch.randn((n, T, 2), generator=generator, device=dev, dtype=src.dtype)
    if out is None:
        if row_index is not None:
            raise ValueError("row_index needs an explicit out= buffer")
        out = torch.empty((n, T, D), device=dev, dtype=torch.float32)
    L.call("idb200_corrupt_from_anchors", L.ptr(src), L.ptr(idx), L.ptr(idx_j), L.ptr(anchor_noise), L.ptr(path_noise),
           L.ptr(None if row_index is None else L.i64c(row_index)), n, K, T, D, float(sigma), float(anchor_sigma),
           int(mode == "dist"), int(bool(clamp_endpoints)), int(bool(recompute_velocity)), L.ptr(out), L.stream(dev))
    return out


def _level_rows(s_idx: torch.Tensor, levels: int):
    """Row indices per level.  One host sync for the level histogram (the reference syncs once per level
    with torch.any, plus the boolean-mask indexing)."""
    order = torch.argsort(s_idx, stable=True)
    counts = torch.bincount(s_idx, minlength=levels + 1).tolist()
    rows, off = {}, 0
    for s in range(levels + 1):
        c = counts[s] if s < len(counts) else 0
        if s >= 1 and c > 0:
            rows[s] = order[off: off + c]
        off += c
    return rows


def build_interp_level_batch(x0: torch.Tensor, K_min: int, levels: int, generator: torch.Generator,
                             recompute_velocity: bool = False, x0_override: Optional[torch.Tensor] = None,
                             masks_levels: Optional[torch.Tensor] = None, idx_levels: Optional[List[torch.Tensor]] = None,
                             s_idx: Optional[torch.Tensor] = None, corrupt_mode: str = "none",
                             corrupt_sigma_max: float = 0.0, corrupt_sigma_min: float = 0.0, corrupt_sigma_pow: float = 1.0,
                             corrupt_anchor_frac: float = 0.0, corrupt_index_jitter_max: int = 0,
                             corrupt_index_jitter_prob: float = 0.0, corrupt_index_jitter_pow: float = 1.0,
                             clamp_endpoints: bool = True, pos_clip: bool = False, pos_clip_min: float = 0.0,
                             pos_clip_max: float = 1.0):
    """train_interp_levels.py:227-291"""
    dev = L.require_cuda(x0)
    B, T, D = x0.shape
    if masks_levels is None or idx_levels is None:
        masks_levels, idx_levels = build_nested_masks_batch(B, T, K_min, levels, generator=generator, device=dev)
    if s_idx is None:
        s_idx = torch.randint(1, levels + 1, (B,), generator=generator, device=dev, dtype=torch.long)
    source = L.f32c(x0_override if x0_override is not None else x0)
    x_s = torch.zeros((B, T, D), device=dev, dtype=torch.float32)
    for s, rows in _level_rows(s_idx, levels).items():
        idx = idx_levels[s][rows]
        K_s = idx.shape[1]
        sigma = anchor_sigma = 0.0
        jitter = 0
        if corrupt_mode != "none":
            sigma = _compute_sigma_for_level(K_s, K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
            anchor_sigma = sigma * float(corrupt_anchor_frac)
            jitter = _compute_jitter_for_level(K_s, K_min, corrupt_index_jitter_max, corrupt_index_jitter_pow)
        _corrupt_from_anchors(source, idx, T, generator, sigma, anchor_sigma, jitter, corrupt_index_jitter_prob,
                              corrupt_mode, clamp_endpoints, recompute_velocity, out=x_s, row_index=rows)
    if pos_clip:
        x_s[..., :2] = x_s[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
    mask_s = masks_levels.gather(1, s_idx.view(B, 1, 1).expand(B, 1, T)).squeeze(1)
    mask_s = mask_s & (s_idx >= 1).view(B, 1)
    return x_s, mask_s, s_idx, masks_levels, idx_levels


def build_interp_adjacent_batch(x0: torch.Tensor, K_min: int, levels: int, generator: torch.Generator,
                                recompute_velocity: bool = False, x0_override: Optional[torch.Tensor] = None,
                                masks_levels: Optional[torch.Tensor] = None, idx_levels: Optional[List[torch.Tensor]] = None,
                                s_idx: Optional[torch.Tensor] = None, corrupt_mode: str = "none",
                                corrupt_sigma_max: float = 0.0, corrupt_sigma_min: float = 0.0,
                                corrupt_sigma_pow: float = 1.0, corrupt_anchor_frac: float = 0.0,
                                corrupt_index_jitter_max: int = 0, corrupt_index_jitter_prob: float = 0.0,
                                corrupt_index_jitter_pow: float = 1.0, clamp_endpoints: bool = True, pos_clip: bool = False,
                                pos_clip_min: float = 0.0, pos_clip_max: float = 1.0):
    """train_interp_levels.py:294-383: x_s = Interp/corrupt at M_s, x_prev at M_{s-1}, per-row level s."""
    dev = L.require_cuda(x0)
    B, T, D = x0.shape
    if masks_levels is None or idx_levels is None:
        masks_levels, idx_levels = build_nested_masks_batch(B, T, K_min, levels, generator=generator, device=dev)
    if s_idx is None:
        s_idx = torch.randint(1, levels + 1, (B,), generator=generator, device=dev, dtype=torch.long)
    source = L.f32c(x0_override if x0_override is not None else x0)
    x_s = torch.zeros((B, T, D), device=dev, dtype=torch.float32)
    x_prev = torch.zeros((B, T, D), device=dev, dtype=torch.float32)
    for s, rows in _level_rows(s_idx, levels).items():
        for target, lvl in ((x_s, s), (x_prev, s - 1)):           # same draw order as :349 then :362
            idx = idx_levels[lvl][rows]
            K_l = idx.shape[1]
            sigma = anchor_sigma = 0.0
            jitter = 0
            if corrupt_mode != "none":
                sigma = _compute_sigma_for_level(K_l, K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
                anchor_sigma = sigma * float(corrupt_anchor_frac)
                jitter = _compute_jitter_for_level(K_l, K_min, corrupt_index_jitter_max, corrupt_index_jitter_pow)
            _corrupt_from_anchors(source, idx, T, generator, sigma, anchor_sigma, jitter, corrupt_index_jitter_prob,
                                  corrupt_mode, clamp_endpoints, recompute_velocity, out=target, row_index=rows)
    if pos_clip:
        x_s[..., :2] = x_s[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
        x_prev[..., :2] = x_prev[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
    valid = (s_idx >= 1).view(B, 1)
    mask_s = masks_levels.gather(1, s_idx.clamp(min=0).view(B, 1, 1).expand(B, 1, T)).squeeze(1) & valid
    mask_prev = masks_levels.gather(1, (s_idx - 1).clamp(min=0).view(B, 1, 1).expand(B, 1, T)).squeeze(1) & valid
    return x_s, x_prev, mask_s, mask_prev, s_idx, masks_levels, idx_levels


def _build_anchor_conf(mask_s: torch.Tensor, student_mask: Optional[torch.Tensor], conf_teacher: float,
                       conf_student: float, conf_endpoints: float, conf_missing: float,
                       clamp_endpoints: bool) -> torch.Tensor:
    """train_interp_levels.py:546-562"""
    conf, _ = anchor_conf_mask_in(mask_s, student_mask, None, 0, 0, "none", conf_teacher, conf_student, conf_endpoints,
                                  conf_missing, clamp_endpoints)
    return conf


def _anneal_conf(conf: torch.Tensor, s_idx: torch.Tensor, levels: int, mode: str) -> torch.Tensor:
    """train_interp_levels.py:565-576 (per-row level)."""
    if conf is None or mode == "none" or levels <= 0:
        return conf
    frac = s_idx.float() / float(levels)
    if mode == "linear":
        lam = 1.0 - frac
    elif mode == "cosine":
        lam = 0.5 * (1.0 + torch.cos(torch.pi * frac))
    else:
        lam = torch.zeros_like(frac)
    return conf + (1.0 - conf) * lam.view(-1, 1)


def _sample_level_indices(B: int, levels: int, generator: torch.Generator, device: torch.device, mode: str,
                          high_prob: float, *, sync_free: bool = False) -> torch.Tensor:
    """train_interp_levels.py:578-596.  Default (parity mode): same draws in the same order (``rand(B)`` then ``randint``
    with the data-dependent count of non-high rows), so the generator stream stays aligned with the reference -- at the price
    of one host sync (the count).  ``sync_free=True`` (speed mode): the low-level candidate is drawn for ALL B rows and
    selected with ``where`` -- the same distribution (s = S w.p. high_prob, else U{1..S}), no ``.item()``, but a different
    number of generator draws, so the stream after this call differs from the reference's."""
    if mode == "uniform" or levels <= 1:
        return torch.randint(1, levels + 1, (B,), generator=generator, device=device, dtype=torch.long)
    high_prob = float(max(0.0, min(1.0, high_prob)))
    draw = torch.rand((B,), generator=generator, device=device)
    high = draw < high_prob
    if sync_free:
        low = torch.randint(1, levels + 1, (B,), generator=generator, device=device, dtype=torch.long)
        return torch.where(high, torch.full_like(low, levels), low)
    s_idx = torch.full((B,), levels, device=device, dtype=torch.long)
    n_low = int((~high).sum().item())
    if n_low > 0:
        s_idx[~high] = torch.randint(1, levels + 1, (n_low,), generator=generator, device=device, dtype=torch.long)
    return s_idx


def level_sigmas(K_list, K_min: int, corrupt_mode: str, corrupt_sigma_max: float, corrupt_sigma_min: float, corrupt_sigma_pow: float,
                 corrupt_anchor_frac: float):
    """Per-level (sigma, anchor_sigma) as :335-341 computes them from K_s (zeros when corrupt_mode == "none")."""
    sig = [0.0 if corrupt_mode == "none" else _compute_sigma_for_level(int(k), K_min, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow)
           for k in K_list]
    return sig, [x * float(corrupt_anchor_frac) for x in sig]


def corrupt_adjacent_fused(x0: torch.Tensor, masks_levels: torch.Tensor, s_idx: torch.Tensor, K_list, K_min: int, *,
                           adjacent: bool = True, recompute_velocity: bool = False, corrupt_mode: str = "none",
                           corrupt_sigma_max: float = 0.0, corrupt_sigma_min: float = 0.0, corrupt_sigma_pow: float = 1.0,
                           corrupt_anchor_frac: float = 0.0, clamp_endpoints: bool = True, pos_clip: bool = False,
                           pos_clip_min: float = 0.0, pos_clip_max: float = 1.0, anchor_noise: Optional[torch.Tensor] = None,
                           path_noise: Optional[torch.Tensor] = None, seed: int = 0, offset: int = 0, export_noise: bool = False):
    """``build_interp_adjacent_batch`` (:294-383; ``adjacent=False``: ``build_interp_level_batch`` :227-291) for the whole batch
    in ONE launch (``idb200_corrupt_adjacent``): per-row level, no per-level loop, no host sync.

    Noise: ``anchor_noise`` [B,2,Kmax,2] / ``path_noise`` [B,2,T,2] given -> parity mode (bit-identical to the per-level path
    when they hold the reference's draws); both None -> Philox in the kernel keyed by (seed, offset).
    Returns (x_s, x_prev | None, mask_s, mask_prev | None[, anchor_noise_drawn, path_noise_drawn])."""
    dev = L.require_cuda(x0, masks_levels, s_idx)
    B, T, D = x0.shape
    n_levels = masks_levels.shape[1]
    if len(K_list) != n_levels:
        raise ValueError("K_list must have one entry per mask level")
    sig, asig = level_sigmas(K_list, K_min, corrupt_mode, corrupt_sigma_max, corrupt_sigma_min, corrupt_sigma_pow, corrupt_anchor_frac)
    Kmax = int(min(T, max(int(k) for k in K_list)))
    import ctypes
    fa = (ctypes.c_float * n_levels)(*sig)
    fb = (ctypes.c_float * n_levels)(*asig)
    src = L.f32c(x0)
    x_s = torch.empty((B, T, D), device=dev, dtype=torch.float32)
    x_prev = torch.empty((B, T, D), device=dev, dtype=torch.float32) if adjacent else None
    mask_s = torch.empty((B, T), device=dev, dtype=torch.bool)
    mask_prev = torch.empty((B, T), device=dev, dtype=torch.bool) if adjacent else None
    an_out = pn_out = None
    if export_noise:
        if anchor_noise is not None:
            raise ValueError("export_noise is for the Philox mode")
        an_out = torch.zeros((B, 2, Kmax, 2), device=dev, dtype=torch.float32)
        pn_out = torch.zeros((B, 2, T, 2), device=dev, dtype=torch.float32)
    if (anchor_noise is None) != (path_noise is None):
        raise ValueError("anchor_noise and path_noise must be given together")
    if anchor_noise is not None:
        if tuple(anchor_noise.shape) != (B, 2, Kmax, 2) or tuple(path_noise.shape) != (B, 2, T, 2):
            raise ValueError(f"noise shapes must be [B,2,{Kmax},2] and [B,2,{T},2]")
        anchor_noise, path_noise = L.f32c(anchor_noise), L.f32c(path_noise)
    L.call("idb200_corrupt_adjacent", L.ptr(src), L.ptr(L.u8c(masks_levels)), L.ptr(L.i64c(s_idx)), B, T, D, n_levels, fa, fb,
           L.ptr(anchor_noise), L.ptr(path_noise), Kmax, int(seed) & (2 ** 64 - 1), int(offset) & (2 ** 64 - 1), L.ptr(an_out), L.ptr(pn_out),
           int(corrupt_mode == "dist"), int(bool(clamp_endpoints)), int(bool(recompute_velocity)), L.ptr(x_s), L.ptr(x_prev),
           L.ptr(mask_s), L.ptr(mask_prev), L.stream(dev))
    if pos_clip:
        x_s[..., :2] = x_s[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
        if x_prev is not None:
            x_prev[..., :2] = x_prev[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
    if export_noise:
        return x_s, x_prev, mask_s, mask_prev, an_out, pn_out
    return x_s, x_prev, mask_s, mask_prev
