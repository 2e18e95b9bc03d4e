"""Mirror of the DP anchor placement of the reference's ``src/selection/epiplexity_dp.py`` (``dp_select_indices`` :171-197,
``dp_select_indices_batch`` :200-228): the producer of ``idx`` when ``kp_index_mode = dp``.  One kernel, one block per sample."""
from __future__ import annotations

import torch

from .. import _lib as L


def dp_select_indices_batch(C: torch.Tensor, K: int) -> torch.Tensor:
    """C fp32 [B, T, T] segment costs (inf where no segment) -> idx i64 [B, min(K, T)], idx[:, 0] = 0, idx[:, -1] = T - 1."""
    if C.dim() != 3:
        raise ValueError("C must be [B,T,T]")
    dev = L.require_cuda(C)
    B, T, _ = C.shape
    if K < 2:
        raise ValueError("K must be >= 2")
    if K > T:
        K = T
    Cc = L.f32c(C)
    idx = torch.empty((B, K), device=dev, dtype=torch.long)
    status = torch.empty((B,), device=dev, dtype=torch.int32)
    L.call("idb200_dp_select", Cc.data_ptr(), B, T, K, idx.data_ptr(), status.data_ptr(), L.stream(dev))
    st = int(status.max().item()) if B > 0 else 0                  # the reference raises here too (:216, :223)
    if st == 1:
        raise RuntimeError("DP failed to find a valid path to T-1 for some samples.")
    if st == 2:
        raise RuntimeError("DP backtrack failed.")
    return idx


def dp_select_indices(C: torch.Tensor, K: int) -> torch.Tensor:
    """Single cost matrix [T, T] -> idx [K] (:171-197)."""
    if C.dim() != 2:
        raise ValueError("C must be [T,T]")
    try:
        return dp_select_indices_batch(C.unsqueeze(0), K)[0]
    except RuntimeError as e:
        raise RuntimeError(str(e).replace(" for some samples", "")) from None


# ----------------------------------------------------------------------------------------------------------------------
# Segment bookkeeping of the DP placement (epiplexity_dp.py:11-168, 231-258).  Index construction is host-side tensor logic
# (vectorised here; the reference builds the same tables with python loops); the segment costs are one kernel.
from dataclasses import dataclass
from typing import Tuple

from ..diffusion.schedules import make_alpha_bars, make_beta_schedule


@dataclass
class SegmentPrecompute:
    seg_i: torch.Tensor          # i64 [S]   S = T (T - 1) / 2 segments (i < j), row-major in (i, j)
    seg_j: torch.Tensor          # i64 [S]
    seg_len: torch.Tensor        # i64 [S]   j - i
    t_idx: torch.Tensor          # i64 [S, n] sample points strictly inside the segment (i when the segment has no interior)
    alpha: torch.Tensor          # f32 [S, n] (t - i) / (j - i)
    weight: torch.Tensor         # f32 [S]   interior / n (0 without interior)
    seg_id: torch.Tensor         # i64 [T, T] segment number of (i, j), -1 elsewhere


def build_snr_weights(schedule: str, n_train: int, s_min: float, s_max: float, gamma: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """:22-34 -> (snr [n_train], clipped-SNR weights)."""
    alpha_bar = make_alpha_bars(make_beta_schedule(schedule, n_train))["alpha_bar"]
    snr = alpha_bar / torch.clamp(1.0 - alpha_bar, min=1e-8)
    return snr, torch.clamp(snr, min=s_min, max=s_max).pow(gamma)


def sample_timesteps_log_snr(snr: torch.Tensor, num_steps: int) -> torch.Tensor:
    """:37-47: timesteps whose log-SNR is closest to a uniform grid between the extremes (sorted, unique)."""
    if num_steps <= 1:
        return torch.tensor([0], dtype=torch.long, device=snr.device)
    log_snr = torch.log(torch.clamp(snr, min=1e-12))
    grid = torch.linspace(log_snr.max(), log_snr.min(), num_steps, device=snr.device)
    idx = torch.unique((log_snr[None, :] - grid[:, None]).abs().argmin(dim=1))
    if idx.numel() < num_steps:
        ends = torch.tensor([0, log_snr.shape[0] - 1], dtype=torch.long, device=snr.device)
        idx = torch.unique(torch.cat([idx, ends]))
    return torch.sort(idx).values


def build_segment_precompute(T: int, samples_per_seg: int, device: torch.device) -> SegmentPrecompute:
    """:50-89, vectorised: all (i < j) pairs in row-major order with ``samples_per_seg`` evenly spread interior sample points."""
    n = int(samples_per_seg)
    pairs = torch.triu_indices(T, T, offset=1)
    seg_i, seg_j = pairs[0], pairs[1]
    gap = seg_j - seg_i
    interior = (gap - 1).clamp(min=0)
    frac = (torch.arange(n, dtype=torch.float32) + 0.5) / n                       # fp32, as the reference computes it
    offs = torch.floor(frac[None, :] * interior[:, None].float()).long()
    has = (gap > 1)[:, None]
    t_idx = torch.where(has, seg_i[:, None] + 1 + offs, seg_i[:, None].expand(-1, n))
    alpha = torch.where(has, (t_idx.float() - seg_i[:, None].float()) / gap[:, None].float(), torch.zeros((1, 1)))
    weight = torch.where(gap > 1, interior.float() / float(n), torch.zeros(()))
    seg_id = torch.full((T, T), -1, dtype=torch.long)
    seg_id[seg_i, seg_j] = torch.arange(seg_i.shape[0])
    to = lambda t: t.to(device)
    return SegmentPrecompute(seg_i=to(seg_i), seg_j=to(seg_j), seg_len=to(gap), t_idx=to(t_idx.contiguous()), alpha=to(alpha.contiguous()),
                             weight=to(weight), seg_id=to(seg_id))


def build_segment_features(T: int, seg_i: torch.Tensor, seg_j: torch.Tensor) -> torch.Tensor:
    """:92-97 -> [S, 3] = (i, j, j - i) / (T - 1)."""
    denom = float(max(1, T - 1))
    i, j = seg_i.float(), seg_j.float()
    return torch.stack([i / denom, j / denom, (j - i) / denom], dim=-1)


def build_segment_features_from_idx(idx: torch.Tensor, T: int, seg_feat_dim: int = 3) -> torch.Tensor:
    """:100-117: features of the K - 1 consecutive segments of each anchor set, truncated / zero-padded to seg_feat_dim."""
    if idx.dim() != 2:
        raise ValueError("idx must be [B, K]")
    B, K = idx.shape
    if seg_feat_dim <= 0:
        return torch.zeros((B, K - 1, 0), device=idx.device)
    feat = build_segment_features(T, idx[:, :-1], idx[:, 1:])
    if seg_feat_dim <= 3:
        return feat[:, :, :seg_feat_dim]
    return torch.nn.functional.pad(feat, (0, seg_feat_dim - 3))


def compute_segment_costs_batch(x_pos: torch.Tensor, precomp: SegmentPrecompute, weight_scale: float) -> torch.Tensor:
    """:120-147 -> cost fp32 [B, S] (idb200_segment_costs: one thread per (sample, segment), no [B, S, n, 2] temporaries)."""
    B, T, D = x_pos.shape
    if D < 2:
        raise ValueError("x_pos must have at least 2 dims")
    dev = L.require_cuda(x_pos, precomp.seg_i)
    x = L.f32c(x_pos)
    S, n = precomp.t_idx.shape
    out = torch.empty((B, S), device=dev, dtype=torch.float32)
    L.call("idb200_segment_costs", x.data_ptr(), B, T, D, L.i64c(precomp.seg_i).data_ptr(), L.i64c(precomp.seg_j).data_ptr(),
           L.i64c(precomp.t_idx).data_ptr(), L.f32c(precomp.alpha).data_ptr(), L.f32c(precomp.weight).data_ptr(), S, n,
           float(weight_scale), out.data_ptr(), L.stream(dev))
    return out


def build_cost_matrix_from_segments(cost_seg: torch.Tensor, precomp: SegmentPrecompute, T: int) -> torch.Tensor:
    """:150-156 -> [T, T], inf where there is no segment."""
    C = torch.full((T, T), float("inf"), device=cost_seg.device)
    C[precomp.seg_i, precomp.seg_j] = cost_seg
    return C


def build_cost_matrix_from_segments_batch(cost_seg: torch.Tensor, precomp: SegmentPrecompute, T: int) -> torch.Tensor:
    """:159-168 -> [B, T, T]."""
    if cost_seg.dim() != 2:
        raise ValueError("cost_seg must be [B, S]")
    C = torch.full((cost_seg.shape[0], T, T), float("inf"), device=cost_seg.device)
    C[:, precomp.seg_i, precomp.seg_j] = cost_seg
    return C


def build_kp_feat_batch(idx: torch.Tensor, T: int) -> torch.Tensor:
    """:246-258 -> [B, K, 3] = (gap to the left anchor, gap to the right anchor, position) / (T - 1)."""
    if idx.dim() != 2:
        raise ValueError("idx must be [B,K]")
    B, K = idx.shape
    denom = float(max(1, T - 1))
    feat = torch.zeros((B, K, 3), dtype=torch.float32, device=idx.device)
    feat[:, :, 2] = idx.float() / denom
    if K > 1:
        gaps = (idx[:, 1:] - idx[:, :-1]).float() / denom
        feat[:, 1:, 0] = gaps
        feat[:, :-1, 1] = gaps
    return feat


def build_kp_feat(idx: torch.Tensor, T: int) -> torch.Tensor:
    """:231-243 (single anchor set [K])."""
    return build_kp_feat_batch(idx.unsqueeze(0), T)[0]
