"""Mirror of the reference's ``src/diffusion/ddpm.py`` on libidb200 kernels (CUDA only).

``ddim_step`` (eta = 0) and ``q_sample`` are single fused launches whose fp32 arithmetic is rounded op by
op in the reference's order with IEEE sqrt / div (``idb200_ddim_step``, ``idb200_q_sample``)."""
from typing import Dict, Optional

import torch

from .. import _lib as L


def _row_shape(x: torch.Tensor, t: torch.Tensor):
    """rows of the gathered coefficient: one per t entry (ddpm.py:6-12 + the unsqueeze loops)."""
    if t.dim() not in (1, 2):
        raise ValueError("t must be 1D or 2D")
    if tuple(x.shape[: t.dim()]) != tuple(t.shape):
        raise ValueError(f"t shape {tuple(t.shape)} does not match the leading dims of {tuple(x.shape)}")
    n_rows = t.numel()
    row_len = x.numel() // max(n_rows, 1)
    return n_rows, row_len


def _table(schedule: Dict[str, torch.Tensor], key: str, dev: torch.device) -> torch.Tensor:
    tab = schedule[key]
    if tab.device != dev or tab.dtype != torch.float32 or not tab.is_contiguous():
        tab = tab.to(device=dev, dtype=torch.float32).contiguous()
    return tab


def _check_timesteps(t: torch.Tensor, t_prev: Optional[torch.Tensor], n_train: int) -> None:
    """The reference's table gather raises IndexError for a timestep outside [0, n_train) (ddpm.py:6-12); the kernels clamp
    the table index instead of reading out of bounds.  ``check_t=True`` restores the reference's error at the cost of one
    host sync (off on the hot path, where the timesteps come from ``_timesteps``)."""
    lo = int(t.min().item()) if t.numel() else 0
    hi = int(t.max().item()) if t.numel() else 0
    if t_prev is not None and t_prev.numel():
        lo, hi = min(lo, int(t_prev.min().item())), max(hi, int(t_prev.max().item()))
    if lo < -n_train or hi >= n_train:
        raise IndexError(f"timestep out of range for a schedule of {n_train} steps (got [{lo}, {hi}])")


def q_sample(r0: torch.Tensor, t: torch.Tensor, schedule: Dict[str, torch.Tensor], noise: Optional[torch.Tensor] = None, *,
             check_t: bool = False):
    """ddpm.py:15-24"""
    dev = L.require_cuda(r0, t)
    if noise is None:
        noise = torch.randn_like(r0)
    r0c, nc, tc = L.f32c(r0), L.f32c(noise), L.i64c(t)
    n_rows, row_len = _row_shape(r0c, tc)
    sab = _table(schedule, "sqrt_alpha_bar", dev)
    s1m = _table(schedule, "sqrt_one_minus_alpha_bar", dev)
    if check_t:
        _check_timesteps(tc, None, sab.numel())
    out = torch.empty_like(r0c)
    L.call("idb200_q_sample", L.ptr(r0c), L.ptr(nc), L.ptr(tc), L.ptr(sab), L.ptr(s1m), sab.numel(), n_rows, row_len,
           L.ptr(out), L.stream(dev))
    return out, noise


def ddim_step(rt: torch.Tensor, eps: torch.Tensor, t: torch.Tensor, t_prev: torch.Tensor,
              schedule: Dict[str, torch.Tensor], eta: float = 0.0, *, known_mask: Optional[torch.Tensor] = None,
              known_values: Optional[torch.Tensor] = None, pos_clip: bool = False, pos_clip_min: float = 0.0,
              pos_clip_max: float = 1.0, out: Optional[torch.Tensor] = None, check_t: bool = False) -> torch.Tensor:
    """ddpm.py:37-57.  eta == 0 (the hot path) is one launch, optionally fused with the known-value clamp of
    sample_generate.py:397-399 (keyword-only extras).  The stochastic branch (eta != 0, unused on the hot
    path) composes the same kernel with the reference's sigma formula."""
    dev = L.require_cuda(rt, eps, t, t_prev)
    rtc, ec, tc, tpc = L.f32c(rt), L.f32c(eps), L.i64c(t), L.i64c(t_prev)
    n_rows, row_len = _row_shape(rtc, tc)
    ab = _table(schedule, "alpha_bar", dev)
    if check_t:
        _check_timesteps(tc, tpc, ab.numel())
    if eta != 0.0:
        # stochastic branch (ddpm.py:50-57; not on the hot path): torch composition with the reference's sigma formula,
        # followed by the same keyword extras the fused kernel applies (known-value clamp, position clip, out=)
        shape = list(tc.shape) + [1] * (rtc.dim() - tc.dim())
        ab_t, ab_p = ab[tc].view(shape), ab[tpc].view(shape)
        x0 = (rtc - torch.sqrt(1.0 - ab_t) * ec) / torch.sqrt(ab_t)
        sigma = eta * torch.sqrt((1.0 - ab_p) / (1.0 - ab_t)) * torch.sqrt(1.0 - ab_t / ab_p)
        noise = torch.randn_like(rtc)
        res = torch.sqrt(ab_p) * x0 + torch.sqrt(1.0 - ab_p - sigma ** 2) * ec + sigma * noise
        if known_mask is not None:
            if known_values is None:
                raise ValueError("known_mask needs known_values")
            res = torch.where(known_mask, L.f32c(known_values), res)
        if pos_clip:
            res[..., :2] = res[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
        if out is not None:
            out.copy_(res)
            return out
        return res
    if out is None:
        out = torch.empty_like(rtc)
    km = L.u8c(known_mask) if known_mask is not None else None
    kv = L.f32c(known_values) if known_values is not None else None
    L.call("idb200_ddim_step", L.ptr(rtc), L.ptr(ec), L.ptr(tc), L.ptr(tpc), L.ptr(ab), ab.numel(), 0.0, 0.0, n_rows,
           row_len, rtc.shape[-1], L.ptr(km), L.ptr(kv), int(bool(pos_clip)), float(pos_clip_min), float(pos_clip_max),
           L.ptr(out), L.stream(dev))
    return out


def ddim_step_scalar(rt: torch.Tensor, eps: torch.Tensor, ab_t: float, ab_prev: float, *,
                     known_mask: Optional[torch.Tensor] = None, known_values: Optional[torch.Tensor] = None,
                     pos_clip: bool = False, pos_clip_min: float = 0.0, pos_clip_max: float = 1.0,
                     out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batch-constant timestep variant used by the generation loop (t is the same for every row at
    sample_generate.py:394-395): the two table entries travel as kernel arguments, so the launch is
    graph-capturable with the step baked in."""
    dev = L.require_cuda(rt, eps)
    rtc, ec = L.f32c(rt), L.f32c(eps)
    if out is None:
        out = torch.empty_like(rtc)
    km = L.u8c(known_mask) if known_mask is not None else None
    kv = L.f32c(known_values) if known_values is not None else None
    L.call("idb200_ddim_step", L.ptr(rtc), L.ptr(ec), None, None, None, 0, float(ab_t), float(ab_prev), 1, rtc.numel(),
           rtc.shape[-1], L.ptr(km), L.ptr(kv), int(bool(pos_clip)), float(pos_clip_min), float(pos_clip_max), L.ptr(out),
           L.stream(dev))
    return out


def _timesteps(n_train: int, steps: int, schedule: str = "linear") -> torch.Tensor:
    """ddpm.py:79-99 (host-side int64 list; same torch op sequence)."""
    device = torch.device("cpu")
    if steps <= 1:
        return torch.tensor([n_train - 1, 0], dtype=torch.long, device=device)
    if steps >= n_train:
        return torch.arange(n_train - 1, -1, -1, dtype=torch.long, device=device)
    if schedule == "quadratic":
        t = torch.linspace(0.0, 1.0, steps, device=device)
        times = (t * t * (n_train - 1)).long()
    elif schedule == "sqrt":
        t = torch.linspace(0.0, 1.0, steps, device=device)
        times = (torch.sqrt(t) * (n_train - 1)).long()
    else:
        times = torch.linspace(0, n_train - 1, steps, device=device).long()
    times = torch.unique(times)
    if times[0].item() != 0:
        times = torch.cat([torch.tensor([0], dtype=torch.long, device=device), times], dim=0)
    if times[-1].item() != n_train - 1:
        times = torch.cat([times, torch.tensor([n_train - 1], dtype=torch.long, device=device)], dim=0)
    return torch.flip(times, dims=[0])
