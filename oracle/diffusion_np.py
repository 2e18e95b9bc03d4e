"""numpy restatement of ``src/diffusion/schedules.py`` and ``src/diffusion/ddpm.py``.

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.
All arithmetic is fp32, one rounding per reference torch op.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np

F32 = np.float32


def _linspace_f32(start: float, end: float, steps: int) -> np.ndarray:
    from .keyframes_np import _linspace_f32 as f
    return f(start, end, steps)


def linear_beta_schedule(n: int, beta_start: float = 1e-4, beta_end: float = 2e-2) -> np.ndarray:
    """schedules.py:7-8"""
    return _linspace_f32(beta_start, beta_end, n)


def cosine_beta_schedule(n: int, s: float = 0.008) -> np.ndarray:
    """schedules.py:11-17"""
    x = _linspace_f32(0.0, float(n), n + 1)
    a = (x / F32(n) + F32(s)) / F32(1 + s)
    a = (a * F32(math.pi)) * F32(0.5)
    ac = np.cos(a).astype(F32)
    ac = ac * ac
    ac = ac / ac[0]
    betas = F32(1.0) - (ac[1:] / ac[:-1])
    return np.clip(betas, F32(1e-8), F32(0.999)).astype(F32)


def make_beta_schedule(name: str, n: int) -> np.ndarray:
    """schedules.py:20-25"""
    if name == "linear":
        return linear_beta_schedule(n)
    if name == "cosine":
        return cosine_beta_schedule(n)
    raise ValueError(f"Unknown schedule {name}")


def make_alpha_bars(betas: np.ndarray) -> Dict[str, np.ndarray]:
    """schedules.py:28-39"""
    betas = betas.astype(F32)
    alphas = (F32(1.0) - betas).astype(F32)
    # torch.cumprod on CPU accumulates in acc_type<float> = double and rounds each output to fp32
    alpha_bar = np.cumprod(alphas.astype(np.float64)).astype(F32)
    return {
        "betas": betas,
        "alphas": alphas,
        "alpha_bar": alpha_bar,
        "sqrt_alpha_bar": np.sqrt(alpha_bar).astype(F32),
        "sqrt_one_minus_alpha_bar": np.sqrt(F32(1.0) - alpha_bar).astype(F32),
    }


def timesteps(n_train: int, steps: int, schedule: str = "linear") -> np.ndarray:
    """ddpm.py:79-99 (``_timesteps``) -> int64, descending, always contains 0 and n_train-1."""
    if steps <= 1:
        return np.array([n_train - 1, 0], dtype=np.int64)
    if steps >= n_train:
        return np.arange(n_train - 1, -1, -1, dtype=np.int64)
    if schedule == "quadratic":
        t = _linspace_f32(0.0, 1.0, steps)
        times = ((t * t) * F32(n_train - 1)).astype(np.int64)      # .long() truncates
    elif schedule == "sqrt":
        t = _linspace_f32(0.0, 1.0, steps)
        times = (np.sqrt(t).astype(F32) * F32(n_train - 1)).astype(np.int64)
    else:
        times = _linspace_f32(0.0, float(n_train - 1), steps).astype(np.int64)
    times = np.unique(times)
    if times[0] != 0:
        times = np.concatenate([[0], times])
    if times[-1] != n_train - 1:
        times = np.concatenate([times, [n_train - 1]])
    return times[::-1].astype(np.int64).copy()


def _gather(vec: np.ndarray, t: np.ndarray, ndim: int) -> np.ndarray:
    """ddpm.py:6-12 + the unsqueeze loops in each caller."""
    out = vec[t]
    if t.ndim == 1:
        out = out.reshape(-1, 1, 1)
    elif t.ndim == 2:
        out = out.reshape(t.shape[0], t.shape[1], 1)
    else:
        raise ValueError("t must be 1D or 2D")
    while out.ndim < ndim:
        out = out[..., None]
    return out


def q_sample(r0: np.ndarray, t: np.ndarray, schedule: Dict[str, np.ndarray],
             noise: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """ddpm.py:15-24 (noise is an input here)."""
    sab = _gather(schedule["sqrt_alpha_bar"], t, r0.ndim)
    som = _gather(schedule["sqrt_one_minus_alpha_bar"], t, r0.ndim)
    return (sab * r0.astype(F32) + som * noise.astype(F32)).astype(F32), noise


def ddim_coefficients(alpha_bar: np.ndarray, t: int, t_prev: int) -> Tuple[F32, F32, F32, F32]:
    """The four fp32 scalars ddpm.py:45-47 computes from the table entries:
    sqrt(1-ab_t), sqrt(ab_t), sqrt(ab_prev), sqrt(1-ab_prev)."""
    ab_t = F32(alpha_bar[t])
    ab_p = F32(alpha_bar[t_prev])
    return (np.sqrt(F32(1.0) - ab_t).astype(F32), np.sqrt(ab_t).astype(F32),
            np.sqrt(ab_p).astype(F32), np.sqrt(F32(1.0) - ab_p).astype(F32))


def ddim_step(rt: np.ndarray, eps: np.ndarray, t: np.ndarray, t_prev: np.ndarray,
              schedule: Dict[str, np.ndarray], eta: float = 0.0,
              noise: Optional[np.ndarray] = None) -> np.ndarray:
    """ddpm.py:37-57.  eta=0 is the hot path; the stochastic branch takes ``noise`` as input."""
    rt = rt.astype(F32)
    eps = eps.astype(F32)
    ab_t = _gather(schedule["alpha_bar"], t, rt.ndim).astype(F32)
    ab_p = _gather(schedule["alpha_bar"], t_prev, rt.ndim).astype(F32)
    x0 = (rt - np.sqrt(F32(1.0) - ab_t) * eps) / np.sqrt(ab_t)
    if eta == 0.0:
        return (np.sqrt(ab_p) * x0 + np.sqrt(F32(1.0) - ab_p) * eps).astype(F32)
    sigma = (F32(eta) * np.sqrt((F32(1.0) - ab_p) / (F32(1.0) - ab_t))) * np.sqrt(F32(1.0) - ab_t / ab_p)
    return (np.sqrt(ab_p) * x0 + np.sqrt(F32(1.0) - ab_p - sigma * sigma) * eps
            + sigma * noise.astype(F32)).astype(F32)
