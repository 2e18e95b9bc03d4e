"""Noise schedules with the surface of the reference's ``src/diffusion/schedules.py`` (``make_beta_schedule``, ``make_alpha_bars``).

The tables (N_train = 1000 entries) are a one-off host computation; the kernels take ``alpha_bar`` (or two of its entries) as
inputs.  Every table must equal the reference's bit for bit (DDIM parity is exact arithmetic on them), so each value is produced
by the same sequence of torch primitives (``linspace`` -> ``cos`` -> square -> normalise -> ratio -> ``clip``; ``cumprod`` ->
``sqrt``); ``tests/test_oracle_golden.py`` pins a few entries against the live reference."""
import math
from typing import Callable, Dict

import torch

_LINEAR_RANGE = (1e-4, 2e-2)
_BETA_CLIP = (1e-8, 0.999)


def linear_beta_schedule(n_timesteps: int, beta_start: float = _LINEAR_RANGE[0], beta_end: float = _LINEAR_RANGE[1]) -> torch.Tensor:
    """schedules.py:7-8: evenly spaced betas."""
    return torch.linspace(beta_start, beta_end, n_timesteps)


def _cosine_signal_level(n_timesteps: int, s: float) -> torch.Tensor:
    """Cumulative signal level cos^2(((t / N) + s) / (1 + s) * pi / 2) on t = 0..N, normalised to 1 at t = 0."""
    grid = torch.linspace(0, n_timesteps, n_timesteps + 1)
    level = torch.cos(((grid / n_timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    return level / level[0]


def cosine_beta_schedule(n_timesteps: int, s: float = 0.008) -> torch.Tensor:
    """schedules.py:11-17 (Nichol & Dhariwal): beta_t = 1 - level_t / level_{t-1}, clipped."""
    level = _cosine_signal_level(n_timesteps, s)
    return torch.clip(1 - (level[1:] / level[:-1]), *_BETA_CLIP)


_BUILDERS: Dict[str, Callable[[int], torch.Tensor]] = {"linear": linear_beta_schedule, "cosine": cosine_beta_schedule}


def make_beta_schedule(name: str, n_timesteps: int) -> torch.Tensor:
    """schedules.py:20-25"""
    builder = _BUILDERS.get(name)
    if builder is None:
        raise ValueError(f"Unknown schedule {name}")
    return builder(n_timesteps)


def make_alpha_bars(betas: torch.Tensor) -> Dict[str, torch.Tensor]:
    """schedules.py:28-39: the five tables the samplers index by timestep."""
    keep = 1.0 - betas
    table = {"betas": betas, "alphas": keep, "alpha_bar": torch.cumprod(keep, dim=0)}
    table["sqrt_alpha_bar"] = torch.sqrt(table["alpha_bar"])
    table["sqrt_one_minus_alpha_bar"] = torch.sqrt(1.0 - table["alpha_bar"])
    return table
