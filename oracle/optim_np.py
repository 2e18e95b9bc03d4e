"""CPU restatement (numpy) of the tail of the reference's Stage-2 training step -- TEST INFRASTRUCTURE ONLY.

Follows ``src/train/train_interp_levels.py:1142-1173``: weighted MSE ``(diff * w).sum() / (w.sum() * D + 1e-8) / grad_accum``
with ``w = w_missing + (w_anchor - w_missing) * conf`` (``anchor_conf`` branch) or ``where(mask, w_anchor, w_missing)``;
``torch.nn.utils.clip_grad_norm_`` (global 2-norm, ``clip_coef = max_norm / (total_norm + 1e-6)`` clamped to 1);
``torch.optim.AdamW`` (decoupled weight decay, torch's single-tensor operation order, defaults betas (0.9, 0.999), eps 1e-8);
``EMA.update`` (``src/utils/ema.py:11-17``).  Pinned against ``tests/golden/optim.npz`` (live torch, ``make_golden_optim.py``).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import numpy as np

F32 = np.float32


def stage2_loss_and_grad(delta_hat: np.ndarray, target: np.ndarray, weight_mask: np.ndarray, anchor_conf: bool = True,
                         w_anchor: float = 0.1, w_missing: float = 1.0, grad_accum: int = 1) -> Tuple[float, np.ndarray]:
    """train_interp_levels.py:1144-1156 and d(loss)/d(delta_hat) (what loss.backward() seeds the model's backward with)."""
    dh, tg = delta_hat.astype(np.float64), target.astype(np.float64)
    D = delta_hat.shape[-1]
    if anchor_conf:
        w = float(w_missing) + (float(w_anchor) - float(w_missing)) * weight_mask.astype(np.float64)
    else:
        w = np.where(weight_mask.astype(bool), float(w_anchor), float(w_missing))
    den = w.sum() * D + 1e-8
    loss = (((dh - tg) ** 2).sum(axis=-1) * w).sum() / den / grad_accum
    grad = 2.0 * (dh - tg) * w[..., None] / den / grad_accum
    return float(loss), grad.astype(F32)


def clip_coef(grads: List[np.ndarray], max_norm: float) -> Tuple[float, float]:
    """clip_grad_norm_: total 2-norm over all tensors, coefficient clamped to 1."""
    total = float(np.sqrt(sum(float((g.astype(np.float64) ** 2).sum()) for g in grads)))
    return total, min(1.0, float(max_norm) / (total + 1e-6))


def adamw_ema_step(p: np.ndarray, g: np.ndarray, m: np.ndarray, v: np.ndarray, ema: Optional[np.ndarray], step: int, lr: float = 2e-4,
                   beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-2, ema_decay: float = 0.999,
                   coef: float = 1.0):
    """One AdamW step (torch/optim/adamw.py _single_tensor_adamw order, fp32) on the clipped gradient g * coef, then EMA.update.
    ``step`` is 1-based.  Returns (p, m, v, ema)."""
    p, g, m, v = p.astype(F32), (g.astype(F32) * F32(coef)).astype(F32), m.astype(F32), v.astype(F32)
    p = p * F32(1.0 - lr * weight_decay)
    m = m + (g - m) * F32(1.0 - beta1)                                   # exp_avg.lerp_(grad, 1 - beta1)
    v = v * F32(beta2) + (g * g) * F32(1.0 - beta2)                      # mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    denom = np.sqrt(v) / F32(np.sqrt(bc2)) + F32(eps)
    p = p - F32(step_size) * (m / denom)
    if ema is not None:
        ema = ema.astype(F32) * F32(ema_decay) + p * F32(1.0 - ema_decay)
    return p.astype(F32), m.astype(F32), v.astype(F32), None if ema is None else ema.astype(F32)
