"""Mirror of the reference's ``src/utils/clamp.py`` on ``idb200_stage2_epilogue`` (CUDA only).
Like the reference, ``clamp_dims == "pos"`` mutates ``x_hat`` in place and returns it."""
from typing import Optional

import torch

from .. import _lib as L

CLAMP_NONE, CLAMP_ENDPOINTS, CLAMP_MASK = 0, 1, 2


def _epilogue(x_in, delta, x_ref, conf, lam, policy, clamp_mask, dims_all, pos_clip, lo, hi, out):
    B, T, D = x_in.shape
    L.call("idb200_stage2_epilogue", L.ptr(x_in), L.ptr(delta), L.ptr(x_ref), L.ptr(conf), float(lam), int(policy),
           L.ptr(clamp_mask), int(dims_all), int(pos_clip), float(lo), float(hi), B, T, D, L.ptr(out),
           L.stream(x_in.device))
    return out


def _inplace_ok(x: torch.Tensor) -> bool:
    return x.dtype == torch.float32 and x.is_contiguous()


def apply_clamp(x_hat: torch.Tensor, x_ref: torch.Tensor, clamp_mask: torch.Tensor, clamp_dims: str) -> torch.Tensor:
    """clamp.py:4-10"""
    if clamp_mask is None:
        return x_hat
    L.require_cuda(x_hat, x_ref, clamp_mask)
    xr, cm = L.f32c(x_ref), L.u8c(clamp_mask)
    if clamp_dims == "pos":
        if _inplace_ok(x_hat):
            return _epilogue(x_hat, None, xr, None, 0.0, CLAMP_MASK, cm, 0, 0, 0.0, 0.0, x_hat)
        res = _epilogue(L.f32c(x_hat), None, xr, None, 0.0, CLAMP_MASK, cm, 0, 0, 0.0, 0.0, torch.empty_like(xr))
        x_hat.copy_(res)
        return x_hat
    xh = L.f32c(x_hat)
    return _epilogue(xh, None, xr, None, 0.0, CLAMP_MASK, cm, 1, 0, 0.0, 0.0, torch.empty_like(xh))


def apply_soft_clamp(x_hat: torch.Tensor, x_ref: torch.Tensor, conf: torch.Tensor, lam: float,
                     clamp_dims: str) -> torch.Tensor:
    """clamp.py:13-32 (conf is [B,T]; a [B,T,1] conf is accepted like the reference's dim()==3 branch)."""
    if conf is None:
        return x_hat
    if lam <= 0.0:
        return x_hat
    L.require_cuda(x_hat, x_ref, conf)
    if conf.dim() == 3:
        if conf.shape[-1] != 1:
            raise ValueError("per-dimension conf is not supported by the fused epilogue; pass conf as [B,T] or [B,T,1]")
        conf = conf[..., 0]
    xr, cf = L.f32c(x_ref), L.f32c(conf)
    if clamp_dims == "pos":
        if _inplace_ok(x_hat):
            return _epilogue(x_hat, None, xr, cf, lam, CLAMP_NONE, None, 0, 0, 0.0, 0.0, x_hat)
        res = _epilogue(L.f32c(x_hat), None, xr, cf, lam, CLAMP_NONE, None, 0, 0, 0.0, 0.0, torch.empty_like(xr))
        x_hat.copy_(res)
        return x_hat
    xh = L.f32c(x_hat)
    return _epilogue(xh, None, xr, cf, lam, CLAMP_NONE, None, 1, 0, 0.0, 0.0, torch.empty_like(xh))


def stage2_epilogue(x_in: torch.Tensor, delta: Optional[torch.Tensor], x_ref: torch.Tensor, conf: Optional[torch.Tensor],
                    lam: float, clamp_policy: str, clamp_mask: Optional[torch.Tensor], clamp_dims: str,
                    pos_clip: bool = False, pos_clip_min: float = 0.0, pos_clip_max: float = 1.0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The whole Stage-2 tail of sample_generate.py:1252-1285 / :1160-1204 in one launch:
    x = x_in + delta; soft clamp; hard clamp by ``clamp_policy`` in {none, endpoints, all_anchors}; pos_clip."""
    L.require_cuda(x_in, delta, x_ref, conf, clamp_mask)
    if clamp_policy == "all_anchors":
        policy = CLAMP_MASK
        if clamp_mask is None:
            raise ValueError("clamp_policy=all_anchors needs the anchor mask")
    elif clamp_policy == "endpoints":
        policy = CLAMP_ENDPOINTS
    elif clamp_policy == "none":
        policy = CLAMP_NONE
    else:
        raise ValueError(f"Unknown clamp_policy: {clamp_policy}")
    xi = L.f32c(x_in)
    if out is None:
        out = torch.empty_like(xi)
    return _epilogue(xi, None if delta is None else L.f32c(delta), L.f32c(x_ref), None if conf is None else L.f32c(conf),
                     lam, policy, None if policy != CLAMP_MASK else L.u8c(clamp_mask), int(clamp_dims != "pos"),
                     int(bool(pos_clip)), pos_clip_min, pos_clip_max, out)
