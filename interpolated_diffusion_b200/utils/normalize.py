"""Mirror of the reference's ``src/utils/normalize.py`` (logit / sigmoid on the position dims 0:2)."""
import torch

from .. import _lib as L


def logit_pos(x: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """normalize.py:4-11"""
    if x.shape[-1] < 2:
        return x
    dev = L.require_cuda(x)
    xc = L.f32c(x)
    out = torch.empty_like(xc)
    L.call("idb200_logit_pos", L.ptr(xc), xc.numel() // xc.shape[-1], xc.shape[-1], float(eps), L.ptr(out), L.stream(dev))
    return out


def sigmoid_pos(x: torch.Tensor) -> torch.Tensor:
    """normalize.py:14-20"""
    if x.shape[-1] < 2:
        return x
    dev = L.require_cuda(x)
    xc = L.f32c(x)
    out = torch.empty_like(xc)
    L.call("idb200_sigmoid_pos", L.ptr(xc), xc.numel() // xc.shape[-1], xc.shape[-1], L.ptr(out), L.stream(dev))
    return out
