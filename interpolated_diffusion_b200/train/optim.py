"""Tail of the reference's Stage-2 training step on libidb200 kernels (``src/train/train_interp_levels.py:1142-1173``):

* ``stage2_loss``      -- the weighted MSE of :1144-1156 and its gradient with respect to the denoiser output (the seed of the
                          backward pass; the network's backward itself is not part of this library yet);
* ``FlatAdamW``        -- ``clip_grad_norm_`` + ``torch.optim.AdamW.step`` + ``EMA.update`` (:1162-1173, ``src/utils/ema.py``) as
                          two launches over ONE flat fp32 arena holding every parameter (36 B of HBM traffic per parameter).

CUDA only.  Same hyper-parameter names and defaults as the reference's call sites (lr 2e-4, weight_decay 1e-2, betas (0.9, 0.999),
eps 1e-8, EMA decay 0.999, grad_clip 1.0)."""
from __future__ import annotations

from typing import Iterable, List, Optional, Tuple

import torch

from .. import _lib as L



def _scratch_doubles() -> int:
    """fp64 partial-sum slots the loss / grad-norm reductions need: asked of the library (their grids are clamped to it), not
    derived from an assumed SM count."""
    return int(L.lib().idb200_tail_scratch_doubles())


def stage2_loss(delta_hat: torch.Tensor, target: torch.Tensor, weight_mask: torch.Tensor, *, anchor_conf: bool = True,
                w_anchor: float = 0.1, w_missing: float = 1.0, grad_accum: int = 1, want_grad: bool = True
                ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """loss (0-dim fp32 tensor, no host sync) and d loss / d delta_hat [B,T,D] (train_interp_levels.py:1144-1156).
    ``weight_mask``: conf fp32 [B,T] when ``anchor_conf`` else the bool anchor mask [B,T]."""
    dev = L.require_cuda(delta_hat, target, weight_mask)
    dh, tg = L.f32c(delta_hat), L.f32c(target)
    if dh.shape != tg.shape or dh.dim() != 3:
        raise ValueError("delta_hat and target must both be [B,T,D]")
    B, T, D = dh.shape
    if tuple(weight_mask.shape) != (B, T):
        raise ValueError("weight_mask must be [B,T]")
    conf = L.f32c(weight_mask) if anchor_conf else None
    mask = None if anchor_conf else L.u8c(weight_mask)
    scratch = torch.empty((_scratch_doubles(),), device=dev, dtype=torch.float64)
    scal = torch.empty((2,), device=dev, dtype=torch.float32)
    grad = torch.empty_like(dh) if want_grad else None
    L.call("idb200_stage2_loss", dh.data_ptr(), tg.data_ptr(), L.ptr(conf), L.ptr(mask), float(w_anchor), float(w_missing),
           float(grad_accum), B, T, D, scratch.data_ptr(), scal.data_ptr(), L.ptr(grad), L.stream(dev))
    return scal[0], grad


class FlatAdamW:
    """``torch.optim.AdamW`` + ``clip_grad_norm_`` + ``EMA`` over one flat fp32 arena.  The parameters are re-pointed at views
    of the arena (so the model keeps working on them); ``step()`` gathers ``p.grad`` (or takes a flat gradient, e.g. the output
    of a reduce-scatter / all-reduce) and runs two launches: global-norm clip coefficient, fused AdamW + EMA."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, ema_decay: Optional[float] = 0.999, max_grad_norm: Optional[float] = 1.0,
                 early: Optional[Iterable[torch.nn.Parameter]] = None):
        """``early``: parameters to place FIRST in the arena (``[0, n_early)``), e.g. those whose gradients are final before the
        rest of the backward -- a data-parallel step can all-reduce that slice while the backward's tail still runs.  Only the
        arena layout changes: ``params`` (and with it ``state_dict()``) keeps the order given."""
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = L.require_cuda(*[p.data for p in self.params])
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), float(weight_decay)
        self.ema_decay, self.max_grad_norm = ema_decay, max_grad_norm
        early_ids = {id(p) for p in early} if early is not None else set()
        self.offsets, n = [0] * len(self.params), 0
        for want_early in (True, False):
            for i, p in enumerate(self.params):
                if (id(p) in early_ids) == want_early:
                    self.offsets[i] = n
                    n += (p.numel() + 3) // 4 * 4           # 16-byte aligned slices
            if want_early:
                self.n_early = n
        self.n = n
        self.flat = torch.zeros((n,), device=dev, dtype=torch.float32)
        for p, o in zip(self.params, self.offsets):
            self.flat[o:o + p.numel()].copy_(p.data.detach().float().reshape(-1))
            p.data = self.flat[o:o + p.numel()].view(p.shape)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema = self.flat.clone() if ema_decay is not None else None
        self.grad = torch.zeros_like(self.flat)
        self.scratch = torch.empty((_scratch_doubles(),), device=dev, dtype=torch.float64)
        self.norm_coef = torch.ones((2,), device=dev, dtype=torch.float32)
        self.step_count = 0

    def sync_replicas(self, group=None, src: int = 0) -> None:
        """Data parallel: broadcast rank ``src``'s parameters, EMA, Adam moments and step count to every rank (call after
        construction -- the trainers do -- and after ``load_state_dict`` / a checkpoint load on any subset of ranks)."""
        from .. import parallel as P
        if P.world_size(group) == 1:
            return
        self.step_count = P.broadcast_replica_state([self.flat, self.ema, self.exp_avg, self.exp_avg_sq], self.step_count, group, src)
        L.PARAM_EPOCH += 1                                  # the parameters may have changed behind torch's version counters

    def views(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    @property
    def ema_shadow(self) -> Optional[List[torch.Tensor]]:
        """The EMA copies, one per parameter (``EMA.shadow`` of src/utils/ema.py)."""
        return None if self.ema is None else self.views(self.ema)

    def gather_grads(self) -> torch.Tensor:
        """Copies ``p.grad`` into the flat arena.  A parameter whose grad is None contributes ZEROS, so the fused step still
        applies weight decay / moment decay to it and its step count advances with the arena's -- ``torch.optim.AdamW`` would
        skip it.  On this path every parameter of both denoisers receives a gradient each step (tested:
        test_gpu_backward), so the two agree; a model with unused parameters should exclude them from ``params``."""
        for p, g in zip(self.params, self.views(self.grad)):
            if p.grad is None:
                g.zero_()
            else:
                g.copy_(p.grad)
        return self.grad

    @torch.no_grad()
    def step(self, flat_grad: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
        """Returns the total gradient norm (device tensor) when clipping is on, like ``clip_grad_norm_``."""
        g = self.gather_grads() if flat_grad is None else flat_grad
        if g.numel() != self.n or g.dtype != torch.float32 or not g.is_contiguous():
            raise ValueError("flat_grad must be a contiguous fp32 tensor with the arena's layout")
        dev = self.flat.device
        self.step_count += 1
        coef = None
        if self.max_grad_norm is not None:
            L.call("idb200_grad_clip_coef", g.data_ptr(), self.n, float(self.max_grad_norm), self.scratch.data_ptr(),
                   self.norm_coef.data_ptr(), L.stream(dev))
            coef = self.norm_coef
        L.call("idb200_adamw_ema_step", self.flat.data_ptr(), g.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
               L.ptr(self.ema), self.n, self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count,
               float(self.ema_decay if self.ema_decay is not None else 0.0), L.ptr(coef), L.stream(dev))
        L.PARAM_EPOCH += 1                                  # the kernel wrote the parameters behind torch's version counters
        return None if coef is None else self.norm_coef[0]

    # ---- torch.optim.AdamW-compatible state (src/utils/checkpoint.py:21-22, 41-42 save / restore ``optimizer.state_dict()``) -------
    def state_dict(self) -> dict:
        """Same structure as ``torch.optim.AdamW(params).state_dict()``: per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``
        (copies) and one param group, so a checkpoint written here resumes under the reference's optimizer and vice versa."""
        group = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(1))], lr=self.lr, betas=self.betas, eps=self.eps,
                                  weight_decay=self.weight_decay).state_dict()["param_groups"][0]
        group["params"] = list(range(len(self.params)))
        state = {}
        if self.step_count > 0:
            for i, (m, v) in enumerate(zip(self.views(self.exp_avg), self.views(self.exp_avg_sq))):
                state[i] = {"step": torch.tensor(float(self.step_count)), "exp_avg": m.clone(), "exp_avg_sq": v.clone()}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        groups = sd["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("loaded state dict has a different number of parameter groups / parameters")
        g = groups[0]
        self.lr, self.betas, self.eps, self.weight_decay = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"]), float(g["weight_decay"])
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        steps = set()
        for i, (m, v) in enumerate(zip(self.views(self.exp_avg), self.views(self.exp_avg_sq))):
            st = sd["state"].get(i)
            if st is None:
                continue
            m.copy_(st["exp_avg"])
            v.copy_(st["exp_avg_sq"])
            steps.add(int(float(st["step"])))
        if len(steps) > 1:
            raise ValueError("per-parameter step counts differ; the fused step keeps one count for the whole arena")
        self.step_count = steps.pop() if steps else 0

    @property
    def ema_state(self) -> "EMAState":
        """An object with the reference EMA's ``state_dict`` / ``load_state_dict`` / ``copy_to`` (src/utils/ema.py:19-35)."""
        if self.ema is None:
            raise ValueError("this optimizer was built without an EMA")
        return EMAState(self)

    def zero_grad(self, set_to_none: bool = True):
        for p in self.params:
            p.grad = None if set_to_none else (p.grad.zero_() if p.grad is not None else None)


class EMAState:
    """View of the EMA arena with the interface of ``src/utils/ema.py``'s ``EMA`` (what save_checkpoint / load_checkpoint use)."""

    def __init__(self, opt: FlatAdamW):
        self.opt = opt

    @property
    def decay(self) -> float:
        return float(self.opt.ema_decay)

    @property
    def shadow(self) -> List[torch.Tensor]:
        return self.opt.ema_shadow

    def copy_to(self, parameters: Iterable[torch.nn.Parameter]) -> None:
        for p, s in zip([p for p in parameters if p.requires_grad], self.shadow):
            p.data.copy_(s)
        L.PARAM_EPOCH += 1

    def state_dict(self) -> dict:
        return {"decay": self.decay, "shadow": [t.clone() for t in self.shadow]}

    def load_state_dict(self, state: dict) -> None:
        self.opt.ema_decay = state["decay"]
        for dst, src in zip(self.shadow, state["shadow"]):
            dst.copy_(src)
