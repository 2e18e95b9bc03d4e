"""Stage-1 (keypoint denoiser) training step of the reference (``src/train/train_keypoints.py:505-556``) on libidb200:
keypoint gather + known-endpoint masks (``_build_keypoint_batch`` :122-142), ``q_sample``, known-value clamp of z_t / eps
(:528-530), KeypointDenoiser forward, masked eps-MSE (:532-537), hand-written backward (``train/backward.py``), data-parallel
all-reduce, clip, fused AdamW + EMA.  Names and defaults follow the reference's CLI (:30-90); the selector / DP-index / d_phi
feature branches (:470-521) are input producers outside this path and arrive through ``idx_override`` / ``cond['kp_feat']``."""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import _lib as L
from .. import parallel as P
from ..corruptions.keyframes import sample_fixed_k_indices_batch
from ..diffusion.ddpm import q_sample
from ..diffusion.schedules import make_alpha_bars, make_beta_schedule
from ..sample.sample_generate import _build_known_mask_values
from ..utils.normalize import logit_pos
from .backward import KeypointBackprop
from .optim import FlatAdamW, stage2_loss


def _gather_keypoints(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """train_keypoints.py:93-96"""
    B, K = idx.shape
    return x.gather(1, idx.unsqueeze(-1).expand(B, K, x.shape[-1]))


def _build_keypoint_batch(x0: torch.Tensor, K: int, cond: dict, generator: torch.Generator, logit_space: bool, logit_eps: float,
                          clamp_endpoints: bool, idx_override: Optional[torch.Tensor] = None):
    """train_keypoints.py:122-142 -> (z0, idx, known_mask, known_values)."""
    B, T, D = x0.shape
    if idx_override is None:
        idx, _ = sample_fixed_k_indices_batch(B, T, K, generator=generator, device=x0.device, ensure_endpoints=True)
    else:
        idx = idx_override
    z0 = _gather_keypoints(x0, idx)
    known_mask, known_values = _build_known_mask_values(idx, cond, D, T, clamp_endpoints)
    if logit_space:
        z0 = logit_pos(z0, eps=logit_eps)
        known_values = logit_pos(known_values, eps=logit_eps)
    return z0, idx, known_mask, known_values


class Stage1Trainer:
    def __init__(self, model, *, T: int = 64, K: int = 8, N_train: int = 1000, schedule: str = "cosine", logit_space: bool = True,
                 logit_eps: float = 1e-5, clamp_endpoints: bool = True, lr: float = 2e-4, weight_decay: float = 1e-2,
                 grad_clip: Optional[float] = 1.0, ema: bool = True, ema_decay: float = 0.999, process_group=None,
                 cuda_graph: bool = False):
        self.model = model
        self.T, self.K, self.N_train = T, K, N_train
        self.logit_space, self.logit_eps, self.clamp_endpoints = bool(logit_space), logit_eps, bool(clamp_endpoints)
        dev = next(model.parameters()).device
        self.schedule = {k: v.to(dev) for k, v in make_alpha_bars(make_beta_schedule(schedule, N_train)).items()}
        self.opt = FlatAdamW(model.parameters(), lr=lr, weight_decay=weight_decay, ema_decay=ema_decay if ema else None,
                             max_grad_norm=grad_clip)
        self.bp = KeypointBackprop(model)
        self.flat_grad = torch.zeros_like(self.opt.flat)
        by_id = {id(p): g for p, g in zip(self.opt.params, self.opt.views(self.flat_grad))}
        self.grads: Dict[str, torch.Tensor] = {n: by_id[id(p)] for n, p in model.named_parameters() if id(p) in by_id}
        self.pg = process_group
        self.sync_replicas()                                   # DDP semantics: every replica starts from rank 0's state
        self.last_grad_norm: Optional[torch.Tensor] = None
        self.cuda_graph = bool(cuda_graph)                     # forward + loss + backward replayed as one CUDA graph
        self._graph = None
        self._side = None                                      # prefetch(): next batch built on a side stream
        self._next = None

    def sync_replicas(self, src: int = 0) -> None:
        """Broadcast rank ``src``'s parameters / EMA / Adam moments / step count to all ranks of the process group (no-op for
        one process).  Called at construction; call it again after loading a checkpoint."""
        self.opt.sync_replicas(self.pg, src)

    def build_batch(self, x0: torch.Tensor, cond: dict, gen: torch.Generator, idx_override: Optional[torch.Tensor] = None,
                    noise: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, ...]:
        """train_keypoints.py:505-530 -> (z_t, t, idx, known_mask, eps_target)."""
        dev = L.require_cuda(x0)
        z0, idx, known_mask, known_values = _build_keypoint_batch(x0, self.K, cond, gen, self.logit_space, self.logit_eps,
                                                                  self.clamp_endpoints, idx_override=idx_override)
        t = torch.randint(0, self.N_train, (x0.shape[0],), device=dev, dtype=torch.long)      # global RNG, as :528
        z_t, eps = q_sample(z0, t, self.schedule, noise=noise)
        z_t = torch.where(known_mask, known_values, z_t)
        eps = eps * (~known_mask)
        return z_t, t, idx, known_mask, eps

    def loss_and_grads(self, z_t, t, idx, known_mask, cond, eps) -> torch.Tensor:
        """eps-MSE over the unknown entries (:532-537) = the weighted-MSE kernel with per-ELEMENT weights (known -> 0, else 1)."""
        B, K, D = z_t.shape
        eps_hat = self.bp.forward(z_t, t, idx, known_mask, cond, self.T)
        world = P.world_size(self.pg)
        loss, dgrad = stage2_loss(eps_hat.view(B, K * D, 1), eps.reshape(B, K * D, 1), known_mask.reshape(B, K * D), anchor_conf=False,
                                  w_anchor=0.0, w_missing=1.0, grad_accum=world)
        self.bp.backward(dgrad.view(B, K, D), self.grads)
        return loss * world if world > 1 else loss

    def prefetch(self, x0: torch.Tensor, cond: Dict[str, torch.Tensor], gen: torch.Generator) -> None:
        """Build the NEXT step's batch on a side stream while the current step runs (see ``Stage2Trainer.prefetch``: x0 / cond
        must already be complete on the device; the following ``step()`` must get the same x0 / cond)."""
        if self._side is None:
            self._side = torch.cuda.Stream()
        with torch.cuda.stream(self._side):
            batch = self.build_batch(x0, cond, gen)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._next = (batch, ev)

    def step(self, x0: torch.Tensor, cond: Dict[str, torch.Tensor], gen: torch.Generator) -> torch.Tensor:
        if self._next is not None:
            batch, ev = self._next
            self._next = None
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            for tns in batch:
                tns.record_stream(main)
            z_t, t, idx, known_mask, eps = batch
        else:
            z_t, t, idx, known_mask, eps = self.build_batch(x0, cond, gen)
        if self.cuda_graph:
            if self._graph is None:
                from .stage2_step import GraphedStep
                self._graph = GraphedStep(lambda z, tt, ii, km, ee, c: self.loss_and_grads(z, tt, ii, km, c, ee))
            loss = self._graph((z_t, t, idx, known_mask, eps), cond)
        else:
            loss = self.loss_and_grads(z_t, t, idx, known_mask, cond, eps)
        P.all_reduce_sum_(self.flat_grad, self.pg)
        self.last_grad_norm = self.opt.step(self.flat_grad)
        return loss
