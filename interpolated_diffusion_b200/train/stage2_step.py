"""One optimisation step of the reference's Stage-2 trainer (``src/train/train_interp_levels.py:1034-1173``) on libidb200:
on-device corruption (``build_interp_adjacent_batch`` / ``build_interp_level_batch``), confidence channels, denoiser forward,
weighted-MSE loss, hand-written backward (``train/backward.py``), data-parallel gradient all-reduce, global-norm clip,
fused AdamW + EMA (``train/optim.py``).  Argument names and defaults are the reference's CLI defaults (:40-140).

Data parallel: one process per GPU, every rank holds the full model and a B/N slice of the batch; the flat fp32
gradient arena is all-reduced once per step (``torch.distributed`` over NCCL/NVLink) and divided by the world size, i.e. the
mean of the per-rank losses' gradients -- DDP semantics.  (The reference normalises each rank's loss by its own sum of
weights, :1154; the mean over ranks of those ratios is what DDP would produce for it too.)"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from .. import _lib as L
from .. import parallel as P
from . import train_interp_levels as TI
from .backward import InterpLevelBackprop
from .optim import FlatAdamW, stage2_loss


class GraphedStep:
    """Replays ``fn(*tensors)`` (a forward + loss + backward that writes into fixed gradient buffers and returns the loss) as one
    CUDA graph per input signature: inputs are copied into static buffers, ~700 launches become one replay.

    ``split_owner`` (an object with a ``split_hook`` attribute that ``fn`` calls once, e.g. a ``_DenoiserBackprop``): the capture
    is cut at that call into TWO graphs sharing one memory pool, and ``__call__(..., between=f)`` runs ``f()`` between the two
    replays -- the data-parallel trainer launches the all-reduce of the gradients that are final at the cut there."""

    def __init__(self, fn, split_owner=None):
        self.fn = fn
        self.state = None
        self.split_owner = split_owner

    def __call__(self, tensors, cond, between=None):
        keys = sorted(cond)
        args = list(tensors) + [cond[k] for k in keys]
        sig = tuple((tuple(a.shape), a.dtype) for a in args)
        if self.state is None or self.state["sig"] != sig:
            static = [a.clone() for a in args]
            n = len(tensors)
            scond = {k: static[n + i] for i, k in enumerate(keys)}
            stream = torch.cuda.Stream()
            stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(stream):                         # warm-up on a side stream (kernel attributes, workspaces)
                for _ in range(2):
                    self.fn(*static[:n], scond)
            torch.cuda.current_stream().wait_stream(stream)
            from ..models import _engine as E
            E.note_graph_captured()                                 # workspaces retire (never free) buffers from here on
            g = torch.cuda.CUDAGraph()
            g2 = None
            if self.split_owner is None:
                with torch.cuda.graph(g):
                    loss = self.fn(*static[:n], scond)
            else:
                g2 = torch.cuda.CUDAGraph()

                def cut():                                          # called by fn at the split point, inside the capture
                    g.capture_end()
                    g2.capture_begin(pool=g.pool())

                torch.cuda.synchronize()
                cap = torch.cuda.Stream()
                cap.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(cap):
                    self.split_owner.split_hook = cut
                    try:
                        g.capture_begin()
                        loss = self.fn(*static[:n], scond)
                        g2.capture_end()
                    finally:
                        self.split_owner.split_hook = None
                torch.cuda.current_stream().wait_stream(cap)
            self.state = {"sig": sig, "g": g, "g2": g2, "static": static, "loss": loss}
        for dst, src in zip(self.state["static"], args):
            dst.copy_(src)
        self.state["g"].replay()
        if self.state["g2"] is not None:
            if between is not None:
                between()
            self.state["g2"].replay()
        return self.state["loss"]


class Stage2Trainer:
    def __init__(self, model, *, K_min: int = 8, levels: int = 3, stage2_mode: str = "adj", kp_index_mode: str = "random_nested",
                 k_schedule: str = "doubling", k_geom_gamma: Optional[float] = None, anchor_conf: bool = True,
                 anchor_conf_teacher: float = 0.95, anchor_conf_student: float = 0.5, anchor_conf_endpoints: float = 1.0,
                 anchor_conf_missing: float = 0.0, anchor_conf_anneal: bool = True, anchor_conf_anneal_mode: str = "linear",
                 corrupt_mode: str = "dist", corrupt_sigma_max: float = 0.08, corrupt_sigma_min: float = 0.012,
                 corrupt_sigma_pow: float = 0.75, corrupt_anchor_frac: float = 0.25, corrupt_index_jitter_max: int = 0,
                 corrupt_index_jitter_prob: float = 0.0, corrupt_index_jitter_pow: float = 1.0, clamp_endpoints: bool = True,
                 recompute_vel: bool = True, pos_clip: bool = False, pos_clip_min: float = 0.0, pos_clip_max: float = 1.0,
                 level_sampling: str = "high", level_high_prob: float = 0.5, w_anchor: float = 0.1, w_missing: float = 1.0,
                 lr: float = 2e-4, weight_decay: float = 1e-2, grad_clip: Optional[float] = 1.0, ema: bool = True,
                 ema_decay: float = 0.999, process_group=None, cuda_graph: bool = False, bootstrap_model=None,
                 bootstrap_schedule: Optional[Dict[str, torch.Tensor]] = None, bootstrap_logit: bool = False,
                 bootstrap_logit_eps: float = 1e-5, bootstrap_ddim_steps: int = 5, bootstrap_ddim_schedule: str = "quadratic",
                 bootstrap_prob_start: float = 0.0, bootstrap_prob_end: float = 0.3, bootstrap_warmup_steps: int = 5000,
                 bootstrap_prob_cap: float = 0.5, bootstrap_mode: str = "batch", bootstrap_replace_prob: float = 0.5,
                 clamp_endpoints_kp: Optional[bool] = None, selector_model=None, selector_level_mode: str = "k_norm",
                 batch_mode: str = "reference_draws", overlap_allreduce: bool = True):
        if stage2_mode not in ("adj", "x0"):
            raise ValueError("stage2_mode must be 'adj' or 'x0'")
        # mask policies of train_interp_levels.py:890-967.  The CLI default "random" is only reachable through --mask_policy_mix,
        # where it means "random_nested" (:893-894; on its own it falls through to "Unknown kp_index_mode", :958); "selector"
        # ranks the interior positions by the logits of a frozen KeypointSelector (:911-947; per level when it is level-conditioned).
        if kp_index_mode == "random":
            kp_index_mode = "random_nested"
        if kp_index_mode not in ("random_nested", "uniform", "dp_precomputed", "selector"):
            raise ValueError(f"Unknown kp_index_mode: {kp_index_mode}")
        if kp_index_mode == "selector" and selector_model is None:
            raise ValueError("kp_index_mode=selector but selector model not loaded")
        self.selector_model = selector_model
        self.selector_level_mode = selector_level_mode
        self.kp_index_mode, self.k_schedule, self.k_geom_gamma = kp_index_mode, k_schedule, k_geom_gamma
        self.model = model
        self.cfg = dict(K_min=K_min, levels=levels, stage2_mode=stage2_mode, anchor_conf=bool(anchor_conf),
                        conf=(anchor_conf_teacher, anchor_conf_student, anchor_conf_endpoints, anchor_conf_missing),
                        anneal=bool(anchor_conf_anneal), anneal_mode=anchor_conf_anneal_mode, clamp_endpoints=bool(clamp_endpoints),
                        recompute_vel=bool(recompute_vel), level_sampling=level_sampling, level_high_prob=level_high_prob,
                        w_anchor=w_anchor, w_missing=w_missing)
        self.corrupt = dict(corrupt_mode=corrupt_mode, corrupt_sigma_max=corrupt_sigma_max, corrupt_sigma_min=corrupt_sigma_min,
                            corrupt_sigma_pow=corrupt_sigma_pow, corrupt_anchor_frac=corrupt_anchor_frac,
                            corrupt_index_jitter_max=corrupt_index_jitter_max, corrupt_index_jitter_prob=corrupt_index_jitter_prob,
                            corrupt_index_jitter_pow=corrupt_index_jitter_pow, clamp_endpoints=bool(clamp_endpoints),
                            pos_clip=bool(pos_clip), pos_clip_min=pos_clip_min, pos_clip_max=pos_clip_max)
        self.bp = InterpLevelBackprop(model)
        # arena layout: the parameters whose gradients are final at the backward's split point (out head, every transformer
        # parameter except the FiLM linears: 89 % of the large model) come first, so that slice can be all-reduced while the
        # backward's tail (FiLM, token assembly, conditioning encoder: ~2.5 ms at the cfg-4 shapes) still runs
        early_names = set(self.bp.early_param_names())
        self.opt = FlatAdamW(model.parameters(), lr=lr, weight_decay=weight_decay, ema_decay=ema_decay if ema else None,
                             max_grad_norm=grad_clip, early=[p for n_, p in model.named_parameters() if n_ in early_names])
        self.overlap_allreduce = bool(overlap_allreduce)
        self._ar_stream = None
        self.flat_grad = torch.zeros_like(self.opt.flat)
        by_id = {id(p): g for p, g in zip(self.opt.params, self.opt.views(self.flat_grad))}
        self.grads: Dict[str, torch.Tensor] = {n: by_id[id(p)] for n, p in model.named_parameters() if id(p) in by_id}
        # bootstrap (train_interp_levels.py:970-1032): a frozen Stage-1 model samples the level-S anchors; with probability
        # bootstrap_replace_prob an interior anchor's position is replaced by the student's prediction (marked in student_mask, which
        # lowers its confidence channel)
        self.boot = None
        if bootstrap_model is not None:
            if bootstrap_schedule is None:
                raise ValueError("bootstrap_model needs bootstrap_schedule (make_alpha_bars of the Stage-1 checkpoint's schedule)")
            if bootstrap_mode not in ("batch", "per_example"):
                raise ValueError("bootstrap_mode must be 'batch' or 'per_example'")
            self.boot = dict(model=bootstrap_model, schedule=bootstrap_schedule, logit=bool(bootstrap_logit), eps=bootstrap_logit_eps,
                             steps=bootstrap_ddim_steps, sched_name=bootstrap_ddim_schedule, p0=bootstrap_prob_start,
                             p1=bootstrap_prob_end, warm=bootstrap_warmup_steps, cap=bootstrap_prob_cap, mode=bootstrap_mode,
                             replace=bootstrap_replace_prob,
                             clamp_kp=bool(clamp_endpoints if clamp_endpoints_kp is None else clamp_endpoints_kp))
        # batch_mode: "reference_draws" (default) builds the batch with the reference's generator calls in the reference's order
        # (per-level noise shapes -> one host sync per batch, 2 launches per level): bit-identical corruption under a shared
        # generator state.  "fused" is the speed mode: level indices without the data-dependent draw count, ONE corruption launch
        # for the whole batch with Philox noise drawn in the kernel (same distributions, a different random stream), no host sync.
        if batch_mode not in ("reference_draws", "fused"):
            raise ValueError("batch_mode must be 'reference_draws' or 'fused'")
        self.batch_mode = batch_mode
        self._noise_calls = 0
        self.step_index = 0
        self.pg = process_group
        self.sync_replicas()                                   # DDP semantics: every replica starts from rank 0's state
        self.last_grad_norm: Optional[torch.Tensor] = None
        # cuda_graph: forward + loss + backward (~700 launches for the 12-layer model) are captured once per batch shape and
        # replayed; batch building (host-side level counts), the all-reduce and the optimiser (host-side step count) stay eager
        self.cuda_graph = bool(cuda_graph)
        self._graph = None
        # prefetch(): the next batch (corruption, masks, confidence channels: launch- and host-sync-bound, ~1.3 ms) is built on a
        # side stream while the current step's graph runs on the main stream
        self._side = None
        self._next = None

    # ------------------------------------------------------------------------------------------------------------------
    def sync_replicas(self, src: int = 0) -> None:
        """Broadcast rank ``src``'s parameters / EMA / Adam moments / step count to all ranks of the process group (no-op for
        one process).  Called at construction; call it again after loading a checkpoint."""
        self.opt.sync_replicas(self.pg, src)

    def build_masks(self, x0: torch.Tensor, gen: torch.Generator, cond: Optional[dict] = None):
        """Nested anchor masks of the batch by ``kp_index_mode`` (train_interp_levels.py:890-967)."""
        from ..corruptions import keyframes as kf
        B, T, _ = x0.shape
        dev = x0.device
        c = self.cfg
        if self.kp_index_mode == "random_nested":
            return kf.build_nested_masks_batch(B, T, c["K_min"], c["levels"], generator=gen, device=dev, k_schedule=self.k_schedule,
                                               k_geom_gamma=self.k_geom_gamma)
        if self.kp_index_mode == "selector":
            if cond is None:
                raise ValueError("kp_index_mode=selector needs the conditioning")
            if getattr(self.selector_model, "use_level", False):
                # level-conditioned selector (:916-937): one ranking per level, conditioned on s / S or K_s / (T - 1)
                k_list = kf._compute_k_schedule(T, c["K_min"], c["levels"], schedule=self.k_schedule, geom_gamma=self.k_geom_gamma)
                per_level = []
                for s_lvl in range(c["levels"] + 1):
                    if self.selector_level_mode == "s_norm":
                        level_val = float(s_lvl) / float(max(1, c["levels"]))
                    else:
                        level_val = float(k_list[s_lvl]) / float(max(1, T - 1))
                    cond_sel = dict(cond)
                    cond_sel["level"] = torch.full((B, 1), level_val, device=dev)
                    per_level.append(self.selector_model(cond_sel))
                return kf.build_nested_masks_from_level_logits(torch.stack(per_level, dim=1), c["K_min"], c["levels"],
                                                               k_schedule=self.k_schedule, k_geom_gamma=self.k_geom_gamma)
            logits = self.selector_model(cond)
            return kf.build_nested_masks_from_logits(logits, c["K_min"], c["levels"], k_schedule=self.k_schedule,
                                                     k_geom_gamma=self.k_geom_gamma)
        if self.kp_index_mode == "dp_precomputed":
            if cond is None or "kp_idx" not in cond:
                raise ValueError("kp_index_mode=dp_precomputed requires kp_idx in dataset")
            idx_base = cond["kp_idx"].to(dev)
        else:
            idx_base, _ = kf.sample_fixed_k_indices_uniform_batch(B, T, c["K_min"], generator=gen, device=dev, ensure_endpoints=True)
        return kf.build_nested_masks_from_base(idx_base, T, c["levels"], generator=gen, device=dev, k_schedule=self.k_schedule,
                                               k_geom_gamma=self.k_geom_gamma)

    def _bootstrap(self, x0: torch.Tensor, cond: dict, idx_levels, gen: torch.Generator, z_T: Optional[torch.Tensor] = None):
        """train_interp_levels.py:970-1032 -> (x0_used, student_mask [B,T] bool)."""
        from ..sample.sample_generate import _build_known_mask_values, _kp_feat_from_idx, _sample_keypoints_ddim
        from ..utils.normalize import logit_pos, sigmoid_pos
        b = self.boot
        B, T, D = x0.shape
        dev = x0.device
        if b["warm"] <= 0:
            p_boot = b["p1"]
        else:
            frac = min(1.0, float(self.step_index) / float(b["warm"]))
            p_boot = b["p0"] + frac * (b["p1"] - b["p0"])
        p_boot = min(p_boot, b["cap"])
        if b["mode"] == "batch":
            use_boot = torch.rand((), generator=gen, device=dev) < p_boot
            use_mask = torch.full((B,), bool(use_boot), device=dev, dtype=torch.bool)
        else:
            use_mask = torch.rand((B,), generator=gen, device=dev) < p_boot
        student_mask = torch.zeros((B, T), device=dev, dtype=torch.bool)
        if not bool(torch.any(use_mask)):
            return x0, student_mask
        idx_s = idx_levels[self.cfg["levels"]]
        known_mask, known_values = _build_known_mask_values(idx_s, cond, D, T, b["clamp_kp"])
        if b["logit"]:
            known_values = logit_pos(known_values, eps=b["eps"])
        cond_boot = cond
        kp_feat_dim = int(getattr(b["model"], "kp_feat_dim", 0))
        if kp_feat_dim > 0:
            cond_boot = dict(cond)
            cond_boot["kp_feat"] = _kp_feat_from_idx(idx_s, T, kp_feat_dim, None, None)
        z_hat = _sample_keypoints_ddim(b["model"], b["schedule"], idx_s, known_mask, known_values, cond_boot, b["steps"], T,
                                       schedule_name=b["sched_name"], z_T=z_T)
        if b["logit"]:
            z_hat = sigmoid_pos(z_hat)
        K = idx_s.shape[1]
        x0_aug = x0.clone()
        replace_mask = torch.rand((B, K), generator=gen, device=dev) < float(b["replace"])
        replace_mask = replace_mask & use_mask.view(-1, 1)
        replace_mask = replace_mask & ~(idx_s == 0) & ~(idx_s == (T - 1))
        student_mask.scatter_(1, idx_s, replace_mask)
        gidx = idx_s.unsqueeze(-1).expand(-1, K, 2)
        vals = x0_aug[:, :, :2].gather(1, gidx)
        vals = torch.where(replace_mask.unsqueeze(-1), z_hat[:, :, :2], vals)
        pos = x0_aug[:, :, :2].scatter(1, gidx, vals)
        x0_aug[:, :, :2] = pos
        x0_used = torch.where(use_mask.view(-1, 1, 1), x0_aug, x0) if bool(torch.any(replace_mask)) else x0
        return x0_used, student_mask

    def build_batch(self, x0: torch.Tensor, gen: torch.Generator, cond: Optional[dict] = None, masks_levels=None, idx_levels=None
                    ) -> Tuple[torch.Tensor, ...]:
        """train_interp_levels.py:890-1135: (x_s, s_idx, mask_in, target, weight_mask).  Draw order
        on ``gen`` as in the reference: masks, then the level indices, then the corruption noise."""
        c = self.cfg
        dev = L.require_cuda(x0)
        B = x0.shape[0]
        if masks_levels is None or idx_levels is None:
            masks_levels, idx_levels = self.build_masks(x0, gen, cond)
        x0_used, student = None, None
        if self.boot is not None:
            if cond is None:
                raise ValueError("the bootstrap branch needs the conditioning (Stage-1 sampling)")
            x0_used, student = self._bootstrap(x0, cond, idx_levels, gen)
        conf_t, conf_st, conf_e, conf_m = c["conf"]
        if self.batch_mode == "fused" and not (self.corrupt["corrupt_index_jitter_max"] > 0 and self.corrupt["corrupt_index_jitter_prob"] > 0.0):
            return self._build_batch_fused(x0, x0_used, student, masks_levels, gen)
        s_idx = TI._sample_level_indices(B, c["levels"], gen, dev, c["level_sampling"], c["level_high_prob"])
        if c["stage2_mode"] == "adj":
            x_s, x_prev, mask_s, mask_prev, s_idx, _, _ = TI.build_interp_adjacent_batch(
                x0, c["K_min"], c["levels"], gen, recompute_velocity=c["recompute_vel"], x0_override=x0_used, masks_levels=masks_levels,
                idx_levels=idx_levels, s_idx=s_idx, **self.corrupt)
            conf_s = TI._build_anchor_conf(mask_s, student, conf_t, conf_st, conf_e, conf_m, c["clamp_endpoints"])
            conf_prev = TI._build_anchor_conf(mask_prev, student, conf_t, conf_st, conf_e, conf_m, c["clamp_endpoints"])
            if c["anneal"]:
                conf_s = TI._anneal_conf(conf_s, s_idx, c["levels"], c["anneal_mode"])
                conf_prev = TI._anneal_conf(conf_prev, torch.clamp(s_idx - 1, min=0), c["levels"], c["anneal_mode"])
            if c["anchor_conf"]:
                mask_in = torch.stack([mask_s.float(), mask_prev.float(), conf_s], dim=-1)
            else:
                mask_in = torch.stack([mask_s, mask_prev], dim=-1)
            return x_s, s_idx, mask_in, x_prev - x_s, (conf_prev if c["anchor_conf"] else mask_prev)
        x_s, mask_s, s_idx, _, _ = TI.build_interp_level_batch(
            x0, c["K_min"], c["levels"], gen, recompute_velocity=c["recompute_vel"], x0_override=x0_used, masks_levels=masks_levels,
            idx_levels=idx_levels, s_idx=s_idx, **self.corrupt)
        conf_s = TI._build_anchor_conf(mask_s, student, conf_t, conf_st, conf_e, conf_m, c["clamp_endpoints"])
        if c["anneal"]:
            conf_s = TI._anneal_conf(conf_s, s_idx, c["levels"], c["anneal_mode"])
        mask_in = torch.stack([mask_s.float(), conf_s], dim=-1) if c["anchor_conf"] else mask_s
        return x_s, s_idx, mask_in, x0 - x_s, (conf_s if c["anchor_conf"] else mask_s)

    def _build_batch_fused(self, x0, x0_used, student, masks_levels, gen):
        """Speed-mode batch (see ``batch_mode``): sync-free level draw, one ``idb200_corrupt_adjacent`` launch, two
        ``idb200_anchor_conf`` launches (confidence + anneal + stacked mask_in), one subtraction."""
        from ..corruptions import keyframes as kf
        from ..sample.sample_generate import anchor_conf_mask_in
        c = self.cfg
        B, T, _ = x0.shape
        dev = x0.device
        S = c["levels"]
        s_idx = TI._sample_level_indices(B, S, gen, dev, c["level_sampling"], c["level_high_prob"], sync_free=True)
        K_list = kf._compute_k_schedule(T, c["K_min"], S, schedule=self.k_schedule, geom_gamma=self.k_geom_gamma)
        cr = self.corrupt
        adj = c["stage2_mode"] == "adj"
        self._noise_calls += 1
        x_s, x_prev, mask_s, mask_prev = TI.corrupt_adjacent_fused(
            x0 if x0_used is None else x0_used, masks_levels, s_idx, K_list, c["K_min"], adjacent=adj, recompute_velocity=c["recompute_vel"],
            corrupt_mode=cr["corrupt_mode"], corrupt_sigma_max=cr["corrupt_sigma_max"], corrupt_sigma_min=cr["corrupt_sigma_min"],
            corrupt_sigma_pow=cr["corrupt_sigma_pow"], corrupt_anchor_frac=cr["corrupt_anchor_frac"], clamp_endpoints=cr["clamp_endpoints"],
            pos_clip=cr["pos_clip"], pos_clip_min=cr["pos_clip_min"], pos_clip_max=cr["pos_clip_max"], seed=gen.initial_seed(),
            offset=self._noise_calls)
        conf_t, conf_st, conf_e, conf_m = c["conf"]
        mode = c["anneal_mode"] if c["anneal"] else "none"
        ac = dict(conf_teacher=conf_t, conf_student=conf_st, conf_endpoints=conf_e, conf_missing=conf_m, clamp_endpoints=c["clamp_endpoints"])
        if adj:
            if c["anchor_conf"]:
                _, mask_in = anchor_conf_mask_in(mask_s, student, mask_prev, s_idx, S, mode, want_conf=False, channels=3, **ac)
                conf_prev, _ = anchor_conf_mask_in(mask_prev, student, None, torch.clamp(s_idx - 1, min=0), S, mode, **ac)
                return x_s, s_idx, mask_in, x_prev - x_s, conf_prev
            return x_s, s_idx, torch.stack([mask_s, mask_prev], dim=-1), x_prev - x_s, mask_prev
        if c["anchor_conf"]:
            conf_s, mask_in = anchor_conf_mask_in(mask_s, student, None, s_idx, S, mode, channels=2, **ac)
            return x_s, s_idx, mask_in, x0 - x_s, conf_s
        return x_s, s_idx, mask_s, x0 - x_s, mask_s

    def loss_and_grads(self, x_s, s_idx, mask_in, cond, target, weight_mask) -> torch.Tensor:
        """Forward + loss + backward; the parameter gradients land in ``self.flat_grad`` (views: ``self.grads``)."""
        c = self.cfg
        delta_hat = self.bp.forward(x_s, s_idx, mask_in, cond)
        # the 1 / world factor of the data-parallel mean rides on the loss kernel's grad_accum divisor
        world = P.world_size(self.pg)
        loss, dgrad = stage2_loss(delta_hat, target, weight_mask, anchor_conf=c["anchor_conf"], w_anchor=c["w_anchor"],
                                  w_missing=c["w_missing"], grad_accum=world)
        self.bp.backward(dgrad, self.grads)
        return loss * world if world > 1 else loss

    def _graphed_loss_and_grads(self, x_s, s_idx, mask_in, cond, target, weight_mask, between=None) -> torch.Tensor:
        if self._graph is None:
            self._graph = GraphedStep(lambda xs, si, mi, tg, wm, c: self.loss_and_grads(xs, si, mi, c, tg, wm),
                                      split_owner=self.bp if self.overlap_allreduce else None)
        return self._graph((x_s, s_idx, mask_in, target, weight_mask), cond, between=between)

    def reduce_gradients(self) -> None:
        """Data-parallel all-reduce of the flat gradient arena (each rank's gradient already carries 1 / world)."""
        P.all_reduce_sum_(self.flat_grad, self.pg)

    def _reduce_early_async(self) -> None:
        """Between the two graphs of a split step: all-reduce ``flat_grad[:n_early]`` on a side stream (NCCL runs under the
        backward's tail); ``step()`` reduces the rest on the main stream and joins."""
        if P.world_size(self.pg) == 1:
            return
        if self._ar_stream is None:
            self._ar_stream = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        with torch.cuda.stream(self._ar_stream):
            self._ar_stream.wait_event(ev)
            P.all_reduce_sum_(self.flat_grad[: self.opt.n_early], self.pg)

    def prefetch(self, x0: torch.Tensor, cond: Dict[str, torch.Tensor], gen: torch.Generator) -> None:
        """Build the batch of the NEXT step now, on a side stream (call right after ``step()``: the step's kernels are still
        running).  The following ``step()`` consumes it instead of building its own; its x0 / cond arguments must be the ones
        given here.  x0 / cond must already be complete on the device (e.g. loaded before the previous step was enqueued): the
        side stream does NOT wait for the main stream -- that wait would put the build behind the running step again."""
        if self._side is None:
            self._side = torch.cuda.Stream()
        with torch.cuda.stream(self._side):
            batch = self.build_batch(x0, gen, cond)
            ev = torch.cuda.Event()
            ev.record(self._side)
        self._next = (batch, ev)

    def step(self, x0: torch.Tensor, cond: Dict[str, torch.Tensor], gen: torch.Generator) -> torch.Tensor:
        """One full training step; returns the (local) loss as a 0-dim device tensor (no host sync).  With ``cuda_graph`` the
        returned tensor is the graph's static output: read it (``float(loss)``) or clone it before the next ``step()``."""
        if self._next is not None:
            batch, ev = self._next
            self._next = None
            main = torch.cuda.current_stream()
            main.wait_event(ev)
            for t in batch:
                t.record_stream(main)                        # allocated on the side stream, consumed here
            x_s, s_idx, mask_in, target, weight_mask = batch
        else:
            x_s, s_idx, mask_in, target, weight_mask = self.build_batch(x0, gen, cond)
        if self.cuda_graph and self.overlap_allreduce:
            loss = self._graphed_loss_and_grads(x_s, s_idx, mask_in, cond, target, weight_mask, between=self._reduce_early_async)
            if P.world_size(self.pg) > 1:
                P.all_reduce_sum_(self.flat_grad[self.opt.n_early:], self.pg)
                if self._ar_stream is not None:
                    torch.cuda.current_stream().wait_stream(self._ar_stream)
        else:
            fn = self._graphed_loss_and_grads if self.cuda_graph else self.loss_and_grads
            loss = fn(x_s, s_idx, mask_in, cond, target, weight_mask)
            self.reduce_gradients()
        self.last_grad_norm = self.opt.step(self.flat_grad)
        self.step_index += 1
        return loss
