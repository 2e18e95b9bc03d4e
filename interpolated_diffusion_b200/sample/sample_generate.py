"""Mirror of the hot-path functions of the reference's ``src/sample/sample_generate.py`` (lines 260-404
and the per-batch body 944-1285) on libidb200 kernels.  CLI, checkpoint reconciliation, plotting and
per-sample metric loops of the reference file are out of scope."""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

from .. import _lib as L
from ..diffusion.ddpm import _timesteps, ddim_step_scalar

ANNEAL = {"none": 0, "linear": 1, "cosine": 2}


def _build_known_mask_values(idx: torch.Tensor, cond: dict, D: int, T: int, clamp_endpoints: bool = True, *,
                             logit_space: bool = False, logit_eps: float = 1e-5) -> Tuple[torch.Tensor, torch.Tensor]:
    """sample_generate.py:260-280 -> (known_mask bool [B,K,D], known_values fp32 [B,K,D]).
    ``logit_space=True`` also applies ``logit_pos`` (sample_generate.py:1021-1022) in the same launch."""
    dev = L.require_cuda(idx)
    B, K = idx.shape
    sg = None
    if clamp_endpoints:
        if "start_goal" not in cond:
            raise ValueError("clamp_endpoints=True but start_goal missing from cond")
        sg = L.f32c(cond["start_goal"])
        L.require_cuda(sg)
    known_mask = torch.empty((B, K, D), device=dev, dtype=torch.bool)
    known_values = torch.empty((B, K, D), device=dev, dtype=torch.float32)
    L.call("idb200_known_mask_values", L.ptr(L.i64c(idx)), L.ptr(sg), B, K, D, T, int(bool(clamp_endpoints)),
           int(bool(logit_space)), float(logit_eps), L.ptr(known_mask), L.ptr(known_values), L.stream(dev))
    return known_mask, known_values


def _kp_feat_from_idx(idx: torch.Tensor, T: int, kp_feat_dim: int, left_diff: Optional[torch.Tensor] = None,
                      right_diff: Optional[torch.Tensor] = None) -> torch.Tensor:
    """sample_generate.py:283-307: [left_gap, right_gap, t_norm(, left_diff, right_diff)] per keypoint."""
    B, K = idx.shape
    feat = torch.zeros((B, K, kp_feat_dim), device=idx.device, dtype=torch.float32)
    if kp_feat_dim <= 0:
        return feat
    denom = float(max(1, T - 1))
    if K > 1:
        gaps = (idx[:, 1:] - idx[:, :-1]).float() / denom
        feat[:, 1:, 0] = gaps
        feat[:, :-1, 1] = gaps
    if kp_feat_dim >= 3:
        feat[:, :, 2] = idx.float() / denom
    if kp_feat_dim >= 5 and left_diff is not None and right_diff is not None:
        feat[:, :, 3] = left_diff
        feat[:, :, 4] = right_diff
    return feat


def anchor_conf_mask_in(mask_s: torch.Tensor, student_mask: Optional[torch.Tensor], mask_prev: Optional[torch.Tensor],
                        s, levels: int, anneal_mode: str, conf_teacher: float, conf_student: float,
                        conf_endpoints: float, conf_missing: float, clamp_endpoints: bool, *, want_conf: bool = True,
                        channels: int = 0):
    """``idb200_anchor_conf``: conf [B,T] (+ optional anneal) and, if ``channels`` in {2,3}, the stacked
    ``mask_in`` of sample_generate.py:1163-1178 / :1255-1256 in one launch.  ``s`` is an int or int64 [B]."""
    dev = L.require_cuda(mask_s, student_mask, mask_prev)
    B, T = mask_s.shape
    ms = L.u8c(mask_s)
    st = L.u8c(student_mask) if student_mask is not None else None
    mp = L.u8c(mask_prev) if mask_prev is not None else None
    conf = torch.empty((B, T), device=dev, dtype=torch.float32) if want_conf else None
    mask_in = torch.empty((B, T, channels), device=dev, dtype=torch.float32) if channels else None
    s_row, s_scalar = (L.i64c(s), 0) if isinstance(s, torch.Tensor) else (None, int(s))
    L.call("idb200_anchor_conf", L.ptr(ms), L.ptr(st), L.ptr(mp), L.ptr(s_row), s_scalar, int(levels),
           ANNEAL.get(anneal_mode, 0) if levels > 0 else 0, float(conf_teacher), float(conf_student),
           float(conf_endpoints), float(conf_missing), int(bool(clamp_endpoints)), B, T, max(channels, 2), L.ptr(conf),
           L.ptr(mask_in), L.stream(dev))
    return conf, mask_in


def _build_anchor_conf(mask_s: torch.Tensor, student_mask: Optional[torch.Tensor], use_student: bool,
                       conf_teacher: float, conf_student: float, conf_endpoints: float, conf_missing: float,
                       clamp_endpoints: bool) -> torch.Tensor:
    """sample_generate.py:319-336"""
    st = student_mask if (student_mask is not None and use_student) else None
    conf, _ = anchor_conf_mask_in(mask_s, st, None, 0, 0, "none", conf_teacher, conf_student, conf_endpoints,
                                  conf_missing, clamp_endpoints)
    return conf


def _soft_clamp_lambda(s: int, levels: int, schedule: str, max_val: float) -> float:
    """sample_generate.py:339-347"""
    if levels <= 0:
        return float(max_val)
    frac = float(s) / float(levels)
    if schedule == "linear":
        return float(max_val) * frac
    if schedule == "cosine":
        return float(max_val) * 0.5 * (1.0 + math.cos(math.pi * (1.0 - frac)))
    return float(max_val)


def _anneal_conf(conf: torch.Tensor, s: int, levels: int, mode: str) -> torch.Tensor:
    """sample_generate.py:350-360 (a scalar-lambda axpy on a [B,T] tensor; the fused form is
    ``anchor_conf_mask_in``)."""
    if conf is None or mode == "none" or levels <= 0:
        return conf
    frac = float(s) / float(levels)
    if mode == "linear":
        lam = 1.0 - frac
    elif mode == "cosine":
        lam = 0.5 * (1.0 + math.cos(math.pi * frac))
    else:
        lam = 0.0
    return conf + (1.0 - conf) * float(lam)


def _compute_sigma_for_level(K_s: int, K_min: int, sigma_max: float, sigma_min: float, sigma_pow: float) -> float:
    """train_interp_levels.py:386-401 (imported by the sampler for s2 noise)."""
    if sigma_max <= 0.0:
        return 0.0
    K_s = max(1, int(K_s))
    K_min = max(1, int(K_min))
    ratio = float(K_min) / float(K_s)
    sigma = float(sigma_max) * (ratio ** float(sigma_pow))
    sigma = min(float(sigma_max), sigma)
    return max(float(sigma_min), sigma)


def _sample_keypoints_ddim(model, schedule, idx: torch.Tensor, known_mask: torch.Tensor, known_values: torch.Tensor,
                           cond: dict, steps: int, T: int, schedule_name: str = "linear",
                           return_intermediates: bool = False, pos_clip: bool = False, pos_clip_min: float = 0.0,
                           pos_clip_max: float = 1.0, *, z_T: Optional[torch.Tensor] = None, cond_vec: Optional[torch.Tensor] = None):
    """sample_generate.py:363-404.  The DDIM update and the known-value ``torch.where`` (:397-399) are one
    launch per step with the step's two alpha-bar entries as kernel arguments (no per-row table gather,
    no host sync).  ``z_T=`` injects the initial noise (the reference draws it from the global RNG, :389).
    Everything the reference recomputes inside every denoiser call although it does not change across steps -- the
    conditioning encoder, the FiLM tables, the cond_proj row, the timestep MLP -- is computed once before the loop
    (``cond_vec=`` lets a caller hoist the encoder further, e.g. across the chunks of the causal sampler)."""
    dev = L.require_cuda(idx, known_mask, known_values)
    B, K = idx.shape
    D = known_values.shape[-1]
    alpha_bar = schedule["alpha_bar"].detach().to("cpu", torch.float32)
    n_train = alpha_bar.shape[0]
    times = _timesteps(n_train, steps, schedule=schedule_name).tolist()
    z = torch.randn((B, K, D), device=dev) if z_T is None else L.f32c(z_T).clone()
    z = torch.where(known_mask, known_values, z)
    if pos_clip:
        z[..., :2] = z[..., :2].clamp(min=pos_clip_min, max=pos_clip_max)
    intermediates = [z.detach().clone()] if return_intermediates else None
    hoist = {}
    if hasattr(model, "encode_cond") and hasattr(model, "timestep_vector") and len(times) > 1:
        if cond_vec is None and cond and model.cond_enc is not None:
            cond_vec = model.encode_cond(cond)
        if cond_vec is not None:
            hoist = {"cond_vec": cond_vec, "film": model.transformer.packed().film_params(cond_vec, K, model.precision),
                     "row_b": _cond_row(model, cond_vec, T, dev)}
        t_vecs = model.timestep_vector(torch.tensor(times[:-1], device=dev, dtype=torch.long))
    for i in range(len(times) - 1):
        t = torch.full((B,), int(times[i]), device=dev, dtype=torch.long)
        if hoist:
            eps = model(z, t, idx, known_mask, cond, T, t_vec=t_vecs[i:i + 1], **hoist)
        else:
            eps = model(z, t, idx, known_mask, cond, T)
        z = ddim_step_scalar(z, eps, float(alpha_bar[times[i]]), float(alpha_bar[times[i + 1]]), known_mask=known_mask,
                             known_values=known_values, pos_clip=pos_clip, pos_clip_min=pos_clip_min,
                             pos_clip_max=pos_clip_max)
        if return_intermediates:
            intermediates.append(z.detach().clone())
    if return_intermediates:
        return z, intermediates
    return z


# ------------------------------------------------------------------------------------------------------------
# The batched generation hot loop (sample_generate.py:974-1285; exact recipe in SURVEY.md 3.5).
# ------------------------------------------------------------------------------------------------------------
class GenerationConfig:
    """The hot-path flags of the reference CLI (sample_generate.py:38-155) with the reference defaults."""

    def __init__(self, T: int = 64, K_min: int = 8, levels: int = 3, data_dim: int = 2, ddim_steps: int = 20,
                 ddim_schedule: str = "quadratic", n_train: int = 1000, beta_schedule: str = "cosine", stage2_mode: str = "x0",
                 clamp_policy: str = "endpoints", clamp_dims: str = "pos", logit_space: bool = True, logit_eps: float = 1e-5,
                 anchor_conf: bool = True, anchor_conf_teacher: float = 0.95, anchor_conf_student: float = 0.5,
                 anchor_conf_endpoints: float = 1.0, anchor_conf_missing: float = 0.0, anchor_conf_anneal_mode: str = "linear",
                 soft_anchor_clamp: bool = True, soft_clamp_schedule: str = "linear", soft_clamp_max: float = 1.0,
                 recompute_vel: bool = True, clamp_endpoints: bool = True, pos_clip: bool = False, pos_clip_min: float = 0.0,
                 pos_clip_max: float = 1.0, kp_index_mode: str = "uniform", k_schedule: str = "doubling"):
        if clamp_policy not in ("none", "endpoints", "all_anchors"):
            raise ValueError(f"Unknown clamp_policy: {clamp_policy}")
        if clamp_dims not in ("pos", "all"):
            raise ValueError(f"Unknown clamp_dims: {clamp_dims}")
        if stage2_mode not in ("x0", "adj"):
            raise ValueError(f"Unknown stage2_mode: {stage2_mode}")
        self.__dict__.update({k: v for k, v in locals().items() if k != "self"})


def generate(kp_model, interp_model, cond: Dict[str, torch.Tensor], cfg: Optional[GenerationConfig] = None, *,
             z_T: Optional[torch.Tensor] = None, idx: Optional[torch.Tensor] = None, masks_levels: Optional[torch.Tensor] = None,
             generator: Optional[torch.Generator] = None, return_all: bool = False, out: Optional[torch.Tensor] = None):
    """Stage-1 DDIM over K keypoints -> sigmoid -> Interp to T -> Stage-2 (one-step ``x0`` jump or ``adj`` chain)
    -> soft / hard clamp by ``clamp_policy``.  Everything runs on the current CUDA stream without host syncs, so
    the whole call is CUDA-graph capturable (see :class:`GenerationGraph`).

    Loop-invariant work the reference repeats every step is hoisted: the conv conditioning encoder, FiLM
    parameters and cond_proj run once per model, and the per-step timestep vectors are one batched product."""
    from ..corruptions import keyframes as kf
    from ..utils.clamp import stage2_epilogue
    from ..utils.normalize import sigmoid_pos

    cfg = cfg or GenerationConfig()
    sg = cond["start_goal"]
    dev = L.require_cuda(sg, cond["occ"])
    B = sg.shape[0]
    T, K, S, D = cfg.T, cfg.K_min, cfg.levels, cfg.data_dim
    # 1. anchors
    if idx is None:
        if cfg.kp_index_mode == "uniform":
            idx, masks = _uniform_idx_cache(B, T, K, dev)       # batch-invariant row: built once per shape
        elif cfg.kp_index_mode == "random":
            idx, masks = kf.sample_fixed_k_indices_batch(B, T, K, generator=generator, device=dev)
        else:
            raise ValueError(f"kp_index_mode={cfg.kp_index_mode!r} needs explicit idx= (selector models are out of scope)")
    else:
        masks = torch.zeros((B, T), device=dev, dtype=torch.bool)
        masks.scatter_(1, idx, True)
    K = idx.shape[1]
    # 2. known endpoints (logit space)
    known_mask, known_values = _build_known_mask_values(idx, cond, D, T, cfg.clamp_endpoints, logit_space=cfg.logit_space,
                                                        logit_eps=cfg.logit_eps)
    # 3. initial noise
    z = torch.randn((B, K, D), device=dev) if z_T is None else L.f32c(z_T).clone()
    z = torch.where(known_mask, known_values, z)
    clip1 = bool(cfg.pos_clip) and not bool(cfg.logit_space)
    if clip1:
        z[..., :2] = z[..., :2].clamp(min=cfg.pos_clip_min, max=cfg.pos_clip_max)
    # 4. Stage-1 DDIM
    sched = _schedule_cache(cfg.beta_schedule, cfg.n_train)
    times = _timesteps(cfg.n_train, cfg.ddim_steps, schedule=cfg.ddim_schedule).tolist()
    ab = sched["alpha_bar_host"]
    cond_vec = kp_model.encode_cond(cond)
    pk = kp_model.transformer.packed()
    film = pk.film_params(cond_vec, K, kp_model.precision)
    row_b = _cond_row(kp_model, cond_vec, T, dev)
    t_vecs = _cached(kp_model, ("t_vecs", tuple(times), str(dev)), [kp_model.t_embed[0].weight, kp_model.t_embed[0].bias, kp_model.t_embed[2].weight, kp_model.t_embed[2].bias],
                     lambda: kp_model.timestep_vector(torch.tensor(times[:-1], device=dev, dtype=torch.long)))
    eps = torch.empty((B, K, D), device=dev, dtype=torch.float32)
    for i in range(len(times) - 1):
        kp_model(z, None, idx, known_mask, None, T, cond_vec=cond_vec, film=film, t_vec=t_vecs[i:i + 1], row_b=row_b, out=eps)
        ddim_step_scalar(z, eps, float(ab[times[i]]), float(ab[times[i + 1]]), known_mask=known_mask, known_values=known_values,
                         pos_clip=clip1, pos_clip_min=cfg.pos_clip_min, pos_clip_max=cfg.pos_clip_max, out=z)
    # 5-6. back to position space, interpolate to T
    z_pred = sigmoid_pos(z) if cfg.logit_space else z
    x_pred = kf.interpolate_from_indices(idx, z_pred, T, recompute_velocity=bool(cfg.recompute_vel))
    # 7-10. Stage 2
    cond_vec2 = interp_model.encode_cond(cond)
    pk2 = interp_model.transformer.packed()
    film2 = pk2.film_params(cond_vec2, T, interp_model.precision)
    row_b2 = _cond_row(interp_model, cond_vec2, T, dev)
    ac = dict(conf_teacher=cfg.anchor_conf_teacher, conf_student=cfg.anchor_conf_student,
              conf_endpoints=cfg.anchor_conf_endpoints, conf_missing=cfg.anchor_conf_missing,
              clamp_endpoints=cfg.clamp_endpoints)
    x_hat = out if out is not None else torch.empty((B, T, D), device=dev, dtype=torch.float32)
    if cfg.stage2_mode == "x0":
        level_vec = _level_vectors(interp_model, S, dev)[S:S + 1]
        conf_pred = None
        if cfg.anchor_conf:
            # conf_pred (student == mask) and the annealed copy fed to the model; anneal at s == S is the identity
            conf_pred, _ = anchor_conf_mask_in(masks, masks, None, S, S, "none", **ac)
            _, mask_in = anchor_conf_mask_in(masks, masks, None, S, S, cfg.anchor_conf_anneal_mode, want_conf=False, channels=2, **ac)
        else:
            mask_in = masks
        delta = interp_model(x_pred, None, mask_in, None, cond_vec=cond_vec2, film=film2, level_vec=level_vec, row_b=row_b2)
        lam = _soft_clamp_lambda(S, S, cfg.soft_clamp_schedule, cfg.soft_clamp_max) if (cfg.soft_anchor_clamp and conf_pred is not None) else 0.0
        stage2_epilogue(x_pred, delta, x_pred, conf_pred if lam > 0.0 else None, lam, cfg.clamp_policy, masks, cfg.clamp_dims, out=x_hat)
    else:
        if masks_levels is None:
            masks_levels, _ = kf.build_nested_masks_from_base(idx, T, S, generator=generator, k_schedule=cfg.k_schedule)
        level_vecs = _level_vectors(interp_model, S, dev)
        x_curr = x_pred
        for s in range(S, 0, -1):
            m_s, m_prev = masks_levels[:, s].contiguous(), masks_levels[:, s - 1].contiguous()
            if cfg.anchor_conf:
                conf_s, mask_in = anchor_conf_mask_in(m_s, None, m_prev, s, S, cfg.anchor_conf_anneal_mode, channels=3, **ac)
            else:
                conf_s, mask_in = None, torch.stack([m_s, m_prev], dim=-1)
            delta = interp_model(x_curr, None, mask_in, None, cond_vec=cond_vec2, film=film2, level_vec=level_vecs[s:s + 1], row_b=row_b2)
            lam = _soft_clamp_lambda(s, S, cfg.soft_clamp_schedule, cfg.soft_clamp_max) if (cfg.soft_anchor_clamp and conf_s is not None) else 0.0
            tgt = x_hat if s == 1 else torch.empty_like(x_pred)
            x_curr = stage2_epilogue(x_curr, delta, x_pred, conf_s if lam > 0.0 else None, lam, cfg.clamp_policy, m_s, cfg.clamp_dims,
                                     pos_clip=cfg.pos_clip, pos_clip_min=cfg.pos_clip_min, pos_clip_max=cfg.pos_clip_max, out=tgt)
    if return_all:
        return {"x_hat": x_hat, "x_pred": x_pred, "z": z, "z_pred": z_pred, "idx": idx, "masks": masks}
    return x_hat


_SCHEDULES: Dict = {}
_UNIFORM_IDX: Dict = {}


def _uniform_idx_cache(B: int, T: int, K: int, dev):
    from ..corruptions import keyframes as kf
    key = (B, T, K, str(dev))
    if key not in _UNIFORM_IDX:
        # one entry per shape, never evicted: a captured GenerationGraph reads these tensors on every replay (8 B x K per
        # trajectory; evicting on a new shape would free memory an older graph still points at)
        _UNIFORM_IDX[key] = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device=dev)
    return _UNIFORM_IDX[key]


def _cached(model, key, params, fn):
    """Per-model cache of small derived tensors (timestep / level vectors), invalidated when the weights change.
    Keeps host->device copies and tiny GEMMs out of the per-call (and graph-captured) path."""
    from ..models import _engine as E
    sig = E._sig(params)
    hit = model._cache.get(key)
    if hit is None or hit[0] != sig:
        hit = (sig, fn())
        model._cache[key] = hit
    return hit[1]


def _level_vectors(interp_model, S: int, dev):
    params = [interp_model.level_emb.weight, interp_model.level_proj[0].weight, interp_model.level_proj[0].bias,
              interp_model.level_proj[2].weight, interp_model.level_proj[2].bias]
    return _cached(interp_model, ("level_vecs", S, str(dev)), params,
                   lambda: interp_model.level_vector(torch.arange(0, S + 1, device=dev, dtype=torch.long)))


def _schedule_cache(name: str, n_train: int):
    from ..diffusion.schedules import make_alpha_bars, make_beta_schedule
    key = (name, n_train)
    if key not in _SCHEDULES:
        sch = make_alpha_bars(make_beta_schedule(name, n_train))
        sch["alpha_bar_host"] = sch["alpha_bar"].to(torch.float32).numpy()
        _SCHEDULES[key] = sch
    return _SCHEDULES[key]


def _cond_row(model, cond_vec: torch.Tensor, T: int, dev) -> torch.Tensor:
    from ..models import _engine as E
    der = model._derived(T, dev)
    return E.sgemm(cond_vec, model.cond_proj.weight.detach().float().contiguous(), der["bias_b"])


def _captured_tensors(*roots):
    """Every tensor reachable from the models' derived-tensor caches / packed encoders / workspaces (references, not copies)."""
    seen, out, stack = set(), [], list(roots)
    while stack:
        o = stack.pop()
        if o is None or id(o) in seen:
            continue
        seen.add(id(o))
        if isinstance(o, torch.Tensor):
            out.append(o)
        elif isinstance(o, dict):
            stack.extend(o.values())
        elif isinstance(o, (list, tuple)):
            stack.extend(o)
        elif isinstance(o, torch.nn.Module):
            for name in ("_cache", "_ws", "_packed", "_gemm_ws", "_w1_packed"):
                stack.append(getattr(o, name, None))
            stack.extend(o.children())
        elif hasattr(o, "__dict__") and type(o).__module__.startswith("interpolated_diffusion_b200"):
            stack.extend(v for k, v in vars(o).items() if k != "enc")
    return out


class GenerationGraph:
    """The whole generation call captured once as a single CUDA graph for a fixed batch shape (19 Stage-1
    evaluations + DDIM updates, sigmoid, interpolation, Stage-2, clamp: ~600 kernel launches replayed with one
    host call).  Inputs are copied into static buffers, ``run`` replays the graph and returns the static output."""

    def __init__(self, kp_model, interp_model, B: int, cfg: Optional[GenerationConfig] = None, *, occ_shape=(1, 21, 21),
                 use_sdf: bool = False, device=None):
        self.cfg = cfg or GenerationConfig()
        self.kp, self.il = kp_model, interp_model
        dev = L.resolve_device(device)
        self.dev = dev
        K, D, T = self.cfg.K_min, self.cfg.data_dim, self.cfg.T
        self.cond = {"occ": torch.zeros((B,) + tuple(occ_shape), device=dev), "start_goal": torch.full((B, 4), 0.5, device=dev)}
        if use_sdf:
            self.cond["sdf"] = torch.zeros((B,) + tuple(occ_shape), device=dev)
        self.z_T = torch.zeros((B, K, D), device=dev)
        self.x_hat = torch.empty((B, T, D), device=dev)
        self.masks_levels = None
        if self.cfg.stage2_mode == "adj":
            from ..corruptions import keyframes as kf
            idx, _ = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device=dev)
            self.masks_levels, _ = kf.build_nested_masks_from_base(idx, T, self.cfg.levels)
        self.graph = None
        self.launches = 0
        self._plist = [p for m in (kp_model, interp_model) for p in m.parameters()]
        self._sig = None
        self._owned = None

    def _weights_sig(self):
        from ..models import _engine as E
        return E._sig(self._plist)

    def _body(self):
        generate(self.kp, self.il, self.cond, self.cfg, z_T=self.z_T, masks_levels=self.masks_levels, out=self.x_hat)

    def capture(self):
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):                     # warm-up: packs weights, sizes workspaces, sets kernel attributes
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        from ..models import _engine as E
        E.note_graph_captured()                    # from here on workspaces retire (never free) buffers they outgrow
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()
        # The graph reads tensors it does not allocate: packed bf16 weights, folded FiLM weights, timestep / level vectors,
        # position tables, the uniform anchor rows.  They were produced by the warm-up (cache hits during capture), so the
        # graph is only valid for the parameter version it was captured at: remember it (run() re-captures on a mismatch) and
        # keep the captured tensors alive for as long as this graph exists.
        self._sig = self._weights_sig()
        self._owned = _captured_tensors(self.kp, self.il, E._CONV_WS, _UNIFORM_IDX.get((self.z_T.shape[0], self.cfg.T, self.cfg.K_min, str(self.dev))))
        return self

    def run(self, cond: Optional[Dict[str, torch.Tensor]] = None, z_T: Optional[torch.Tensor] = None) -> torch.Tensor:
        if self.graph is None or self._sig != self._weights_sig():
            # first use, or the weights changed since capture (optimizer step, EMA copy_to, load_state_dict): the captured
            # launches point at packed copies of the OLD weights -- capture again instead of silently sampling stale weights
            self.capture()
        if cond is not None:
            for k, buf in self.cond.items():
                buf.copy_(cond[k], non_blocking=True)
        if z_T is not None:
            self.z_T.copy_(z_T, non_blocking=True)
        self.graph.replay()
        return self.x_hat
