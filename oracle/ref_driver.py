"""TEST / BASELINE INFRASTRUCTURE -- never imported by the product package.

Drives the UNMODIFIED reference (``baseline/_ref/src`` vendored by ``oracle/vendor_reference.py``, or ``/root/reference`` in the
build container) through the per-batch body of ``src/sample/sample_generate.py:974-1285`` with its own component functions:
uniform anchors -> ``_build_known_mask_values`` -> ``logit_pos`` -> ``_sample_keypoints_ddim`` (20 quadratic timesteps) ->
``sigmoid_pos`` -> ``interpolate_from_indices`` -> ``_build_anchor_conf`` / ``_anneal_conf`` -> ``InterpLevelDenoiser`` one-step
(``x0``) -> ``apply_soft_clamp`` -> ``apply_clamp`` (clamp_policy=endpoints, clamp_dims=pos).  ``main()`` of the reference is not
used: it refuses ``--dataset particle`` and needs checkpoints.  Used by ``bench.py`` (``--impl reference`` and
the ``cpu_baseline`` leg) and by ``tests/golden/make_golden_cfg1.py``.
"""
from __future__ import annotations

import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    """Directory holding the reference's ``src`` package, or None."""
    for cand in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("IDB200_REFERENCE", "/root/reference")):
        if cand and os.path.isfile(os.path.join(cand, "src", "sample", "sample_generate.py")):
            return cand
    return None


def import_reference():
    """Import the reference's modules (matplotlib stubbed: ``src/eval/visualize.py:3`` imports it)."""
    root = reference_root()
    if root is None:
        raise ImportError("reference sources not found (baseline/_ref not vendored and /root/reference absent)")
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
        try:
            __import__(name)
        except Exception:
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
    mp = sys.modules.get("matplotlib.patches")
    if mp is not None and not hasattr(mp, "Polygon"):
        mp.Polygon = object
        mp.Rectangle = object
    if root not in sys.path:
        sys.path.insert(0, root)
    import src.corruptions.keyframes as rk
    import src.diffusion.schedules as rs
    import src.sample.sample_generate as rg
    import src.utils.clamp as rc
    import src.utils.normalize as rn
    from src.models.denoiser_interp_levels import InterpLevelDenoiser
    from src.models.denoiser_keypoints import KeypointDenoiser
    return types.SimpleNamespace(rk=rk, rs=rs, rg=rg, rc=rc, rn=rn, KeypointDenoiser=KeypointDenoiser,
                                 InterpLevelDenoiser=InterpLevelDenoiser, root=root)


def build_models(R, D: int = 2, levels: int = 3, **kw):
    import torch
    torch.manual_seed(0)
    kp = R.KeypointDenoiser(data_dim=D, **kw).eval()
    il = R.InterpLevelDenoiser(data_dim=D, max_levels=levels, mask_channels=2, **kw).eval()
    return kp, il


def generate(R, kp, il, cond, *, T: int = 64, K: int = 8, S: int = 3, D: int = 2, ddim_steps: int = 20, seed: int = 123,
             return_all: bool = False):
    """One batch through the reference's generation path (x0 one-step).  The reference draws z_T from the
    global RNG inside ``_sample_keypoints_ddim`` (:389): ``torch.manual_seed(seed)`` right before it."""
    import torch
    rk, rs, rg, rc, rn = R.rk, R.rs, R.rg, R.rc, R.rn
    B = cond["start_goal"].shape[0]
    schedule = rs.make_alpha_bars(rs.make_beta_schedule("cosine", 1000))
    with torch.no_grad():
        idx, masks = rk.sample_fixed_k_indices_uniform_batch(B, T, K)
        km, kv = rg._build_known_mask_values(idx, cond, D, T, True)
        kv = rn.logit_pos(kv, eps=1e-5)
        torch.manual_seed(seed)
        out = rg._sample_keypoints_ddim(kp, schedule, idx, km, kv, cond, ddim_steps, T, schedule_name="quadratic",
                                        return_intermediates=return_all)
        z, inter = out if return_all else (out, None)
        z_pred = rn.sigmoid_pos(z)
        x_pred = rk.interpolate_from_indices(idx, z_pred, T, recompute_velocity=True)
        conf_pred = rg._build_anchor_conf(masks, masks, True, 0.95, 0.5, 1.0, 0.0, True)
        s_level = torch.full((B,), S, dtype=torch.long)
        conf_s = rg._anneal_conf(conf_pred, S, S, "linear")
        mask_in = torch.stack([masks.float(), conf_s], dim=-1)
        delta = il(x_pred, s_level, mask_in, cond)
        x_hat = x_pred + delta
        x_hat = rc.apply_soft_clamp(x_hat, x_pred, conf_pred, rg._soft_clamp_lambda(S, S, "linear", 1.0), "pos")
        cm = torch.zeros_like(masks)
        cm[:, 0] = True
        cm[:, -1] = True
        x_hat = rc.apply_clamp(x_hat, x_pred, cm, "pos")
    if return_all:
        return {"idx": idx, "masks": masks, "known_mask": km, "known_values": kv, "z_inter": torch.stack(inter), "z": z,
                "x_pred": x_pred, "conf_pred": conf_pred, "mask_in": mask_in, "delta": delta, "x_hat": x_hat}
    return x_hat


def timed_rate(cond_fn, sample_B: int, steps: int = 1, warmup: int = 0):
    """trajectories/s of the reference on the host cores (all threads).  cond_fn(B) -> cond dict (CPU tensors)."""
    import torch
    R = import_reference()
    torch.set_num_threads(os.cpu_count() or 1)
    kp, il = build_models(R)
    cond = cond_fn(sample_B)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        generate(R, kp, il, cond)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return sample_B / dt, dt, torch.get_num_threads()
