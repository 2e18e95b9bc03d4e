#!/usr/bin/env python
"""Generate the golden fixtures in this directory from the LIVE reference.

Run in the build container only (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_golden.py

Every ``*.npz`` written here is the output of the reference's own functions
(EquilibriaW/Interpolated_Diffusion, imported read-only from /root/reference) on seeded
inputs, together with those inputs and every random draw the reference made, so that
``oracle/`` can be pinned against them without the reference being present.
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("IDB200_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))

# sample_generate imports matplotlib via src/eval/visualize.py; stub it.
for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches"):
    if name not in sys.modules:
        sys.modules[name] = types.ModuleType(name)
sys.modules["matplotlib.patches"].Polygon = object
sys.modules["matplotlib.patches"].Rectangle = object
sys.path.insert(0, REF)

from src.corruptions import keyframes as rk  # noqa: E402
from src.diffusion import ddpm as rd  # noqa: E402
from src.diffusion import schedules as rs  # noqa: E402
from src.models.denoiser_interp_levels import InterpLevelDenoiser  # noqa: E402
from src.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser  # noqa: E402
from src.models.denoiser_keypoints import KeypointDenoiser  # noqa: E402
from src.sample import sample_generate as rg  # noqa: E402
from src.train import train_interp_levels as rt  # noqa: E402
from src.utils import clamp as rc  # noqa: E402
from src.utils import normalize as rn  # noqa: E402

torch.set_num_threads(4)
torch.use_deterministic_algorithms(False)


def npy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: npy(v) for k, v in d.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(d)} arrays")


# --------------------------------------------------------------------------- #
def gold_keyframes():
    g = {}
    # k schedules
    cases = [(64, 8, 3), (256, 32, 4), (128, 8, 3), (16, 3, 3), (10, 4, 2), (5, 2, 1), (64, 8, 1), (33, 5, 4)]
    rows = []
    for (T, K, S) in cases:
        for sched in ("doubling", "linear", "geom"):
            ks = rk._compute_k_schedule(T, K, S, schedule=sched)
            rows.append([T, K, S, {"doubling": 0, "linear": 1, "geom": 2}[sched]] + ks + [-1] * (6 - len(ks)))
    g["ksched"] = np.array(rows, dtype=np.int64)

    # uniform indices (no jitter): many (T,K)
    uni = []
    for T in (8, 16, 33, 64, 100, 128, 256):
        for K in (2, 3, 5, 8, 16, 32):
            if K > T:
                continue
            idx, mask = rk.sample_fixed_k_indices_uniform_batch(2, T, K)
            uni.append(np.concatenate([[T, K], npy(idx[0]), -np.ones(32 - K, dtype=np.int64)]))
    g["uniform"] = np.array(uni, dtype=np.int64)
    # uniform with jitter: record u
    for tag, (B, T, K, jit) in {"a": (16, 64, 8, 0.5), "b": (8, 256, 32, 1.0)}.items():
        gen = torch.Generator().manual_seed(11)
        idx, mask = rk.sample_fixed_k_indices_uniform_batch(B, T, K, generator=gen, jitter=jit)
        u = torch.rand((B, K), generator=torch.Generator().manual_seed(11))
        g[f"unij_{tag}_u"] = u
        g[f"unij_{tag}_idx"] = idx
        g[f"unij_{tag}_mask"] = mask
        g[f"unij_{tag}_cfg"] = np.array([B, T, K, jit], dtype=np.float64)

    # random fixed-K indices: record scores
    for tag, (B, T, K, ee) in {"a": (32, 64, 8, True), "b": (8, 10, 4, True), "c": (8, 12, 5, False),
                               "d": (4, 2, 2, True), "e": (4, 16, 2, True)}.items():
        gen = torch.Generator().manual_seed(5)
        idx, mask = rk.sample_fixed_k_indices_batch(B, T, K, generator=gen, ensure_endpoints=ee)
        n = T - 2 if ee else T
        sc = torch.rand((B, n), generator=torch.Generator().manual_seed(5)) if (n > 0 and (not ee or (T > 2 and K > 2))) else torch.zeros((B, max(n, 0)))
        g[f"fixk_{tag}_scores"] = sc
        g[f"fixk_{tag}_idx"] = idx
        g[f"fixk_{tag}_mask"] = mask
        g[f"fixk_{tag}_cfg"] = np.array([B, T, K, int(ee)], dtype=np.int64)

    # nested masks: record scores; includes forced ties (quantised scores)
    orig_rand = torch.rand
    for tag, (B, T, K, S, sched, quant) in {
        "a": (64, 64, 8, 3, "doubling", 0), "b": (16, 256, 32, 4, "doubling", 0),
        "c": (8, 16, 3, 3, "doubling", 0), "d": (8, 64, 8, 3, "linear", 0),
        "e": (8, 128, 8, 3, "geom", 0), "t": (64, 64, 8, 3, "doubling", 8),
        "f": (4, 5, 2, 1, "doubling", 0), "g": (4, 3, 2, 2, "doubling", 0),
    }.items():
        scores = orig_rand((B, T - 2), generator=torch.Generator().manual_seed(1234))
        if quant:
            scores = torch.floor(scores * quant) / quant        # many exact ties

        def fake_rand(shape, generator=None, device=None, _s=scores):
            assert tuple(shape) == tuple(_s.shape)
            return _s.clone()

        # torch.argsort on CPU is NOT stable for rows longer than 16 (measured: 11 of 200 000 natural
        # fp32 rows of 62 order a tie differently from stable=True), and the CUDA sort differs again, so
        # tie order is implementation-defined in the reference.  The contract of this repo is the stable
        # order (lower index first); the forced-tie case is therefore recorded with stable=True.
        orig_argsort = torch.argsort
        torch.rand = fake_rand
        if quant:
            torch.argsort = lambda x, dim=-1, **kw: orig_argsort(x, dim=dim, stable=True)
        try:
            masks, idxs = rk.build_nested_masks_batch(B, T, K, S, k_schedule=sched)
        finally:
            torch.rand = orig_rand
            torch.argsort = orig_argsort
        g[f"nest_{tag}_scores"] = scores
        g[f"nest_{tag}_masks"] = masks
        for s, ix in enumerate(idxs):
            g[f"nest_{tag}_idx{s}"] = ix
        g[f"nest_{tag}_cfg"] = np.array([B, T, K, S, {"doubling": 0, "linear": 1, "geom": 2}[sched]], dtype=np.int64)

    # from_base: record the randperm sequence
    for tag, (B, T, K, S) in {"a": (6, 64, 8, 3), "b": (3, 16, 3, 2)}.items():
        idx_base, _ = rk.sample_fixed_k_indices_uniform_batch(B, T, K)
        perms = []
        orig_rp = torch.randperm

        def rec_rp(n, generator=None, device=None):
            p = orig_rp(n, generator=generator)
            perms.append(p.clone())
            return p

        torch.randperm = rec_rp
        try:
            masks, idxs = rk.build_nested_masks_from_base(idx_base, T, S, generator=torch.Generator().manual_seed(3))
        finally:
            torch.randperm = orig_rp
        g[f"base_{tag}_idx_base"] = idx_base
        g[f"base_{tag}_masks"] = masks
        for s, ix in enumerate(idxs):
            g[f"base_{tag}_idx{s}"] = ix
        g[f"base_{tag}_nperm"] = np.array([len(perms)])
        for i, p in enumerate(perms):
            g[f"base_{tag}_perm{i}"] = p
        g[f"base_{tag}_cfg"] = np.array([B, T, K, S], dtype=np.int64)

    # from_logits (with ties) and from_level_logits (no ties)
    gen = torch.Generator().manual_seed(9)
    logits = torch.randn((6, 32), generator=gen)
    logits[0, 5] = logits[0, 9]
    logits[1] = torch.round(logits[1] * 2) / 2
    orig_argsort = torch.argsort          # rows 0 and 1 hold exact ties: record the stable order (see above)
    torch.argsort = lambda x, dim=-1, descending=False, **kw: orig_argsort(x, dim=dim, descending=descending, stable=True)
    try:
        masks, idxs = rk.build_nested_masks_from_logits(logits, 4, 2)
    finally:
        torch.argsort = orig_argsort
    g["logit_logits"] = logits
    g["logit_masks"] = masks
    for s, ix in enumerate(idxs):
        g[f"logit_idx{s}"] = ix
    ll = torch.randn((5, 3, 32), generator=gen)
    masks, idxs = rk.build_nested_masks_from_level_logits(ll, 4, 2)
    g["lvl_logits"] = ll
    g["lvl_masks"] = masks
    for s, ix in enumerate(idxs):
        g[f"lvl_idx{s}"] = ix

    # interpolate_from_indices
    gen = torch.Generator().manual_seed(21)
    for tag, (B, T, K, D, vel) in {"a": (16, 64, 8, 2, False), "b": (16, 64, 8, 4, True), "c": (8, 256, 32, 4, True),
                                   "d": (8, 64, 32, 4, False), "e": (4, 100, 7, 4, True), "f": (4, 64, 64, 2, False),
                                   "g": (4, 20, 2, 3, False)}.items():
        idx, _ = rk.sample_fixed_k_indices_batch(B, T, K, generator=gen)
        vals = torch.rand((B, K, D), generator=gen) * 2 - 0.5
        y = rk.interpolate_from_indices(idx, vals, T, recompute_velocity=vel)
        g[f"interp_{tag}_idx"] = idx
        g[f"interp_{tag}_vals"] = vals
        g[f"interp_{tag}_y"] = y
        g[f"interp_{tag}_cfg"] = np.array([B, T, K, D, int(vel)], dtype=np.int64)
    # extrapolation (idx[0]>0, idx[-1]<T-1) and duplicate indices
    idx = torch.tensor([[2, 5, 9, 12], [0, 4, 4, 15], [3, 3, 8, 8]], dtype=torch.long)
    vals = torch.rand((3, 4, 2), generator=gen)
    g["interp_x_idx"] = idx
    g["interp_x_vals"] = vals
    g["interp_x_y"] = rk.interpolate_from_indices(idx, vals, 16)
    # known answer from SURVEY 8c
    g["interp_ka_y"] = rk.interpolate_from_indices(torch.tensor([[0, 3, 6, 7]]), torch.tensor([[[0.0], [3.0], [12.0], [7.0]]]), 8)
    # interpolate_from_mask (legacy loop) vs the idx path: store inputs + outputs
    x = torch.rand((5, 24, 4), generator=gen)
    m = torch.zeros((5, 24), dtype=torch.bool)
    for b in range(5):
        m[b] = rk.sample_fixed_k_mask(24, 3 + b, generator=gen)
    m[4] = False
    m[4, 3] = True
    m[4, 17] = True          # anchors not at the ends: legacy loop keeps x outside [3,17]
    g["imask_x"] = x
    g["imask_m"] = m
    g["imask_y"] = rk.interpolate_from_mask(x, m, recompute_velocity=False)
    g["imask_yv"] = rk.interpolate_from_mask(x, m, recompute_velocity=True)
    save("keyframes.npz", g)


# --------------------------------------------------------------------------- #
def gold_diffusion():
    g = {}
    for name in ("linear", "cosine"):
        for n in (10, 200, 1000):
            sch = rs.make_alpha_bars(rs.make_beta_schedule(name, n))
            for k, v in sch.items():
                g[f"sched_{name}_{n}_{k}"] = v
    ts = []
    for (n, steps, sched) in [(1000, 20, "quadratic"), (1000, 20, "linear"), (1000, 20, "sqrt"), (1000, 50, "quadratic"),
                              (200, 10, "quadratic"), (1000, 1, "linear"), (10, 20, "linear"), (1000, 5, "sqrt"),
                              (1000, 100, "quadratic")]:
        t = npy(rd._timesteps(n, steps, sched))
        g[f"ts_{n}_{steps}_{sched}"] = t
    gen = torch.Generator().manual_seed(2)
    sch = rs.make_alpha_bars(rs.make_beta_schedule("cosine", 1000))
    z = torch.randn((8, 8, 4), generator=gen)
    eps = torch.randn((8, 8, 4), generator=gen)
    pairs = [(999, 896), (896, 799), (334, 276), (2, 0), (11, 2)]
    for i, (t, tp) in enumerate(pairs):
        out = rd.ddim_step(z, eps, torch.full((8,), t), torch.full((8,), tp), sch)
        g[f"ddim_{i}_out"] = out
    g["ddim_pairs"] = np.array(pairs, dtype=np.int64)
    g["ddim_z"] = z
    g["ddim_eps"] = eps
    # per-sample t, and [B,K] t
    t = torch.randint(1, 1000, (8,), generator=gen)
    tp = torch.clamp(t - torch.randint(1, 100, (8,), generator=gen), min=0)
    g["ddim_v_t"] = t
    g["ddim_v_tp"] = tp
    g["ddim_v_out"] = rd.ddim_step(z, eps, t, tp, sch)
    # stochastic branch: record the randn_like draw
    torch.manual_seed(77)
    out = rd.ddim_step(z, eps, t, tp, sch, eta=0.5)
    torch.manual_seed(77)
    g["ddim_eta_noise"] = torch.randn_like(z)
    g["ddim_eta_out"] = out
    # q_sample
    noise = torch.randn((8, 8, 4), generator=gen)
    rt_, _ = rd.q_sample(z, t, sch, noise=noise)
    g["q_noise"] = noise
    g["q_out"] = rt_
    save("diffusion.npz", g)


# --------------------------------------------------------------------------- #
def gold_sampling():
    g = {}
    gen = torch.Generator().manual_seed(31)
    B, T, D = 6, 16, 4
    x_hat = torch.randn((B, T, D), generator=gen)
    x_ref = torch.randn((B, T, D), generator=gen)
    mask = torch.rand((B, T), generator=gen) < 0.3
    conf = torch.rand((B, T), generator=gen)
    g["cl_x_hat"], g["cl_x_ref"], g["cl_mask"], g["cl_conf"] = x_hat, x_ref, mask, conf
    g["cl_hard_pos"] = rc.apply_clamp(x_hat.clone(), x_ref, mask, "pos")
    g["cl_hard_all"] = rc.apply_clamp(x_hat.clone(), x_ref, mask, "all")
    g["cl_soft_pos"] = rc.apply_soft_clamp(x_hat.clone(), x_ref, conf, 0.7, "pos")
    g["cl_soft_all"] = rc.apply_soft_clamp(x_hat.clone(), x_ref, conf, 0.7, "all")
    p = torch.rand((B, T, D), generator=gen) * 1.2 - 0.1
    g["nz_p"] = p
    g["nz_logit"] = rn.logit_pos(p, eps=1e-5)
    g["nz_sigmoid"] = rn.sigmoid_pos(x_hat * 3)
    # known mask / values
    idx = torch.tensor([[0, 3, 6, 15], [0, 1, 2, 15], [2, 3, 4, 15], [0, 5, 9, 12], [0, 7, 8, 15], [0, 2, 4, 15]])
    sg = torch.rand((B, 4), generator=gen)
    for Dk in (2, 4):
        km, kv = rg._build_known_mask_values(idx, {"start_goal": sg}, Dk, T, True)
        g[f"kn_mask_{Dk}"], g[f"kn_vals_{Dk}"] = km, kv
    g["kn_idx"], g["kn_sg"] = idx, sg
    # anchor conf (+ anneal, scalar and per-row)
    st = torch.rand((B, T), generator=gen) < 0.5
    g["cf_student"] = st
    g["cf_a"] = rg._build_anchor_conf(mask, mask, True, 0.95, 0.5, 1.0, 0.0, True)
    g["cf_b"] = rg._build_anchor_conf(mask, None, False, 0.95, 0.5, 1.0, 0.0, True)
    g["cf_c"] = rg._build_anchor_conf(mask, st, True, 0.9, 0.4, 0.8, 0.1, False)
    for mode in ("linear", "cosine", "none"):
        for s in (1, 2, 3):
            g[f"an_{mode}_{s}"] = rg._anneal_conf(g["cf_b"].clone(), s, 3, mode)
    s_idx = torch.tensor([1, 2, 3, 3, 1, 2])
    for mode in ("linear", "cosine"):
        g[f"anv_{mode}"] = rt._anneal_conf(g["cf_b"].clone(), s_idx, 3, mode)
    g["an_s_idx"] = s_idx
    lam = []
    for sched in ("linear", "cosine", "const"):
        for s in (0, 1, 2, 3):
            lam.append(rg._soft_clamp_lambda(s, 3, sched, 0.8))
    g["lam"] = np.array(lam, dtype=np.float64)
    sig = []
    for K in (8, 16, 32, 64):
        sig.append(rt._compute_sigma_for_level(K, 8, 0.08, 0.012, 0.75))
        sig.append(float(rt._compute_jitter_for_level(K, 8, 3, 1.0)))
    g["sigma_jitter"] = np.array(sig, dtype=np.float64)

    # _distance_alpha / _corrupt_from_anchors: record every generator draw
    B, T, D, K = 8, 32, 4, 6
    src = torch.rand((B, T, D), generator=gen)
    idx, _ = rk.sample_fixed_k_indices_batch(B, T, K, generator=gen)
    g["co_src"], g["co_idx"] = src, idx
    g["co_alpha"] = rt._distance_alpha(idx, T)

    def run_corrupt(tag, jitter, jprob, mode, vel, sigma, asig):
        draws = []
        o_randn, o_rand, o_randint = torch.randn, torch.rand, torch.randint

        def w_randn(*a, **k):
            r = o_randn(*a, **k); draws.append(r.clone()); return r

        def w_rand(*a, **k):
            r = o_rand(*a, **k); draws.append(r.clone()); return r

        def w_randint(*a, **k):
            r = o_randint(*a, **k); draws.append(r.clone()); return r

        torch.randn, torch.rand, torch.randint = w_randn, w_rand, w_randint
        try:
            out = rt._corrupt_from_anchors(src, idx, T, torch.Generator().manual_seed(41), sigma, asig, jitter, jprob,
                                           mode, True, vel)
        finally:
            torch.randn, torch.rand, torch.randint = o_randn, o_rand, o_randint
        g[f"co_{tag}_out"] = out
        g[f"co_{tag}_n"] = np.array([len(draws)])
        for i, d_ in enumerate(draws):
            g[f"co_{tag}_draw{i}"] = d_
        g[f"co_{tag}_cfg"] = np.array([jitter, jprob, {"dist": 0, "const": 1}[mode], int(vel), sigma, asig], dtype=np.float64)

    run_corrupt("a", 0, 0.0, "dist", True, 0.08, 0.02)
    run_corrupt("b", 0, 0.0, "const", False, 0.05, 0.0)
    run_corrupt("c", 2, 0.5, "dist", True, 0.03, 0.01)
    run_corrupt("d", 0, 0.0, "dist", False, 0.0, 0.0)

    # build_interp_adjacent_batch / level_batch, given masks + s_idx, modes none & dist
    B, T, D, K_min, S = 12, 64, 2, 8, 3
    x0 = torch.rand((B, T, D), generator=gen)
    masks_levels, idx_levels = rk.build_nested_masks_batch(B, T, K_min, S, generator=gen)
    s_idx = torch.tensor([1, 2, 3, 3, 3, 2, 1, 3, 3, 2, 3, 1])
    g["ad_x0"], g["ad_masks"], g["ad_s_idx"] = x0, masks_levels, s_idx
    for s, ix in enumerate(idx_levels):
        g[f"ad_idx{s}"] = ix
    xs, xp, ms, mp, *_ = rt.build_interp_adjacent_batch(x0, K_min, S, torch.Generator().manual_seed(1), masks_levels=masks_levels,
                                                         idx_levels=idx_levels, s_idx=s_idx)
    g["ad_none_xs"], g["ad_none_xp"], g["ad_none_ms"], g["ad_none_mp"] = xs, xp, ms, mp
    draws = []
    o_randn = torch.randn

    def w_randn(*a, **k):
        r = o_randn(*a, **k); draws.append(r.clone()); return r

    torch.randn = w_randn
    try:
        xs, xp, ms, mp, *_ = rt.build_interp_adjacent_batch(
            x0, K_min, S, torch.Generator().manual_seed(1), masks_levels=masks_levels, idx_levels=idx_levels, s_idx=s_idx,
            corrupt_mode="dist", corrupt_sigma_max=0.08, corrupt_sigma_min=0.012, corrupt_sigma_pow=0.75,
            corrupt_anchor_frac=0.25)
    finally:
        torch.randn = o_randn
    g["ad_dist_xs"], g["ad_dist_xp"] = xs, xp
    g["ad_dist_n"] = np.array([len(draws)])
    for i, d_ in enumerate(draws):
        g[f"ad_dist_draw{i}"] = d_
    xs, ms, *_ = rt.build_interp_level_batch(x0, K_min, S, torch.Generator().manual_seed(1), masks_levels=masks_levels,
                                             idx_levels=idx_levels, s_idx=s_idx)
    g["lv_none_xs"], g["lv_none_ms"] = xs, ms
    save("sampling.npz", g)


# --------------------------------------------------------------------------- #
TINY = dict(d_model=64, n_layers=2, n_heads=2, d_ff=128, d_cond=32, maze_channels=(8, 16))


def _cond(B, gen, sdf=False, hw=(21, 21)):
    c = {"occ": (torch.rand((B, 1, *hw), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    if sdf:
        c["sdf"] = torch.rand((B, 1, *hw), generator=gen)
    return c


def gold_models():
    g = {}
    gen = torch.Generator().manual_seed(101)
    B, T, K = 3, 16, 5
    # Stage 1, D=2, kp_feat_dim=0
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=2, **TINY).eval()
    for k, v in kp.state_dict().items():
        g["kp2/" + k] = v
    cond = _cond(B, gen)
    idx, _ = rk.sample_fixed_k_indices_batch(B, T, K, generator=gen)
    z = torch.randn((B, K, 2), generator=gen)
    t = torch.tensor([999, 500, 3])
    km, _ = rg._build_known_mask_values(idx, cond, 2, T, True)
    with torch.no_grad():
        g["kp2_eps"] = kp(z, t, idx, km, cond, T)
        g["kp2_condvec"] = kp.cond_enc(cond)
    g["kp2_z"], g["kp2_t"], g["kp2_idx"], g["kp2_km"] = z, t, idx, km
    g["kp2_occ"], g["kp2_sg"] = cond["occ"], cond["start_goal"]
    # Stage 1, D=4, kp_feat_dim=3, use_sdf
    torch.manual_seed(1)
    kp4 = KeypointDenoiser(data_dim=4, kp_feat_dim=3, use_sdf=True, **TINY).eval()
    for k, v in kp4.state_dict().items():
        g["kp4/" + k] = v
    cond4 = _cond(B, gen, sdf=True, hw=(9, 12))
    cond4["kp_feat"] = rg._kp_feat_from_idx(idx, T, 3)
    z4 = torch.randn((B, K, 4), generator=gen)
    km4, _ = rg._build_known_mask_values(idx, cond4, 4, T, True)
    with torch.no_grad():
        g["kp4_eps"] = kp4(z4, t, idx, km4, cond4, T)
    g["kp4_z"], g["kp4_km"], g["kp4_kpfeat"] = z4, km4, cond4["kp_feat"]
    g["kp4_occ"], g["kp4_sg"], g["kp4_sdf"] = cond4["occ"], cond4["start_goal"], cond4["sdf"]
    # Stage 2 bidirectional, C=2 and C=3, and causal C=1
    for tag, cls, C, D in (("il2", InterpLevelDenoiser, 2, 2), ("il3", InterpLevelDenoiser, 3, 4),
                           ("ic1", InterpLevelCausalDenoiser, 1, 2)):
        torch.manual_seed(2)
        m = cls(data_dim=D, max_levels=3, mask_channels=C, **TINY).eval()
        for k, v in m.state_dict().items():
            g[f"{tag}/" + k] = v
        x = torch.rand((B, T, D), generator=gen)
        s = torch.tensor([3, 1, 2])
        if C == 1:
            mask = torch.rand((B, T), generator=gen) < 0.4
        else:
            mask = torch.rand((B, T, C), generator=gen)
        with torch.no_grad():
            g[f"{tag}_out"] = m(x, s, mask, cond)
        g[f"{tag}_x"], g[f"{tag}_s"], g[f"{tag}_mask"] = x, s, mask
    save("models_tiny.npz", g)


def gold_generate():
    """cfg-1 style pipeline with tiny random-init models, driven through the reference's
    component functions exactly as sample_generate.py:974-1285 does (x0 one-step and adj chain)."""
    g = {}
    gen = torch.Generator().manual_seed(7)
    B, T, K, S, D = 4, 64, 8, 3, 2
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D, **TINY).eval()
    il = InterpLevelDenoiser(data_dim=D, max_levels=S, mask_channels=2, **TINY).eval()
    il3 = InterpLevelDenoiser(data_dim=D, max_levels=S, mask_channels=3, **TINY).eval()
    for k, v in kp.state_dict().items():
        g["kp/" + k] = v
    for k, v in il.state_dict().items():
        g["il/" + k] = v
    for k, v in il3.state_dict().items():
        g["il3/" + k] = v
    cond = _cond(B, gen)
    g["occ"], g["sg"] = cond["occ"], cond["start_goal"]
    schedule = rs.make_alpha_bars(rs.make_beta_schedule("cosine", 1000))
    with torch.no_grad():
        idx, masks = rk.sample_fixed_k_indices_uniform_batch(B, T, K)
        km, kv = rg._build_known_mask_values(idx, cond, D, T, True)
        kv = rn.logit_pos(kv, eps=1e-5)
        torch.manual_seed(123)
        z, inter = rg._sample_keypoints_ddim(kp, schedule, idx, km, kv, cond, 20, T, schedule_name="quadratic",
                                             return_intermediates=True)
        torch.manual_seed(123)
        g["z_T"] = torch.randn((B, K, D))
        g["z_inter"] = torch.stack(inter)
        g["z"] = z
        z_pred = rn.sigmoid_pos(z)
        x_pred = rk.interpolate_from_indices(idx, z_pred, T, recompute_velocity=True)
        g["x_pred"] = x_pred
        conf_pred = rg._build_anchor_conf(masks, masks, True, 0.95, 0.5, 1.0, 0.0, True)
        # x0 one-step (sample_generate.py:1252-1285), all three clamp policies
        s_level = torch.full((B,), S, dtype=torch.long)
        conf_s = rg._anneal_conf(conf_pred, S, S, "linear")
        mask_in = torch.stack([masks.float(), conf_s], dim=-1)
        delta = il(x_pred, s_level, mask_in, cond)
        g["delta"] = delta
        for pol in ("none", "endpoints", "all_anchors"):
            for dims in ("pos", "all"):
                x_hat = x_pred + delta
                lam = rg._soft_clamp_lambda(S, S, "linear", 1.0)
                x_hat = rc.apply_soft_clamp(x_hat, x_pred, conf_pred, lam, dims)
                if pol == "all_anchors":
                    cm = masks
                elif pol == "endpoints":
                    cm = torch.zeros_like(masks); cm[:, 0] = True; cm[:, -1] = True
                else:
                    cm = None
                if cm is not None:
                    x_hat = rc.apply_clamp(x_hat, x_pred, cm, dims)
                g[f"x_hat_x0_{pol}_{dims}"] = x_hat
        # adj chain (sample_generate.py:1133-1204) with from_base masks
        masks_levels, _ = rk.build_nested_masks_from_base(idx, T, S, generator=torch.Generator().manual_seed(5))
        g["adj_masks_levels"] = masks_levels
        x_curr = x_pred
        for s in range(S, 0, -1):
            m_s, m_prev = masks_levels[:, s], masks_levels[:, s - 1]
            conf = rg._build_anchor_conf(m_s, None, False, 0.95, 0.5, 1.0, 0.0, True)
            conf = rg._anneal_conf(conf, s, S, "linear")
            mi = torch.stack([m_s.float(), m_prev.float(), conf], dim=-1)
            x_curr = x_curr + il3(x_curr, torch.full((B,), s, dtype=torch.long), mi, cond)
            x_curr = rc.apply_soft_clamp(x_curr, x_pred, conf, rg._soft_clamp_lambda(s, S, "linear", 1.0), "pos")
            cm = torch.zeros_like(m_s); cm[:, 0] = True; cm[:, -1] = True
            x_curr = rc.apply_clamp(x_curr, x_pred, cm, "pos")
        g["x_hat_adj"] = x_curr
    save("generate_tiny.npz", g)


if __name__ == "__main__":
    gold_keyframes()
    gold_diffusion()
    gold_sampling()
    gold_models()
    gold_generate()
