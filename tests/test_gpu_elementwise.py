"""GPU parity: K2 / K2b elementwise kernels (DDIM step, q_sample, known mask, conf, clamps, logit/sigmoid)
through the C ABI vs golden vectors and the oracle.  Bars: bit-exact for the pure fp32 arithmetic
(DDIM given identical z / eps, clamps, conf); <= 4 ulp where expf / logf is involved."""
import numpy as np
import pytest
import torch

from oracle import diffusion_np as odf
from oracle import sampling_np as osp

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def eq(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a, b), f"mismatch: {np.sum(a != b)} of {a.size}, max {np.abs(a.astype(np.float64) - b).max()}"


def ulp_close(a, b, ulps):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a, np.float32)
    b = np.asarray(b, np.float32)
    tol = ulps * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32))
    assert np.all(np.abs(a - b) <= tol), f"max diff {np.abs(a - b).max()}"


def _sched(g):
    keys = ("betas", "alphas", "alpha_bar", "sqrt_alpha_bar", "sqrt_one_minus_alpha_bar")
    return {k: g[f"sched_cosine_1000_{k}"] for k in keys}


def test_ddim_step_golden(golden):
    from interpolated_diffusion_b200.diffusion import ddpm
    g = golden("diffusion")
    np_s = _sched(g)
    sch = {k: torch.from_numpy(v) for k, v in np_s.items()}
    z, eps = dev(g["ddim_z"]), dev(g["ddim_eps"])
    for i, (t, tp) in enumerate(g["ddim_pairs"]):
        tt = torch.full((8,), int(t), dtype=torch.long, device="cuda")
        tpp = torch.full((8,), int(tp), dtype=torch.long, device="cuda")
        eq(ddpm.ddim_step(z, eps, tt, tpp, sch), g[f"ddim_{i}_out"])
        eq(ddpm.ddim_step_scalar(z, eps, float(np_s["alpha_bar"][t]), float(np_s["alpha_bar"][tp])), g[f"ddim_{i}_out"])
    # per-row timesteps: the oracle uses IEEE sqrt like the kernel (torch's AVX512 CPU sqrt is 1 ulp off on
    # 0.6 % of inputs, so the golden tensor is compared at 2 ulp of the result scale and the oracle exactly)
    out = ddpm.ddim_step(z, eps, dev(g["ddim_v_t"]), dev(g["ddim_v_tp"]), sch)
    eq(out, odf.ddim_step(g["ddim_z"], g["ddim_eps"], g["ddim_v_t"], g["ddim_v_tp"], np_s))
    np.testing.assert_allclose(out.cpu().numpy(), g["ddim_v_out"], rtol=3e-6, atol=1e-5)
    # [B,K] timesteps
    t2 = torch.randint(1, 1000, (8, 8))
    tp2 = torch.clamp(t2 - 7, min=0)
    out = ddpm.ddim_step(z, eps, t2.cuda(), tp2.cuda(), sch)
    eq(out, odf.ddim_step(g["ddim_z"], g["ddim_eps"], t2.numpy(), tp2.numpy(), np_s))
    # fused known clamp + pos clip
    km = torch.rand(8, 8, 4) < 0.3
    kv = torch.randn(8, 8, 4)
    out = ddpm.ddim_step_scalar(z, eps, float(np_s["alpha_bar"][334]), float(np_s["alpha_bar"][276]), known_mask=km.cuda(),
                                known_values=kv.cuda(), pos_clip=True, pos_clip_min=-0.5, pos_clip_max=0.5)
    ref = odf.ddim_step(g["ddim_z"], g["ddim_eps"], np.full((8,), 334), np.full((8,), 276), np_s)
    ref = np.where(km.numpy(), kv.numpy(), ref)
    ref[..., :2] = np.clip(ref[..., :2], -0.5, 0.5)
    eq(out, ref)
    # q_sample
    o, _ = ddpm.q_sample(z, dev(g["ddim_v_t"]), sch, noise=dev(g["q_noise"]))
    eq(o, g["q_out"])


def test_clamps_golden(golden):
    from interpolated_diffusion_b200.utils import clamp, normalize
    g = golden("sampling")
    xh, xr, m, c = g["cl_x_hat"], g["cl_x_ref"], g["cl_mask"], g["cl_conf"]
    x = dev(xh)
    r = clamp.apply_clamp(x, dev(xr), dev(m), "pos")
    assert r.data_ptr() == x.data_ptr()          # in-place, like the reference
    eq(r, g["cl_hard_pos"])
    eq(clamp.apply_clamp(dev(xh), dev(xr), dev(m), "all"), g["cl_hard_all"])
    eq(clamp.apply_soft_clamp(dev(xh), dev(xr), dev(c), 0.7, "pos"), g["cl_soft_pos"])
    eq(clamp.apply_soft_clamp(dev(xh), dev(xr), dev(c), 0.7, "all"), g["cl_soft_all"])
    assert clamp.apply_clamp(x, dev(xr), None, "pos") is x
    assert clamp.apply_soft_clamp(x, dev(xr), dev(c), 0.0, "pos") is x
    ulp_close(normalize.logit_pos(dev(g["nz_p"])), g["nz_logit"], 4)
    ulp_close(normalize.sigmoid_pos(dev(xh * 3)), g["nz_sigmoid"], 4)
    # reference tests/test_interp_system.py:104-115
    x_ref = torch.zeros(1, 5, 4).cuda()
    x_hat = torch.ones(1, 5, 4).cuda()
    cm = torch.zeros(1, 5, dtype=torch.bool).cuda()
    cm[:, 0] = True
    cm[:, -1] = True
    out = clamp.apply_clamp(x_hat.clone(), x_ref, cm, "pos")
    assert (out[:, 0, :2] == 0).all() and (out[:, -1, :2] == 0).all()
    assert (out[:, 1:-1, :2] == 1).all() and (out[:, :, 2:] == 1).all()


def test_stage2_epilogue_fused_vs_oracle(golden):
    from interpolated_diffusion_b200.utils import clamp
    gen = torch.Generator().manual_seed(3)
    B, T, D = 33, 64, 4
    x_pred = torch.rand((B, T, D), generator=gen)
    delta = torch.randn((B, T, D), generator=gen) * 0.1
    conf = torch.rand((B, T), generator=gen)
    mask = torch.rand((B, T), generator=gen) < 0.2
    for pol in ("none", "endpoints", "all_anchors"):
        for dims in ("pos", "all"):
            for lam in (0.0, 1.0, 0.6):
                out = clamp.stage2_epilogue(x_pred.cuda(), delta.cuda(), x_pred.cuda(), conf.cuda(), lam, pol, mask.cuda(), dims)
                ref = (x_pred.numpy() + delta.numpy()).astype(np.float32)
                ref = osp.apply_soft_clamp(ref, x_pred.numpy(), conf.numpy(), lam, dims)
                if pol == "all_anchors":
                    ref = osp.apply_clamp(ref, x_pred.numpy(), mask.numpy(), dims)
                elif pol == "endpoints":
                    cm = np.zeros((B, T), bool); cm[:, 0] = cm[:, -1] = True
                    ref = osp.apply_clamp(ref, x_pred.numpy(), cm, dims)
                eq(out, ref)


def test_known_conf_anneal_golden(golden):
    from interpolated_diffusion_b200.sample import sample_generate as sg
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    g = golden("sampling")
    for D in (2, 4):
        km, kv = sg._build_known_mask_values(dev(g["kn_idx"]), {"start_goal": dev(g["kn_sg"])}, D, 16, True)
        assert km.dtype == torch.bool
        eq(km, g[f"kn_mask_{D}"])
        eq(kv, g[f"kn_vals_{D}"])
        _, kvl = sg._build_known_mask_values(dev(g["kn_idx"]), {"start_goal": dev(g["kn_sg"])}, D, 16, True, logit_space=True)
        ulp_close(kvl, osp.logit_pos(g[f"kn_vals_{D}"]), 4)
    # reference tests/test_interp_system.py:78-90
    idx = torch.tensor([[0, 3, 6, 7]]).cuda()
    cond = {"start_goal": torch.tensor([[1.0, 2.0, 3.0, 4.0]]).cuda()}
    km, kv = sg._build_known_mask_values(idx, cond, 4, 8)
    assert km[0, 0, :2].all() and not km[0, 0, 2:].any() and km[0, -1, :2].all() and not km[0, 1:-1].any()
    assert torch.equal(kv[0, 0, :2], cond["start_goal"][0, :2]) and torch.equal(kv[0, -1, :2], cond["start_goal"][0, 2:])
    with pytest.raises(ValueError, match="start_goal missing"):
        sg._build_known_mask_values(idx, {}, 4, 8)
    m = dev(g["cl_mask"])
    eq(sg._build_anchor_conf(m, m, True, 0.95, 0.5, 1.0, 0.0, True), g["cf_a"])
    eq(sg._build_anchor_conf(m, None, False, 0.95, 0.5, 1.0, 0.0, True), g["cf_b"])
    eq(sg._build_anchor_conf(m, dev(g["cf_student"]), True, 0.9, 0.4, 0.8, 0.1, False), g["cf_c"])
    eq(tr._build_anchor_conf(m, None, 0.95, 0.5, 1.0, 0.0, True), g["cf_b"])
    for mode in ("linear", "cosine", "none"):
        for s in (1, 2, 3):
            eq(sg._anneal_conf(dev(g["cf_b"]), s, 3, mode), g[f"an_{mode}_{s}"])
            conf, mask_in = sg.anchor_conf_mask_in(m, None, m, s, 3, mode, 0.95, 0.5, 1.0, 0.0, True, channels=3)
            eq(conf, g[f"an_{mode}_{s}"])
            eq(mask_in[..., 0], g["cl_mask"].astype(np.float32))
            eq(mask_in[..., 2], g[f"an_{mode}_{s}"])
    s_idx = dev(g["an_s_idx"])
    eq(tr._anneal_conf(dev(g["cf_b"]), s_idx, 3, "linear"), g["anv_linear"])
    conf, _ = sg.anchor_conf_mask_in(m, None, None, s_idx, 3, "linear", 0.95, 0.5, 1.0, 0.0, True)
    eq(conf, g["anv_linear"])
    conf, _ = sg.anchor_conf_mask_in(m, None, None, s_idx, 3, "cosine", 0.95, 0.5, 1.0, 0.0, True)
    ulp_close(conf, g["anv_cosine"], 8)
    lam = [sg._soft_clamp_lambda(s, 3, sc, 0.8) for sc in ("linear", "cosine", "const") for s in (0, 1, 2, 3)]
    np.testing.assert_array_equal(np.array(lam), g["lam"])
