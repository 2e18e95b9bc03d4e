"""GPU parity of the generation loop (Stage-1 DDIM -> sigmoid -> interp -> Stage-2 -> clamp) against the golden
outputs of the live reference (tiny random-init models) and the CPU oracle (BASELINE small model).

Protocol: the cosine schedule's first DDIM step multiplies (z - eps) by ~3243, so free-running
rollouts of non-contractive random-init nets amplify any rounding difference.  The gate is therefore teacher-forced
per step (identical z_t into both sides, z_{t-1} compared relative to its scale); free-running agreement is checked
where it is meaningful (fp32 check mode on the tiny models, where the reference itself is stable)."""
import numpy as np
import pytest
import torch

from oracle import generate as og
from oracle import diffusion_np as odf
from oracle import keyframes_np as okf
from oracle import sampling_np as osp

pytestmark = pytest.mark.gpu
TINY = dict(d_model=64, n_layers=2, n_heads=2, d_ff=128, d_cond=32, maze_channels=(8, 16))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def _models(g, precision):
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    kp = KeypointDenoiser(data_dim=2, **TINY).cuda()
    kp.load_state_dict(_sd(g, "kp/"))
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2, **TINY).cuda()
    il.load_state_dict(_sd(g, "il/"))
    il3 = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=3, **TINY).cuda()
    il3.load_state_dict(_sd(g, "il3/"))
    for m in (kp, il, il3):
        m.precision = precision
    return kp, il, il3


@pytest.mark.parametrize("precision,tol", [("fp32", 3e-4), ("bf16", 3e-2)])
def test_stage1_teacher_forced_golden(golden, precision, tol):
    from interpolated_diffusion_b200.diffusion import schedules
    from interpolated_diffusion_b200.diffusion.ddpm import _timesteps, ddim_step_scalar
    from interpolated_diffusion_b200.sample import sample_generate as sg
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    g = golden("generate_tiny")
    kp, _, _ = _models(g, precision)
    B, T, K, D = 4, 64, 8, 2
    cond = {"occ": torch.from_numpy(g["occ"]).cuda(), "start_goal": torch.from_numpy(g["sg"]).cuda()}
    idx, masks = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device="cuda")
    km, kv = sg._build_known_mask_values(idx, cond, D, T, True, logit_space=True)
    sch = schedules.make_alpha_bars(schedules.make_beta_schedule("cosine", 1000))
    ab = sch["alpha_bar"].numpy()
    times = _timesteps(1000, 20, "quadratic").tolist()
    inter = g["z_inter"]
    assert len(times) - 1 == inter.shape[0] - 1 == 19
    for i in range(19):
        z = torch.from_numpy(inter[i]).cuda()
        t = torch.full((B,), times[i], device="cuda", dtype=torch.long)
        eps = kp(z, t, idx, km, cond, T)
        z2 = ddim_step_scalar(z, eps, float(ab[times[i]]), float(ab[times[i + 1]]), known_mask=km, known_values=kv)
        ref = inter[i + 1]
        scale = max(1.0, float(np.abs(ref).max()))
        assert np.abs(z2.cpu().numpy() - ref).max() / scale < tol, (i, np.abs(z2.cpu().numpy() - ref).max(), scale)


def test_generate_free_running_fp32_golden(golden):
    """fp32 check mode reproduces the reference's own free-running pipeline on the tiny models: x_pred and x_hat for
    every clamp_policy x clamp_dims, and the adj chain."""
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, generate
    g = golden("generate_tiny")
    kp, il, il3 = _models(g, "fp32")
    cond = {"occ": torch.from_numpy(g["occ"]).cuda(), "start_goal": torch.from_numpy(g["sg"]).cuda()}
    z_T = torch.from_numpy(g["z_T"]).cuda()
    for pol in ("none", "endpoints", "all_anchors"):
        for dims in ("pos", "all"):
            out = generate(kp, il, cond, GenerationConfig(clamp_policy=pol, clamp_dims=dims), z_T=z_T, return_all=True)
            np.testing.assert_allclose(out["x_pred"].cpu().numpy(), g["x_pred"], atol=3e-3, rtol=0)
            np.testing.assert_allclose(out["x_hat"].cpu().numpy(), g[f"x_hat_x0_{pol}_{dims}"], atol=3e-3, rtol=0)
            if pol == "endpoints":          # hard clamp: endpoints equal x_pred exactly
                assert torch.equal(out["x_hat"][:, [0, -1], :2], out["x_pred"][:, [0, -1], :2])
    out = generate(kp, il3, cond, GenerationConfig(stage2_mode="adj"), z_T=z_T,
                   masks_levels=torch.from_numpy(g["adj_masks_levels"]).cuda(), return_all=True)
    np.testing.assert_allclose(out["x_hat"].cpu().numpy(), g["x_hat_adj"], atol=3e-3, rtol=0)


def test_generate_stage2_given_x_pred_vs_oracle():
    """BASELINE small model: everything after Stage 1 (interp, conf channel, Stage-2 jump, soft + hard clamp) from the
    SAME keypoints, bf16 path vs the CPU oracle: <= 2e-2 on x_hat (delta is the only rounded quantity)."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph, generate
    B, T, K, S, D = 32, 64, 8, 3, 2
    gen = torch.Generator().manual_seed(11)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D)
    il = InterpLevelDenoiser(data_dim=D, max_levels=S, mask_channels=2)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    sd_il = {k: v.clone() for k, v in il.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    ccond = {k: v.cuda() for k, v in cond.items()}
    z_T = torch.randn((B, K, D), generator=gen)
    cfg = GenerationConfig()
    out = generate(kp, il, ccond, cfg, z_T=z_T.cuda(), return_all=True)
    # oracle continuation from the CUDA path's own keypoints z (teacher-forced at the stage boundary)
    z = out["z"].cpu().numpy()
    idx, masks = okf.sample_fixed_k_indices_uniform_batch(B, T, K)
    assert np.array_equal(out["idx"].cpu().numpy(), idx)
    z_pred = osp.sigmoid_pos(z)
    x_pred = okf.interpolate_from_indices(idx, z_pred, T, True)
    np.testing.assert_allclose(out["x_pred"].cpu().numpy(), x_pred, atol=1e-6, rtol=0)
    conf = osp.build_anchor_conf(masks, masks, True, 0.95, 0.5, 1.0, 0.0, True)
    mask_in = np.stack([masks.astype(np.float32), osp.anneal_conf(conf, S, S, "linear")], axis=-1)
    from oracle import denoiser_torch as odn
    delta = odn.interp_level_denoiser(sd_il, 8, torch.from_numpy(x_pred), torch.full((B,), S), torch.from_numpy(mask_in), cond).numpy()
    x_hat = (x_pred + delta).astype(np.float32)
    x_hat = osp.apply_soft_clamp(x_hat, x_pred, conf, 1.0, "pos")
    cm = np.zeros_like(masks); cm[:, 0] = cm[:, -1] = True
    x_hat = osp.apply_clamp(x_hat, x_pred, cm, "pos")
    assert np.abs(out["x_hat"].cpu().numpy() - x_hat).max() < 2e-2
    # Stage 1, teacher-forced on the first and a late step against the oracle eps
    sched = odf.make_alpha_bars(odf.make_beta_schedule("cosine", 1000))
    km, kv = osp.build_known_mask_values(idx, cond["start_goal"].numpy(), D, T, True)
    for tval in (999, 44):
        zz = torch.randn((B, K, D), generator=gen)
        eps_ref = odn.keypoint_denoiser(sd_kp, 8, zz, torch.full((B,), tval), torch.from_numpy(idx), torch.from_numpy(km), cond, T)
        eps = kp(zz.cuda(), torch.full((B,), tval, device="cuda"), out["idx"], torch.from_numpy(km).cuda(), ccond, T)
        assert (eps.cpu() - eps_ref).abs().max().item() < 2e-2
    # the CUDA-graph replay equals the eager call bit for bit
    gg = GenerationGraph(kp, il, B, cfg).capture()
    xg = gg.run(ccond, z_T.cuda())
    torch.cuda.synchronize()
    assert torch.equal(xg, out["x_hat"])
    xg2 = gg.run(ccond, z_T.cuda())
    torch.cuda.synchronize()
    assert torch.equal(xg2, out["x_hat"])


def test_sample_keypoints_ddim_mirror(golden):
    """The drop-in `_sample_keypoints_ddim(model, schedule, idx, known_mask, known_values, cond, steps, T, ...)`."""
    from interpolated_diffusion_b200.diffusion import schedules
    from interpolated_diffusion_b200.sample import sample_generate as sg
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    g = golden("generate_tiny")
    kp, _, _ = _models(g, "fp32")
    B, T, K, D = 4, 64, 8, 2
    cond = {"occ": torch.from_numpy(g["occ"]).cuda(), "start_goal": torch.from_numpy(g["sg"]).cuda()}
    idx, _ = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device="cuda")
    km, kv = sg._build_known_mask_values(idx, cond, D, T, True, logit_space=True)
    sch = schedules.make_alpha_bars(schedules.make_beta_schedule("cosine", 1000))
    z, inter = sg._sample_keypoints_ddim(kp, sch, idx, km, kv, cond, 20, T, schedule_name="quadratic", return_intermediates=True,
                                         z_T=torch.from_numpy(g["z_T"]).cuda())
    assert len(inter) == 20
    ref = g["z_inter"]
    for i in (1, 5, 19):
        scale = max(1.0, float(np.abs(ref[i]).max()))
        assert np.abs(inter[i].cpu().numpy() - ref[i]).max() / scale < 5e-3, i


def test_generation_graph_equals_eager_and_is_batch_invariant():
    """Size-independent properties of the production path (BASELINE small models, bf16, whole-encoder kernel as CTA pairs,
    tcgen05 conv encoder, CUDA graph): (1) the graph replay equals the eager call bit for bit, (2) a trajectory's sample does
    not depend on what else is in the batch or where it sits in it (tile / CTA-pair position): generating two halves
    separately gives the same bits as generating the whole batch -- every op on the path is row-independent in B
    (SURVEY 8e), which is also what makes the multi-GPU sharding exact -- (3) the clamp policy holds on the result."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph, generate
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=2).cuda()
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2).cuda()
    cfg = GenerationConfig()
    B = 2 * 148 * 16 + 37 * 16 + 5                       # several tiles per CTA pair in both stages, ragged last tile
    gen = torch.Generator().manual_seed(9)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
    z_T = torch.randn((B, cfg.K_min, 2), generator=gen).cuda()
    whole = generate(kp, il, cond, cfg, z_T=z_T).clone()
    assert torch.isfinite(whole).all()
    graph = GenerationGraph(kp, il, B, cfg)
    replay = graph.run(cond, z_T).clone()
    assert torch.equal(replay, whole)
    assert torch.equal(graph.run(cond, z_T), whole)       # idempotent across replays
    cut = 1000                                            # not a multiple of the 16 trajectories of a Stage-1 tile
    parts = [generate(kp, il, {k: v[s] for k, v in cond.items()}, cfg, z_T=z_T[s]) for s in (slice(0, cut), slice(cut, B))]
    assert torch.equal(torch.cat(parts, dim=0), whole)
    # clamp_policy = endpoints, clamp_dims = pos: first / last positions are the start / goal of the conditioning
    sg = cond["start_goal"]
    # (through logit -> DDIM known-value clamp -> sigmoid: a few ulp of round trip)
    assert (whole[:, 0, :2] - sg[:, :2]).abs().max().item() < 2e-5 and (whole[:, -1, :2] - sg[:, 2:]).abs().max().item() < 2e-5


def test_generate_long_horizon_causal_cfg5():
    """BASELINE configs[4] shape: T = 256, K = 32, levels = 4 (K schedule [256,256,128,64,32]), causal Stage-2 denoiser, tiny
    random-init models.  fp32 check mode free-running against the CPU oracle's whole pipeline; bf16 mode teacher-forced from
    the same keypoints (Stage-2 jump + clamps within 2e-2)."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, generate
    from oracle import denoiser_torch as odn
    B, T, K, S, D = 3, 256, 32, 4, 2
    gen = torch.Generator().manual_seed(13)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(3)
    kp = KeypointDenoiser(data_dim=D, **TINY)
    il = InterpLevelCausalDenoiser(data_dim=D, max_levels=S, mask_channels=2, **TINY)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    sd_il = {k: v.clone() for k, v in il.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    ccond = {k: v.cuda() for k, v in cond.items()}
    z_T = torch.randn((B, K, D), generator=gen)
    cfg = GenerationConfig(T=T, K_min=K, levels=S, ddim_steps=6)
    ref = og.generate(sd_kp, sd_il, 2, cond, z_T.numpy(), T=T, K_min=K, levels=S, D=D, ddim_steps=6, causal=True)
    kp.precision = il.precision = "fp32"
    out = generate(kp, il, ccond, cfg, z_T=z_T.cuda(), return_all=True)
    assert np.array_equal(out["idx"].cpu().numpy(), ref["idx"])
    np.testing.assert_allclose(out["x_pred"].cpu().numpy(), ref["x_pred"], atol=3e-3, rtol=0)
    np.testing.assert_allclose(out["x_hat"].cpu().numpy(), ref["x_hat"], atol=3e-3, rtol=0)
    # bf16: Stage-2 continuation from the CUDA path's own x_pred
    kp.precision = il.precision = "bf16"
    out = generate(kp, il, ccond, cfg, z_T=z_T.cuda(), return_all=True)
    x_pred = out["x_pred"].cpu().numpy()
    idx, masks = okf.sample_fixed_k_indices_uniform_batch(B, T, K)
    conf = osp.build_anchor_conf(masks, masks, True, 0.95, 0.5, 1.0, 0.0, True)
    mask_in = np.stack([masks.astype(np.float32), osp.anneal_conf(conf, S, S, "linear")], axis=-1)
    delta = odn.interp_level_denoiser(sd_il, 2, torch.from_numpy(x_pred), torch.full((B,), S), torch.from_numpy(mask_in), cond, causal=True).numpy()
    x_hat = (x_pred + delta).astype(np.float32)
    x_hat = osp.apply_soft_clamp(x_hat, x_pred, conf, 1.0, "pos")
    cm = np.zeros_like(masks); cm[:, 0] = cm[:, -1] = True
    x_hat = osp.apply_clamp(x_hat, x_pred, cm, "pos")
    assert np.abs(out["x_hat"].cpu().numpy() - x_hat).max() < 2e-2


@pytest.mark.parametrize("T,chunk,K,policy", [(64, 16, 8, "endpoints"), (50, 16, 8, "all_anchors"), (40, 12, 6, "none")])
def test_causal_chunked_generation_vs_oracle(T, chunk, K, policy):
    """Batched chunk loop of sample_generate_causal.py:485-583 (Stage-1 on each chunk, causal Stage-2 over prefix + chunk):
    fp32 check mode against the CPU restatement with the same anchors and DDIM noise per chunk (tiny models)."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate_causal import chunk_plan, generate_causal_chunked
    B, D, S = 3, 2, 3
    gen = torch.Generator().manual_seed(19)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": 0.1 + 0.8 * torch.rand((B, 4), generator=gen)}
    torch.manual_seed(4)
    kp = KeypointDenoiser(data_dim=D, **TINY)
    il = InterpLevelCausalDenoiser(data_dim=D, max_levels=S, mask_channels=1, **TINY)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    sd_il = {k: v.clone() for k, v in il.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    kp.precision = il.precision = "fp32"
    plan = chunk_plan(T, chunk, K)
    assert plan == og.causal_chunk_plan(T, chunk, K)
    idx_chunks, z_chunks = [], []
    for (_, _, local_T, k) in plan:
        # random strictly-increasing anchors with both endpoints
        rows = [np.sort(np.concatenate([[0, local_T - 1], 1 + torch.randperm(local_T - 2, generator=gen)[:k - 2].numpy()])) for _ in range(B)]
        idx_chunks.append(np.stack(rows).astype(np.int64))
        z_chunks.append(torch.randn((B, k, D), generator=gen))
    ref = og.generate_causal_chunked(sd_kp, sd_il, 2, cond, T=T, chunk=chunk, K_min=K, levels=S, idx_chunks=idx_chunks,
                                     z_T_chunks=[z.numpy() for z in z_chunks], D=D, ddim_steps=5, clamp_policy=policy, logit_space=True)
    got = generate_causal_chunked(kp, il, {k: v.cuda() for k, v in cond.items()}, T=T, chunk=chunk, K_min=K, levels=S, data_dim=D,
                                  ddim_steps=5, clamp_policy=policy, logit_space=True, idx_chunks=[torch.from_numpy(i).cuda() for i in idx_chunks],
                                  z_T_chunks=[z.cuda() for z in z_chunks])
    assert got.shape == (B, T, D)
    np.testing.assert_allclose(got.cpu().numpy(), ref, atol=3e-3, rtol=0)
    # the free-running batched call (own anchors / noise) runs and respects the endpoint clamp of the first chunk
    out = generate_causal_chunked(kp, il, {k: v.cuda() for k, v in cond.items()}, T=T, chunk=chunk, K_min=K, levels=S, ddim_steps=3,
                                  logit_space=True, generator=torch.Generator(device="cuda").manual_seed(1))
    assert torch.isfinite(out).all() and torch.equal(out[:, 0, :2].cpu(), cond["start_goal"][:, :2])
