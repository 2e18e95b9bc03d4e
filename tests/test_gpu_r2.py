"""Round-2 GPU parity tests.

* cfg 1 literally (BASELINE.json configs[0]): ``ParticleMazeDataset(16, T=64, seed=123)`` conditioning, DEFAULT-size random-init
  models (seeded init == the reference's weights, checksummed), every DDIM step teacher-forced on the reference's own z_t
  (``tests/golden/cfg1.npz`` from the live reference): bf16 2e-2, fp32 check mode 1e-4 -- both relative to the step's
  magnitude, because a random-init rollout blows |z_t| up to ~8e4 (first DDIM step multiplies by ~3243).
* K1c single-launch corruption (``idb200_corrupt_adjacent``): bit-exact against the live-reference golden of
  ``build_interp_adjacent_batch`` with its recorded draws scattered into the per-row layout; Philox mode == parity mode on
  the exported noise, moments of the in-kernel normals; the fused trainer batch.
* a GenerationGraph re-captures when the weights change and survives a larger eager call on the same models.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.asarray(a)).cuda()


def _checksum(sd):
    rows = []
    for v in sd.values():
        a = v.detach().cpu().numpy().astype(np.float64).reshape(-1)
        rows.append([a.sum(), (a * a).sum(), a[0], a[-1]])
    return np.array(rows, dtype=np.float64)


def _cfg1_models(g):
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=2)
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2)
    assert np.array_equal(_checksum(kp.state_dict()), g["kp_checksum"]) and np.array_equal(_checksum(il.state_dict()), g["il_checksum"])
    return kp.cuda(), il.cuda()


@pytest.mark.parametrize("precision,tol", [("bf16", 2e-2), ("fp32", 1e-4)])
def test_cfg1_default_models_teacher_forced(golden, precision, tol):
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    from interpolated_diffusion_b200.diffusion.ddpm import _timesteps
    from interpolated_diffusion_b200.sample import sample_generate as sg
    g = golden("cfg1")
    kp, il = _cfg1_models(g)
    kp.precision = il.precision = precision
    B, T, K, D, S = 16, 64, 8, 2, 3
    cond = {"occ": dev(g["occ"]), "start_goal": dev(g["start_goal"])}
    idx, masks = kf.sample_fixed_k_indices_uniform_batch(B, T, K, device="cuda")
    assert torch.equal(idx.cpu(), torch.from_numpy(g["idx"])) and torch.equal(masks.cpu(), torch.from_numpy(g["masks"]))
    km, kv = sg._build_known_mask_values(idx, cond, D, T, True, logit_space=True)
    assert torch.equal(km.cpu(), torch.from_numpy(g["known_mask"]))
    assert np.abs(kv.cpu().numpy() - g["known_values"]).max() <= 2e-6
    times = _timesteps(1000, 20, "quadratic").tolist()
    assert times == g["times"].tolist()
    worst = 0.0
    for i in range(19):
        t = torch.full((B,), times[i], device="cuda", dtype=torch.long)
        eps = kp(dev(g["z_inter"][i]), t, idx, km, cond, T)
        scale = max(1.0, float(np.abs(g["eps"][i]).max()))
        err = float(np.abs(eps.cpu().numpy() - g["eps"][i]).max()) / scale
        worst = max(worst, err)
        assert err < tol, (i, err, scale)
    # interpolation of the reference's keypoints, Stage-2 one-step on the reference's x_pred, clamp
    from interpolated_diffusion_b200.utils.normalize import sigmoid_pos
    x_pred = kf.interpolate_from_indices(idx, sigmoid_pos(dev(g["z"])), T, recompute_velocity=True)
    assert np.abs(x_pred.cpu().numpy() - g["x_pred"]).max() <= 1.2e-7
    conf, mask_in = sg.anchor_conf_mask_in(masks, masks, None, S, S, "linear", 0.95, 0.5, 1.0, 0.0, True, channels=2)
    assert np.array_equal(mask_in.cpu().numpy(), g["mask_in"])
    delta = il(dev(g["x_pred"]), torch.full((B,), S, device="cuda", dtype=torch.long), mask_in, cond)
    assert np.abs(delta.cpu().numpy() - g["delta"]).max() < tol * max(1.0, float(np.abs(g["delta"]).max()))
    from interpolated_diffusion_b200.utils.clamp import stage2_epilogue
    x_hat = stage2_epilogue(dev(g["x_pred"]), dev(g["delta"]), dev(g["x_pred"]), dev(g["conf_pred"]), 1.0, "endpoints", masks, "pos")
    assert np.abs(x_hat.cpu().numpy() - g["x_hat"]).max() <= 1.2e-7


# ---------------------------------------------------------------------------------------------------------------------------
def _scatter_reference_draws(g, s_idx, K_list, T, sig_on=True):
    """The live reference's per-level draws (order: for s = 1..S over rows s_idx == s: x_s anchor, x_s path, x_prev anchor,
    x_prev path; train_interp_levels.py:328-372) scattered into the per-row layout of idb200_corrupt_adjacent."""
    B = s_idx.shape[0]
    Kmax = max(K_list)
    an = np.zeros((B, 2, Kmax, 2), np.float32)
    pn = np.zeros((B, 2, T, 2), np.float32)
    n = int(g["ad_dist_n"][0])
    draws = [g[f"ad_dist_draw{i}"] for i in range(n)]
    pos = 0
    for s in range(1, len(K_list)):
        rows = np.nonzero(s_idx == s)[0]
        if rows.size == 0:
            continue
        for slot, lvl in ((0, s), (1, s - 1)):
            a = draws[pos]; pos += 1
            assert a.shape == (rows.size, K_list[lvl], 2), (a.shape, rows.size, K_list[lvl])
            an[rows, slot, :K_list[lvl]] = a
            p = draws[pos]; pos += 1
            assert p.shape == (rows.size, T, 2)
            pn[rows, slot] = p
    assert pos == n
    return an, pn


def test_corrupt_adjacent_single_launch_equals_live_reference(golden):
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    g = golden("sampling")
    x0, masks, s_idx = dev(g["ad_x0"]), dev(g["ad_masks"]), dev(g["ad_s_idx"])
    B, T, D = x0.shape
    K_list = [int(g[f"ad_idx{s}"].shape[1]) for s in range(4)]
    # no corruption: pure Interp(x0 | M_s), Interp(x0 | M_{s-1})
    xs, xp, ms, mp = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, 8)
    assert np.array_equal(xs.cpu().numpy(), g["ad_none_xs"]) and np.array_equal(xp.cpu().numpy(), g["ad_none_xp"])
    assert np.array_equal(ms.cpu().numpy(), g["ad_none_ms"]) and np.array_equal(mp.cpu().numpy(), g["ad_none_mp"])
    xs1, none, ms1, _ = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, 8, adjacent=False)
    assert none is None and np.array_equal(xs1.cpu().numpy(), g["lv_none_xs"]) and np.array_equal(ms1.cpu().numpy(), g["lv_none_ms"])
    # dist corruption with the reference's recorded draws
    an, pn = _scatter_reference_draws(g, g["ad_s_idx"], K_list, T)
    kw = dict(corrupt_mode="dist", corrupt_sigma_max=0.08, corrupt_sigma_min=0.012, corrupt_sigma_pow=0.75, corrupt_anchor_frac=0.25)
    xs, xp, ms, mp = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, 8, anchor_noise=dev(an), path_noise=dev(pn), **kw)
    assert np.array_equal(xs.cpu().numpy(), g["ad_dist_xs"]), np.abs(xs.cpu().numpy() - g["ad_dist_xs"]).max()
    assert np.array_equal(xp.cpu().numpy(), g["ad_dist_xp"])


@pytest.mark.parametrize("T,K_min,S,D", [(64, 8, 3, 2), (64, 8, 3, 4), (256, 32, 4, 2), (33, 5, 2, 2)])
def test_corrupt_adjacent_philox_mode(T, K_min, S, D):
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    B = 4096 if T <= 64 else 512
    gen = torch.Generator(device="cuda").manual_seed(5)
    x0 = torch.rand((B, T, D), device="cuda", generator=gen)
    masks, idx_levels = kf.build_nested_masks_batch(B, T, K_min, S, generator=gen, device="cuda")
    K_list = kf._compute_k_schedule(T, K_min, S)
    s_idx = tr._sample_level_indices(B, S, gen, torch.device("cuda"), "high", 0.5, sync_free=True)
    kw = dict(corrupt_mode="dist", corrupt_sigma_max=0.08, corrupt_sigma_min=0.012, corrupt_sigma_pow=0.75, corrupt_anchor_frac=0.25,
              recompute_velocity=True)
    xs, xp, ms, mp, an, pn = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, K_min, seed=1234, offset=7, export_noise=True, **kw)
    # (1) parity mode on the exported noise reproduces the Philox run bit for bit
    xs2, xp2, ms2, mp2 = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, K_min, anchor_noise=an, path_noise=pn, **kw)
    assert torch.equal(xs, xs2) and torch.equal(xp, xp2) and torch.equal(ms, ms2) and torch.equal(mp, mp2)
    # (2) ... and parity mode is the per-level reference-order path given the same noise (generator replay)
    rows = torch.arange(B, device="cuda")
    assert torch.equal(ms, masks[rows, s_idx]) and torch.equal(mp, masks[rows, s_idx - 1])
    from oracle import sampling_np as osp
    x0n, idxn = x0.cpu().numpy(), [i.cpu().numpy() for i in idx_levels]
    sn = s_idx.cpu().numpy()
    ann, pnn = an.cpu().numpy(), pn.cpu().numpy()
    for b in range(0, B, max(1, B // 64)):
        for slot, out in ((0, xs), (1, xp)):
            lvl = int(sn[b]) - slot
            K = K_list[lvl]
            sigma = osp.compute_sigma_for_level(K, K_min, 0.08, 0.012, 0.75)
            tape = osp.NoiseTape([ann[b:b + 1, slot, :K], pnn[b:b + 1, slot]])
            ref = osp.corrupt_from_anchors(x0n[b:b + 1], idxn[lvl][b:b + 1], T, tape, sigma, sigma * 0.25, 0, 0.0, "dist", True, True)
            assert np.array_equal(out[b:b + 1].cpu().numpy(), ref), (b, slot)
    # (3) the in-kernel normals: N(0,1) moments, independent across slots / seeds, deterministic
    used = pn[:, :, :, :].reshape(-1).double()
    assert abs(float(used.mean())) < 5e-3 and abs(float(used.var()) - 1.0) < 1e-2
    assert abs(float((used ** 3).mean())) < 3e-2 and abs(float((used ** 4).mean()) - 3.0) < 1e-1
    c = float((pn[:, 0].reshape(-1) * pn[:, 1].reshape(-1)).mean())
    assert abs(c) < 1e-2
    again = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, K_min, seed=1234, offset=7, **kw)
    other = tr.corrupt_adjacent_fused(x0, masks, s_idx, K_list, K_min, seed=1234, offset=8, **kw)
    assert torch.equal(again[0], xs) and not torch.equal(other[0], xs)


def test_stage2_trainer_fused_batch_mode():
    """batch_mode='fused': same batch structure as the reference-draw mode (masks, level marginals, targets consistent with the
    masks), no host sync inside build_batch, and a training step runs on it."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    SMALL = dict(d_model=128, n_layers=2, n_heads=4, d_ff=256, maze_channels=(32, 64))   # conv widths the training GEMMs take
    torch.manual_seed(0)
    model = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=3, **SMALL).cuda()
    tr = Stage2Trainer(model, batch_mode="fused")
    B, T = 2048, 64
    g = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.rand((B, T, 2), device="cuda", generator=g)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device="cuda", generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), device="cuda", generator=g)}
    x_s, s_idx, mask_in, target, wm = tr.build_batch(x0, g, cond)
    assert x_s.shape == (B, T, 2) and mask_in.shape == (B, T, 3) and target.shape == (B, T, 2) and wm.shape == (B, T)
    p = torch.bincount(s_idx, minlength=4).float() / B
    assert abs(float(p[3]) - (0.5 + 0.5 / 3)) < 0.05 and float(p[0]) == 0.0
    m_s, m_prev = mask_in[..., 0] > 0.5, mask_in[..., 1] > 0.5
    K_list = [64, 32, 16, 8]
    assert torch.equal(m_s.sum(1), torch.tensor(K_list, device="cuda")[s_idx]) and torch.equal(m_prev.sum(1), torch.tensor(K_list, device="cuda")[s_idx - 1])
    assert bool((m_prev | ~m_s).all())                                 # nested: M_s subset of M_{s-1}
    # corruption noise is bounded by a few sigma: x_s stays close to the clean interpolation at the anchors of M_s
    assert float((x_s - x0)[m_s].abs().max()) < 0.5
    # endpoints carry conf_endpoints = 1 -> weight w = 1 + (0.1 - 1) * conf uses wm = conf_prev: endpoints 1.0
    assert torch.all(wm[:, 0] == 1.0) and torch.all(wm[:, -1] == 1.0)
    l0 = float(tr.step(x0, cond, g))
    l1 = float(tr.step(x0, cond, g))
    assert np.isfinite(l0) and np.isfinite(l1)


# ---------------------------------------------------------------------------------------------------------------------------
def test_generation_graph_recaptures_on_weight_change_and_survives_growth():
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph, generate
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=2).cuda()
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2).cuda()
    cfg = GenerationConfig()
    B = 64
    g = torch.Generator().manual_seed(2)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=g).cuda()}
    z_T = torch.randn((B, 8, 2), generator=g).cuda()
    graph = GenerationGraph(kp, il, B, cfg)
    a = graph.run(cond, z_T).clone()
    assert torch.equal(a, generate(kp, il, cond, cfg, z_T=z_T))
    # a larger eager call on the same models grows every workspace: the captured graph must keep replaying correctly
    B2 = 4 * B
    cond2 = {"occ": (torch.rand((B2, 1, 21, 21), generator=g) < 0.2).float().cuda(), "start_goal": torch.rand((B2, 4), generator=g).cuda()}
    big = generate(kp, il, cond2, cfg, z_T=torch.randn((B2, 8, 2), generator=g).cuda())
    junk = [torch.randn((1 << 20,), device="cuda") for _ in range(8)]      # re-use whatever memory a wrongly freed buffer left
    assert torch.isfinite(big).all()
    assert torch.equal(graph.run(cond, z_T), a)
    del junk
    # weights change in place (what an optimizer step / EMA copy_to / load_state_dict does): run() must not replay the stale capture
    with torch.no_grad():
        for p in il.parameters():
            p.mul_(1.01)
        for p in kp.parameters():
            p.add_(1e-3)
    b = graph.run(cond, z_T).clone()
    assert not torch.equal(a, b)
    assert torch.equal(b, generate(kp, il, cond, cfg, z_T=z_T))


# ---------------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,K", [(32768, 1536, 384), (1000, 1024, 256), (257, 64, 64), (4096, 192, 128)])
def test_gemm_silu_epilogues_equal_the_separate_passes(M, N, K):
    """idb200_gemm_bf16_aux (epilogue 4: u and SiLU(u) from one launch; 5: (A W^T) * SiLU'(u)) == token GEMM + silu_bf16 kernels,
    bit for bit, and against torch on the bf16-rounded operands."""
    from interpolated_diffusion_b200.models import _engine as E
    from interpolated_diffusion_b200.train import backward as bw
    g = torch.Generator(device="cuda").manual_seed(4)
    A = (torch.randn((M, K), device="cuda", generator=g)).bfloat16()
    W = (torch.randn((N, K), device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn((N,), device="cuda", generator=g)
    u = torch.empty((M, N), device="cuda", dtype=torch.bfloat16)
    f = torch.empty_like(u)
    bw.gemm_bf16_aux(A, W, bias, u, f, bw.EPI_BF16_SILU_DUAL, fused=True)
    u_ref = torch.empty_like(u)
    E.gemm_bf16(A, W, bias, u_ref, E.EPI_BF16)
    f_ref = bw.silu_bf16(u_ref, torch.empty_like(u))
    assert torch.equal(u, u_ref) and torch.equal(f, f_ref)
    t = A.float() @ W.float().t() + bias
    assert (u.float() - t).abs().max().item() < 2e-2 * max(1.0, t.abs().max().item())
    assert (f.float() - torch.nn.functional.silu(u.float())).abs().max().item() < 2e-2
    # backward form: dU = (dY W2) * SiLU'(u)
    du = torch.empty_like(u)
    bw.gemm_bf16_aux(A, W, None, du, u, bw.EPI_BF16_DSILU, fused=True)
    dF = torch.empty_like(u)
    E.gemm_bf16(A, W, None, dF, E.EPI_BF16)
    du_ref = bw.silu_bf16(u, dF, g=dF)
    assert torch.equal(du, du_ref)
    # the form the training step uses: the same dU + per-warp column sums (column sums of colpart = the ff.0 bias gradient)
    from interpolated_diffusion_b200 import _lib as L
    du2 = torch.full_like(du, float("nan"))
    part = torch.full((8 * ((M + 255) // 256), N), float("nan"), device="cuda")          # 4 rows per 128-row block, blocks rounded up to CTA pairs
    L.call("idb200_gemm_bf16_dsilu_sums", A.data_ptr(), W.data_ptr(), du2.data_ptr(), u.data_ptr(), part.data_ptr(), M, N, K, L.stream(A.device))
    assert torch.equal(du2, du)
    want_cs = du.double().sum(0)
    assert (part.double().sum(0) - want_cs).abs().max().item() <= 1e-4 * max(1.0, want_cs.abs().max().item())
    x = u.float()
    s = torch.sigmoid(x)
    want = (A.float() @ W.float().t()).bfloat16().float() * (s * (1 + x * (1 - s)))
    assert (du.float() - want).abs().max().item() < 2e-2 * max(1.0, want.abs().max().item())


def test_split_graph_step_equals_single_graph_step():
    """The data-parallel overlap machinery on one GPU: the backward captured as TWO graphs cut where the transformer gradients
    are final (and the arena laid out so that those gradients are one leading slice) gives the same losses and parameters as the
    single-graph step, bit for bit."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    SMALL = dict(d_model=128, n_layers=2, n_heads=4, d_ff=256, maze_channels=(32, 64))
    B, T = 256, 64
    g0 = torch.Generator(device="cuda").manual_seed(1)
    x0 = torch.rand((B, T, 2), device="cuda", generator=g0)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device="cuda", generator=g0) < 0.2).float(), "start_goal": torch.rand((B, 4), device="cuda", generator=g0)}
    out = []
    for overlap in (True, False):
        torch.manual_seed(0)
        model = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=3, **SMALL).cuda()
        tr = Stage2Trainer(model, cuda_graph=True, overlap_allreduce=overlap)
        names = [n for n, _ in model.named_parameters()]
        early = set(tr.bp.early_param_names())
        assert all(n.startswith("out.") or n.startswith("transformer.") for n in early) and not any(".film" in n for n in early)
        # arena: early parameters occupy [0, n_early), the others [n_early, n)
        for n_, off, p in zip(names, tr.opt.offsets, tr.opt.params):
            assert (off < tr.opt.n_early) == (n_ in early), n_
        assert 0 < tr.opt.n_early < tr.opt.n
        gen = torch.Generator(device="cuda").manual_seed(23)
        losses = [float(tr.step(x0, cond, gen)) for _ in range(3)]
        out.append((losses, torch.cat([p.detach().reshape(-1) for p in model.parameters()]).clone()))
    assert out[0][0] == out[1][0], (out[0][0], out[1][0])
    assert torch.equal(out[0][1], out[1][1])


@pytest.mark.parametrize("M,Lq,d", [(4096 + 24, 8, 384), (8192, 64, 256), (5000, 5, 384), (4104, 24, 256), (4800, 8, 320), (6144, 256, 512)])
def test_ln_film_bulk_kernel_vs_torch(M, Lq, d):
    """The TMA-staged LayerNorm+FiLM kernel (M >= 4096 selects it): ragged last chunk, trajectory lengths that do not divide the
    16-row chunk (incremental phase), L < 8 (FiLM rows from global memory), widths with and without the exact-width instantiation."""
    from interpolated_diffusion_b200.models import _engine as E
    M = (M // Lq) * Lq
    g = torch.Generator(device="cuda").manual_seed(M + d)
    h = torch.randn((M, d), generator=g, device="cuda") * 3 + 1
    w, b = torch.randn((d,), generator=g, device="cuda"), torch.randn((d,), generator=g, device="cuda")
    gb = torch.randn((M // Lq, 2 * d), generator=g, device="cuda")
    ref = torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)
    ref = (ref.view(-1, Lq, d) * (1 + gb[:, None, :d]) + gb[:, None, d:]).view(M, d)
    assert float((E.ln_film(h, w, b, gb, torch.empty_like(h), Lq) - ref).abs().max()) < 5e-5
    ob = E.ln_film(h, w, b, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq)
    assert ((ob.float() - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-5).all()
    plain = torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)
    assert float((E.ln_film(h, w, b, None, torch.empty_like(h), Lq) - plain).abs().max()) < 5e-5
    # training-forward form (idb200_ln_film_save): same output bit for bit + an exact copy of the rows it read
    hc = torch.full_like(h, float("nan"))
    ob2 = E.ln_film(h, w, b, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq, h_copy=hc)
    assert torch.equal(ob2, ob) and torch.equal(hc, h)


@pytest.mark.parametrize("M,Lq,d", [(512, 8, 256), (1000, 5, 384)])
def test_ln_film_save_small_batches(M, Lq, d):
    """idb200_ln_film_save below the TMA-staged kernel's threshold (the register kernel writes the copy)."""
    from interpolated_diffusion_b200.models import _engine as E
    g = torch.Generator(device="cuda").manual_seed(3)
    h = torch.randn((M, d), generator=g, device="cuda")
    w, b = torch.randn((d,), generator=g, device="cuda"), torch.randn((d,), generator=g, device="cuda")
    gb = torch.randn((M // Lq, 2 * d), generator=g, device="cuda")
    want = E.ln_film(h, w, b, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq)
    hc = torch.full_like(h, float("nan"))
    got = E.ln_film(h, w, b, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq, h_copy=hc)
    assert torch.equal(got, want) and torch.equal(hc, h)


@pytest.mark.parametrize("segs,M,N,dt", [(24, 512, 1152, torch.float32), (3, 777, 100, torch.float32), (5, 4096, 768, torch.bfloat16), (1, 64, 33, torch.bfloat16)])
def test_colsum_segments_vs_torch(segs, M, N, dt):
    """idb200_colsum_segments: column sums of `segs` stacked [M, N] matrices in two launches (vector and scalar paths, ragged slices)."""
    from interpolated_diffusion_b200.train import backward as bw
    g = torch.Generator(device="cuda").manual_seed(segs * 131 + N)
    src = torch.randn((segs, M, N), generator=g, device="cuda").to(dt)
    out = torch.full((segs, N), float("nan"), device="cuda")
    bw._Scratch().colsum_segments(src, out)
    want = src.double().sum(1)
    assert (out.double() - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item()) * (8 if dt == torch.float32 else 1)
    one = torch.empty((N,), device="cuda")
    sc = bw._Scratch()
    for k in (0, segs - 1):                                     # the segmented form equals the plain call on each slice, bit for bit
        assert torch.equal(sc.colsum(src[k], one), out[k])


@pytest.mark.parametrize("B,Hh,Ww,C,act", [(3, 21, 21, 128, True), (5, 9, 12, 32, False), (2, 21, 21, 8, True), (4, 7, 5, 64, True)])
def test_im2col_scatter_form_vs_unfold(B, Hh, Ww, C, act):
    """Patch matrix of the training conv stack (scatter form: one thread per input channel group): every entry, including the zero
    taps of border pixels and the K padding, against torch's unfold of the activated input."""
    from interpolated_diffusion_b200.models import _engine as E
    g = torch.Generator(device="cuda").manual_seed(B * 100 + C)
    u = torch.randn((B, Hh * Ww, C), generator=g, device="cuda").to(torch.bfloat16)
    Kpad = ((9 * C + 63) // 64) * 64
    col = torch.full((B * Hh * Ww, Kpad), 7.0, device="cuda", dtype=torch.bfloat16)
    E.im2col3x3(u, B, Hh, Ww, C, act, col)
    a = u.float()
    if act:
        a = torch.nn.functional.silu(a).to(torch.bfloat16).float()
    x = a.view(B, Hh, Ww, C).permute(0, 3, 1, 2)                                  # NCHW
    pat = torch.nn.functional.unfold(x, 3, padding=1)                              # [B, C * 9, P], row = c * 9 + tap
    pat = pat.view(B, C, 9, Hh * Ww).permute(0, 3, 2, 1).reshape(B * Hh * Ww, 9 * C)   # [(b, p), tap * C + c]
    got = col.float()
    assert float((got[:, :9 * C] - pat).abs().max()) <= 2.0 ** -7 * float(pat.abs().max())   # silu rounding differences only
    assert torch.equal(got[:, :9 * C] == 0, pat == 0) or not act
    assert float(got[:, 9 * C:].abs().max()) == 0.0 if Kpad > 9 * C else True


def _pack_groups(w, b, d):
    order = torch.cat([torch.arange(64) + part * d + g * 64 for g in range(d // 64) for part in range(3)]).to(w.device)
    return w[order].contiguous(), b[order].contiguous()


@pytest.mark.parametrize("B,Lq,d,causal", [(40, 8, 384, False), (33, 64, 384, False), (16, 64, 256, True), (24, 16, 384, True), (7, 128, 384, False),
                                           (1, 8, 384, False), (129, 4, 256, False), (50, 32, 384, True)])
def test_qkv_attention_fused_vs_torch(B, Lq, d, causal):
    """idb200_qkv_attention (in_proj + attention in one kernel, qkv never in HBM) against fp32 torch on the same bf16 inputs: odd tile
    counts (the pair's dead tile), a ragged last tile, L = 4 .. 128, causal, both widths; and in place (o aliases a)."""
    from interpolated_diffusion_b200.models import _engine as E
    H, M = d // 32, B * Lq
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Lq)
    a = torch.randn((M, d), generator=g, device="cuda").to(torch.bfloat16)
    w = (torch.randn((3 * d, d), generator=g, device="cuda") / d ** 0.5).to(torch.bfloat16)
    b = torch.randn((3 * d,), generator=g, device="cuda") * 0.1
    qkv = (a.float() @ w.float().t() + b).to(torch.bfloat16).float()            # the per-op path rounds the projections to bf16 too
    q, k, v = (t.view(B, Lq, H, 32).transpose(1, 2) for t in qkv.split(d, dim=-1))
    ref = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(M, d)
    wg, bg = _pack_groups(w, b, d)
    out = E.qkv_attention(a, wg, bg, torch.empty_like(a), Lq, H, causal)
    tol = 2e-2 * max(1.0, float(ref.abs().max()))
    assert float((out.float() - ref).abs().max()) <= tol
    inplace = a.clone()
    E.qkv_attention(inplace, wg, bg, inplace, Lq, H, causal)
    assert torch.equal(inplace, out)


@pytest.mark.parametrize("M,d,ff", [(640, 384, 1536), (128, 384, 128), (1000, 384, 384), (4096 + 64, 256, 1024), (300, 384, 2048), (129, 256, 640)])
def test_mlp_pair_vs_torch(M, d, ff):
    """idb200_mlp_pair (FF1 + SiLU + FF2 + residual in one pair-mode kernel, hidden activation never in HBM) against fp32 torch on the
    same bf16 operands (hidden activation rounded to bf16 like the kernel's MMA operand): odd tile counts (the pair's dead tile),
    ragged last tiles, odd and even chunk counts, both widths."""
    from interpolated_diffusion_b200.models import _engine as E
    g = torch.Generator(device="cuda").manual_seed(M + ff)
    a = torch.randn((M, d), generator=g, device="cuda").to(torch.bfloat16)
    w1 = (torch.randn((ff, d), generator=g, device="cuda") / d ** 0.5).to(torch.bfloat16)
    b1 = torch.randn((ff,), generator=g, device="cuda") * 0.1
    w2 = (torch.randn((d, ff), generator=g, device="cuda") / ff ** 0.5).to(torch.bfloat16)
    b2 = torch.randn((d,), generator=g, device="cuda") * 0.1
    h0 = torch.randn((M, d), generator=g, device="cuda")
    hid = torch.nn.functional.silu(a.float() @ w1.float().t() + b1).to(torch.bfloat16).float()
    ref = h0 + hid @ w2.float().t() + b2
    h = h0.clone()
    E.mlp_pair(a, w1, b1, w2[E.mlp_pair_w2_order(d, "cuda")].contiguous(), b2, h)
    assert float((h - ref).abs().max()) <= 2e-2 * max(1.0, float((ref - h0).abs().max()))


@pytest.mark.parametrize("B,Lq,d,causal,film", [(40, 8, 384, False, True), (33, 64, 384, False, True), (16, 64, 256, True, False), (129, 4, 256, False, True),
                                                (1, 8, 384, False, True), (50, 32, 384, True, True)])
def test_ln_qkv_attention_equals_ln_film_then_qkv_attention(B, Lq, d, causal, film):
    """idb200_ln_qkv_attention (LayerNorm + FiLM prologue in the kernel) against idb200_ln_film followed by idb200_qkv_attention
    (the prologue sums a row in a different order: the bf16 operand may differ by one rounding)."""
    from interpolated_diffusion_b200.models import _engine as E
    H, M = d // 32, B * Lq
    g = torch.Generator(device="cuda").manual_seed(B * 77 + Lq)
    h = torch.randn((M, d), generator=g, device="cuda") * 2 + 0.5
    lw, lb = torch.randn((d,), generator=g, device="cuda"), torch.randn((d,), generator=g, device="cuda")
    big = torch.randn((B, 6, 2 * d), generator=g, device="cuda") * 0.3
    gb = big[:, 3] if film else None                                      # a strided view, as the encoder passes it
    w = (torch.randn((3 * d, d), generator=g, device="cuda") / d ** 0.5).to(torch.bfloat16)
    b = torch.randn((3 * d,), generator=g, device="cuda") * 0.1
    wg, bg = _pack_groups(w, b, d)
    a = E.ln_film(h, lw, lb, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq)
    ref = E.qkv_attention(a, wg, bg, torch.empty_like(a), Lq, H, causal)
    out = E.ln_qkv_attention(h, lw, lb, gb, wg, bg, torch.empty_like(a), Lq, H, causal)
    assert float((out.float() - ref.float()).abs().max()) <= 2e-2 * max(1.0, float(ref.float().abs().max()))


@pytest.mark.parametrize("B,Lq,d,ff,film", [(40, 8, 384, 1536, True), (33, 64, 384, 1536, True), (65, 64, 256, 1024, False), (1, 8, 384, 128, True), (50, 32, 384, 384, True)])
def test_ln_mlp_pair_vs_ln_film_then_mlp_pair(B, Lq, d, ff, film):
    """idb200_ln_mlp_pair (LayerNorm + FiLM prologue in the kernel, h both normalised and updated) against idb200_ln_film followed by
    idb200_mlp_pair (the prologue sums a row in a different order: the bf16 operand may differ by one rounding)."""
    from interpolated_diffusion_b200.models import _engine as E
    M = B * Lq
    g = torch.Generator(device="cuda").manual_seed(B * 31 + ff)
    h0 = torch.randn((M, d), generator=g, device="cuda") * 2 + 0.5
    lw, lb = torch.randn((d,), generator=g, device="cuda"), torch.randn((d,), generator=g, device="cuda")
    big = torch.randn((B, 4, 2 * d), generator=g, device="cuda") * 0.3
    gb = big[:, 1] if film else None
    w1 = (torch.randn((ff, d), generator=g, device="cuda") / d ** 0.5).to(torch.bfloat16)
    b1 = torch.randn((ff,), generator=g, device="cuda") * 0.1
    w2 = (torch.randn((d, ff), generator=g, device="cuda") / ff ** 0.5).to(torch.bfloat16)
    b2 = torch.randn((d,), generator=g, device="cuda") * 0.1
    w2p = w2[E.mlp_pair_w2_order(d, "cuda")].contiguous()
    a = E.ln_film(h0, lw, lb, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq)
    ref = E.mlp_pair(a, w1, b1, w2p, b2, h0.clone())
    out = E.ln_mlp_pair(h0.clone(), lw, lb, gb, w1, b1, w2p, b2, Lq)
    assert float((out - ref).abs().max()) <= 2e-2 * max(1.0, float((ref - h0).abs().max()))


def test_multi_copy_and_cast_weights_batches():
    """idb200_multi_copy_f32 (more segments than one launch takes, ragged sizes, a non-contiguous pair through the fallback) and
    idb200_cast_weights_bf16 (ragged matrices: bf16 copy and bf16 transpose == torch's casts, bit for bit)."""
    import ctypes
    from interpolated_diffusion_b200 import _lib as L
    from interpolated_diffusion_b200.train import backward as bw
    g = torch.Generator(device="cuda").manual_seed(9)
    sizes = [1, 7, 384, 1000, 147456] + [33 + 5 * i for i in range(120)]
    srcs = [torch.randn((n,), generator=g, device="cuda") for n in sizes]
    dsts = [torch.full((n,), float("nan"), device="cuda") for n in sizes]
    strided_src = torch.randn((16, 8), generator=g, device="cuda")[:, ::2]
    strided_dst = torch.full((16, 4), float("nan"), device="cuda")
    bw.multi_copy(list(zip(dsts, srcs)) + [(strided_dst, strided_src)])
    assert all(torch.equal(d, s) for d, s in zip(dsts, srcs)) and torch.equal(strided_dst, strided_src)
    shapes = [(1152, 384), (384, 1536), (70, 130), (1, 64), (65, 1)]
    mats = [torch.randn(s, generator=g, device="cuda") for s in shapes]
    d0 = [torch.empty(s, device="cuda", dtype=torch.bfloat16) for s in shapes]
    d1 = [torch.empty((s[1], s[0]), device="cuda", dtype=torch.bfloat16) for s in shapes]
    n = len(mats)
    L.call("idb200_cast_weights_bf16", (ctypes.c_void_p * n)(*[m.data_ptr() for m in mats]), (ctypes.c_void_p * n)(*[t.data_ptr() for t in d0]),
           (ctypes.c_void_p * n)(*[t.data_ptr() for t in d1]), (ctypes.c_int * n)(*[s[0] for s in shapes]), (ctypes.c_int * n)(*[s[1] for s in shapes]),
           n, L.stream(mats[0].device))
    for m, a, b in zip(mats, d0, d1):
        assert torch.equal(a, m.bfloat16()) and torch.equal(b, m.t().contiguous().bfloat16())


def test_grouped_weight_gradients_equal_the_per_layer_form(monkeypatch):
    """EncoderBackprop.backward_layers: one split-K GEMM per weight kind over the stacked layers (default) against the per-layer
    form (IDB200_TRAIN_DW_GROUPED=0): same gradients up to the fp32 summation order of the token reduction; every other gradient
    (LayerNorm, biases, dh) is bit-identical."""
    from interpolated_diffusion_b200.models.transformer import TransformerEncoder
    from interpolated_diffusion_b200.train import backward as bw
    torch.manual_seed(5)
    B, Lq, d, nl = 64, 64, 128, 3
    enc = TransformerEncoder(d_model=d, n_layers=nl, n_heads=4, d_ff=256, cond_dim=64).cuda()
    h0 = torch.randn((B * Lq, d), device="cuda")
    cond = torch.randn((B, 64), device="cuda")
    dh0 = torch.randn((B * Lq, d), device="cuda") * 0.1
    out = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("IDB200_TRAIN_DW_GROUPED", mode)
        bp = bw.EncoderBackprop(enc)
        bp.forward(h0.clone(), B, Lq, cond)
        grads = {"transformer." + n: torch.zeros_like(p) for n, p in enc.named_parameters()}
        dh, dh16 = dh0.clone(), dh0.bfloat16()
        bp.backward_layers(dh, dh16, grads)
        torch.cuda.synchronize()
        out[mode] = (grads, dh, dh16)
    ga, gb = out["1"][0], out["0"][0]
    assert torch.equal(out["1"][1], out["0"][1]) and torch.equal(out["1"][2], out["0"][2])
    for k in ga:
        if ".film" in k:
            continue
        if k.endswith(("ff.0.weight", "ff.2.weight", "in_proj_weight", "out_proj.weight")):
            den = max(1e-12, float(gb[k].norm()))
            assert float((ga[k] - gb[k]).norm()) / den < 1e-5, k
        else:
            assert torch.equal(ga[k], gb[k]), k
