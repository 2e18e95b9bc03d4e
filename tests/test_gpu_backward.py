"""Backward of the Stage-2 denoiser (train/backward.py + csrc/train_bwd.cu) against torch autograd on the CPU fp32 functional
oracle (oracle/denoiser_torch.py): what ``loss.backward()`` computes in src/train/train_interp_levels.py:1142-1161.

Tolerances: the single kernels that work in fp32 (LayerNorm+FiLM backward, column sums, narrow outer products, strided fp32
GEMM) are checked at 2e-4 relative; everything that passes through bf16 operands (token GEMMs, attention backward, conv stack)
at a relative L2 error of 3e-2 per gradient tensor -- the bf16 tolerance of the north star (2e-2 max-abs on outputs) carried
to gradients."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import denoiser_torch as O  # noqa: E402


def _rel(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda", 0)


def test_transpose_colsum_reduce(dev):
    from interpolated_diffusion_b200.train import backward as BW
    g = torch.Generator(device="cpu").manual_seed(0)
    for M, N, dt in ((192, 96, torch.float32), (130, 40, torch.bfloat16), (4096, 384, torch.float32), (4096, 1536, torch.bfloat16),
                     (512, 96, torch.bfloat16)):
        x = torch.randn((M, N), generator=g).to(dt).to(dev)
        out = torch.empty((N, M), device=dev, dtype=torch.bfloat16)
        BW.transpose_bf16(x, out)
        assert torch.equal(out.cpu(), x.cpu().float().t().to(torch.bfloat16))
        sc = BW._Scratch()
        cs = torch.empty((N,), device=dev, dtype=torch.float32)
        sc.colsum(x, cs)
        ref = x.cpu().double().sum(0)
        assert float((cs.cpu().double() - ref).abs().max()) <= 2e-4 * max(1.0, float(ref.abs().max()))
        sc.colsum(x, cs, scale=0.5, accumulate=True)
        assert float((cs.cpu().double() - 1.5 * ref).abs().max()) <= 4e-4 * max(1.0, float(ref.abs().max()))


def test_splitk_weight_gradient_gemm(dev):
    from interpolated_diffusion_b200.train import backward as BW
    g = torch.Generator(device="cpu").manual_seed(1)
    for M, n_out, k_in in ((4096, 384, 128), (8192, 96, 64), (1024, 1152, 384), (2048, 32, 320), (4096, 1536, 384), (640, 200, 192)):
        dy = torch.randn((M, n_out), generator=g).to(torch.bfloat16)
        x = torch.randn((M, k_in), generator=g).to(torch.bfloat16)
        sc = BW._Scratch()
        tdy = BW.transpose_bf16(dy.to(dev), torch.empty((n_out, M), device=dev, dtype=torch.bfloat16))
        tx = BW.transpose_bf16(x.to(dev), torch.empty((k_in, M), device=dev, dtype=torch.bfloat16))
        out = torch.empty((n_out, k_in), device=dev, dtype=torch.float32)
        sc.dweight_t(tdy, tx, out)                           # K-major operands (transposed copies)
        ref = dy.double().t() @ x.double()
        assert _rel(out, ref) < 1e-5, (M, n_out, k_in)
        out2 = torch.empty_like(out)
        sc.dweight_t(tdy, tx, out2)
        assert torch.equal(out, out2)                       # fixed reduction order
        if k_in % 64 == 0:
            out3 = torch.empty_like(out)
            sc.dweight(dy.to(dev), x.to(dev), out3)         # MN-major operands, no transposes
            assert _rel(out3, ref) < 1e-5, ("mn", M, n_out, k_in)


def test_small_fp32_backward_kernels(dev):
    from interpolated_diffusion_b200 import _lib as L
    from interpolated_diffusion_b200.train import backward as BW
    g = torch.Generator(device="cpu").manual_seed(2)
    sc = BW._Scratch()
    # narrow outer product (out head / in_proj weight gradients)
    M, n, K = 4096, 5, 384
    A, X = torch.randn((M, n), generator=g), torch.randn((M, K), generator=g)
    out = torch.empty((n, K), device=dev)
    sc.narrow_outer(A.to(dev), X.to(dev), out)
    assert _rel(out, A.double().t() @ X.double()) < 2e-6
    # strided fp32 GEMM in all four transpose combinations
    a, b = torch.randn((70, 33), generator=g), torch.randn((50, 33), generator=g)
    for a_t in (False, True):
        for b_t in (False, True):
            Am = a.t().contiguous() if a_t else a
            Bm = b.t().contiguous() if b_t else b
            o = torch.empty((70, 50), device=dev)
            BW.sgemm_strided(Am.to(dev), a_t, Bm.to(dev), b_t, o)
            assert _rel(o, a.double() @ b.double().t()) < 2e-6
    # token sums and out-head backward
    src = torch.randn((8, 16, 128), generator=g)
    o = torch.empty((8, 128), device=dev)
    src_d = src.to(dev)                                     # keep device copies alive across the raw-pointer calls
    L.call("idb200_token_sum", src_d.data_ptr(), 8, 16, 128, o.data_ptr(), L.stream(dev))
    assert _rel(o, src.sum(1)) < 2e-6
    dy, W = torch.randn((300, 2), generator=g), torch.randn((2, 128), generator=g)
    dh = torch.empty((300, 128), device=dev)
    dh16 = torch.empty((300, 128), device=dev, dtype=torch.bfloat16)
    dy_d, W_d = dy.to(dev), W.to(dev)
    L.call("idb200_head_bwd", dy_d.data_ptr(), W_d.data_ptr(), 300, 128, 2, dh.data_ptr(), dh16.data_ptr(), L.stream(dev))
    assert _rel(dh, dy @ W) < 2e-6 and _rel(dh16, dy @ W) < 4e-3
    # SiLU forward / backward
    u = torch.randn((1000,), generator=g) * 3
    gg = torch.randn((1000,), generator=g)
    ur = u.clone().requires_grad_()
    torch.nn.functional.silu(ur).backward(gg)
    assert _rel(BW.silu_f32(u.to(dev)), torch.nn.functional.silu(u)) < 2e-6
    assert _rel(BW.silu_f32(u.to(dev), g=gg.to(dev)), ur.grad) < 2e-6


@pytest.mark.parametrize("two_pass", [False, True])
@pytest.mark.parametrize("d,Lseq,with_film", [(128, 8, True), (256, 64, True), (384, 64, True), (256, 16, False)])
def test_ln_film_backward_matches_autograd(dev, d, Lseq, with_film, two_pass):
    from interpolated_diffusion_b200 import _lib as L
    g = torch.Generator(device="cpu").manual_seed(3)
    B = 6
    h = (torch.randn((B, Lseq, d), generator=g) * 1.5 + 0.3).requires_grad_()
    w = (1.0 + 0.2 * torch.randn((d,), generator=g)).requires_grad_()
    b = (0.1 * torch.randn((d,), generator=g)).requires_grad_()
    gb = (0.3 * torch.randn((B, 2 * d), generator=g)).requires_grad_()
    da = torch.randn((B, Lseq, d), generator=g)
    dh0 = torch.randn((B, Lseq, d), generator=g)
    a = torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)
    if with_film:
        a = a * (1.0 + gb[:, None, :d]) + gb[:, None, d:]
    a.backward(da)
    dh = dh0.clone().to(dev).view(B * Lseq, d)
    dh16 = torch.empty((B * Lseq, d), device=dev, dtype=torch.bfloat16)
    dgb = torch.zeros((B, 2 * d), device=dev) if with_film else None
    dwb = torch.empty((B, 2 * d), device=dev)
    gbd = gb.detach().to(dev) if with_film else None
    da_d, h_d, w_d, b_d = da.to(dev), h.detach().to(dev), w.detach().to(dev), b.detach().to(dev)
    stats = torch.empty((B * Lseq, 4), device=dev) if two_pass else None
    L.call("idb200_ln_film_bwd", da_d.data_ptr(), h_d.data_ptr(), w_d.data_ptr(), b_d.data_ptr(), L.ptr(gbd), 2 * d if with_film else 0, B, Lseq, d, dh.data_ptr(), dh16.data_ptr(), L.ptr(dgb),
           2 * d if with_film else 0, dwb.data_ptr(), L.ptr(stats), L.stream(dev))
    assert _rel(dh.view(B, Lseq, d), dh0 + h.grad) < 2e-5
    assert _rel(dh16.view(B, Lseq, d), dh0 + h.grad) < 4e-3
    assert _rel(dwb.sum(0)[:d], w.grad) < 2e-5 and _rel(dwb.sum(0)[d:], b.grad) < 2e-5
    if with_film:
        assert _rel(dgb, gb.grad) < 2e-5


@pytest.mark.parametrize("da_bf16", [False, True])
@pytest.mark.parametrize("d,Lseq,with_film", [(256, 64, True), (384, 64, True), (128, 8, False)])
def test_ln_film_backward2_bf16_da_and_dh_column_sums(dev, d, Lseq, with_film, da_bf16):
    """idb200_ln_film_bwd2: gradient w.r.t. the LayerNorm output given in bf16, and the third block of dwb_part = per-trajectory sums
    of the UPDATED dh (the bias gradient of the GEMM below the LayerNorm) against autograd."""
    from interpolated_diffusion_b200 import _lib as L
    g = torch.Generator(device="cpu").manual_seed(5)
    B = 7
    h = (torch.randn((B, Lseq, d), generator=g) * 1.5 + 0.3).requires_grad_()
    w = (1.0 + 0.2 * torch.randn((d,), generator=g)).requires_grad_()
    b = (0.1 * torch.randn((d,), generator=g)).requires_grad_()
    gb = (0.3 * torch.randn((B, 2 * d), generator=g)).requires_grad_()
    da = torch.randn((B, Lseq, d), generator=g)
    if da_bf16:
        da = da.to(torch.bfloat16).float()                  # the reference sees the same (rounded) upstream gradient
    dh0 = torch.randn((B, Lseq, d), generator=g)
    a = torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)
    if with_film:
        a = a * (1.0 + gb[:, None, :d]) + gb[:, None, d:]
    a.backward(da)
    dh = dh0.clone().to(dev).view(B * Lseq, d)
    dh16 = torch.empty((B * Lseq, d), device=dev, dtype=torch.bfloat16)
    dgb = torch.zeros((B, 2 * d), device=dev) if with_film else None
    dwb = torch.empty((B, 3 * d), device=dev)
    gbd = gb.detach().to(dev) if with_film else None
    da_d = da.to(dev).to(torch.bfloat16) if da_bf16 else da.to(dev)
    h_d, w_d, b_d = h.detach().to(dev), w.detach().to(dev), b.detach().to(dev)
    stats = torch.empty((B * Lseq, 4), device=dev)
    L.call("idb200_ln_film_bwd2", da_d.data_ptr(), int(da_bf16), h_d.data_ptr(), w_d.data_ptr(), b_d.data_ptr(), L.ptr(gbd), 2 * d if with_film else 0,
           B, Lseq, d, dh.data_ptr(), dh16.data_ptr(), L.ptr(dgb), 2 * d if with_film else 0, dwb.data_ptr(), 1, stats.data_ptr(), L.stream(dev))
    want = dh0 + h.grad
    assert _rel(dh.view(B, Lseq, d), want) < 2e-5
    assert _rel(dwb[:, :d].sum(0), w.grad) < 2e-5 and _rel(dwb[:, d:2 * d].sum(0), b.grad) < 2e-5
    assert _rel(dwb[:, 2 * d:], want.sum(1)) < 2e-5
    if with_film:
        assert _rel(dgb, gb.grad) < 2e-5


@pytest.mark.parametrize("Lseq,H,causal,simt", [(64, 4, False, 0), (64, 12, True, 0), (64, 4, False, 1), (64, 12, True, 1), (48, 3, True, 0),
                                                 (40, 2, False, 0), (8, 8, False, 0), (32, 4, True, 0), (16, 2, False, 0), (5, 2, False, 0)])
def test_attention_backward_matches_autograd(dev, Lseq, H, causal, simt):
    from interpolated_diffusion_b200 import _lib as L
    g = torch.Generator(device="cpu").manual_seed(4)
    B, d = 3, 32 * H
    qkv = (torch.randn((B, Lseq, 3 * d), generator=g)).to(torch.bfloat16)
    dO = torch.randn((B, Lseq, d), generator=g).to(torch.bfloat16)
    x = qkv.float().requires_grad_()
    q, k, v = x.split(d, dim=-1)
    sh = lambda t: t.view(B, Lseq, H, 32).transpose(1, 2)
    sc = (sh(q) / math.sqrt(32.0)) @ sh(k).transpose(-1, -2)
    if causal:
        sc = sc + torch.triu(torch.full((Lseq, Lseq), float("-inf")), diagonal=1)
    o = (torch.softmax(sc, -1) @ sh(v)).transpose(1, 2).reshape(B, Lseq, d)
    o.backward(dO.float())
    dqkv = torch.empty((B * Lseq, 3 * d), device=dev, dtype=torch.bfloat16)
    qkv_d, dO_d = qkv.to(dev), dO.to(dev)
    L.call("idb200_attention_bwd", qkv_d.data_ptr(), dO_d.data_ptr(), dqkv.data_ptr(), B, Lseq, H, int(causal), simt, L.stream(dev))
    # fp32 kernel: bf16 rounding of the output only; tensor-core kernel (L > 32): P and dS are bf16 operands as well
    assert _rel(dqkv.view(B, Lseq, 3 * d), x.grad) < (6e-3 if (simt or Lseq <= 32) else 1e-2)
    if Lseq > 32 and not simt:
        # the form the training step uses: the same gradient + per-trajectory column sums of the bf16 rows (in_proj bias gradient)
        dqkv2 = torch.full_like(dqkv, float("nan"))
        sums = torch.full((B, 3 * d), float("nan"), device=dev)
        L.call("idb200_attention_bwd_sums", qkv_d.data_ptr(), dO_d.data_ptr(), dqkv2.data_ptr(), sums.data_ptr(), B, Lseq, H, int(causal), L.stream(dev))
        assert torch.equal(dqkv2, dqkv)
        want = dqkv.float().view(B, Lseq, 3 * d).sum(1)
        assert (sums - want).abs().max().item() <= 1e-5 * max(1.0, want.abs().max().item())


def _make_model(dev, d, nl, H, ff, maze_channels, C, causal=False):
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    torch.manual_seed(0)
    cls = InterpLevelCausalDenoiser if causal else InterpLevelDenoiser
    m = cls(d_model=d, n_layers=nl, n_heads=H, d_ff=ff, data_dim=2, max_levels=3, mask_channels=C, maze_channels=maze_channels)
    # random-init out / FiLM layers are small; scale a few up so every gradient is well away from the bf16 noise floor
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    return m.to(dev)


def _batch(B, T, C, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x_s = torch.rand((B, T, 2), generator=g)
    mask = (torch.rand((B, T, C), generator=g) < 0.3).float() if C > 1 else (torch.rand((B, T), generator=g) < 0.3)
    s = torch.randint(1, 4, (B,), generator=g)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=g)}
    target = 0.1 * torch.randn((B, T, 2), generator=g)
    conf = torch.rand((B, T), generator=g)
    return x_s, mask, s, cond, target, conf


@pytest.mark.parametrize("d,nl,H,ff,chan,C,causal", [(128, 2, 4, 256, (32, 64), 3, False), (256, 2, 8, 512, (32, 64, 64), 2, True),
                                                     (384, 2, 12, 1536, (32, 64, 128, 128), 3, False)])   # cfg-4 widths, 2 of 12 layers
def test_stage2_backward_matches_autograd(dev, d, nl, H, ff, chan, C, causal):
    """Every parameter gradient of InterpLevelDenoiser for the Stage-2 loss (train_interp_levels.py:1142-1159)."""
    from interpolated_diffusion_b200.train.backward import InterpLevelBackprop
    from interpolated_diffusion_b200.train.optim import stage2_loss
    B, T = 64, 64
    model = _make_model(dev, d, nl, H, ff, chan, C, causal)
    x_s, mask, s, cond, target, conf = _batch(B, T, C, 11)
    # oracle: fp32 autograd on the functional restatement
    sd = {k: v.detach().cpu().float().clone().requires_grad_() for k, v in model.state_dict().items()}
    ref = O.interp_level_denoiser(sd, H, x_s, s, mask, cond, causal=causal)
    w = 1.0 + (0.1 - 1.0) * conf
    ref_loss = (w[..., None] * (ref - target) ** 2).sum() / (w.sum() * 2 + 1e-8)
    ref_loss.backward()

    bp = InterpLevelBackprop(model)
    to = lambda t: t.to(dev)
    delta = bp.forward(to(x_s), to(s), to(mask), {k: to(v) for k, v in cond.items()})
    assert float((delta.cpu() - ref.detach()).abs().max()) < 2e-2
    loss, dgrad = stage2_loss(delta, to(target), to(conf), anchor_conf=True, w_anchor=0.1, w_missing=1.0)
    assert abs(float(loss) - float(ref_loss.detach())) < 2e-2 * abs(float(ref_loss.detach())) + 1e-6
    grads = bp.new_grads()
    bp.backward(dgrad, grads)
    torch.cuda.synchronize()
    bad = []
    for name, gref in ((k, v.grad) for k, v in sd.items()):
        r = _rel(grads[name], gref)
        if not (r < 3e-2):
            bad.append((name, r, float(gref.norm())))
    assert not bad, bad
    # deterministic
    grads2 = bp.new_grads()
    delta2 = bp.forward(to(x_s), to(s), to(mask), {k: to(v) for k, v in cond.items()})
    _, dgrad2 = stage2_loss(delta2, to(target), to(conf), anchor_conf=True, w_anchor=0.1, w_missing=1.0)
    bp.backward(dgrad2, grads2)
    assert all(torch.equal(grads[k], grads2[k]) for k in grads)


def test_stage2_trainer_step(dev):
    """A whole optimisation step (train_interp_levels.py:1034-1173): corruption -> forward -> loss -> backward -> clip -> AdamW
    + EMA.  (1) the gradients the optimiser consumes are the backward's; (2) the parameter update equals torch.optim.AdamW's
    on the autograd gradients up to the bf16 noise of the gradients (|delta| <= lr per element on step 1); (3) the loss falls
    on a fixed batch."""
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    B, T = 64, 64
    model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
    sd0 = {k: v.detach().cpu().float().clone() for k, v in model.state_dict().items()}
    tr = Stage2Trainer(model, lr=2e-4)
    g = torch.Generator(device="cpu").manual_seed(21)
    x0 = torch.rand((B, T, 2), generator=g).cumsum(1) / T
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((B, 4), generator=g).to(dev)}
    gen = torch.Generator(device=dev).manual_seed(23)
    x_s, s_idx, mask_in, target, wm = tr.build_batch(x0.to(dev), gen)
    assert mask_in.shape == (B, T, 3) and target.shape == (B, T, 2)
    loss0 = tr.loss_and_grads(x_s, s_idx, mask_in, cond, target, wm)
    # oracle step: autograd gradients + torch clip + torch AdamW
    sd = {k: v.clone().requires_grad_() for k, v in sd0.items()}
    cpu = lambda t: t.detach().cpu()
    ref = O.interp_level_denoiser(sd, 4, cpu(x_s), cpu(s_idx), cpu(mask_in), {k: cpu(v) for k, v in cond.items()})
    w = 1.0 + (0.1 - 1.0) * cpu(wm)
    ref_loss = (w[..., None] * (ref - cpu(target)) ** 2).sum() / (w.sum() * 2 + 1e-8)
    ref_loss.backward()
    assert abs(float(loss0) - float(ref_loss.detach())) < 2e-2 * float(ref_loss.detach())
    params = list(sd.values())
    ref_norm = torch.nn.utils.clip_grad_norm_(params, 1.0)
    torch.optim.AdamW(params, lr=2e-4, weight_decay=1e-2).step()
    tr.reduce_gradients()
    norm = tr.opt.step(tr.flat_grad)
    assert abs(float(norm) - float(ref_norm)) < 3e-2 * float(ref_norm)
    worst = 0.0
    for k, v in model.state_dict().items():
        worst = max(worst, float((v.detach().cpu().float() - sd[k].detach()).abs().max()))
    assert worst <= 2.0 * 2e-4 * 1.05, worst                      # a sign flip of a noise-level gradient moves 2 lr at most
    moved = sum(float((v.detach().cpu().float() - sd0[k]).abs().sum()) for k, v in model.state_dict().items())
    assert moved > 0.0
    # the loss goes down on a fixed batch
    losses = [float(loss0)]
    for _ in range(6):
        losses.append(float(tr.loss_and_grads(x_s, s_idx, mask_in, cond, target, wm)))
        tr.opt.step(tr.flat_grad)
    assert losses[-1] < 0.9 * losses[0], losses
    # and a full step() with fresh corruption runs end to end
    l2 = tr.step(x0.to(dev), cond, gen)
    assert math.isfinite(float(l2))


def test_stage1_backward_and_trainer(dev):
    """KeypointDenoiser gradients for the Stage-1 loss (train_keypoints.py:526-540) against autograd on the oracle, then a few
    optimisation steps on a fixed batch."""
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.train.train_keypoints import Stage1Trainer
    B, T, K = 64, 64, 8
    torch.manual_seed(0)
    model = KeypointDenoiser(d_model=128, n_layers=2, n_heads=4, d_ff=256, data_dim=2, kp_feat_dim=3)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    model = model.to(dev)
    sd = {k: v.detach().cpu().float().clone().requires_grad_() for k, v in model.state_dict().items()}
    tr = Stage1Trainer(model, T=T, K=K)
    g = torch.Generator(device="cpu").manual_seed(31)
    x0 = (0.1 + 0.8 * torch.rand((B, T, 2), generator=g)).to(dev)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((B, 4), generator=g).to(dev),
            "kp_feat": torch.rand((B, K, 3), generator=g).to(dev)}
    gen = torch.Generator(device=dev).manual_seed(33)
    z_t, t, idx, km, eps = tr.build_batch(x0, cond, gen)
    assert z_t.shape == (B, K, 2) and bool((eps[km] == 0).all())
    loss = tr.loss_and_grads(z_t, t, idx, km, cond, eps)
    cpu = lambda v: v.detach().cpu()
    ref = O.keypoint_denoiser(sd, 4, cpu(z_t), cpu(t), cpu(idx), cpu(km), {k: cpu(v) for k, v in cond.items()}, T)
    valid = (~cpu(km)).float()
    ref_loss = (((ref - cpu(eps)) ** 2) * valid).sum() / (valid.sum() + 1e-8)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss.detach())) < 2e-2 * float(ref_loss.detach())
    bad = [(k, _rel(tr.grads[k], v.grad)) for k, v in sd.items() if not (_rel(tr.grads[k], v.grad) < 3e-2)]
    assert not bad, bad
    losses = [float(loss)]
    for _ in range(6):
        tr.opt.step(tr.flat_grad)
        losses.append(float(tr.loss_and_grads(z_t, t, idx, km, cond, eps)))
    assert losses[-1] < 0.95 * losses[0], losses
    assert math.isfinite(float(tr.step(x0, cond, gen)))


def test_inference_caches_follow_the_optimizer(dev):
    """The fused AdamW kernel updates parameters in place behind torch's version counters; the packed-weight caches of the
    inference modules must still notice (PARAM_EPOCH in their key)."""
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    model = _make_model(dev, 256, 2, 8, 512, (32, 64), 3)
    x_s, mask, s, cond, target, conf = _batch(64, 64, 3, 5)
    to = lambda t: t.to(dev)
    args = (to(x_s), to(s), to(mask), {k: to(v) for k, v in cond.items()})
    before = model(*args).clone()
    tr = Stage2Trainer(model, lr=2e-3)
    for _ in range(3):
        tr.loss_and_grads(args[0], args[1], args[2], args[3], to(target), to(conf))
        tr.opt.step(tr.flat_grad)
    after = model(*args)
    sd = {k: v.detach().cpu().float() for k, v in model.state_dict().items()}
    ref = O.interp_level_denoiser(sd, 8, x_s, s, mask, cond)
    assert float((after.cpu() - ref).abs().max()) < 2e-2 * max(1.0, float(ref.abs().max()))
    assert float((after - before).abs().max()) > 1e-3


def test_stage2_trainer_cuda_graph_equals_eager(dev):
    """forward + loss + backward replayed as a CUDA graph give bit-identical gradients to the eager launches, also after the
    parameters changed (the graph re-reads them) and for a second batch."""
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
    tr = Stage2Trainer(model, cuda_graph=True)
    to = lambda t: t.to(dev)
    for seed in (5, 6):
        x_s, mask, s, cond, target, conf = _batch(64, 64, 3, seed)
        args = (to(x_s), to(s), to(mask), {k: to(v) for k, v in cond.items()}, to(target), to(conf))
        l_e = float(tr.loss_and_grads(*args))
        g_e = tr.flat_grad.clone()
        l_g = float(tr._graphed_loss_and_grads(*args))
        assert l_e == l_g and torch.equal(g_e, tr.flat_grad)
        tr.opt.step(tr.flat_grad)


def test_stage1_trainer_cuda_graph_steps(dev):
    """Stage-1 trainer with the graphed forward + backward: identical losses to the eager trainer over a few steps (same seeds)."""
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.train.train_keypoints import Stage1Trainer
    losses = {}
    for graph in (False, True):
        torch.manual_seed(0)
        model = KeypointDenoiser(d_model=128, n_layers=2, n_heads=4, d_ff=256, data_dim=2, kp_feat_dim=3).to(dev)
        tr = Stage1Trainer(model, T=64, K=8, cuda_graph=graph)
        g = torch.Generator(device="cpu").manual_seed(41)
        x0 = (0.1 + 0.8 * torch.rand((64, 64, 2), generator=g)).to(dev)
        cond = {"occ": (torch.rand((64, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((64, 4), generator=g).to(dev),
                "kp_feat": torch.rand((64, 8, 3), generator=g).to(dev)}
        gen = torch.Generator(device=dev).manual_seed(43)
        torch.manual_seed(7)                                         # t / noise come from the global RNG like the reference
        losses[graph] = []
        for _ in range(4):
            losses[graph].append(float(tr.step(x0, cond, gen)))
            if graph:
                tr.prefetch(x0, cond, gen)                           # graphed trainer also prefetches the next batch
    assert losses[False] == losses[True], losses


def test_stage2_trainer_mask_policies(dev):
    """kp_index_mode of train_interp_levels.py:890-967: 'uniform' nests every level around the uniform K_min anchors; 'random'
    is the mix-only alias of 'random_nested'; unknown modes raise like the reference."""
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
    with pytest.raises(ValueError, match="Unknown kp_index_mode"):
        Stage2Trainer(model, kp_index_mode="selector_only")
    tr = Stage2Trainer(model, kp_index_mode="uniform")
    assert Stage2Trainer.__init__.__kwdefaults__["kp_index_mode"] == "random_nested"
    B, T = 64, 64
    x0 = torch.rand((B, T, 2), device=dev)
    gen = torch.Generator(device=dev).manual_seed(3)
    masks, idxs = tr.build_masks(x0, gen)
    uni, umask = kf.sample_fixed_k_indices_uniform_batch(B, T, 8, device=dev)
    assert torch.equal(idxs[3], uni) and torch.equal(masks[:, 3], umask)
    for s in range(3):
        assert bool((masks[:, s] | ~masks[:, s + 1]).all())          # M_{s+1} subset of M_s
        assert int(masks[:, s].sum(1).min()) == int(masks[:, s].sum(1).max()) == idxs[s].shape[1]
    x_s, s_idx, mask_in, target, wm = tr.build_batch(x0, gen)
    top = s_idx == 3
    assert torch.equal(mask_in[top][..., 0].bool(), umask[top])


def test_stage2_trainer_bootstrap_branch(dev):
    """train_interp_levels.py:970-1032: a frozen Stage-1 model replaces interior level-S anchors by its own samples; those
    positions are flagged in student_mask (lower confidence), everything else of x0 is untouched."""
    from interpolated_diffusion_b200.diffusion.schedules import make_alpha_bars, make_beta_schedule
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.sample.sample_generate import _build_known_mask_values, _sample_keypoints_ddim
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    from interpolated_diffusion_b200.utils.normalize import logit_pos, sigmoid_pos
    model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
    torch.manual_seed(1)
    kp = KeypointDenoiser(d_model=128, n_layers=2, n_heads=4, d_ff=256, data_dim=2).to(dev)
    sched = {k: v.to(dev) for k, v in make_alpha_bars(make_beta_schedule("cosine", 1000)).items()}
    tr = Stage2Trainer(model, corrupt_mode="none", bootstrap_model=kp, bootstrap_schedule=sched, bootstrap_logit=True,
                       bootstrap_prob_end=1.0, bootstrap_warmup_steps=0, bootstrap_prob_cap=1.0, bootstrap_mode="per_example",
                       bootstrap_replace_prob=1.0, bootstrap_ddim_steps=3)
    B, T = 64, 64
    g = torch.Generator(device="cpu").manual_seed(9)
    x0 = (0.1 + 0.8 * torch.rand((B, T, 2), generator=g)).to(dev)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((B, 4), generator=g).to(dev)}
    gen = torch.Generator(device=dev).manual_seed(5)
    masks, idxs = tr.build_masks(x0, gen)
    z_T = torch.randn((B, 8, 2), generator=g).to(dev)
    x0_used, student = tr._bootstrap(x0, cond, idxs, gen, z_T=z_T)
    idx_s = idxs[3]
    interior = torch.zeros((B, T), device=dev, dtype=torch.bool).scatter_(1, idx_s, (idx_s != 0) & (idx_s != T - 1))
    assert torch.equal(student, interior)                                   # replace_prob = 1, every sample bootstrapped
    km, kv = _build_known_mask_values(idx_s, cond, 2, T, True)
    z = sigmoid_pos(_sample_keypoints_ddim(kp, sched, idx_s, km, logit_pos(kv), cond, 3, T, schedule_name="quadratic", z_T=z_T))
    got = x0_used.gather(1, idx_s.unsqueeze(-1).expand(-1, -1, 2))
    inner = ((idx_s != 0) & (idx_s != T - 1)).unsqueeze(-1)
    assert torch.equal(torch.where(inner, got, z), z)                        # student predictions at the interior anchors
    assert torch.equal(x0_used[~student], x0[~student])                      # nothing else moved
    # the confidence channel sees the student anchors (0.5 instead of 0.95 before annealing) and a full step runs
    x_s, s_idx, mask_in, target, wm = tr.build_batch(x0, gen, cond)
    assert mask_in.shape == (B, T, 3) and math.isfinite(float(tr.step(x0, cond, gen)))


def test_stage2_trainer_selector_policies(dev):
    """kp_index_mode = selector (train_interp_levels.py:911-947): masks ranked by a frozen KeypointSelector, plain and
    level-conditioned; nested, right sizes, endpoints kept."""
    from interpolated_diffusion_b200.models.keypoint_selector import KeypointSelector
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
    B, T = 64, 64
    g = torch.Generator(device="cpu").manual_seed(2)
    x0 = torch.rand((B, T, 2), generator=g).to(dev)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((B, 4), generator=g).to(dev)}
    gen = torch.Generator(device=dev).manual_seed(1)
    for use_level in (False, True):
        torch.manual_seed(5)
        sel = KeypointSelector(T=T, d_model=64, n_heads=2, d_ff=128, use_level=use_level).to(dev)
        tr = Stage2Trainer(model, kp_index_mode="selector", selector_model=sel)
        masks, idxs = tr.build_masks(x0, gen, cond)
        assert masks.shape == (B, 4, T) and [i.shape[1] for i in idxs] == [64, 32, 16, 8]
        assert bool(masks[:, :, 0].all()) and bool(masks[:, :, -1].all())
        if not use_level:                                            # one ranking: levels are nested by construction
            for s in range(3):
                assert bool((masks[:, s] | ~masks[:, s + 1]).all())
        assert math.isfinite(float(tr.step(x0, cond, gen)))
    with pytest.raises(ValueError, match="selector model not loaded"):
        Stage2Trainer(model, kp_index_mode="selector")


def test_stage2_trainer_prefetch_equals_plain_steps(dev):
    """step() + prefetch() (next batch built on a side stream) walks exactly the same losses as plain step() calls: same
    generator draws in the same order, same data."""
    from interpolated_diffusion_b200.train.stage2_step import Stage2Trainer
    losses = {}
    for mode in ("plain", "prefetch"):
        model = _make_model(dev, 128, 2, 4, 256, (32, 64), 3)
        tr = Stage2Trainer(model, cuda_graph=True)
        g = torch.Generator(device="cpu").manual_seed(77)
        x0 = torch.rand((64, 64, 2), generator=g).to(dev)
        cond = {"occ": (torch.rand((64, 1, 21, 21), generator=g) < 0.2).float().to(dev), "start_goal": torch.rand((64, 4), generator=g).to(dev)}
        gen = torch.Generator(device=dev).manual_seed(9)
        out = []
        for _ in range(5):
            out.append(tr.step(x0, cond, gen))
            if mode == "prefetch":
                tr.prefetch(x0, cond, gen)
        losses[mode] = [float(v) for v in out]
    assert losses["plain"] == losses["prefetch"], losses
