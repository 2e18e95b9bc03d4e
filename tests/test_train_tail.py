"""Tail of the Stage-2 training step (loss + clip + AdamW + EMA): oracle vs golden outputs of live torch /
the reference's EMA (CPU), CUDA kernels vs the same goldens (GPU)."""
import numpy as np
import pytest
import torch

from oracle import optim_np as oo

SHAPES = 5


def test_oracle_loss_and_optimizer_vs_golden(golden):
    g = golden("optim")
    for n, ac in (("conf", True), ("mask", False)):
        loss, grad = oo.stage2_loss_and_grad(g[f"loss_{n}_delta_hat"], g[f"loss_{n}_target"], g[f"loss_{n}_weight"], ac, 0.1, 1.0,
                                             int(g[f"loss_{n}_grad_accum"]))
        assert abs(loss - float(g[f"loss_{n}_value"])) < 1e-6 * abs(loss)
        assert np.abs(grad - g[f"loss_{n}_grad"]).max() < 1e-8
    P = [g[f"opt_p0_{i}"] for i in range(SHAPES)]
    M, V, E = [np.zeros_like(p) for p in P], [np.zeros_like(p) for p in P], [p.copy() for p in P]
    clipped = []
    for step in (1, 2, 3):
        G = [g[f"opt_g{step}_{i}"] for i in range(SHAPES)]
        total, coef = oo.clip_coef(G, 1.0)
        clipped.append(coef < 1.0)
        assert abs(total - float(g[f"opt_norm{step}"])) < 1e-5 * total
        for i in range(SHAPES):
            P[i], M[i], V[i], E[i] = oo.adamw_ema_step(P[i], G[i], M[i], V[i], E[i], step, coef=coef)
            for got, key in ((P[i], "p"), (M[i], "m"), (V[i], "v"), (E[i], "ema")):
                assert np.abs(got - g[f"opt_{key}{step}_{i}"]).max() < 1e-7, (step, i, key)
    assert clipped == [False, True, False]


@pytest.mark.gpu
def test_cuda_loss_clip_adamw_ema_vs_golden(golden):
    from interpolated_diffusion_b200.train.optim import FlatAdamW, stage2_loss
    g = golden("optim")
    for n, ac in (("conf", True), ("mask", False)):
        loss, grad = stage2_loss(torch.from_numpy(g[f"loss_{n}_delta_hat"]).cuda(), torch.from_numpy(g[f"loss_{n}_target"]).cuda(),
                                 torch.from_numpy(g[f"loss_{n}_weight"]).cuda(), anchor_conf=ac, grad_accum=int(g[f"loss_{n}_grad_accum"]))
        ref = float(g[f"loss_{n}_value"])
        assert abs(float(loss) - ref) < 2e-6 * abs(ref)
        assert np.abs(grad.cpu().numpy() - g[f"loss_{n}_grad"]).max() < 1e-8
    params = [torch.nn.Parameter(torch.from_numpy(g[f"opt_p0_{i}"]).cuda()) for i in range(SHAPES)]
    opt = FlatAdamW(params, lr=2e-4, weight_decay=1e-2, ema_decay=0.999, max_grad_norm=1.0)
    for step in (1, 2, 3):
        for i, p in enumerate(params):
            p.grad = torch.from_numpy(g[f"opt_g{step}_{i}"]).cuda()
        norm = opt.step()
        opt.zero_grad()
        assert abs(float(norm) - float(g[f"opt_norm{step}"])) < 1e-5 * float(norm)
        for i, p in enumerate(params):
            for got, key in ((p.data, "p"), (opt.views(opt.exp_avg)[i], "m"), (opt.views(opt.exp_avg_sq)[i], "v"), (opt.ema_shadow[i], "ema")):
                ref = g[f"opt_{key}{step}_{i}"]              # one fp32 ulp: the clip coefficient is reduced in a different order
                assert (np.abs(got.cpu().numpy() - ref) <= 2.4e-7 * np.maximum(1.0, np.abs(ref))).all(), (step, i, key)
    # larger random case against the oracle, flat-gradient entry (what a gradient all-reduce hands over), no clipping
    gen = torch.Generator().manual_seed(5)
    big = [torch.nn.Parameter(torch.randn((1 << 20) + 3, generator=gen).cuda())]
    p0 = big[0].detach().cpu().numpy().copy()
    opt2 = FlatAdamW(big, max_grad_norm=None)
    fg = torch.randn((opt2.n,), generator=gen).cuda() * 0.1
    opt2.step(fg)
    pr, mr, vr, er = oo.adamw_ema_step(p0, fg.cpu().numpy()[:p0.size], np.zeros_like(p0), np.zeros_like(p0), p0.copy(), 1)
    assert (np.abs(big[0].detach().cpu().numpy() - pr) <= 2.4e-7 * np.maximum(1.0, np.abs(pr))).all()
    assert (np.abs(opt2.ema_shadow[0].cpu().numpy() - er) <= 2.4e-7 * np.maximum(1.0, np.abs(er))).all()
    with pytest.raises(ValueError):
        stage2_loss(torch.zeros((2, 3, 2)).cuda(), torch.zeros((2, 3, 2)).cuda(), torch.zeros((2, 4)).cuda())


@pytest.mark.gpu
def test_flat_adamw_state_dict_interchanges_with_torch_adamw(tmp_path):
    """FlatAdamW.state_dict() has torch.optim.AdamW's structure: resume a torch AdamW from it (and back) and take one more step
    on both sides with the same gradients -> same parameters; EMA state round-trips through save/load_checkpoint."""
    import torch
    from interpolated_diffusion_b200.train.optim import FlatAdamW
    from interpolated_diffusion_b200.utils.checkpoint import load_checkpoint, save_checkpoint
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 5)).to(dev)
    ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 5)).to(dev)
    ref.load_state_dict(net.state_dict())
    opt = FlatAdamW(net.parameters(), lr=1e-3, max_grad_norm=None)
    g = torch.Generator(device=dev).manual_seed(1)
    for _ in range(3):
        for p in net.parameters():
            p.grad = torch.randn(p.shape, device=dev, generator=g)
        opt.step()
    p = str(tmp_path / "c.pt")
    save_checkpoint(p, net, opt, 3, ema=opt.ema_state)
    topt = torch.optim.AdamW(ref.parameters(), lr=5.0)
    assert load_checkpoint(p, ref, topt) == 3                       # torch's AdamW accepts the state
    assert topt.param_groups[0]["lr"] == 1e-3 and float(topt.state[next(iter(ref.parameters()))]["step"]) == 3.0
    grads = [torch.randn(q.shape, device=dev, generator=g) for q in net.parameters()]
    for q, gr in zip(net.parameters(), grads):
        q.grad = gr.clone()
    for q, gr in zip(ref.parameters(), grads):
        q.grad = gr.clone()
    opt.step()
    topt.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        assert float((a - b).abs().max()) <= 2.4e-7 * max(1.0, float(b.abs().max()))
    # and back: a FlatAdamW resumes from torch's state dict; EMA shadow restored
    net2 = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.SiLU(), torch.nn.Linear(32, 5)).to(dev)
    opt2 = FlatAdamW(net2.parameters(), lr=9.0, max_grad_norm=None)
    save_checkpoint(p, ref, topt, 4, ema=opt.ema_state)
    assert load_checkpoint(p, net2, opt2, ema=opt2.ema_state) == 4
    assert opt2.step_count == 4 and opt2.lr == 1e-3
    assert all(torch.equal(a, b) for a, b in zip(opt2.ema_shadow, opt.ema_shadow))
    assert all(float((a - b).abs().max()) <= 2.4e-7 * max(1.0, float(b.abs().max())) for a, b in zip(opt2.views(opt2.exp_avg), opt.views(opt.exp_avg)))
