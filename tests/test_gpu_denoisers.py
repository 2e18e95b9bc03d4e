"""GPU parity of the denoiser path (conv encoder, embeddings, LN+FiLM, tcgen05 GEMMs, attention, output head)
against the CPU oracle on identical weights and inputs, and against the golden outputs of the live reference.

Tolerances (BASELINE.json north_star): bf16 path <= 2e-2 max-abs on eps / delta (O(1) outputs), fp32 check
mode <= 1e-4."""
import numpy as np
import pytest
import torch

from oracle import denoiser_torch as odn

pytestmark = pytest.mark.gpu

TINY = dict(d_model=64, n_layers=2, n_heads=2, d_ff=128, d_cond=32, maze_channels=(8, 16))


def _sd(g, prefix):
    return {k[len(prefix):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(prefix)}


def _cuda(d):
    return {k: v.cuda() for k, v in d.items()}


def _maxabs(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()


def test_attention_kernels_vs_torch():
    from interpolated_diffusion_b200.models import _engine as E
    g = torch.Generator(device="cuda").manual_seed(0)
    for (B, L, H, causal) in [(5, 8, 8, False), (3, 64, 8, False), (2, 64, 8, True), (2, 256, 4, True), (3, 16, 2, False),
                              (2, 32, 4, True), (7, 8, 2, True), (1, 256, 8, False), (2, 128, 2, False), (3, 5, 2, False),
                              (16, 8, 8, False), (24, 8, 8, True), (6, 8, 8, True), (12, 8, 2, False), (64, 8, 4, False),
                              # round 2: lengths that are not a multiple of 16 run on the tcgen05 path too (key masking, no SIMT fallback)
                              (7, 33, 2, False), (7, 33, 2, True), (5, 100, 2, False), (3, 129, 2, True), (3, 200, 8, False), (37, 8, 12, False),
                              (9, 48, 12, True), (1, 1, 2, False), (130, 2, 2, True)]:
        d = H * 32
        qkv = torch.randn((B * L, 3 * d), generator=g, device="cuda")
        q, k, v = [t.view(B, L, H, 32).transpose(1, 2) for t in qkv.split(d, dim=-1)]
        ref32 = torch.nn.functional.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(B * L, d)
        out32 = E.attention(qkv, torch.empty((B * L, d), device="cuda"), B, L, H, causal)
        assert _maxabs(out32, ref32) < 2e-5, (B, L, H, causal)
        qb = qkv.bfloat16()
        qf, kf, vf = [t.float().view(B, L, H, 32).transpose(1, 2) for t in qb.split(d, dim=-1)]
        refb = torch.nn.functional.scaled_dot_product_attention(qf, kf, vf, is_causal=causal).transpose(1, 2).reshape(B * L, d)
        # default bf16 path: tcgen05 (S = Q K^T and O = P V as tcgen05.mma, scores / output accumulators in tensor memory)
        outb = E.attention(qb, torch.full((B * L, d), float("nan"), device="cuda", dtype=torch.bfloat16), B, L, H, causal)
        assert _maxabs(outb, refb) < 2.5e-2, (B, L, H, causal, "tcgen05")
        # legacy mma.sync path (force_simt = 2), kept as a cross-check
        outl = E.attention(qb, torch.empty((B * L, d), device="cuda", dtype=torch.bfloat16), B, L, H, causal, force_simt=2)
        assert _maxabs(outl, refb) < 2.5e-2, (B, L, H, causal, "mma" if (L % 16 == 0 or L == 8) else "simt")
        outs = E.attention(qb, torch.empty((B * L, d), device="cuda", dtype=torch.bfloat16), B, L, H, causal, force_simt=True)
        assert _maxabs(outs, refb) < 1e-2, (B, L, H, causal, "simt-bf16")


def test_small_kernels_vs_torch():
    from interpolated_diffusion_b200.models import _engine as E
    g = torch.Generator(device="cuda").manual_seed(1)
    # sgemm
    for (M, N, K) in [(100, 70, 33), (1, 256, 128), (513, 512, 4), (64, 64, 256)]:
        A = torch.randn((M, K), generator=g, device="cuda")
        W = torch.randn((N, K), generator=g, device="cuda")
        b = torch.randn((N,), generator=g, device="cuda")
        ref = A.double() @ W.double().t() + b.double()
        assert _maxabs(E.sgemm(A, W, b), ref.float()) < 1e-4
        assert _maxabs(E.sgemm(A, W, b, act=1), torch.nn.functional.silu(ref).float()) < 1e-4
        o = torch.ones((M, N), device="cuda")
        assert _maxabs(E.sgemm(A, W, None, o, accumulate=True), (A.double() @ W.double().t() + 1).float()) < 1e-4
    # LN + FiLM
    for d in (64, 256, 384):
        M, Lq = 77 * 8, 8
        h = torch.randn((M, d), generator=g, device="cuda") * 3 + 1
        w, b = torch.randn((d,), generator=g, device="cuda"), torch.randn((d,), generator=g, device="cuda")
        gb = torch.randn((M // Lq, 2 * d), generator=g, device="cuda")
        ref = torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)
        ref = (ref.view(-1, Lq, d) * (1 + gb[:, None, :d]) + gb[:, None, d:]).view(M, d)
        assert _maxabs(E.ln_film(h, w, b, gb, torch.empty_like(h), Lq), ref) < 2e-5
        ob = E.ln_film(h, w, b, gb, torch.empty((M, d), device="cuda", dtype=torch.bfloat16), Lq)
        assert ((ob.float() - ref).abs() <= 2.0 ** -8 * ref.abs() + 1e-6).all()      # one bf16 rounding
        assert _maxabs(E.ln_film(h, w, b, None, torch.empty_like(h), Lq), torch.nn.functional.layer_norm(h, (d,), w, b, 1e-5)) < 2e-5
    # conv encoder
    for (chans, H, W) in [((1, 32, 64), 21, 21), ((2, 8, 16), 9, 12), ((1, 32, 64, 128, 128), 21, 21), ((2, 16, 32), 9, 12), ((1, 64, 64), 12, 9)]:
        B = 5
        x = (torch.rand((B, chans[0], H, W), generator=g, device="cuda") < 0.3).float()
        ws = [torch.randn((chans[i + 1], chans[i], 3, 3), generator=g, device="cuda") * (chans[i] * 9) ** -0.5 for i in range(len(chans) - 1)]
        bs = [torch.randn((c,), generator=g, device="cuda") * 0.1 for c in chans[1:]]
        ref = x.double()                  # fp64 reference (cuDNN fp32 convolutions default to TF32)
        for w_, b_ in zip(ws, bs):
            ref = torch.nn.functional.silu(torch.nn.functional.conv2d(ref, w_.double(), b_.double(), padding=1))
        ref = ref.mean(dim=[2, 3]).float()
        got = E.conv_encoder(x[:, 0:1].contiguous(), x[:, 1:2].contiguous() if chans[0] == 2 else None, ws, bs)
        assert _maxabs(got, ref) < 2e-5, chans
        if len(chans) == 3 and chans[1] % 16 == 0 and chans[2] in (32, 64):
            # tensor-core implicit-GEMM path: bf16 activations / weights of the second conv, fp32 accumulate
            w1p = ws[1].permute(0, 2, 3, 1).reshape(chans[2], -1).to(torch.bfloat16).contiguous()
            got_tc = E.conv_encoder_tc(x[:, 0:1].contiguous(), x[:, 1:2].contiguous() if chans[0] == 2 else None, ws[0], bs[0], w1p, bs[1])
            assert _maxabs(got_tc, ref) < 3e-3, (chans, _maxabs(got_tc, ref))
    # tcgen05 conv encoder (maze_channels (32, 64)): several mazes per CTA (persistent loop, both accumulator buffers), sdf channel
    # (the activation buffer is handed over per MMA tile: shapes with one pass / junk tiles / the widest padded row / two channels)
    for cin, B, H, W in [(1, 5, 21, 21), (1, 148 * 3 + 7, 21, 21), (2, 301, 21, 21), (1, 9, 12, 9), (1, 300, 21, 22), (1, 311, 5, 7),
                         (2, 200, 16, 14), (1, 151, 3, 30), (1, 1, 21, 21)]:
        x = torch.rand((B, cin, H, W), generator=g, device="cuda")
        x[:, 0] = (x[:, 0] < 0.3).float()
        ws = [torch.randn((32, cin, 3, 3), generator=g, device="cuda") * (cin * 9) ** -0.5, torch.randn((64, 32, 3, 3), generator=g, device="cuda") * (32 * 9) ** -0.5]
        bs = [torch.randn((32,), generator=g, device="cuda") * 0.1, torch.randn((64,), generator=g, device="cuda") * 0.1]
        ref = x.double()
        for w_, b_ in zip(ws, bs):
            ref = torch.nn.functional.silu(torch.nn.functional.conv2d(ref, w_.double(), b_.double(), padding=1))
        ref = ref.mean(dim=[2, 3]).float()
        w1p = ws[1].permute(0, 2, 3, 1).reshape(64, -1).to(torch.bfloat16).contiguous()
        occ_, sdf_ = x[:, 0:1].contiguous(), (x[:, 1:2].contiguous() if cin == 2 else None)
        got5 = E.conv_encoder_tc5(occ_, sdf_, ws[0], bs[0], w1p, bs[1])
        assert torch.isfinite(got5).all()
        assert _maxabs(got5, ref) < 3e-3, (cin, B, H, W, _maxabs(got5, ref))
        assert torch.equal(got5, E.conv_encoder_tc5(occ_, sdf_, ws[0], bs[0], w1p, bs[1]))      # deterministic


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_denoisers_tiny_golden(golden, precision, tol):
    """Outputs of the LIVE reference on the tiny random-init models (tests/golden/models_tiny.npz)."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    g = golden("models_tiny")
    t = lambda k: torch.from_numpy(g[k]).cuda()
    cond = {"occ": t("kp2_occ"), "start_goal": t("kp2_sg")}
    m = KeypointDenoiser(data_dim=2, **TINY).cuda()
    m.load_state_dict(_sd(g, "kp2/"))
    m.precision = precision
    assert _maxabs(m.cond_enc(cond), torch.from_numpy(g["kp2_condvec"])) < 1e-5
    eps = m(t("kp2_z"), t("kp2_t"), t("kp2_idx"), t("kp2_km"), cond, 16)
    assert _maxabs(eps, torch.from_numpy(g["kp2_eps"])) < tol
    m4 = KeypointDenoiser(data_dim=4, kp_feat_dim=3, use_sdf=True, **TINY).cuda()
    m4.load_state_dict(_sd(g, "kp4/"))
    m4.precision = precision
    cond4 = {"occ": t("kp4_occ"), "start_goal": t("kp4_sg"), "sdf": t("kp4_sdf"), "kp_feat": t("kp4_kpfeat")}
    eps = m4(t("kp4_z"), t("kp2_t"), t("kp2_idx"), t("kp4_km"), cond4, 16)
    assert _maxabs(eps, torch.from_numpy(g["kp4_eps"])) < tol
    for tag, cls, C, D in (("il2", InterpLevelDenoiser, 2, 2), ("il3", InterpLevelDenoiser, 3, 4), ("ic1", InterpLevelCausalDenoiser, 1, 2)):
        mi = cls(data_dim=D, max_levels=3, mask_channels=C, **TINY).cuda()
        mi.load_state_dict(_sd(g, tag + "/"))
        mi.precision = precision
        out = mi(t(f"{tag}_x"), t(f"{tag}_s"), t(f"{tag}_mask"), cond)
        assert _maxabs(out, torch.from_numpy(g[f"{tag}_out"])) < tol, tag


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_denoisers_full_size_vs_oracle(precision, tol):
    """BASELINE small model (d=256, 8 layers, 8 heads, ff 1024, maze 32-64), random init: Stage-1 (K=8) and
    Stage-2 (T=64) single evaluations on identical inputs (ladder L2 of SURVEY 7.3-1)."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    B, T, K, D = 24, 64, 8, 2
    gen = torch.Generator().manual_seed(5)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D)
    il = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=2)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    sd_il = {k: v.clone() for k, v in il.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    kp.precision = il.precision = precision
    idx = torch.tensor([0, 9, 18, 27, 36, 45, 54, 63]).repeat(B, 1)
    z = torch.randn((B, K, D), generator=gen)
    km = torch.zeros((B, K, D), dtype=torch.bool)
    km[:, 0] = km[:, -1] = True
    for tval in (999, 500, 2):
        t = torch.full((B,), tval, dtype=torch.long)
        ref = odn.keypoint_denoiser(sd_kp, 8, z, t, idx, km, cond, T)
        got = kp(z.cuda(), t.cuda(), idx.cuda(), km.cuda(), _cuda(cond), T)
        assert _maxabs(got, ref) < tol, (tval, _maxabs(got, ref))
    x = torch.rand((B, T, D), generator=gen)
    mask_in = torch.rand((B, T, 2), generator=gen)
    s = torch.full((B,), 3, dtype=torch.long)
    ref = odn.interp_level_denoiser(sd_il, 8, x, s, mask_in, cond)
    got = il(x.cuda(), s.cuda(), mask_in.cuda(), _cuda(cond))
    assert _maxabs(got, ref) < tol, _maxabs(got, ref)
    # hoisted loop-invariants give the same result as the plain call
    cv = il.encode_cond(_cuda(cond))
    got2 = il(x.cuda(), s.cuda(), mask_in.cuda(), None, cond_vec=cv, film=il.transformer.packed().film_params(cv, T, il.precision),
              level_vec=il.level_vector(torch.tensor([3]).cuda()))
    assert _maxabs(got2, got) < 1e-6


def test_causal_long_horizon_vs_oracle():
    """BASELINE config 5 shape: T=256 causal Stage-2 denoiser (small model), bf16 path."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    B, T, D = 3, 256, 2
    gen = torch.Generator().manual_seed(6)
    cond = {"occ": (torch.rand((B, 1, 9, 12), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(1)
    m = InterpLevelCausalDenoiser(data_dim=D, max_levels=4, mask_channels=1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    x = torch.rand((B, T, D), generator=gen)
    mask = torch.rand((B, T), generator=gen) < 0.2
    s = torch.tensor([4, 2, 1])
    ref = odn.interp_level_denoiser(sd, 8, x, s, mask, cond, causal=True)
    got = m(x.cuda(), s.cuda(), mask.cuda(), _cuda(cond))
    assert _maxabs(got, ref) < 2e-2, _maxabs(got, ref)
    m.precision = "fp32"
    got = m(x.cuda(), s.cuda(), mask.cuda(), _cuda(cond))
    assert _maxabs(got, ref) < 1e-4, _maxabs(got, ref)


@pytest.mark.parametrize("Lseq,B,causal", [(8, 37, False), (64, 5, False), (64, 3, True), (16, 9, False), (32, 4, True),
                                           (128, 3, False), (4, 70, False), (1, 130, False)])
def test_fused_transformer_blocks_vs_oracle(Lseq, B, causal):
    """idb200_attn_block + idb200_mlp_block (two kernels per layer, LN/FiLM/QKV/attention/out-proj and LN/FiLM/MLP fused)
    against the CPU oracle of transformer.py:35-46 and against the unfused kernel sequence; ragged last tile
    (B*L not a multiple of 128), FiLM on, bidirectional and causal."""
    from interpolated_diffusion_b200.models.transformer import TransformerEncoder
    torch.manual_seed(7 + Lseq)
    enc = TransformerEncoder(d_model=256, n_layers=2, n_heads=8, d_ff=1024, cond_dim=128, causal=causal)
    with torch.no_grad():
        for l in enc.layers:      # non-trivial LayerNorm affine / biases (default init is 1 / 0)
            for p_ in (l.norm1.weight, l.norm2.weight):
                p_.add_(0.1 * torch.randn_like(p_))
            for p_ in (l.norm1.bias, l.norm2.bias, l.attn.in_proj_bias, l.attn.out_proj.bias):
                p_.add_(0.1 * torch.randn_like(p_))
    sd = {"transformer.layers." + k[len("layers."):]: v.clone() for k, v in enc.state_dict().items()}
    gen = torch.Generator().manual_seed(11)
    x = torch.randn((B, Lseq, 256), generator=gen)
    cv = torch.randn((B, 128), generator=gen)
    ref = odn.transformer_encoder(x, cv, sd, 8, causal)
    enc = enc.cuda()
    pk = enc.packed()
    assert pk.fuse_blocks and pk.fuse_encoder
    whole = enc(x.cuda(), cv.cuda())                 # idb200_encoder_fused: one kernel for all layers
    pk.fuse_encoder = False
    got = enc(x.cuda(), cv.cuda())                   # attn_block + mlp_block per layer
    pk.fuse_blocks = False
    unfused = enc(x.cuda(), cv.cuda())
    pk.fuse_blocks = pk.fuse_encoder = True
    scale = ref.abs().max().item()
    assert _maxabs(whole, ref) < 2e-2 * max(1.0, scale / 4), (_maxabs(whole, ref), scale)
    assert _maxabs(got, ref) < 2e-2 * max(1.0, scale / 4), (_maxabs(got, ref), scale)
    assert _maxabs(got, unfused) < 2e-2 * max(1.0, scale / 4), _maxabs(got, unfused)
    assert _maxabs(whole, unfused) < 2e-2 * max(1.0, scale / 4), _maxabs(whole, unfused)
    # no FiLM
    enc2 = TransformerEncoder(d_model=256, n_layers=1, n_heads=8, d_ff=1024, cond_dim=None, causal=causal)
    sd2 = {"transformer.layers." + k[len("layers."):]: v.clone() for k, v in enc2.state_dict().items()}
    ref2 = odn.transformer_encoder(x, None, sd2, 8, causal)
    got2 = enc2.cuda()(x.cuda(), None)
    assert _maxabs(got2, ref2) < 2e-2 * max(1.0, ref2.abs().max().item() / 4), _maxabs(got2, ref2)


@pytest.mark.parametrize("Lseq,B,n_layers,ff,causal", [(8, 16 * 148 * 2 + 37, 8, 1024, False), (64, 2 * 148 + 3, 3, 512, True)])
def test_encoder_fused_many_tiles(Lseq, B, n_layers, ff, causal):
    """idb200_encoder_fused with several tiles per CTA (persistent loop, parameter / weight rings wrapping across tiles and
    layers, ragged last tile) against the unfused kernel sequence, and idempotent across repeated launches."""
    from interpolated_diffusion_b200.models.transformer import TransformerEncoder
    torch.manual_seed(3)
    enc = TransformerEncoder(d_model=256, n_layers=n_layers, n_heads=8, d_ff=ff, cond_dim=128, causal=causal)
    with torch.no_grad():
        for l in enc.layers:
            for p_ in (l.norm1.bias, l.norm2.bias, l.attn.in_proj_bias, l.attn.out_proj.bias, l.ff[2].bias):
                p_.add_(0.1 * torch.randn_like(p_))
    enc = enc.cuda()
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn((B, Lseq, 256), generator=gen, device="cuda")
    cv = torch.randn((B, 128), generator=gen, device="cuda")
    pk = enc.packed()
    whole = enc(x, cv)
    again = enc(x, cv)
    pk.fuse_encoder = pk.fuse_blocks = False
    unfused = enc(x, cv)
    assert torch.isfinite(whole).all()
    assert torch.equal(whole, again)
    scale = unfused.abs().max().item()
    assert _maxabs(whole, unfused) < 2e-2 * max(1.0, scale / 4), (_maxabs(whole, unfused), scale)


def test_encoder_fused_many_tiles_vs_oracle():
    """Round 2: the many-tile regime of the whole-encoder kernel held to the ORACLE (CPU fp32), not to another CUDA path: rows are
    sampled from every region of a launch that gives each CTA pair several tiles (first / middle / ragged last tiles), so ring
    wrap-around, the pair's dead tile and the parameter reloads are all covered by an independent reference."""
    from interpolated_diffusion_b200.models.transformer import TransformerEncoder
    for (Lseq, B, causal) in [(8, 16 * 148 * 2 + 37, False), (64, 2 * 148 * 2 + 3, True)]:
        torch.manual_seed(11)
        enc = TransformerEncoder(d_model=256, n_layers=4, n_heads=8, d_ff=1024, cond_dim=128, causal=causal)
        sd = {"transformer.layers." + k[len("layers."):]: v.clone() for k, v in enc.state_dict().items()}
        gen = torch.Generator().manual_seed(6)
        x = torch.randn((B, Lseq, 256), generator=gen)
        cv = torch.randn((B, 128), generator=gen)
        enc = enc.cuda()
        assert enc.packed().fused_path(Lseq)
        got = enc(x.cuda(), cv.cuda()).cpu()
        per_tile = 128 // Lseq
        n_tiles = (B + per_tile - 1) // per_tile
        # trajectories from the first tiles, a wrap-around region in the middle, and the ragged tail
        pick = sorted(set(list(range(0, 3 * per_tile)) + list(range((n_tiles // 2) * per_tile, (n_tiles // 2 + 2) * per_tile))
                          + list(range(B - per_tile - 5, B))))
        pick = [b for b in pick if 0 <= b < B]
        ref = odn.transformer_encoder(x[pick], cv[pick], sd, 8, causal)
        scale = ref.abs().max().item()
        assert _maxabs(got[pick], ref) < 2e-2 * max(1.0, scale / 4), (Lseq, _maxabs(got[pick], ref), scale)


@pytest.mark.parametrize("which", ["keypoints", "interp"])
def test_denoiser_fused_io_matches_separate_kernels(which):
    """idb200_denoiser_fused (token assembly + encoder + out head in one launch, h only in tensor memory) against the
    three-launch sequence embed_tokens -> encoder_fused -> out_head on the same weights: the residual stream entering and
    leaving the encoder is bit-identical (same fp32 operation order), only the 256-term head dot product is re-associated."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    torch.manual_seed(11)
    gen = torch.Generator().manual_seed(12)
    B, T, K, D = 37, 64, 8, 2
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
    if which == "keypoints":
        m = KeypointDenoiser(data_dim=D).cuda()
        idx = torch.sort(torch.stack([torch.randperm(T, generator=gen)[:K] for _ in range(B)]), dim=1).values.cuda()
        args = (torch.randn((B, K, D), generator=gen).cuda(), torch.full((B,), 500, dtype=torch.long).cuda(), idx,
                (torch.rand((B, K, D), generator=gen) < 0.3).cuda(), cond, T)
    else:
        m = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=2).cuda()
        args = (torch.rand((B, T, D), generator=gen).cuda(), torch.full((B,), 2, dtype=torch.long).cuda(), torch.rand((B, T, 2), generator=gen).cuda(), cond)
    m.fuse_io = True
    fused = m(*args)
    m.fuse_io = False
    assert m.fuse_head
    head_only = m(*args)                     # embed_tokens kernel -> encoder + head in one launch (the default)
    m.fuse_head = False
    separate = m(*args)                      # embed_tokens -> encoder_fused -> out_head
    assert torch.isfinite(fused).all()
    tol = 1e-5 * max(1.0, separate.abs().max().item())
    assert _maxabs(fused, separate) < tol, _maxabs(fused, separate)
    assert _maxabs(head_only, separate) < tol, _maxabs(head_only, separate)
    assert torch.equal(head_only, fused)     # same kernel, bit-identical residual stream


@pytest.mark.parametrize("B,T,K,kp_feat_dim,per_sample_t", [(37, 64, 8, 0, True), (300, 32, 16, 3, False), (5, 64, 32, 0, False), (129, 48, 8, 4, True)])
def test_staged_prologue_shapes(B, T, K, kp_feat_dim, per_sample_t):
    """Token assembly as the whole-encoder kernel's tile prologue with TMA-staged operands (table <= 64 rows, <= 8 features, K >= 8)
    against the separate idb200_embed_tokens kernel on the same weights: bit-identical residual stream, so identical eps.  Covers
    7-8 features (the predicated feature loop), tables shorter than 64 rows (zero-filled boxes), row_a per sample (global) and
    batch-constant (staged), dead rows / dead pair tiles, and a single tile."""
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    torch.manual_seed(21)
    gen = torch.Generator().manual_seed(22)
    D = 2
    m = KeypointDenoiser(data_dim=D, kp_feat_dim=kp_feat_dim).cuda()
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
    if kp_feat_dim:
        cond["kp_feat"] = torch.randn((B, K, kp_feat_dim), generator=gen).cuda()
    idx = torch.sort(torch.stack([torch.randperm(T, generator=gen)[:K] for _ in range(B)]), dim=1).values.cuda()
    z = torch.randn((B, K, D), generator=gen).cuda()
    km = (torch.rand((B, K, D), generator=gen) < 0.3).cuda()
    t = torch.randint(0, 1000, (B,), generator=gen).cuda() if per_sample_t else torch.full((B,), 123, dtype=torch.long).cuda()
    kw = {} if per_sample_t else {"t_vec": m.timestep_vector(t[:1])}
    m.fuse_io = True
    fused = m(z, t, idx, km, cond, T, **kw)
    m.fuse_io = False
    separate = m(z, t, idx, km, cond, T, **kw)
    assert torch.isfinite(fused).all()
    assert torch.equal(fused, separate)


def test_half_film_table_as_accurate_as_fp32_table(monkeypatch):
    """The folded FiLM table of the whole-encoder kernel stored as IEEE half [scale | shift] (default) against the fp32 table, both
    measured against the fp32 check mode: the half table does not add error (2^-12 rounding of the scale; a bf16 table measured
    1.2e-2 between the two tables)."""
    from interpolated_diffusion_b200.models import _engine as E
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    torch.manual_seed(31)
    gen = torch.Generator().manual_seed(32)
    B, T, D = 40, 64, 2
    m = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=2).cuda()
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
    args = (torch.rand((B, T, D), generator=gen).cuda(), torch.full((B,), 1, dtype=torch.long).cuda(), torch.rand((B, T, 2), generator=gen).cuda(), cond)
    m.precision = "fp32"
    ref = m(*args).clone()
    m.precision = "bf16"
    assert E.FILM_F16
    y16 = m(*args).clone()
    film = m.transformer.packed().film_params(m.encode_cond(cond), T, m.precision)
    assert film.t.dtype == torch.float16 and film.code() == 2
    monkeypatch.setattr(E, "FILM_F16", False)
    y32 = m(*args).clone()
    assert m.transformer.packed().film_params(m.encode_cond(cond), T, m.precision).code() == 1
    e16, e32 = _maxabs(y16, ref), _maxabs(y32, ref)
    scale = max(1.0, ref.abs().max().item())
    assert e32 < 2e-2 * scale and e16 < 2e-2 * scale, (e16, e32)
    assert e16 < e32 + 2e-3 * scale, (e16, e32)


@pytest.mark.parametrize("chan,use_sdf,B", [((32, 64, 128, 128), False, 37), ((32, 64, 64), True, 5), ((64,), False, 4100),
                                            ((64, 128), False, 9), ((32, 64, 128, 128), False, 8300)])
def test_deep_conv_stack_gemm_vs_oracle(chan, use_sdf, B):
    """The trainer-default conditioning encoder (maze_channels 32,64,128,128) and other depths (encoders.py:8-25): the tap-shifted
    implicit GEMM (default for stacks it takes; a batch larger than one chunk included) and the im2col + tcgen05 GEMM path."""
    from interpolated_diffusion_b200.models.encoders import MazeConditionEncoder
    gen = torch.Generator().manual_seed(9)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    if use_sdf:
        cond["sdf"] = torch.rand((B, 1, 21, 21), generator=gen)
    torch.manual_seed(2)
    m = MazeConditionEncoder(use_sdf=use_sdf, d_cond=128, maze_channels=chan)
    sd = {"cond_enc." + k: v.clone() for k, v in m.state_dict().items()}
    ref = odn.cond_encoder(sd, cond)
    m = m.cuda()
    got = m(_cuda(cond))
    assert _maxabs(got, ref) < 2e-2 * max(1.0, ref.abs().max().item()), _maxabs(got, ref)
    m.maze.use_implicit = False                  # im2col + GEMM form of the same stack
    got2 = m(_cuda(cond))
    assert _maxabs(got2, ref) < 2e-2 * max(1.0, ref.abs().max().item()), _maxabs(got2, ref)
    m.precision = "fp32"
    if B <= 64:
        got = m(_cuda(cond))
        assert _maxabs(got, ref) < 1e-4, _maxabs(got, ref)


@pytest.mark.parametrize("T", [17, 49, 100, 145])
def test_causal_denoiser_padded_to_fused_path(T):
    """A causal Stage-2 model at a length that does not divide 128 is right-padded to one that does and runs through the
    whole-encoder kernel: every real token equals the unpadded generic path (causal attention never looks right) and the oracle."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    B, D = 5, 2
    gen = torch.Generator().manual_seed(T)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(2)
    m = InterpLevelCausalDenoiser(data_dim=D, max_levels=3, mask_channels=1)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    m = m.cuda()
    x = torch.rand((B, T, D), generator=gen)
    mask = torch.rand((B, T), generator=gen) < 0.3
    s = torch.full((B,), 3)
    ref = odn.interp_level_denoiser(sd, 8, x, s, mask, cond, causal=True)
    got = m(x.cuda(), s.cuda(), mask.cuda(), _cuda(cond))
    assert got.shape == (B, T, D) and _maxabs(got, ref) < 2e-2, _maxabs(got, ref)
    m.pad_causal = False
    plain = m(x.cuda(), s.cuda(), mask.cuda(), _cuda(cond))
    assert _maxabs(got, plain) < 2e-2


def test_denoisers_large_model_vs_oracle():
    """The trainer-default ("large") widths -- d_model 384, 12 heads, ff 1536, maze 32-64-128-128 -- with 3 of the 12 layers:
    Stage-1 (K = 8) and Stage-2 (T = 64) single evaluations through the generic tcgen05 GEMM (pair mode, TMA-store epilogue),
    the packed block-diagonal attention kernel and the im2col + GEMM conv stack, bf16, against the CPU oracle."""
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    kw = dict(d_model=384, n_layers=3, n_heads=12, d_ff=1536, maze_channels=(32, 64, 128, 128))
    B, T, K, D = 40, 64, 8, 2
    gen = torch.Generator().manual_seed(15)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D, **kw)
    il = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=3, **kw)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    sd_il = {k: v.clone() for k, v in il.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    idx = torch.tensor([0, 9, 18, 27, 36, 45, 54, 63]).repeat(B, 1)
    z = torch.randn((B, K, D), generator=gen)
    km = torch.zeros((B, K, D), dtype=torch.bool)
    km[:, 0] = km[:, -1] = True
    t = torch.full((B,), 500, dtype=torch.long)
    ref = odn.keypoint_denoiser(sd_kp, 12, z, t, idx, km, cond, T)
    got = kp(z.cuda(), t.cuda(), idx.cuda(), km.cuda(), _cuda(cond), T)
    assert _maxabs(got, ref) < 2e-2, _maxabs(got, ref)
    x = torch.rand((B, T, D), generator=gen)
    mask_in = torch.rand((B, T, 3), generator=gen)
    s = torch.randint(1, 4, (B,), generator=gen)
    ref = odn.interp_level_denoiser(sd_il, 12, x, s, mask_in, cond)
    got = il(x.cuda(), s.cuda(), mask_in.cuda(), _cuda(cond))
    assert _maxabs(got, ref) < 2e-2, _maxabs(got, ref)
