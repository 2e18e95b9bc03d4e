"""One warm-up + two launches of ONE hot kernel at its bench shape, for `ncu --set full -k regex:<kernel>`:
    python tools/profile_targets.py <target>
targets: qkv_attn mlp_pair encoder_L8 encoder_L64 attn_L8 attn_L64 attn_L256 conv_tap ln_film ln_bwd attn_bwd im2col colsum embed sgemm interp_T256 interp_T64 gemm_qkv384 gemm_silu_dual gemm_dsilu dw_grouped corrupt_adj"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from interpolated_diffusion_b200 import _lib as L  # noqa: E402
from interpolated_diffusion_b200.models import _engine as E  # noqa: E402

t = sys.argv[1]
torch.manual_seed(0)
dev = "cuda"


def run(fn, n=3):
    for _ in range(n):
        fn()
    torch.cuda.synchronize()


if t in ("encoder_L8", "encoder_L64"):
    from interpolated_diffusion_b200.models.transformer import TransformerEncoder
    Lq = 8 if t == "encoder_L8" else 64
    B = 65536
    enc = TransformerEncoder(d_model=256, n_layers=8, n_heads=8, d_ff=1024, cond_dim=128).cuda()
    pk = enc.packed()
    h = torch.randn((B * Lq, 256), device=dev)
    film = pk.film_params(torch.randn((B, 128), device=dev), Lq)
    run(lambda: pk.forward(h, B, Lq, film))
elif t in ("attn_L8", "attn_L64", "attn_L256"):
    B, Lq, H, causal = {"attn_L8": (65536, 8, 12, 0), "attn_L64": (16384, 64, 12, 0), "attn_L256": (8192, 256, 8, 1)}[t]
    d = H * 32
    qkv = torch.randn((B * Lq, 3 * d), device=dev).bfloat16()
    out = torch.empty((B * Lq, d), device=dev, dtype=torch.bfloat16)
    run(lambda: E.attention(qkv, out, B, Lq, H, bool(causal)))
elif t == "conv_tap":
    from interpolated_diffusion_b200.models.encoders import MazeEncoder
    m = MazeEncoder(1, 128, channels=(32, 64, 128, 128)).cuda()
    x = (torch.rand((8192, 1, 21, 21), device=dev) < 0.2).float()
    run(lambda: m(x))
elif t == "ln_film":
    M, d, Lq = 16384 * 64, 384, 64
    h = torch.randn((M, d), device=dev)
    w, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
    gb = torch.randn((M // Lq, 2 * d), device=dev)
    out = torch.empty((M, d), device=dev, dtype=torch.bfloat16)
    run(lambda: E.ln_film(h, w, b, gb, out, Lq))
elif t in ("embed", "sgemm"):
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2).cuda()
    B = 65536
    x = torch.rand((B, 64, 2), device=dev)
    mask = torch.rand((B, 64, 2), device=dev)
    cond = {"occ": (torch.rand((B, 1, 21, 21), device=dev) < 0.2).float(), "start_goal": torch.rand((B, 4), device=dev)}
    s = torch.full((B,), 3, device=dev, dtype=torch.long)
    run(lambda: il(x, s, mask, cond), n=2)
elif t in ("interp_T256", "interp_T64"):
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    T, K, S, B = (256, 32, 4, 1 << 18) if t == "interp_T256" else (64, 8, 3, 1 << 20)
    sc = torch.rand((B, T - 2), device=dev)
    x0 = torch.rand((B, T, 4), device=dev)
    KL = kf._compute_k_schedule(T, K, S)
    run(lambda: kf.nested_masks_interp(sc, T, KL, x0=x0, levels_out=(1, S), want_idx=False))
elif t == "gemm_qkv384":
    M = 16384 * 64
    A = torch.randn((M, 384), device=dev).bfloat16()
    W = torch.randn((1152, 384), device=dev).bfloat16()
    bias = torch.randn((1152,), device=dev)
    out = torch.empty((M, 1152), device=dev, dtype=torch.bfloat16)
    run(lambda: E.gemm_bf16(A, W, bias, out, E.EPI_BF16))
elif t in ("gemm_silu_dual", "gemm_dsilu"):
    # the training step's ff.0 forward (u and SiLU(u) from one launch) / dU = (dY W2) . SiLU'(u) with per-warp column sums
    M, N, K = 4096 * 64, 1536, 384
    A = torch.randn((M, K), device=dev).bfloat16()
    W = (torch.randn((N, K), device=dev) / K ** 0.5).bfloat16()
    bias = torch.randn((N,), device=dev)
    u = torch.randn((M, N), device=dev).bfloat16()
    out = torch.empty((M, N), device=dev, dtype=torch.bfloat16)
    part = torch.empty((8 * ((M + 255) // 256), N), device=dev)
    st = L.stream(torch.device(dev))
    if t == "gemm_silu_dual":
        run(lambda: L.call("idb200_gemm_bf16_aux", A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), u.data_ptr(), M, N, K, 4, st))
    else:
        run(lambda: L.call("idb200_gemm_bf16_dsilu_sums", A.data_ptr(), W.data_ptr(), out.data_ptr(), u.data_ptr(), part.data_ptr(), M, N, K, st))
elif t == "dw_grouped":
    # weight gradients of all 12 layers as one split-K launch over the stacked token axis (cfg-4 per-GPU shape: M = 32 768 per layer)
    nl, M, n_out, k_in = 12, 32768, 384, 1536
    dy = torch.randn((nl, M, n_out), device=dev).bfloat16()
    x = torch.randn((nl, M, k_in), device=dev).bfloat16()
    part = torch.empty((nl, n_out, k_in), device=dev)
    run(lambda: L.call("idb200_gemm_bf16_nn_splitk", dy.data_ptr(), x.data_ptr(), part.data_ptr(), n_out, k_in, nl * M, nl, L.stream(torch.device(dev))))
elif t == "corrupt_adj":
    from interpolated_diffusion_b200.corruptions import keyframes as kf
    from interpolated_diffusion_b200.train import train_interp_levels as tr
    B, T = 1 << 18, 64
    x0 = torch.rand((B, T, 2), device=dev)
    masks, _ = kf.build_nested_masks_batch(B, T, 8, 3, device=dev)
    s_idx = torch.randint(1, 4, (B,), device=dev)
    kw = dict(corrupt_mode="dist", corrupt_sigma_max=0.08, corrupt_sigma_min=0.012, corrupt_sigma_pow=0.75, corrupt_anchor_frac=0.25)
    run(lambda: tr.corrupt_adjacent_fused(x0, masks, s_idx, [64, 32, 16, 8], 8, seed=1, offset=1, **kw))
elif t == "qkv_attn":
    B, Lq, d = 16384, 8, 384
    M = B * Lq
    a = torch.randn((M, d), device=dev).bfloat16()
    w = (torch.randn((3 * d, d), device=dev) / d ** 0.5).bfloat16()
    b = torch.randn((3 * d,), device=dev) * 0.1
    order = torch.cat([torch.arange(64) + part * d + g * 64 for g in range(d // 64) for part in range(3)]).to(dev)
    wg, bg = w[order].contiguous(), b[order].contiguous()
    out = torch.empty_like(a)
    run(lambda: E.qkv_attention(a, wg, bg, out, Lq, d // 32, False))
elif t == "mlp_pair":
    M, d, ff = 16384 * 8, 384, 1536
    a = torch.randn((M, d), device=dev).bfloat16()
    w1 = (torch.randn((ff, d), device=dev) / d ** 0.5).bfloat16()
    b1 = torch.randn((ff,), device=dev) * 0.1
    w2 = (torch.randn((d, ff), device=dev) / ff ** 0.5).bfloat16()
    b2 = torch.randn((d,), device=dev) * 0.1
    h = torch.randn((M, d), device=dev)
    w2p = w2[E.mlp_pair_w2_order(d, dev)].contiguous()
    run(lambda: E.mlp_pair(a, w1, b1, w2p, b2, h))
elif t == "ln_bwd":
    B, Lq, d = 4096, 64, 384
    M = B * Lq
    h = torch.randn((M, d), device=dev)
    da = torch.randn((M, d), device=dev).bfloat16()
    dh = torch.randn((M, d), device=dev)
    dh16 = torch.empty((M, d), device=dev, dtype=torch.bfloat16)
    w, b = torch.randn(d, device=dev), torch.randn(d, device=dev)
    gb = torch.randn((B, 2 * d), device=dev)
    dgb = torch.empty((B, 2 * d), device=dev)
    dwb = torch.empty((B, 3 * d), device=dev)
    stats = torch.empty((M, 4), device=dev)
    run(lambda: L.call("idb200_ln_film_bwd2", da.data_ptr(), 1, h.data_ptr(), w.data_ptr(), b.data_ptr(), gb.data_ptr(), 2 * d, B, Lq, d,
                       dh.data_ptr(), dh16.data_ptr(), dgb.data_ptr(), 2 * d, dwb.data_ptr(), 1, stats.data_ptr(), L.stream(torch.device(dev))))
elif t == "attn_bwd":
    B, Lq, H = 4096, 64, 12
    d = 32 * H
    qkv = torch.randn((B * Lq, 3 * d), device=dev).bfloat16()
    dO = torch.randn((B * Lq, d), device=dev).bfloat16()
    dqkv = torch.empty_like(qkv)
    sums = torch.empty((B, 3 * d), device=dev)
    run(lambda: L.call("idb200_attention_bwd_sums", qkv.data_ptr(), dO.data_ptr(), dqkv.data_ptr(), sums.data_ptr(), B, Lq, H, 0, L.stream(torch.device(dev))))
elif t == "im2col":
    B, C = 4096, 128
    u = torch.randn((B, 441, C), device=dev).bfloat16()
    col = torch.empty((B * 441, 9 * C), device=dev, dtype=torch.bfloat16)
    run(lambda: E.im2col3x3(u, B, 21, 21, C, True, col))
elif t == "colsum":
    from interpolated_diffusion_b200.train import backward as BW
    x = torch.randn((262144, 1536), device=dev).bfloat16()
    out = torch.empty((1536,), device=dev)
    sc = BW._Scratch()
    run(lambda: sc.colsum(x, out))
else:
    raise SystemExit(f"unknown target {t}")
print("ok", t)
