"""Dev: idb200_mlp_pair against the two-GEMM path (and mlp_fused at d = 256), and their timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E

torch.manual_seed(0)
dev = "cuda"
for (M, d, ff) in [(640, 384, 1536), (128, 384, 128), (1000, 384, 384), (4096 + 64, 256, 1024), (131072, 384, 1536), (1048576, 384, 1536), (524288, 256, 1024)]:
    a = torch.randn((M, d), device=dev).bfloat16()
    w1 = (torch.randn((ff, d), device=dev) / d ** 0.5).bfloat16()
    b1 = torch.randn((ff,), device=dev) * 0.1
    w2 = (torch.randn((d, ff), device=dev) / ff ** 0.5).bfloat16()
    b2 = torch.randn((d,), device=dev) * 0.1
    h0 = torch.randn((M, d), device=dev)
    w2p = w2[E.mlp_pair_w2_order(d, dev)].contiguous()
    f = torch.empty((M, ff), device=dev, dtype=torch.bfloat16)
    href, hout = h0.clone(), h0.clone()

    def unfused(h=href):
        E.gemm_bf16(a, w1, b1, f, E.EPI_SILU_BF16)
        E.gemm_bf16(f, w2, b2, h, E.EPI_RESID_F32)

    def fused(h=hout):
        E.mlp_pair(a, w1, b1, w2p, b2, h)

    unfused(); fused()
    torch.cuda.synchronize()
    err = float((hout - href).abs().max())
    upd = float((href - h0).abs().max())
    line = f"M={M} d={d} ff={ff}: max|diff|={err:.4f} (max|update|={upd:.2f})"
    if M >= 100000:
        for name, fn in (("unfused", unfused), ("fused", fused)):
            for _ in range(2): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            line += f"  {name} {ms:.3f} ms ({4.0 * M * d * ff / ms / 1e9:.0f} TF/s)"
        if d == 256:
            hm = h0.clone()
            for _ in range(2): E.mlp_fused(a, w1, b1, w2, b2, hm)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): E.mlp_fused(a, w1, b1, w2, b2, hm)
            e1.record(); torch.cuda.synchronize()
            line += f"  mlp_fused(single CTA) {e0.elapsed_time(e1) / 10:.3f} ms"
    print(line, flush=True)
    assert err <= 0.03 * max(1.0, upd), "mismatch"
print("ok")
