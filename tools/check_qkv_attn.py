"""Dev: idb200_qkv_attention against the two-launch path (token GEMM + idb200_attention), and their timings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E

torch.manual_seed(0)
dev = "cuda"


def pack(w, b, d):
    order = torch.cat([torch.arange(64) + part * d + g * 64 for g in range(d // 64) for part in range(3)]).to(w.device)
    return w[order].contiguous(), b[order].contiguous()


for (B, Lq, d, causal) in [(40, 8, 384, False), (33, 64, 384, False), (16, 64, 256, True), (24, 16, 384, True), (7, 128, 384, False), (16384, 8, 384, False), (16384, 64, 384, False)]:
    H = d // 32
    M = B * Lq
    a = torch.randn((M, d), device=dev).bfloat16()
    w = (torch.randn((3 * d, d), device=dev) / d ** 0.5).bfloat16()
    b = torch.randn((3 * d,), device=dev) * 0.1
    wg, bg = pack(w, b, d)
    qkv = torch.empty((M, 3 * d), device=dev, dtype=torch.bfloat16)
    ref = torch.empty((M, d), device=dev, dtype=torch.bfloat16)
    out = torch.empty((M, d), device=dev, dtype=torch.bfloat16)

    def unfused():
        E.gemm_bf16(a, w, b, qkv, E.EPI_BF16)
        E.attention(qkv, ref, B, Lq, H, causal)

    def fused():
        E.qkv_attention(a, wg, bg, out, Lq, H, causal)

    unfused(); fused()
    torch.cuda.synchronize()
    err = float((out.float() - ref.float()).abs().max())
    mag = float(ref.float().abs().max())
    line = f"B={B} L={Lq} d={d} causal={int(causal)}: max|diff|={err:.4f} (max|ref|={mag:.2f})"
    if M >= 100000:
        for name, fn in (("unfused", unfused), ("fused", fused)):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10): fn()
            e1.record(); torch.cuda.synchronize()
            line += f"  {name} {e0.elapsed_time(e1) / 10:.3f} ms"
    print(line, flush=True)
    assert err <= 0.03 * max(1.0, mag), "mismatch"
print("ok")
