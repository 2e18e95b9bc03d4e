"""Dev micro-benchmark of the individual denoiser kernels at the bench shapes (B = 65536 small model)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
d, ff, H = 256, 1024, 8
dev = "cuda"
res = {}

def timeit(name, fn, it=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    res[name] = round(e0.elapsed_time(e1) / it, 4)

for L in (8, 64):
    M = B * L
    a = torch.randn((M, d), device=dev).bfloat16()
    h = torch.randn((M, d), device=dev)
    W1 = (torch.randn((ff, d), device=dev) * 0.06).bfloat16(); W2 = (torch.randn((d, ff), device=dev) * 0.03).bfloat16()
    b1 = torch.zeros(ff, device=dev); b2 = torch.zeros(d, device=dev)
    timeit(f"mlp_fused M={M}", lambda: E.mlp_fused(a, W1, b1, W2, b2, h))
    f = torch.empty((M, ff), device=dev, dtype=torch.bfloat16)
    timeit(f"ff1+ff2 unfused M={M}", lambda: (E.gemm_bf16(a, W1, b1, f, 1), E.gemm_bf16(f, W2, b2, h, 2)))
    del f
    qkv = torch.randn((M, 3 * d), device=dev).bfloat16()
    o = torch.empty((M, d), device=dev, dtype=torch.bfloat16)
    timeit(f"attention L={L} M={M}", lambda: E.attention(qkv, o, B, L, H, False))
    if L == 8:
        timeit(f"attention simt L={L} M={M}", lambda: E.attention(qkv, o, B, L, H, False, force_simt=True), it=2)
    w = torch.ones(d, device=dev); bb = torch.zeros(d, device=dev); gb = torch.randn((B, 2 * d), device=dev)
    timeit(f"ln_film M={M}", lambda: E.ln_film(h, w, bb, gb, a, L))
    Wq = (torch.randn((3 * d, d), device=dev) * 0.06).bfloat16(); bq = torch.zeros(3 * d, device=dev)
    timeit(f"qkv gemm M={M}", lambda: E.gemm_bf16(a, Wq, bq, qkv, 0))
    Wo = (torch.randn((d, d), device=dev) * 0.06).bfloat16()
    timeit(f"outproj gemm M={M}", lambda: E.gemm_bf16(a, Wo, b2, h, 2))
    timeit(f"attn_block L={L} M={M}", lambda: E.attn_block(h, w, bb, gb, Wq, bq, Wo, b2, L, H, False))
    timeit(f"mlp_block L={L} M={M}", lambda: E.mlp_block(h, w, bb, gb, W1, b1, W2, b2, L))
    del a, h, qkv, o
occ = (torch.rand((B, 1, 21, 21), device=dev) < 0.2).float()
ws = [torch.randn((32, 1, 3, 3), device=dev) * 0.3, torch.randn((64, 32, 3, 3), device=dev) * 0.06]
bs = [torch.zeros(32, device=dev), torch.zeros(64, device=dev)]
timeit("conv_encoder", lambda: E.conv_encoder(occ, None, ws, bs), it=2)
w1p = ws[1].permute(0, 2, 3, 1).reshape(64, 9 * 32).contiguous().bfloat16()
timeit("conv_encoder_tc", lambda: E.conv_encoder_tc(occ, None, ws[0], bs[0], w1p, bs[1]), it=2)
print(json.dumps(res, indent=1))
