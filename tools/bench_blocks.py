"""Dev: time the fused transformer-block kernels alone (IDB200_DBG ablations are read once per process)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
which = sys.argv[2] if len(sys.argv) > 2 else "both"
Ls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [8, 64]
d, ff, H = 256, 1024, 8
dev = "cuda"
for L in Ls:
    M = B * L
    h = torch.randn((M, d), device=dev)
    w = torch.ones(d, device=dev); bb = torch.zeros(d, device=dev); gb = torch.randn((B, 2 * d), device=dev) * 0.1
    Wq = (torch.randn((3 * d, d), device=dev) * 0.06).bfloat16(); bq = torch.zeros(3 * d, device=dev)
    Wo = (torch.randn((d, d), device=dev) * 0.06).bfloat16(); b2 = torch.zeros(d, device=dev)
    W1 = (torch.randn((ff, d), device=dev) * 0.06).bfloat16(); W2 = (torch.randn((d, ff), device=dev) * 0.03).bfloat16()
    b1 = torch.zeros(ff, device=dev)
    fns = {"attn_block": lambda: E.attn_block(h, w, bb, gb, Wq, bq, Wo, b2, L, H, False),
           "mlp_block": lambda: E.mlp_block(h, w, bb, gb, W1, b1, W2, b2, L)}
    for name, fn in fns.items():
        if which not in ("both", name): continue
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = (2.0 * d * 3 * d + 2.0 * d * d + 4.0 * L * d) if name == "attn_block" else 4.0 * d * ff
        print(f"{name} L={L} M={M} dbg={os.environ.get('IDB200_DBG','0')} ms={ms:.4f} TF/s={fl*M/ms/1e9:.0f} us_per_tile={ms*1e3/((M/128)/148):.2f}")
    del h
