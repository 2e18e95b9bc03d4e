import os, sys
sys.path.insert(0, "/root/repo")
import torch
from interpolated_diffusion_b200.models import _engine as E
dev="cuda"; d=384; M=131072*2
for ff in (128, 256, 512, 1024, 1536, 2048):
    a = torch.randn((M, d), device=dev).bfloat16()
    w1 = (torch.randn((ff, d), device=dev) / d ** 0.5).bfloat16(); b1 = torch.randn((ff,), device=dev) * 0.1
    w2 = (torch.randn((d, ff), device=dev) / ff ** 0.5).bfloat16(); b2 = torch.randn((d,), device=dev) * 0.1
    h = torch.randn((M, d), device=dev)
    w2p = w2[E.mlp_pair_w2_order(d, dev)].contiguous()
    for _ in range(2): E.mlp_pair(a, w1, b1, w2p, b2, h)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): E.mlp_pair(a, w1, b1, w2p, b2, h)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tiles_per_sm = M / 128 / 148
    print(f"ff={ff}: {ms:.3f} ms  {ms*1e3/tiles_per_sm:.2f} us per tile  ({ms*1e3/tiles_per_sm*1.9e3/ (ff/64):.0f} cycles per chunk incl. fixed)", flush=True)
