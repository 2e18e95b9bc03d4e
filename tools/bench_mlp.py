"""Dev: time / profile the fused MLP kernel alone."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E
M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
d, ff = 256, 1024
a = torch.randn((M, d), device="cuda").bfloat16(); h = torch.randn((M, d), device="cuda")
W1 = (torch.randn((ff, d), device="cuda") * 0.06).bfloat16(); W2 = (torch.randn((d, ff), device="cuda") * 0.03).bfloat16()
b1 = torch.zeros(ff, device="cuda"); b2 = torch.zeros(d, device="cuda")
for _ in range(3): E.mlp_fused(a, W1, b1, W2, b2, h)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): E.mlp_fused(a, W1, b1, W2, b2, h)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"M={M} ms={ms:.4f} TFLOPs={4.0*M*d*ff/ms/1e9:.1f} us_per_tile_per_sm={ms*1e3/((M/128)/148):.1f}")
