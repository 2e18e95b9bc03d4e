"""Dev: time the conv encoders (mma.sync vs tcgen05) on B mazes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
g = torch.Generator(device="cuda").manual_seed(0)
occ = (torch.rand((B, 1, 21, 21), generator=g, device="cuda") < 0.2).float()
w0 = torch.randn((32, 1, 3, 3), generator=g, device="cuda") / 3; b0 = torch.randn(32, generator=g, device="cuda") * 0.1
w1 = torch.randn((64, 32, 3, 3), generator=g, device="cuda") / 17; b1 = torch.randn(64, generator=g, device="cuda") * 0.1
w1p = w1.permute(0, 2, 3, 1).reshape(64, -1).to(torch.bfloat16).contiguous()
for name, fn in (("tc (mma.sync)", E.conv_encoder_tc), ("tc5 (tcgen05)", E.conv_encoder_tc5)):
    for _ in range(2): out = fn(occ, None, w0, b0, w1p, b1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = fn(occ, None, w0, b0, w1p, b1)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{name}: {ms:.3f} ms for {B} mazes = {ms*1e6/B:.1f} ns/maze, {B*16.5e6/ms/1e9:.0f} TF/s", flush=True)
    if name.startswith("tc ("): ref = out
print("max diff tc5 vs tc:", (out - ref).abs().max().item())
