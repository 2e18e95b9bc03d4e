"""Dev: the trainer-default ("large") models through the same generation path (generic per-op kernels: d_model = 384 is outside
the whole-encoder kernel's specialisation)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
kw = dict(d_model=384, n_layers=12, n_heads=12, d_ff=1536, maze_channels=(32, 64, 128, 128))
torch.manual_seed(0)
kp = KeypointDenoiser(data_dim=2, **kw).cuda()
il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2, **kw).cuda()
cfg = GenerationConfig()
gen = torch.Generator().manual_seed(1)
cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
z = torch.randn((B, cfg.K_min, 2), generator=gen).cuda()
g = GenerationGraph(kp, il, B, cfg)
out = g.run(cond, z)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): g.run(cond, z)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"large models: B={B} {ms:.1f} ms per generation = {B/ms*1e3:.0f} trajectories/s, {B*9.58e9/ms/1e9:.0f} TF/s; finite={bool(torch.isfinite(out).all())}")
