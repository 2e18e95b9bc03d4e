"""Dev: small invocations of the hot kernels for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, generate
from interpolated_diffusion_b200.eval.metrics import compute_metrics_batch
torch.manual_seed(0)
kp = KeypointDenoiser(data_dim=2, n_layers=2).cuda()
il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2, n_layers=2).cuda()
B = 37
gen = torch.Generator().manual_seed(1)
cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
cfg = GenerationConfig(ddim_steps=4)
x = generate(kp, il, cond, cfg, z_T=torch.randn((B, cfg.K_min, 2), generator=gen).cuda())
m = compute_metrics_batch(cond["occ"][:, 0], x, cond["start_goal"][:, 2:])
torch.cuda.synchronize()
print("ok", float(x.abs().max()), float(m["path_length"].mean()))
