"""Dev: replay the whole generation path many times and check the samples are bit-identical (races show up as flips)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
from interpolated_diffusion_b200.sample.sample_generate import GenerationConfig, GenerationGraph
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384 + 333
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 30
torch.manual_seed(0)
large = len(sys.argv) > 3 and sys.argv[3] == "large"       # trainer-default models: the pair-mode block kernels (qkv_attention, ln_mlp_pair)
kw = dict(d_model=384, n_layers=12, n_heads=12, d_ff=1536, maze_channels=(32, 64, 128, 128)) if large else {}
kp = KeypointDenoiser(data_dim=2, **kw).cuda()
il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2, **kw).cuda()
cfg = GenerationConfig()
gen = torch.Generator().manual_seed(1)
cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
z = torch.randn((B, cfg.K_min, 2), generator=gen).cuda()
g = GenerationGraph(kp, il, B, cfg)
ref = g.run(cond, z).clone()
bad = 0
for i in range(reps):
    out = g.run(cond, z)
    if not torch.equal(out, ref):
        bad += 1
        d = (out != ref).any(dim=2).any(dim=1)
        print(f"rep {i}: {int(d.sum())} trajectories differ, first {int(d.nonzero()[0])}, max |diff| {float((out - ref).abs().max()):.3e}", flush=True)
print(f"B={B} reps={reps} mismatching replays: {bad}; finite: {bool(torch.isfinite(ref).all())}")
