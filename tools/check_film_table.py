"""Dev: error of the bf16-mode denoiser outputs against the fp32 check mode, with the folded FiLM table as IEEE half [scale | shift] and as fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from interpolated_diffusion_b200.models import _engine as E
from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
for seed in (1, 2, 3):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(100 + seed)
    B, T, K, D = 256, 64, 8, 2
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float().cuda(), "start_goal": torch.rand((B, 4), generator=gen).cuda()}
    il = InterpLevelDenoiser(data_dim=D, max_levels=3, mask_channels=2).cuda()
    kp = KeypointDenoiser(data_dim=D).cuda()
    a_il = (torch.rand((B, T, D), generator=gen).cuda(), torch.full((B,), 1, dtype=torch.long).cuda(), torch.rand((B, T, 2), generator=gen).cuda(), cond)
    idx = torch.sort(torch.stack([torch.randperm(T, generator=gen)[:K] for _ in range(B)]), dim=1).values.cuda()
    a_kp = (torch.randn((B, K, D), generator=gen).cuda(), torch.full((B,), 500, dtype=torch.long).cuda(), idx, (torch.rand((B, K, D), generator=gen) < 0.3).cuda(), cond, T)
    for name, m, a in (("interp", il, a_il), ("keypoints", kp, a_kp)):
        m.precision = "fp32"; ref = m(*a).clone(); m.precision = "bf16"
        E.FILM_F16 = True; y16 = m(*a).clone()
        E.FILM_F16 = False; y32 = m(*a).clone()
        E.FILM_F16 = True
        print(f"seed {seed} {name}: |ref| max {ref.abs().max():.3f}  err half-table {(y16-ref).abs().max():.5f}  err fp32-table {(y32-ref).abs().max():.5f}  table diff {(y16-y32).abs().max():.5f}", flush=True)
