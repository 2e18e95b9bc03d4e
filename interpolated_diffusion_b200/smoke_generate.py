"""One tiny invocation of the generation hot path on cuda:0, checked against the CPU oracle (used by
__graft_entry__.smoke(); the oracle import lives here only for that purpose)."""
import numpy as np
import torch


def run():
    from oracle import denoiser_torch as odn
    from .models.denoiser_interp_levels import InterpLevelDenoiser
    from .models.denoiser_keypoints import KeypointDenoiser
    from .sample.sample_generate import GenerationConfig, generate

    B, T, K, S, D = 8, 64, 8, 3, 2
    gen = torch.Generator().manual_seed(3)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=D)
    il = InterpLevelDenoiser(data_dim=D, max_levels=S, mask_channels=2)
    sd_kp = {k: v.clone() for k, v in kp.state_dict().items()}
    kp, il = kp.cuda(), il.cuda()
    ccond = {k: v.cuda() for k, v in cond.items()}
    out = generate(kp, il, ccond, GenerationConfig(), z_T=torch.randn((B, K, D), generator=gen).cuda(), return_all=True)
    assert torch.isfinite(out["x_hat"]).all()
    assert torch.equal(out["x_hat"][:, [0, -1], :2], out["x_pred"][:, [0, -1], :2]), "endpoint clamp"
    z = torch.randn((B, K, D), generator=gen)
    t = torch.full((B,), 500)
    km = torch.zeros((B, K, D), dtype=torch.bool)
    km[:, 0] = km[:, -1] = True
    ref = odn.keypoint_denoiser(sd_kp, 8, z, t, out["idx"].cpu(), km, cond, T)
    eps = kp(z.cuda(), t.cuda(), out["idx"], km.cuda(), ccond, T)
    err = (eps.cpu() - ref).abs().max().item()
    assert err < 2e-2, f"denoiser parity {err}"
