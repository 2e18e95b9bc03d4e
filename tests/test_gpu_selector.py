"""KeypointSelector mirror (models/keypoint_selector.py: conv stack, cross attention over the spatial memory, token GEMMs on
tcgen05) against logits of the live reference (tests/golden/selector.npz) and the CPU oracle: bf16 path, 2e-2 on the logits."""
import os

import numpy as np
import pytest
import torch

from oracle import selector_torch as osel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "selector.npz")
CFG = {
    "default": dict(T=64, d_model=64, n_heads=2, d_ff=128, n_layers=2, pos_dim=64, maze_channels=(32, 64)),
    "full": dict(T=48, d_model=64, n_heads=2, d_ff=128, n_layers=2, pos_dim=32, use_sdf=True, use_goal_dist_token=True, use_level=True,
                 sg_map_sigma=2.0, maze_channels=(32, 32, 64)),
    # round 2: query bias from the mean memory token + one-hot start / goal maps (:101-111, :129-139, :170-175)
    "cbmem": dict(T=32, d_model=64, n_heads=2, d_ff=128, n_layers=1, pos_dim=32, use_cond_bias=True, cond_bias_mode="memory",
                  sg_map_sigma=0.0, use_goal_dist_token=True, maze_channels=(32, 64)),
    # query bias from a MazeConditionEncoder of its own
    "cbenc": dict(T=32, d_model=64, n_heads=2, d_ff=128, n_layers=1, pos_dim=32, use_cond_bias=True, cond_bias_mode="encoder",
                  maze_channels=(32, 64)),
}


@pytest.mark.parametrize("name", ["default", "full", "cbmem", "cbenc"])
def test_selector_matches_reference_golden(name):
    from interpolated_diffusion_b200.models.keypoint_selector import KeypointSelector, select_topk_indices
    g = np.load(GOLD)
    m = KeypointSelector(**CFG[name])
    sd = {k[len(name) + 4:]: torch.from_numpy(g[k]) for k in g.files if k.startswith(name + "/sd/")}
    m.load_state_dict(sd)                                             # same parameter tree as the reference
    m = m.cuda()
    cond = {k[len(name) + 6:]: torch.from_numpy(g[k]).cuda() for k in g.files if k.startswith(name + "/cond/")}
    logits = m(cond)
    ref = g[name + "/logits"]
    assert logits.shape == ref.shape
    assert np.abs(logits.cpu().numpy() - ref).max() < 2e-2 * max(1.0, np.abs(ref).max())
    assert select_topk_indices(logits, 8).shape == (ref.shape[0], 8)


def test_selector_full_size_vs_oracle_and_cross_attention():
    """Reference-default size (d_model 256, 8 heads, 21 x 21 maze -> 441 + 1 memory tokens) against the CPU oracle; the cross
    attention kernel alone against torch on bf16-rounded inputs."""
    from interpolated_diffusion_b200 import _lib as L
    from interpolated_diffusion_b200.models.keypoint_selector import KeypointSelector
    torch.manual_seed(3)
    m = KeypointSelector(T=64)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    gen = torch.Generator().manual_seed(8)
    B = 7
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    ref = osel.keypoint_selector(sd, cond, T=64, n_heads=8, pos_dim=64)
    got = m.cuda()({k: v.cuda() for k, v in cond.items()})
    assert (got.cpu() - ref).abs().max().item() < 2e-2 * max(1.0, ref.abs().max().item())
    # kernel alone
    Lq, La, Lb, H = 32, 441, 2, 4
    d = 32 * H
    q = torch.randn((B, Lq, d), generator=gen).bfloat16()
    ka = torch.randn((B, La, 2 * d), generator=gen).bfloat16()
    kb = torch.randn((B, Lb, 2 * d), generator=gen).bfloat16()
    kv = torch.cat([ka, kb], dim=1).float()
    sh = lambda t, n: t.view(B, n, H, 32).transpose(1, 2)
    p = torch.softmax(sh(q.float(), Lq) @ sh(kv[..., :d].contiguous(), La + Lb).transpose(-1, -2) / 32 ** 0.5, -1)
    o_ref = (p @ sh(kv[..., d:].contiguous(), La + Lb)).transpose(1, 2).reshape(B, Lq, d)
    qd, kad, kbd = q.cuda(), ka.cuda(), kb.cuda()
    out = torch.empty((B, Lq, d), device="cuda", dtype=torch.bfloat16)
    L.call("idb200_cross_attention", qd.data_ptr(), kad.data_ptr(), kbd.data_ptr(), out.data_ptr(), B, Lq, La, Lb, H, L.stream(out.device))
    assert (out.float().cpu() - o_ref).abs().max().item() < 2e-2
    with pytest.raises(RuntimeError, match="multiple of 16"):
        L.call("idb200_cross_attention", qd.data_ptr(), kad.data_ptr(), kbd.data_ptr(), out.data_ptr(), B, 24, La, Lb, H, L.stream(out.device))
