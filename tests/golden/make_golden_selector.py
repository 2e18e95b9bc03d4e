"""Golden vectors for the KeypointSelector from the LIVE reference (run in the build container, where /root/reference exists):
    python tests/golden/make_golden_selector.py
Two small random-init configurations: the defaults (start/goal maps + sg token) and every optional input the CUDA mirror
supports (sdf channel, goal-distance token, level embedding, three conv layers)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from src.models.keypoint_selector import KeypointSelector, select_topk_indices  # noqa: E402

out = {}
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
for name, kw in CFG.items():
    torch.manual_seed(11)
    m = KeypointSelector(**kw).eval()
    with torch.no_grad():
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(0.05 * torch.randn_like(p))
    g = torch.Generator().manual_seed(5)
    B = 5
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=g)}
    if kw.get("use_sdf"):
        cond["sdf"] = torch.rand((B, 1, 21, 21), generator=g)
    if kw.get("use_level"):
        cond["level"] = torch.rand((B, 1), generator=g)
    with torch.no_grad():
        logits = m(cond)
    for k, v in m.state_dict().items():
        out[f"{name}/sd/{k}"] = v.numpy()
    for k, v in cond.items():
        out[f"{name}/cond/{k}"] = v.numpy()
    out[f"{name}/logits"] = logits.numpy()
    out[f"{name}/idx_k8"] = select_topk_indices(logits, 8).numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "selector.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith(("logits", "idx_k8"))})
