"""Golden vectors for the DP-placement bookkeeping and the segment-cost predictor from the LIVE reference:
    python tests/golden/make_golden_segments.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from src.models.segment_cost import SegmentCostPredictor  # noqa: E402
from src.selection import epiplexity_dp as R  # noqa: E402

out = {}
g = torch.Generator().manual_seed(3)
for T, n in ((16, 1), (16, 4), (33, 3)):
    pc = R.build_segment_precompute(T, n, torch.device("cpu"))
    tag = f"pc_T{T}_n{n}"
    for k in ("seg_i", "seg_j", "seg_len", "t_idx", "alpha", "weight", "seg_id"):
        out[f"{tag}/{k}"] = getattr(pc, k).numpy()
    x = torch.rand((6, T, 4), generator=g).cumsum(1) / T
    out[f"{tag}/x"] = x.numpy()
    out[f"{tag}/cost"] = R.compute_segment_costs_batch(x, pc, 1.0).numpy()
    out[f"{tag}/cost_scaled"] = R.compute_segment_costs_batch(x, pc, 2.5).numpy()
    out[f"{tag}/feat"] = R.build_segment_features(T, pc.seg_i, pc.seg_j).numpy()
    C = R.build_cost_matrix_from_segments_batch(out[f"{tag}/cost"].__class__ and torch.from_numpy(out[f"{tag}/cost"]), pc, T)
    out[f"{tag}/C"] = C.numpy()
idx = torch.tensor([[0, 3, 7, 12, 15], [0, 1, 2, 9, 15]])
out["kp_feat"] = R.build_kp_feat_batch(idx, 16).numpy()
out["seg_feat_idx3"] = R.build_segment_features_from_idx(idx, 16, 3).numpy()
out["seg_feat_idx5"] = R.build_segment_features_from_idx(idx, 16, 5).numpy()
snr, w = R.build_snr_weights("cosine", 1000, 0.01, 100.0, 0.5)
out["snr_w"] = w.numpy()
out["ts_log_snr"] = R.sample_timesteps_log_snr(snr, 12).numpy()
torch.manual_seed(21)
m = SegmentCostPredictor(hidden_dim=128, n_layers=3).eval()
B, T = 4, 16
pc = R.build_segment_precompute(T, 1, torch.device("cpu"))
seg_feat = R.build_segment_features(T, pc.seg_i, pc.seg_j)
cond = {"occ": (torch.rand((B, 1, 21, 21), generator=g) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=g)}
with torch.no_grad():
    out["dphi/pred_shared"] = m(cond, seg_feat).numpy()
    sf3 = seg_feat.unsqueeze(0).expand(B, -1, -1) + 0.01 * torch.rand((B, seg_feat.shape[0], 3), generator=g)
    out["dphi/pred_batched"] = m(cond, sf3).numpy()
out["dphi/seg_feat"] = seg_feat.numpy()
out["dphi/seg_feat_b"] = sf3.numpy()
for k, v in m.state_dict().items():
    out[f"dphi/sd/{k}"] = v.numpy()
for k, v in cond.items():
    out[f"dphi/cond/{k}"] = v.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "segments.npz"), **out)
print(len(out), "arrays")
