"""Golden vectors for the DP anchor placement from the LIVE reference (run in the build container, where /root/reference exists):
    python tests/golden/make_golden_dp.py
Cost matrices: upper-triangular random costs (inf elsewhere, like build_cost_matrix_from_segments_batch), one set quantised to a
few values so that ties exercise torch.argmin's first-minimum rule, one with a maximum segment length (more inf)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from src.selection.epiplexity_dp import dp_select_indices_batch  # noqa: E402

out = {}
g = torch.Generator().manual_seed(0)
for name, (B, T, K, mode) in {"t64_k8": (16, 64, 8, "rand"), "t64_k8_ties": (16, 64, 8, "ties"), "t256_k32_band": (4, 256, 32, "band"),
                              "t17_k8": (8, 17, 8, "rand"), "t12_k12": (3, 12, 12, "rand")}.items():
    C = torch.full((B, T, T), float("inf"))
    iu = torch.triu_indices(T, T, offset=1)
    vals = torch.rand((B, iu.shape[1]), generator=g)
    if mode == "ties":
        vals = torch.round(vals * 3) / 3
    C[:, iu[0], iu[1]] = vals * (iu[1] - iu[0]).float().pow(1.5)
    if mode == "band":
        C[:, iu[0][(iu[1] - iu[0]) > 24], iu[1][(iu[1] - iu[0]) > 24]] = float("inf")
    out[name + "/C"] = C.numpy()
    out[name + "/K"] = np.int64(K)
    out[name + "/idx"] = dp_select_indices_batch(C, K).numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "dp_select.npz"), **out)
print({k: v.shape for k, v in out.items() if k.endswith("idx")})
