"""DP anchor placement kernel (selection/epiplexity_dp.py mirror) against golden indices from the live reference and the numpy
oracle: bit-exact integer output, including ties and unreachable-segment (inf) patterns, and the reference's error behaviour."""
import os

import numpy as np
import pytest
import torch

from oracle import selection_np as osel

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dp_select.npz")


def test_dp_select_matches_reference_golden():
    from interpolated_diffusion_b200.selection.epiplexity_dp import dp_select_indices, dp_select_indices_batch
    g = np.load(GOLD)
    for n in sorted({k.split("/")[0] for k in g.files}):
        C = torch.from_numpy(g[n + "/C"]).cuda()
        idx = dp_select_indices_batch(C, int(g[n + "/K"]))
        assert idx.dtype == torch.long and np.array_equal(idx.cpu().numpy(), g[n + "/idx"]), n
        assert np.array_equal(dp_select_indices(C[0], int(g[n + "/K"])).cpu().numpy(), g[n + "/idx"][0])


def test_dp_select_large_batch_vs_oracle_and_errors():
    from interpolated_diffusion_b200.selection.epiplexity_dp import dp_select_indices_batch
    gen = torch.Generator().manual_seed(3)
    B, T, K = 4096, 64, 8
    C = torch.full((B, T, T), float("inf"))
    iu = torch.triu_indices(T, T, offset=1)
    C[:, iu[0], iu[1]] = torch.round(torch.rand((B, iu.shape[1]), generator=gen) * 7) / 7 + 0.01 * (iu[1] - iu[0]).float()
    idx = dp_select_indices_batch(C.cuda(), K).cpu().numpy()
    assert np.array_equal(idx, osel.dp_select_indices_batch(C.numpy(), K))
    # K > T clamps to T (every index selected)
    assert np.array_equal(dp_select_indices_batch(C[:2, :10, :10].contiguous().cuda(), 50).cpu().numpy(), np.tile(np.arange(10), (2, 1)))
    with pytest.raises(ValueError):
        dp_select_indices_batch(C[:1].cuda(), 1)
    with pytest.raises(ValueError):
        dp_select_indices_batch(C[0].cuda(), 4)
    bad = C[:3].clone()
    bad[1, :, T - 1] = float("inf")                               # sample 1 cannot reach T-1
    with pytest.raises(RuntimeError, match="valid path"):
        dp_select_indices_batch(bad.cuda(), K)
