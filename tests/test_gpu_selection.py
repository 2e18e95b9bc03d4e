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


def test_segment_costs_and_dphi_match_reference_golden():
    """compute_segment_costs_batch (kernel) and SegmentCostPredictor (d_phi) against the live reference's outputs; then the whole
    DP placement pipeline (costs -> matrix -> dp_select) against the oracle."""
    from interpolated_diffusion_b200.models.segment_cost import SegmentCostPredictor
    from interpolated_diffusion_b200.selection import epiplexity_dp as S
    g = np.load(os.path.join(os.path.dirname(GOLD), "segments.npz"))
    dev = torch.device("cuda", 0)
    for T, n in ((16, 1), (16, 4), (33, 3)):
        tag = f"pc_T{T}_n{n}"
        pc = S.build_segment_precompute(T, n, dev)
        x = torch.from_numpy(g[f"{tag}/x"]).to(dev)
        np.testing.assert_allclose(S.compute_segment_costs_batch(x, pc, 1.0).cpu().numpy(), g[f"{tag}/cost"], rtol=2e-6, atol=1e-10)
        np.testing.assert_allclose(S.compute_segment_costs_batch(x, pc, 2.5).cpu().numpy(), g[f"{tag}/cost_scaled"], rtol=2e-6, atol=1e-10)
    m = SegmentCostPredictor(hidden_dim=128, n_layers=3)
    m.load_state_dict({k[len("dphi/sd/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("dphi/sd/")})
    m = m.to(dev)
    cond = {k[len("dphi/cond/"):]: torch.from_numpy(g[k]).to(dev) for k in g.files if k.startswith("dphi/cond/")}
    for feat, key in (("dphi/seg_feat", "dphi/pred_shared"), ("dphi/seg_feat_b", "dphi/pred_batched")):
        pred = m(cond, torch.from_numpy(g[feat]).to(dev)).cpu().numpy()
        assert np.abs(pred - g[key]).max() < 2e-2 * max(1.0, np.abs(g[key]).max()), key
    # pipeline: true segment costs of random walks -> DP anchors, bit-exact against the numpy oracle
    T, K, B = 64, 8, 256
    pc = S.build_segment_precompute(T, 2, dev)
    gen = torch.Generator().manual_seed(12)
    x = (torch.randn((B, T, 2), generator=gen) * 0.05).cumsum(1).to(dev)
    cost = S.compute_segment_costs_batch(x, pc, 1.0)
    C = S.build_cost_matrix_from_segments_batch(cost, pc, T)
    idx = S.dp_select_indices_batch(C, K).cpu().numpy()
    assert np.array_equal(idx, osel.dp_select_indices_batch(C.cpu().numpy(), K))
    assert (idx[:, 0] == 0).all() and (idx[:, -1] == T - 1).all()
