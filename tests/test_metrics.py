"""Batched trajectory metrics: oracle vs the reference's golden outputs (CPU), CUDA kernel vs oracle (GPU)."""
import numpy as np
import pytest
import torch

from oracle import metrics_np as om

KEYS = ("collision_rate", "goal_dist", "success", "path_length", "smoothness", "mse_to_gt")
EXACT = ("collision_rate", "success")


def _check(got, ref, name):
    for k in ref:
        a, b = np.asarray(got[k]), np.asarray(ref[k])
        if k in EXACT:
            assert np.array_equal(a, b), (name, k)
        else:
            assert np.allclose(a, b, rtol=2e-6, atol=1e-7), (name, k, np.abs(a - b).max())


def test_oracle_metrics_vs_golden(golden):
    g = golden("metrics")
    for n in "abcd":
        ref = {k: g[f"{n}_out_{k}"] for k in KEYS}
        _check(om.compute_metrics_batch(g[f"{n}_occ"], g[f"{n}_traj"], g[f"{n}_goal"], g[f"{n}_gt"]), ref, n)
    ref = {k: g[f"a1_out_{k}"] for k in KEYS[:5]}
    _check(om.compute_metrics_batch(g["a_occ"][0], g["a_traj"], g["a_goal"][0]), ref, "a1")
    with pytest.raises(ValueError):
        om.compute_metrics_batch(g["a_occ"][:3], g["a_traj"], g["a_goal"])


@pytest.mark.gpu
def test_cuda_metrics_vs_golden_and_oracle(golden):
    from interpolated_diffusion_b200.eval.metrics import compute_metrics, compute_metrics_batch
    g = golden("metrics")
    for n in "abcd":
        ref = {k: g[f"{n}_out_{k}"] for k in KEYS}
        got = compute_metrics_batch(torch.from_numpy(g[f"{n}_occ"]).cuda(), torch.from_numpy(g[f"{n}_traj"]).cuda(),
                                    torch.from_numpy(g[f"{n}_goal"]).cuda(), torch.from_numpy(g[f"{n}_gt"]).cuda())
        _check({k: v.cpu().numpy() for k, v in got.items()}, ref, n)
    got = compute_metrics_batch(torch.from_numpy(g["a_occ"][0]).cuda(), torch.from_numpy(g["a_traj"]).cuda(), torch.from_numpy(g["a_goal"][0]).cuda())
    _check({k: v.cpu().numpy() for k, v in got.items()}, {k: g[f"a1_out_{k}"] for k in KEYS[:5]}, "a1")
    one = compute_metrics(torch.from_numpy(g["a_occ"][3]).cuda(), torch.from_numpy(g["a_traj"][3]).cuda(), torch.from_numpy(g["a_goal"][3]).cuda())
    assert abs(one["path_length"] - float(g["a_out_path_length"][3])) < 1e-5
    with pytest.raises(ValueError):
        compute_metrics_batch(torch.from_numpy(g["a_occ"][:3]).cuda(), torch.from_numpy(g["a_traj"]).cuda(), torch.from_numpy(g["a_goal"]).cuda())
    with pytest.raises(RuntimeError):
        compute_metrics_batch(torch.from_numpy(g["a_occ"]), torch.from_numpy(g["a_traj"]), torch.from_numpy(g["a_goal"]))
    # larger random batch against the oracle
    gen = torch.Generator().manual_seed(3)
    B, T = 4099, 64
    occ = (torch.rand((B, 21, 21), generator=gen) < 0.2).float()
    traj = torch.rand((B, 1, 2), generator=gen) + 0.02 * torch.randn((B, T, 2), generator=gen).cumsum(1)
    goal = torch.rand((B, 2), generator=gen)
    got = compute_metrics_batch(occ.cuda(), traj.cuda(), goal.cuda())
    _check({k: v.cpu().numpy() for k, v in got.items()}, om.compute_metrics_batch(occ.numpy(), traj.numpy(), goal.numpy()), "rand")
