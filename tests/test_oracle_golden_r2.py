"""Round-2 pins (CPU only) against fixtures produced by the live reference (``tests/golden/make_golden_r2.py``):

* row a26: ``_sample_level_indices`` -- the host mirror reproduces the reference's draws AND leaves the generator stream where
  the reference leaves it (same draw count), for CPU generators; the sync-free speed variant has the same distribution.
* cfg 1 literally: the oracle's denoisers at the DEFAULT model sizes (weights rebuilt by seeded init of the mirrored module
  tree, proven by per-tensor checksums) against the reference's per-step ``z_t -> eps`` pairs on ParticleMazeDataset(16,T=64).
* row f3: the oracle's causal chunk loop against a run of the reference's own ``sample_generate_causal.main()``.
"""
import numpy as np
import torch

from oracle import denoiser_torch as dn
from oracle import generate as og
from oracle import keyframes_np as kf


def _checksum(sd):
    rows = []
    for v in sd.values():                                      # numpy fp64 sums: independent of torch's thread count
        a = v.detach().cpu().numpy().astype(np.float64).reshape(-1)
        rows.append([a.sum(), (a * a).sum(), a[0], a[-1]])
    return np.array(rows, dtype=np.float64)


def test_sample_level_indices_draws(golden):
    from interpolated_diffusion_b200.train.train_interp_levels import _sample_level_indices
    g = golden("level_indices")
    for i, (seed, B, S, mode, hp) in enumerate(g["cases"]):
        gen = torch.Generator().manual_seed(int(seed))
        s = _sample_level_indices(int(B), int(S), gen, torch.device("cpu"), "high" if mode == 0 else "uniform", float(hp))
        assert s.dtype == torch.int64
        assert np.array_equal(s.numpy(), g[f"s_{i}"]), i
        # same number of generator draws as the reference: the stream after the call is the reference's
        assert np.array_equal(torch.rand((4,), generator=gen).numpy(), g[f"next_{i}"]), i


def test_sample_level_indices_sync_free_distribution():
    from interpolated_diffusion_b200.train.train_interp_levels import _sample_level_indices
    gen = torch.Generator().manual_seed(3)
    B, S, hp = 200_000, 3, 0.5
    s = _sample_level_indices(B, S, gen, torch.device("cpu"), "high", hp, sync_free=True).numpy()
    assert s.min() == 1 and s.max() == S
    p = np.bincount(s, minlength=S + 1)[1:] / B
    want = np.array([(1 - hp) / S, (1 - hp) / S, hp + (1 - hp) / S])     # s = S w.p. hp, else U{1..S}
    assert np.abs(p - want).max() < 5e-3, (p, want)


def _mirrored_models():
    from interpolated_diffusion_b200.models.denoiser_interp_levels import InterpLevelDenoiser
    from interpolated_diffusion_b200.models.denoiser_keypoints import KeypointDenoiser
    torch.manual_seed(0)
    kp = KeypointDenoiser(data_dim=2)
    il = InterpLevelDenoiser(data_dim=2, max_levels=3, mask_channels=2)
    return kp, il


def test_cfg1_seeded_init_equals_reference_weights(golden):
    g = golden("cfg1")
    kp, il = _mirrored_models()
    assert list(kp.state_dict().keys()) == list(g["kp_keys"])
    assert list(il.state_dict().keys()) == list(g["il_keys"])
    assert np.array_equal(_checksum(kp.state_dict()), g["kp_checksum"])
    assert np.array_equal(_checksum(il.state_dict()), g["il_checksum"])
    assert sum(v.numel() for v in kp.state_dict().values()) == 7_618_306
    assert sum(v.numel() for v in il.state_dict().values()) == 7_586_562          # mask_channels = 2 (SURVEY quotes 7 586 818 for C = 3)


def test_cfg1_oracle_per_step_eps_and_stage2(golden):
    """The oracle (CPU fp32) on the reference's own z_t of every DDIM step, default 256x8 models, particle-maze conditioning."""
    g = golden("cfg1")
    kp, il = _mirrored_models()
    sd_kp = {k: v.detach() for k, v in kp.state_dict().items()}
    sd_il = {k: v.detach() for k, v in il.state_dict().items()}
    cond = {"occ": torch.from_numpy(g["occ"]), "start_goal": torch.from_numpy(g["start_goal"])}
    idx, km = torch.from_numpy(g["idx"]), torch.from_numpy(g["known_mask"])
    times = g["times"]
    assert times.tolist() == [999, 896, 799, 708, 622, 542, 467, 398, 334, 276, 224, 177, 135, 99, 69, 44, 24, 11, 2, 0]
    assert np.array_equal(g["idx"][0], [0, 9, 18, 27, 36, 45, 54, 63])
    with torch.no_grad():
        for i in (0, 1, 9, 18):
            t = torch.full((16,), int(times[i]), dtype=torch.long)
            eps = dn.keypoint_denoiser(sd_kp, 8, torch.from_numpy(g["z_inter"][i]), t, idx, km, cond, 64)
            # a random-init rollout blows |z_t| up to ~8e4 after the first step (the cosine schedule's first DDIM step multiplies
            # by ~3243), so |eps| reaches ~1.5e3: the bar is relative to the step's magnitude
            assert np.abs(eps.numpy() - g["eps"][i]).max() < 1e-5 * max(1.0, np.abs(g["eps"][i]).max()), i
        delta = dn.interp_level_denoiser(sd_il, 8, torch.from_numpy(g["x_pred"]), torch.full((16,), 3), torch.from_numpy(g["mask_in"]), cond)
    assert np.abs(delta.numpy() - g["delta"]).max() < 2e-5
    # interpolation of the reference's own keypoints is bit-exact
    from oracle import sampling_np as sp
    x_pred = kf.interpolate_from_indices(g["idx"], sp.sigmoid_pos(g["z"]), 64, recompute_velocity=True)
    assert np.abs(x_pred - g["x_pred"]).max() <= 1.2e-7


def test_causal_chunk_loop_against_live_main(golden):
    """oracle.generate.generate_causal_chunked == the reference's sample_generate_causal.main() (tiny models, T = 24, chunk 8,
    3 samples; the script's own random draws replayed)."""
    g = golden("causal_chunks")
    T, chunk, K_min, levels, n, steps = [int(v) for v in g["cfg"]]
    sd = lambda p: {k[len(p):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(p)}
    cond = {"occ": torch.from_numpy(g["occ"]), "start_goal": torch.from_numpy(g["start_goal"])}
    plan = og.causal_chunk_plan(T, chunk, K_min)
    idx_chunks = [g[f"idx_{c}"] for c in range(len(plan))]
    z_chunks = [g[f"zT_{c}"] for c in range(len(plan))]
    for (cur, end, local_T, K), idx, z in zip(plan, idx_chunks, z_chunks):
        assert idx.shape == (n, K) and z.shape == (n, K, 2)
        assert (idx[:, 0] == 0).all() and (idx[:, -1] == local_T - 1).all()
    with torch.no_grad():
        x = og.generate_causal_chunked(sd("kp/"), sd("il/"), 2, cond, T=T, chunk=chunk, K_min=K_min, levels=levels, idx_chunks=idx_chunks,
                                       z_T_chunks=z_chunks, ddim_steps=steps, logit_space=True)
    assert x.shape == (n, T, 2)
    assert np.abs(x[:, :, :2] - g["x_gen"]).max() < 2e-4, np.abs(x[:, :, :2] - g["x_gen"]).max()
