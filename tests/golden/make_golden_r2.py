#!/usr/bin/env python
"""Round-2 golden fixtures, generated from the LIVE reference (build container only; needs /root/reference):

    python tests/golden/make_golden_r2.py

* ``level_indices.npz``  -- draws of ``_sample_level_indices`` (train_interp_levels.py:578-596) for seeded CPU generators
                            (SURVEY 8a row a26).
* ``cfg1.npz``           -- BASELINE.json configs[0] literally: ``ParticleMazeDataset(16, T=64, with_velocity=False, seed=123)``
                            + DEFAULT-size random-init models (``torch.manual_seed(0)``; KeypointDenoiser 256x8, InterpLevelDenoiser
                            256x8 max_levels=3 mask_channels=2) through the reference's component functions
                            (sample_generate.py:363-404, 1252-1285): per-step ``z_t -> eps`` pairs (teacher forcing), x_pred,
                            Stage-2 delta, x_hat.  The 7.6 M-parameter weights are not stored: seeded init of the mirrored module
                            tree reproduces them, and the fixture carries per-tensor checksums to prove it.
* ``causal_chunks.npz``  -- the chunk loop of ``sample_generate_causal.py:485-583`` pinned by running the reference's own
                            ``main()`` (particle dataset, tiny checkpoints written to a temp dir, plotting stubbed), with every
                            random draw (per-chunk anchor indices, DDIM noise) recorded.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
os.environ.setdefault("IDB200_REFERENCE", "/root/reference")

from oracle import ref_driver  # noqa: E402

# the LIVE reference, not the vendored copy
ref_driver.reference_root = lambda: os.environ["IDB200_REFERENCE"]
R = ref_driver.import_reference()
assert R.root == "/root/reference", R.root

from src.data.dataset import ParticleMazeDataset  # noqa: E402
from src.train import train_interp_levels as rt  # noqa: E402

torch.set_num_threads(4)
TINY = dict(d_model=64, n_layers=2, n_heads=2, d_ff=128, d_cond=32, maze_channels=(8, 16))


def npy(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def save(name, d):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **{k: npy(v) for k, v in d.items()})
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB, {len(d)} arrays")


def gold_level_indices():
    g = {}
    cases = [(0, 64, 3, "high", 0.5), (1, 257, 3, "high", 0.5), (2, 100, 4, "high", 0.9), (3, 33, 3, "uniform", 0.5),
             (4, 50, 1, "high", 0.5), (5, 40, 3, "high", 0.0), (6, 40, 3, "high", 1.0), (7, 512, 3, "high", 0.25)]
    rows = []
    for i, (seed, B, S, mode, hp) in enumerate(cases):
        gen = torch.Generator().manual_seed(seed)
        s = rt._sample_level_indices(B, S, gen, torch.device("cpu"), mode, hp)
        g[f"s_{i}"] = s
        g[f"next_{i}"] = torch.rand((4,), generator=gen)            # the generator stream AFTER the call (draw count parity)
        rows.append([seed, B, S, 0 if mode == "high" else 1, hp])
    g["cases"] = np.array(rows, dtype=np.float64)
    save("level_indices.npz", g)


def checksum(sd):
    """Per-tensor (sum, sum of squares, first, last) in fp64, in state_dict order."""
    rows = []
    for v in sd.values():                                      # numpy fp64 sums: independent of torch's thread count
        a = v.detach().cpu().numpy().astype(np.float64).reshape(-1)
        rows.append([a.sum(), (a * a).sum(), a[0], a[-1]])
    return np.array(rows, dtype=np.float64)


def gold_cfg1():
    g = {}
    ds = ParticleMazeDataset(num_samples=16, T=64, with_velocity=False, seed=123)
    samples = [ds[i] for i in range(16)]
    cond = {k: torch.stack([s["cond"][k] for s in samples]) for k in samples[0]["cond"]}
    g["occ"], g["start_goal"], g["x_gt"] = cond["occ"], cond["start_goal"], torch.stack([s["x"] for s in samples])
    kp, il = ref_driver.build_models(R, D=2, levels=3)              # torch.manual_seed(0); kp then il; default sizes
    g["kp_checksum"], g["il_checksum"] = checksum(kp.state_dict()), checksum(il.state_dict())
    g["kp_keys"] = np.array(list(kp.state_dict().keys()))
    g["il_keys"] = np.array(list(il.state_dict().keys()))
    out = ref_driver.generate(R, kp, il, cond, seed=123, return_all=True)
    torch.manual_seed(123)
    g["z_T"] = torch.randn((16, 8, 2))
    times = R.rg._timesteps(1000, 20, schedule="quadratic")
    g["times"] = times
    # teacher-forced eps of every step: the model evaluated on the reference's own z_t
    eps = []
    with torch.no_grad():
        for i in range(len(times) - 1):
            t = torch.full((16,), int(times[i]), dtype=torch.long)
            eps.append(kp(out["z_inter"][i], t, out["idx"], out["known_mask"], cond, 64))
    g["eps"] = torch.stack(eps)
    for k in ("idx", "masks", "known_mask", "known_values", "z_inter", "z", "x_pred", "conf_pred", "mask_in", "delta", "x_hat"):
        g[k] = out[k]
    save("cfg1.npz", g)


def gold_causal_chunks():
    import src.sample.sample_generate_causal as sc
    from src.models.denoiser_interp_levels_causal import InterpLevelCausalDenoiser
    g = {}
    T, chunk, K_min, levels, n = 24, 8, 4, 3, 3
    torch.manual_seed(0)
    kp = R.KeypointDenoiser(data_dim=2, **TINY).eval()
    il = InterpLevelCausalDenoiser(data_dim=2, max_levels=levels, mask_channels=1, **TINY).eval()
    for k, v in kp.state_dict().items():
        g["kp/" + k] = v
    for k, v in il.state_dict().items():
        g["il/" + k] = v
    ds = ParticleMazeDataset(num_samples=n, T=T, with_velocity=False, use_sdf=False)
    samples = [ds[i] for i in range(n)]
    g["occ"] = torch.stack([s["cond"]["occ"] for s in samples])
    g["start_goal"] = torch.stack([s["cond"]["start_goal"] for s in samples])
    rec = {"idx": [], "randn": [], "pred": []}
    o_fix, o_randn, o_plot = sc.sample_fixed_k_indices_batch, torch.randn, sc.plot_trajectories

    def fix(*a, **k):
        idx, mask = o_fix(*a, **k)
        rec["idx"].append(npy(idx).copy())
        return idx, mask

    def randn(*a, **k):
        r = o_randn(*a, **k)
        rec["randn"].append(npy(r).copy())
        return r

    def plot(occ, trajs, labels, out_path=None, **k):
        if labels == ["pred"]:
            rec["pred"].append(np.asarray(trajs[0]).copy())

    with tempfile.TemporaryDirectory() as td:
        ck, ci = os.path.join(td, "kp.pt"), os.path.join(td, "il.pt")
        torch.save({"model": kp.state_dict(), "meta": {"stage": "keypoints", "T": T, "N_train": 1000, "schedule": "cosine", "K": K_min}}, ck)
        torch.save({"model": il.state_dict(), "meta": {}}, ci)
        argv = ["x", "--ckpt_keypoints", ck, "--ckpt_interp", ci, "--out_dir", os.path.join(td, "out"), "--n_samples", str(n), "--T", str(T),
                "--chunk", str(chunk), "--K_min", str(K_min), "--levels", str(levels), "--ddim_steps", "6", "--logit_space", "1",
                "--use_ema", "0", "--dataset", "d4rl_prepared", "--prepared_path", os.path.join(td, "prep.npz"), "--device", "cpu",
                "--kp_d_model", "64", "--kp_n_layers", "2", "--kp_n_heads", "2", "--kp_d_ff", "128", "--kp_d_cond", "32", "--kp_maze_channels", "8,16",
                "--s2_d_model", "64", "--s2_n_layers", "2", "--s2_n_heads", "2", "--s2_d_ff", "128", "--s2_d_cond", "32", "--s2_maze_channels", "8,16"]
        # main() refuses --dataset particle (:218-219): hand it the same particle-maze samples as a prepared dataset.npz
        np.savez(os.path.join(td, "prep.npz"), x=npy(torch.stack([s["x"] for s in samples])), start_goal=npy(g["start_goal"]), occ=npy(g["occ"][:, 0]))
        os.makedirs(os.path.join(td, "out"), exist_ok=True)
        sc.sample_fixed_k_indices_batch, torch.randn, sc.plot_trajectories = fix, randn, plot
        old_argv = sys.argv
        sys.argv = argv
        try:
            sc.main()
        finally:
            sys.argv = old_argv
            sc.sample_fixed_k_indices_batch, torch.randn, sc.plot_trajectories = o_fix, o_randn, o_plot
    n_chunks = len(rec["idx"]) // n
    assert len(rec["idx"]) == n * n_chunks and len(rec["randn"]) == n * n_chunks and len(rec["pred"]) == n, (len(rec["idx"]), len(rec["randn"]), len(rec["pred"]))
    for c in range(n_chunks):                                       # draws are per sample (B = 1), chunk-major inside a sample
        g[f"idx_{c}"] = np.concatenate([rec["idx"][i * n_chunks + c] for i in range(n)], axis=0)
        g[f"zT_{c}"] = np.concatenate([rec["randn"][i * n_chunks + c] for i in range(n)], axis=0)
    g["x_gen"] = np.stack(rec["pred"])                              # [n, T, 2] (positions: what main() plots as "pred")
    g["cfg"] = np.array([T, chunk, K_min, levels, n, 6], dtype=np.int64)
    save("causal_chunks.npz", g)


if __name__ == "__main__":
    gold_level_indices()
    gold_cfg1()
    gold_causal_chunks()
