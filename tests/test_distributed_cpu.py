"""world_size-2 gloo test (CPU) of the N > 1 path: contiguous trajectory sharding + final all-gather reproduce the
single-process result exactly (the path has no cross-sample coupling), including a ragged split."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "generate_tiny.npz")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full(B):
    sys.path.insert(0, ROOT)
    from oracle import generate as og
    g = dict(np.load(GOLD))
    sd = lambda p: {k[len(p):]: torch.from_numpy(v) for k, v in g.items() if k.startswith(p)}
    gen = torch.Generator().manual_seed(17)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    z_T = torch.randn((B, 8, 2), generator=gen).numpy()
    return og, sd("kp/"), sd("il/"), cond, z_T


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from interpolated_diffusion_b200.parallel import gather_samples, shard_cond, shard_range
        og, sd_kp, sd_il, cond, z_T = _full(B)
        lo, hi = shard_range(B, rank, world)
        local = og.generate(sd_kp, sd_il, 2, shard_cond(cond, rank, world), z_T[lo:hi], T=64, K_min=8, levels=3, D=2)["x_hat"]
        allx = gather_samples(torch.from_numpy(local), B)
        if rank == 0:
            q.put(allx.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [6, 5])
def test_sharded_generation_equals_single_process(B):
    sys.path.insert(0, ROOT)
    from interpolated_diffusion_b200.parallel import shard_range
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    og, sd_kp, sd_il, cond, z_T = _full(B)
    ref = og.generate(sd_kp, sd_il, 2, cond, z_T, T=64, K_min=8, levels=3, D=2)["x_hat"]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got.shape == ref.shape
    # identical rows, identical math: only BLAS blocking may differ with the batch size
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-4)


# ---- data-parallel gradient all-reduce of the training step: flat arena, 1 / world on the loss --------------------
def _tiny_grads(rows, world):
    """Flat gradient of (1 / world) * local Stage-2 loss for the rows [lo, hi), computed by autograd on the oracle."""
    sys.path.insert(0, ROOT)
    from oracle import denoiser_torch as O
    g = dict(np.load(GOLD))
    sd = {k[len("il/"):]: torch.from_numpy(v).clone().requires_grad_() for k, v in g.items() if k.startswith("il/")}
    B, T = 6, 64
    gen = torch.Generator().manual_seed(5)
    x_s = torch.rand((B, T, 2), generator=gen)
    C = sd["in_proj.weight"].shape[1] - 2
    mask = (torch.rand((B, T, C), generator=gen) < 0.3).float()
    s = torch.randint(1, 4, (B,), generator=gen)
    cond = {"occ": (torch.rand((B, 1, 21, 21), generator=gen) < 0.2).float(), "start_goal": torch.rand((B, 4), generator=gen)}
    target = 0.1 * torch.randn((B, T, 2), generator=gen)
    conf = torch.rand((B, T), generator=gen)
    lo, hi = rows
    out = O.interp_level_denoiser(sd, 2, x_s[lo:hi], s[lo:hi], mask[lo:hi], {k: v[lo:hi] for k, v in cond.items()})
    w = 1.0 + (0.1 - 1.0) * conf[lo:hi]
    loss = (w[..., None] * (out - target[lo:hi]) ** 2).sum() / (w.sum() * 2 + 1e-8) / world
    loss.backward()
    return torch.cat([v.grad.reshape(-1) for _, v in sorted(sd.items())])


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from interpolated_diffusion_b200.parallel import all_reduce_sum_, shard_range, world_size
        assert world_size() == world
        flat = _tiny_grads(shard_range(6, rank, world), world)
        all_reduce_sum_(flat, bucket_elems=100_000)               # bucketed async path
        if rank == 0:
            q.put(flat.numpy())
    finally:
        dist.destroy_process_group()


def test_data_parallel_gradient_allreduce_is_the_mean_of_rank_gradients():
    sys.path.insert(0, ROOT)
    from interpolated_diffusion_b200.parallel import shard_range
    ref = sum(_tiny_grads(shard_range(6, r, 2), 2) for r in range(2)).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    np.testing.assert_allclose(got, ref, rtol=1e-5, atol=1e-7)


# ---- replica state broadcast (DDP semantics at trainer construction / after a checkpoint load on some ranks) ------------------
def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        from interpolated_diffusion_b200.parallel import broadcast_replica_state, replicas_in_sync
        torch.manual_seed(100 + rank)                              # every rank builds "its model" with a different seed
        flat, ema, m, v = (torch.randn(1000) for _ in range(4))
        before = replicas_in_sync(flat)
        step = broadcast_replica_state([flat, ema, None, m, v], step_count=7 * (rank + 1))
        after = replicas_in_sync(flat) and replicas_in_sync(ema) and replicas_in_sync(m) and replicas_in_sync(v)
        q.put((rank, before, after, step, float(flat.sum())))
    finally:
        dist.destroy_process_group()


def test_replica_state_broadcast_makes_ranks_identical():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [g[1] for g in got] == [False, False]                   # different seeds: the replicas started out different
    assert [g[2] for g in got] == [True, True]
    assert [g[3] for g in got] == [7, 7]                           # rank 0's step count everywhere
    assert got[0][4] == got[1][4]
