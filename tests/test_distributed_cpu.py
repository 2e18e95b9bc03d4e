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
