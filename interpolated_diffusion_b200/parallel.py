"""Trajectory-batch sharding across the GPUs of one box (one process per GPU).

The generation path is row-independent (no cross-sample statistic: LayerNorm is per token), so the batch is split into
contiguous per-rank ranges with no data-path collective; the only exchange is the final all-gather of the samples.
The training step adds the one real exchange of that path: the data-parallel all-reduce of the flat fp32
gradient arena."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) rows of `rank`; the first (total % world) ranks take one extra row."""
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_cond(cond: dict, rank: int, world: int) -> dict:
    B = next(iter(cond.values())).shape[0]
    lo, hi = shard_range(B, rank, world)
    return {k: v[lo:hi] for k, v in cond.items()}


def gather_samples(x_local: torch.Tensor, total: int | None = None) -> torch.Tensor:
    """All ranks receive the samples of every rank, concatenated in rank order.  Equal shards use one
    `all_gather_into_tensor` (NCCL over NVLink on the GPU box); ragged shards fall back to padded gathers."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return x_local
    world = dist.get_world_size()
    n_local = torch.tensor([x_local.shape[0]], device=x_local.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(sizes, n_local)
    sizes = [int(s.item()) for s in sizes]
    if len(set(sizes)) == 1 and dist.get_backend() == "nccl":
        out = torch.empty((world * sizes[0],) + tuple(x_local.shape[1:]), device=x_local.device, dtype=x_local.dtype)
        dist.all_gather_into_tensor(out, x_local.contiguous())
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(x_local.shape[1:]), device=x_local.device, dtype=x_local.dtype)
    pad[: x_local.shape[0]] = x_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[:n] for p, n in zip(parts, sizes)], dim=0)


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def all_reduce_sum_(flat: torch.Tensor, group=None, bucket_elems: int = 0) -> torch.Tensor:
    """In-place sum over ranks of a flat gradient arena (no-op for one process).  Each rank's gradient already carries the
    1 / world factor (it rides on the loss kernel), so the sum IS the data-parallel mean.  ``bucket_elems`` > 0 issues the
    reduction as async buckets of that many elements (launch-latency sized, not link sized: NVSwitch reduces in the fabric)."""
    if world_size(group) == 1:
        return flat
    if bucket_elems <= 0 or bucket_elems >= flat.numel():
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        return flat
    works = [dist.all_reduce(flat[o:o + bucket_elems], op=dist.ReduceOp.SUM, group=group, async_op=True)
             for o in range(0, flat.numel(), bucket_elems)]
    for w in works:
        w.wait()
    return flat


def broadcast_replica_state(tensors, step_count: int = 0, group=None, src: int = 0) -> int:
    """Make every rank's replica state equal rank ``src``'s: the tensors (parameter arena, EMA, Adam moments) are broadcast in
    place and the broadcast step count is returned.  DDP does this at construction; without it, ranks that built the model
    with different seeds -- or of which only some loaded a checkpoint -- would silently all-reduce gradients taken at different
    weights.  No-op for one process."""
    if world_size(group) == 1:
        return int(step_count)
    for t in tensors:
        if t is not None:
            dist.broadcast(t, src=src, group=group)
    dev = next((t.device for t in tensors if t is not None), torch.device("cpu"))
    sc = torch.tensor([int(step_count)], dtype=torch.int64, device=dev)
    dist.broadcast(sc, src=src, group=group)
    return int(sc.item())


def replicas_in_sync(flat: torch.Tensor, group=None) -> bool:
    """True when every rank holds the same arena (compares an order-independent checksum pair: sum and sum of squares in fp64)."""
    if world_size(group) == 1:
        return True
    f = flat.detach().double()
    mine = torch.stack([f.sum(), (f * f).sum()])
    lo, hi = mine.clone(), mine.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return bool(torch.equal(lo, hi))
