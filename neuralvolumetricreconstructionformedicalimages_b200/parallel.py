"""Host-side plumbing of the multi-GPU path (one process per GPU, torch.distributed).

The NAF step shards over rays: parameters and optimizer state are replicated, every rank renders its
own rays, and ONE exchange step per iteration sums the flat gradient (the summed gradient goes to
Adam unscaled: see combined_loss).  The voxel
query shards over slabs of the outermost lattice index and needs no collective.  Everything here works on
any backend (NCCL on the GPUs; the gloo tests in tests/test_distributed_cpu.py run it on CPU tensors).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_info(group=None):
    """(rank, world_size) of the default / given process group; (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_range(n: int, rank: int, world: int):
    """[i0, i1) of a balanced contiguous split of range(n): the first n % world shards get one extra item.
    Shards are disjoint, ordered by rank and cover range(n) exactly (also when n < world: trailing shards are empty)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"shard_range: rank {rank} outside world of {world}")
    base, extra = divmod(int(n), world)
    i0 = rank * base + min(rank, extra)
    return i0, i0 + base + (1 if rank < extra else 0)


def allreduce_sum_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the flat gradient over all ranks (the step's only exchange); no-op for a single process."""
    if world_info(group)[1] > 1:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat


def broadcast_(flat: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """Make every replica start from rank `src`'s values (parameters and optimizer state)."""
    if world_info(group)[1] > 1:
        dist.broadcast(flat, src=src, group=group)
    return flat


def replica_divergence(flat: torch.Tensor, group=None) -> float:
    """max over ranks of |flat - flat_on_rank0| -- 0.0 when the replicas are bit-identical (they must be: same
    all-reduced gradient, same deterministic Adam).  Diagnostic; costs a broadcast + an all-reduce."""
    rank, world = world_info(group)
    if world == 1:
        return 0.0
    ref = flat.clone()
    dist.broadcast(ref, src=0, group=group)
    d = (flat - ref).abs().max().reshape(1)
    dist.all_reduce(d, op=dist.ReduceOp.MAX, group=group)
    return float(d.item())


def combined_loss(local_loss: torch.Tensor, group=None) -> torch.Tensor:
    """The step's loss over all ranks: the SUM of the per-rank losses.  The reference's loss is itself a sum of chunk means
    (train.py:69-127: one calc_mse_loss per 200-ray chunk, loss.py:26-46 adds them up), so W ranks that each run the
    reference's loop on their own rays compute exactly what ONE GPU computes on the concatenated batch with the same chunk
    boundaries -- the summed gradient is the gradient of this sum, and Adam gets it unscaled (grad_scale = 1)."""
    rank, world = world_info(group)
    out = local_loss.detach().clone().reshape(1)
    if world > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out[0]


def gather_shards(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather of a contiguous split made with shard_range: rank r holds elements [i0_r, i1_r) of a length-n_total vector
    (first dimension); every rank gets the whole vector.  Shards may differ by one element, so they travel padded to the
    largest one.  Used by the sharded evaluation render (reference train.py:235-240: the rays of the chosen view, split over the
    ranks, `acc` gathered)."""
    rank, world = world_info(group)
    if world == 1:
        assert local.shape[0] == n_total
        return local
    base, extra = divmod(int(n_total), world)
    cap = base + (1 if extra else 0)
    i0, i1 = shard_range(n_total, rank, world)
    assert local.shape[0] == i1 - i0, "gather_shards: the local shard does not match shard_range"
    padded = local.new_zeros((cap,) + tuple(local.shape[1:]))
    padded[: i1 - i0] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded, group=group)
    parts = []
    for r in range(world):
        a, b = shard_range(n_total, r, world)
        parts.append(out[r][: b - a])
    return torch.cat(parts, 0)
