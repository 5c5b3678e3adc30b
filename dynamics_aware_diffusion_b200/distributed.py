"""Batch-sharded multi-GPU sampling: one process per GPU, replicated weights, independent per-sample
noise streams, and exactly two collectives -- a weight broadcast at start-up and a gather of the final
trajectories.  The reference has no multi-device code; this is new.

Every sample's trajectory is independent through the whole loop, so there is no per-step exchange.
The Philox subsequence of a sample is its GLOBAL index, which makes the result independent of the
number of ranks.
"""
import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int):
    """Contiguous [start, stop) of `total` items for `rank`; the first (total % world) ranks get one more."""
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def broadcast_module(module: torch.nn.Module, src: int = 0, group=None):
    """Replicate parameters and buffers from rank `src` (NCCL over NVLink on GPUs, gloo on CPU)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return module
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src, group=group)
    return module


def gather_trajectories(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """all_gather of ragged batch shards -> (total, H, T) on every rank, in global sample order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    out = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(out, padded.contiguous(), group=group)
    return torch.cat([o[:b - a] for o, (a, b) in zip(out, sizes)], dim=0)


def sharded_sample(sample_fn, batch_size: int, conditions=None, gather: bool = True, group=None, **kw):
    """Run `sample_fn(batch_size=local_B, conditions=local_conditions, sample_offset=start, **kw)` on this
    rank's shard (e.g. a policy's `sample_loop`) and gather the result."""
    if dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    start, stop = shard_bounds(batch_size, rank, world)
    local_cond = None
    if conditions:
        local_cond = {}
        for h, v in conditions.items():
            v = torch.as_tensor(v)
            local_cond[h] = v[start:stop] if (v.dim() >= 2 and v.shape[0] == batch_size and batch_size > 1) else v
    local = sample_fn(batch_size=stop - start, conditions=local_cond, sample_offset=start, **kw)
    return gather_trajectories(local, batch_size, group) if gather else local
