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
    """Replicate parameters and buffers from rank `src` (NCCL over NVLink on GPUs, gloo on CPU): ONE packed buffer
    per dtype and one collective for it, instead of one broadcast per tensor (160+ for PointMaze)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return module
    groups = {}
    for t in list(module.parameters()) + list(module.buffers()):
        groups.setdefault((t.dtype, t.device), []).append(t.data)
    for ts in groups.values():
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
    # the writes above go through `.data` and bump no autograd version: tell the native handles to re-pack
    for m in module.modules():
        if hasattr(m, "invalidate"):
            m.invalidate()
    return module


def gather_trajectories(local: torch.Tensor, total: int, group=None, out: torch.Tensor = None) -> torch.Tensor:
    """all_gather of batch shards -> (total, H, T) on every rank, in global sample order.
    Equal shards go straight into `out` (or a new tensor) with one all_gather_into_tensor; ragged shards are padded
    to the widest one and trimmed."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    width = max(b - a for a, b in sizes)
    local = local.contiguous()
    if total % world == 0:
        if out is None:
            out = local.new_empty((total,) + tuple(local.shape[1:]))
        dist.all_gather_into_tensor(out, local, group=group)
        return out
    padded = local.new_zeros((width,) + tuple(local.shape[1:]))
    padded[:local.shape[0]] = local
    flat = local.new_empty((world * width,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(flat, padded, group=group)
    return torch.cat([flat[r * width:r * width + (b - a)] for r, (a, b) in enumerate(sizes)], dim=0)


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
    if not gather:
        return local
    nvtx = local.is_cuda
    if nvtx:
        torch.cuda.nvtx.range_push("gather_trajectories")
    try:
        return gather_trajectories(local, batch_size, group)
    finally:
        if nvtx:
            torch.cuda.nvtx.range_pop()
