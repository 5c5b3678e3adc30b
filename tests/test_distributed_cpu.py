"""World-size-2 gloo tests of the batch-sharding helpers (the N>1 path of bench.py / sharded_sample)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dynamics_aware_diffusion_b200.distributed import shard_bounds, gather_trajectories, broadcast_module, sharded_sample


def test_shard_bounds_cover_and_balance():
    for total in (0, 1, 7, 4096, 8191):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class Net(torch.nn.Module):
            """float parameters + an integer buffer (two packed buffers), and the re-pack hook of TemporalUnet"""

            def __init__(self):
                super().__init__()
                self.lin = torch.nn.Linear(3, 2)
                self.register_buffer("steps", torch.full((5,), rank + 7, dtype=torch.long))
                self.invalidated = 0

            def invalidate(self):
                self.invalidated += 1

        lin = Net()
        with torch.no_grad():
            lin.lin.weight.fill_(float(rank + 1))
            lin.lin.bias.fill_(float(10 * (rank + 1)))
        broadcast_module(lin, src=0)
        ok_bcast = (bool((lin.lin.weight == 1.0).all()) and bool((lin.lin.bias == 10.0).all())
                    and bool((lin.steps == 7).all()) and lin.invalidated == 1)

        def sample_fn(batch_size, conditions, sample_offset, **kw):
            # a stand-in sampler: row value = global row index, plus the per-row condition
            rows = torch.arange(sample_offset, sample_offset + batch_size, dtype=torch.float32)
            out = rows[:, None, None].expand(batch_size, 4, 3).clone()
            if conditions:
                out[:, 0] = conditions[0] if conditions[0].shape[0] == 1 else conditions[0][:batch_size]
            return out

        cond = {0: torch.arange(total, dtype=torch.float32)[:, None].expand(total, 3) * 10}
        full = sharded_sample(sample_fn, total, conditions=cond)
        want = torch.arange(total, dtype=torch.float32)[:, None, None].expand(total, 4, 3).clone()
        want[:, 0] = cond[0]
        q.put((rank, ok_bcast, bool(torch.equal(full, want)), tuple(full.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_sample_gloo_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, ok_bcast, ok_gather, shape in res:
        assert ok_bcast, "weights were not replicated from rank 0"
        assert ok_gather and shape == (total, 4, 3)
