"""N > 1 host logic on CPU: world_size-2 gloo run of the batch-parallel shard + sampler-boundary all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lakonlab.parallel.batch_parallel import gather_latents, shard_batch, shard_bounds


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 6, dtype=torch.float32).reshape(total, 3, 2)
        local = shard_batch(full)                      # this rank's images
        local = local * 2.0 + 1.0                      # stand-in for the per-image denoise (no cross-image dependence)
        out = gather_latents(local, total)
        q.put((rank, torch.equal(out, full * 2.0 + 1.0), tuple(out.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 5, 2])
def test_shard_and_gather_world2(total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, shape in res:
        assert ok, f"rank {rank} gathered a wrong tensor"
        assert shape == (total, 3, 2)


def test_shard_bounds_cover_everything():
    for total in (1, 2, 7, 8, 64):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_is_identity():
    x = torch.randn(3, 4, 64)
    assert shard_batch(x, 0, 1) is not None and torch.equal(shard_batch(x, 0, 1), x)
    assert torch.equal(gather_latents(x, 3), x)


def _opt_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from arcflow_b200.optim import FlatAdamW
        o = FlatAdamW.__new__(FlatAdamW)          # host-side logic only: the arena all-reduce (no CUDA on this box)
        o.grads = torch.full((10,), float(rank + 1))
        o.all_reduce_grads()
        q.put((rank, o.grads.tolist()))
    finally:
        dist.destroy_process_group()


def test_gradient_arena_all_reduce_is_a_mean_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_opt_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, g in res:
        assert g == [1.5] * 10
