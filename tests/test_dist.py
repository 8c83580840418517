"""world_size-2 gloo tests (CPU) of the N>1 host logic: sharding + the descriptor all-gather."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dh3d_b200.dist import all_gather_descriptors, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 32, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.arange(total * 4, dtype=torch.float32).view(total, 4)
        lo, hi = shard_range(total, rank, world)
        counts = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
        got = all_gather_descriptors(full[lo:hi].clone(), counts)
        ok = torch.equal(got, full)
        if total % world == 0:
            ok = ok and torch.equal(all_gather_descriptors(full[lo:hi].clone()), full)
        results[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_all_gather_descriptors_world2_gloo(total):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(results[r] for r in range(world))
