"""N>1 host logic on CPU: genome sharding and the all-gather of the frequency matrix over gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kf2vecfsw_b200.dist import OverlappedGather, all_gather_rows, shard_by_size


def test_shard_by_size_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    sizes = rng.integers(1_000_000, 8_000_000, size=1001).tolist()
    for world in (1, 2, 4, 8):
        parts = shard_by_size(sizes, world)
        assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
        loads = [sum(sizes[i] for i in p) for p in parts]
        assert max(loads) - min(loads) <= max(sizes)
        assert all(p == sorted(p) for p in parts)
    assert shard_by_size([5, 5, 5], 8)[3:] == [[]] * 5


def _worker(rank, world, port, sizes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    parts = shard_by_size(sizes, world)
    V = 32
    local = torch.stack([torch.full((V,), float(i)) + torch.arange(V) / 100.0 for i in parts[rank]]) \
        if parts[rank] else torch.empty((0, V))
    full = all_gather_rows(local.float(), parts)
    want = torch.stack([torch.full((V,), float(i)) + torch.arange(V) / 100.0 for i in range(len(sizes))]).float()
    q.put((rank, bool(torch.equal(full, want))))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_files", [7, 2, 1])
def test_all_gather_rows_world2_gloo(n_files):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    sizes = [(i * 37) % 11 + 1 for i in range(n_files)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, sizes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res


def _worker_overlapped(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    og = OverlappedGather(3, 8, torch.float32, "cpu")
    ok = True
    outs = []
    for batch in range(5):
        loc = og.slot()
        loc.copy_(torch.full((3, 8), float(100 * batch + rank)))
        outs.append((batch, og.submit()))
        if batch >= 2:   # the buffer pair of batch - 2 is about to be reused: its result must be complete now
            pass
    og.drain()
    for batch, full in outs[-2:]:   # the last two results are still in their buffers
        want = torch.cat([torch.full((3, 8), float(100 * batch + r)) for r in range(world)])
        ok = ok and bool(torch.equal(full, want))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_overlapped_gather_world2_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_overlapped, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
