"""CPU, world_size 2 over gloo: the N > 1 host logic (frame sharding + gather to rank 0)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bev_b200 import sharding


def test_shard_range_partitions_everything():
    for n in (0, 1, 7, 8, 256, 1000):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                b, e = sharding.shard_range(n, r, world)
                assert 0 <= b <= e <= n
                cover += list(range(b, e))
            assert cover == list(range(n))
            sizes = [sharding.shard_range(n, r, world)[1] - sharding.shard_range(n, r, world)[0]
                     for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    # BASELINE configs[3]: 8 camera streams on 2 / 4 / 8 GPUs
    assert sharding.shard_cameras(8, 1, 2) == [4, 5, 6, 7]
    assert sharding.shard_cameras(8, 3, 4) == [6, 7]
    assert sharding.shard_cameras(8, 5, 8) == [5]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 10
        b, e = sharding.shard_range(n, rank, world)
        frames = torch.arange(n * 6, dtype=torch.uint8).reshape(n, 2, 3)[b:e].contiguous()
        bev = frames + 1                       # stand-in for this rank's warp output
        got = sharding.gather_to_rank0(bev, chunks=3)
        boxes = torch.full((rank + 2, 5), float(rank))   # ragged: tracked boxes per rank
        gb = sharding.gather_ragged_to_rank0(boxes)
        # shards that differ by one frame (n % world != 0) must not hang or corrupt
        b7, e7 = sharding.shard_range(7, rank, world)
        odd = (torch.arange(7 * 4, dtype=torch.int32).reshape(7, 4))[b7:e7].contiguous()
        godd = sharding.gather_to_rank0(odd, chunks=2)

        # pipelined: slices are produced (here: copied) and gathered one after the other
        src = torch.arange(n * 6, dtype=torch.uint8).reshape(n, 2, 3)[b:e].contiguous()
        calls = []

        def produce(lo, hi, out):
            calls.append((lo, hi))
            out[lo:hi] = src[lo:hi] * 2

        local, pg = sharding.pipelined_gather_to_rank0(produce, e - b, (2, 3), torch.uint8, "cpu", chunks=3)
        assert torch.equal(local, src * 2) and len(calls) == 3
        if rank == 0:
            q.put((got.numpy(), gb.numpy(), godd.numpy(), pg.numpy()))
        else:
            assert got is None and gb is None and godd is None and pg is None
            q.put(None)
    finally:
        dist.destroy_process_group()


def test_gather_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got, gb, godd, pg = [r for r in results if r is not None][0]
    expect = (np.arange(60, dtype=np.uint8).reshape(10, 2, 3) + 1)
    assert np.array_equal(got, expect)
    assert gb.shape == (5, 5) and np.array_equal(gb[:, 0], [0, 0, 1, 1, 1])
    assert np.array_equal(godd, np.arange(28, dtype=np.int32).reshape(7, 4))
    assert np.array_equal(pg, np.arange(60, dtype=np.uint8).reshape(10, 2, 3) * 2)


def _subgroup_worker(rank, world, port, q):
    """Gather inside a sub-group whose rank 0 is NOT global rank 0: dst must be translated."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        grp = dist.new_group([1, 2])
        if rank in (1, 2):
            t = torch.full((2, 3), float(rank))
            got = sharding.gather_to_rank0(t, group=grp, chunks=2)
            q.put((rank, None if got is None else got.numpy()))
        else:
            q.put((rank, None))
    finally:
        dist.destroy_process_group()


def test_gather_in_a_subgroup_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_subgroup_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert results[0] is None and results[2] is None
    assert np.array_equal(results[1], np.repeat(np.array([1.0, 1.0, 2.0, 2.0])[:, None], 3, 1))


def test_single_process_is_identity():
    t = torch.arange(12).reshape(3, 4)
    assert sharding.gather_to_rank0(t) is t
    assert sharding.gather_ragged_to_rank0(t) is t
    local, g = sharding.pipelined_gather_to_rank0(lambda b, e, out: out[b:e].copy_(t[b:e]), 3, (4,),
                                                  t.dtype, "cpu", chunks=2)
    assert torch.equal(local, t) and g is local
