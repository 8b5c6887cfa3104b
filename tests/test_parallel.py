"""N>1 host logic on CPU: contiguous pose sharding and the chunked frame all-gather (gloo, world 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_shard_bounds_cover_everything():
    from rust_doom_b200.parallel import shard_bounds
    for n in (0, 1, 7, 8, 9, 1000, 100001):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                s, e, per = shard_bounds(n, r, world)
                assert e - s <= per
                got += list(range(s, e))
            assert got == list(range(n))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rust_doom_b200.parallel import all_gather_frames, shard_bounds
        H, W = 6, 10
        # frame i is filled with a pattern that encodes i
        full = (np.arange(n)[:, None, None] * 7 + np.arange(H)[None, :, None] * 3 + np.arange(W)[None, None, :]) % 251
        s, e, _ = shard_bounds(n, rank, world)
        local = torch.from_numpy(full[s:e].astype(np.uint8))
        out = all_gather_frames(local, n, chunk_frames=chunk)
        q.put((rank, bool(np.array_equal(out.numpy(), full.astype(np.uint8)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,chunk", [(10, 3), (7, 2), (1, 4), (16, 256)])
def test_all_gather_frames_gloo(n, chunk):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, chunk, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}


def _sink_worker(rank, world, port, n, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from rust_doom_b200.parallel import shard_bounds, write_frames_in_order
        H, W = 5, 9
        full = ((np.arange(n)[:, None, None] * 11 + np.arange(H)[None, :, None] * 5 + np.arange(W)[None, None, :]) % 253).astype(np.uint8)
        s, e, _ = shard_bounds(n, rank, world)
        got, firsts = [], []

        def sink(frames, first):
            firsts.append(first)
            got.append(frames.copy())

        written = write_frames_in_order(torch.from_numpy(full[s:e]), n, sink, chunk_frames=chunk)
        if rank == 0:
            ok = written == n and firsts == sorted(firsts) and np.array_equal(np.concatenate(got) if got else full[:0], full)
        else:
            ok = written == 0 and not got
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,chunk,world", [(10, 3, 2), (7, 64, 2), (1, 4, 2), (9, 2, 3)])
def test_frame_sink_on_gather_rank_gloo(n, chunk, world):
    """write_frames_in_order: rank 0 receives every rank's frames in global pose order, chunk by chunk."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sink_worker, args=(r, world, port, n, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(world))
    assert res == {r: True for r in range(world)}
