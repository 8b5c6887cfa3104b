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


def test_sharded_schedule_matches_the_c_plan():
    """b2d_render_sharded's partition, restated in parallel.sharded_schedule: every pose is rendered by exactly one rank,
    padding only repeats the last pose, chunks tile a rank's block, and jobs.global_pose_index inverts the layout."""
    from rust_doom_b200.jobs import global_pose_index, map_assignment
    from rust_doom_b200.parallel import padded_block, sharded_schedule
    for n in (1, 7, 8, 9, 1000, 100000):
        for world in (1, 2, 3, 8):
            for chunk in (1, 3, 256):
                per, plan = sharded_schedule(n, world, chunk)
                assert per * world >= n and (per - 1) * world < n
                assert sum(c for _, c in plan) == per and [f for f, _ in plan] == list(np.cumsum([0] + [c for _, c in plan[:-1]]))
                if n <= 1000:
                    poses = np.arange(n)
                    seen = []
                    for q in range(world):
                        blk = padded_block(poses, q, world)
                        for first, cnt in plan:
                            for j in range(cnt):
                                g = global_pose_index(q, per, first, j)
                                assert blk[first + j] == min(g, n - 1)
                                if g < n:
                                    seen.append(g)
                    assert sorted(seen) == list(range(n))
    assert [len(x) for x in map_assignment(10, 4)] == [3, 3, 2, 2]          # BASELINE.json configs[3]
    assert [len(x) for x in map_assignment(10, 1)] == [10]
    assert sum(map_assignment(9, 8), []) == list(range(9))


def test_frame_checksum_is_position_sensitive():
    import rust_doom_b200 as b2d
    rng = np.random.default_rng(3)
    f = rng.integers(0, 256, (20, 30), dtype=np.uint8)
    g = f.copy()
    g[3, 4], g[3, 5] = f[3, 5], f[3, 4]
    assert f[3, 4] != f[3, 5] and b2d.frame_checksum(f) != b2d.frame_checksum(g)
    z = np.zeros((4, 4), np.uint8)
    assert b2d.frame_checksum(z) == sum((i * 0x9E3779B1 + 0x7F4A7C15) & 0xFFFFFFFF for i in range(16)) & 0xFFFFFFFF


def _c5_worker(rank, world, port, n, chunk, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import rust_doom_b200 as b2d
        from rust_doom_b200.jobs import global_pose_index
        from rust_doom_b200.parallel import padded_block, sharded_gather_emulated, sharded_schedule
        H, W = 6, 10
        frame_of = lambda g: ((g * 7 + np.arange(H)[:, None] * 3 + np.arange(W)[None, :]) % 251).astype(np.uint8)   # noqa: E731
        per, plan = sharded_schedule(n, world, chunk)
        mine = padded_block(np.arange(n), rank, world)
        local = torch.from_numpy(np.stack([frame_of(int(g)) for g in mine]))
        table = np.zeros((world, per), np.uint32)

        def on_chunk(k, first, cnt, gathered):
            for qq in range(world):
                for j in range(cnt):
                    table[qq, first + j] = b2d.frame_checksum(gathered[qq, j].numpy())

        sharded_gather_emulated(local, n, chunk, on_chunk)
        want = np.array([[b2d.frame_checksum(frame_of(min(global_pose_index(qq, per, 0, j), n - 1))) for j in range(per)]
                         for qq in range(world)], np.uint32)
        t = torch.from_numpy(table.view(np.int32).copy())
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        q.put((rank, bool(np.array_equal(table, want)) and bool(torch.equal(lo, hi))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,chunk", [(10, 3), (7, 2), (9, 256)])
def test_c5_gathered_checksum_table_gloo(n, chunk):
    """The c5 job's validation logic on CPU (world 2, gloo): chunk by chunk in b2d_render_sharded's buffer layout, every
    rank ends with the same [world, per] checksum table and it equals the checksums of the frames by global pose index."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_c5_worker, args=(r, 2, port, n, chunk, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
