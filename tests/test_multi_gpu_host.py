"""CPU, world_size 2 over gloo: the N>1 host logic — contiguous query shards, no data-path collective, results
gathered on the control path — with a stand-in for Planner.findPaths (the CUDA path itself needs a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nq, out_dir):
    sys.path.insert(0, ROOT)
    import astar_planners_b200  # noqa: F401
    from astar_planners_b200 import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)
    starts, goals = rng.random((nq, 3)), rng.random((nq, 3))
    dt = np.dtype([("solved", "i4"), ("g_final", "f8"), ("rank", "i4")])
    calls = []

    def fake_find_paths(s, g):
        calls.append(len(s))
        r = np.zeros(len(s), dtype=dt)
        r["solved"] = 1
        r["g_final"] = np.linalg.norm(s - g, axis=1)
        r["rank"] = rank
        return r

    res = shard.find_paths_sharded(fake_find_paths, starts, goals, rank, world, dist)
    lo, hi = shard.query_shard(nq, rank, world)
    assert calls == [hi - lo]
    np.save(os.path.join(out_dir, f"r{rank}.npy"), res)
    dist.barrier()
    dist.destroy_process_group()


def test_query_shards_cover_everything_once():
    sys.path.insert(0, ROOT)
    import astar_planners_b200  # noqa: F401
    from astar_planners_b200 import shard
    for nq in (0, 1, 7, 10000, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard.query_shard(nq, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == nq
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_queries_two_ranks_gloo(tmp_path):
    nq, world = 1001, 2
    mp.spawn(_worker, args=(world, _free_port(), nq, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(r0, r1) and len(r0) == nq                      # every rank ends with the full result set
    rng = np.random.default_rng(0)
    starts, goals = rng.random((nq, 3)), rng.random((nq, 3))
    assert np.array_equal(r0["g_final"], np.linalg.norm(starts - goals, axis=1))   # order preserved
    assert (r0["rank"][:500] == 0).all() and (r0["rank"][500:] == 1).all()         # contiguous shards


def _slab_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import astar_planners_b200  # noqa: F401
    from astar_planners_b200 import shard
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    grid = (rng.random((12, 10, 48)) < 0.02).astype(np.uint8)
    grid[:, :, 16:32] = 0 if world == 3 else grid[:, :, 16:32]   # middle slab empty: gaps must skip over it
    nz = 48 // world
    mine = grid[:, :, rank * nz:(rank + 1) * nz]
    has = mine.any(axis=2)
    lo = np.where(has, mine.argmax(axis=2), -1).astype(np.int32)
    hi = np.where(has, nz - 1 - mine[:, :, ::-1].argmax(axis=2), -1).astype(np.int32)
    parts = [None] * world
    dist.all_gather_object(parts, (lo, hi))            # the one exchange step of the slab EDT (tiny)
    below, above = shard.slab_gaps([p[0] for p in parts], [p[1] for p in parts], [nz] * world, rank)
    np.savez(os.path.join(out_dir, f"s{rank}.npz"), below=below, above=above)
    dist.barrier()
    dist.destroy_process_group()


def test_slab_halo_exchange_three_ranks_gloo(tmp_path):
    world = 3
    mp.spawn(_slab_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(3)
    grid = (rng.random((12, 10, 48)) < 0.02).astype(np.uint8)
    grid[:, :, 16:32] = 0
    nz = 16
    for r in range(world):
        got = np.load(tmp_path / f"s{r}.npz")
        z_first, z_last = r * nz, (r + 1) * nz - 1
        below = np.full(grid.shape[:2], -1)
        above = np.full(grid.shape[:2], -1)
        for x in range(grid.shape[0]):
            for y in range(grid.shape[1]):
                zs = np.nonzero(grid[x, y])[0]
                lower, upper = zs[zs < z_first], zs[zs > z_last]
                if len(lower):
                    below[x, y] = z_first - lower.max()
                if len(upper):
                    above[x, y] = upper.min() - z_last
        assert np.array_equal(got["below"], below) and np.array_equal(got["above"], above)


def test_planner_pool_schedules_batches_round_robin_and_in_order():
    """Host logic of b200.PlannerPool.find_paths_many without a GPU: batch i runs on lane i % lanes, the batches of a lane in
    submission order, results come back in batch order, and a failing lane's exception reaches the caller."""
    import threading
    import astar_planners_b200 as b200

    class FakePlanner:
        def __init__(self, lane, log, fail_on=None):
            self.lane, self.log, self.fail_on = lane, log, fail_on

        def findPaths(self, starts, goals, results=None, **kw):
            if self.fail_on is not None and starts == self.fail_on:
                raise RuntimeError("lane failure")
            self.log.append((self.lane, starts, threading.get_ident()))
            return ("res", starts, goals, results)

        def close(self):
            pass

    log = []
    pool = object.__new__(b200.PlannerPool)  # the constructor needs a device; the scheduling does not
    pool.lanes, pool.streams = 3, []
    pool.planners = [FakePlanner(l, log) for l in range(3)]
    batches = [(i, -i) if i % 2 else (i, -i, f"out{i}") for i in range(8)]
    out = pool.find_paths_many(batches)
    assert [o[1] for o in out] == list(range(8)) and [o[2] for o in out] == [-i for i in range(8)]
    assert [o[3] for o in out] == [None if i % 2 else f"out{i}" for i in range(8)]
    for lane in range(3):
        mine = [b for (l, b, _) in log if l == lane]
        assert mine == [i for i in range(8) if i % 3 == lane]            # right lane, submission order
        assert len({t for (l, _, t) in log if l == lane}) == 1            # one host thread per lane
    assert pool.find_paths_many([]) == []
    pool.planners = [FakePlanner(l, log, fail_on=4) for l in range(3)]
    with pytest.raises(RuntimeError):
        pool.find_paths_many(batches)
    pool.planners = []
