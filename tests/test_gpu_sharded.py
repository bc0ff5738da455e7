"""-m gpu: the N > 1 query path on real kernels — `shard.find_paths_sharded` run by two "ranks" on ONE GPU (two map
replicas + two planner handles, each on its own CUDA stream, the ranks in two threads joined by a stand-in for
torch.distributed.all_gather_object).  The assembled result must equal the unsharded run and the oracle."""
import threading

import numpy as np
import pytest

from util import assert_results_equal, make_pair

pytestmark = pytest.mark.gpu


class _ThreadDist:
    """all_gather_object for `world` threads of one process (what torch.distributed does across ranks)."""

    def __init__(self, world):
        self.world, self.slots = world, {}
        self.barrier = threading.Barrier(world)
        self.tls = threading.local()

    def all_gather_object(self, out, obj):
        self.slots[self.tls.rank] = obj
        self.barrier.wait()
        for r in range(self.world):
            out[r] = self.slots[r]
        self.barrier.wait()


@pytest.mark.parametrize("algo,sem", [("lazythetastar", "canonical"), ("thetastar", "faithful")])
def test_sharded_queries_on_two_handles_equal_the_unsharded_run(b200, po, synth, algo, sem):
    import torch
    from astar_planners_b200 import shard
    world, nq = 2, 600
    maps, planners, streams = [], [], []
    om = c = None
    for r in range(world):
        gm, om, c = make_pair(b200, po, synth, "C3")       # every rank rebuilds the replica from the same cloud
        st = torch.cuda.Stream()
        gm.set_stream(st.cuda_stream)
        pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=200000, semantics=sem)
        pl.setGridMap(gm)
        maps.append(gm); planners.append(pl); streams.append(st)
    s, g = synth.random_queries(9, om.get_inflate_occupancy, c["origin"], c["size"], nq, min_dist=3.0)
    dist = _ThreadDist(world)
    out, errs = [None] * world, []

    def rank_main(r):
        try:
            dist.tls.rank = r
            out[r] = shard.find_paths_sharded(lambda a, b: planners[r].findPaths(a, b)[0], s, g, r, world, dist)
        except Exception as e:  # surface failures of a rank thread in the test
            errs.append(e)
            dist.barrier.abort()

    th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    whole, _ = planners[0].findPaths(s, g)
    mode = po.FAITHFUL if sem == "faithful" else po.CANONICAL
    ro, _ = po.OraclePlanner(algo, om, 0.1, 5.0, 200000, mode=mode).find_paths(s, g, nthreads=8)
    for r in range(world):
        assert len(out[r]) == nq
        for f in ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords"):
            assert np.array_equal(out[r][f], whole[f]), (r, f)
        assert_results_equal(out[r], ro)


@pytest.mark.parametrize("algo,sem", [("lazythetastar", "canonical"), ("astar", "faithful")])
def test_planners_on_their_own_streams_overlap_on_one_map(b200, po, synth, algo, sem):
    """b200_planner_set_stream: three planner handles on ONE map, each with its own CUDA stream and host thread, each
    running the same batch twice (paths included) — every result equals the single-stream run and the oracle."""
    import torch
    gm, om, c = make_pair(b200, po, synth, "C3")
    nq = 1500
    s, g = synth.random_queries(11, om.get_inflate_occupancy, c["origin"], c["size"], nq, min_dist=3.0)
    mk = lambda: b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=200000, semantics=sem)
    base = mk()
    base.setGridMap(gm)
    want, want_path = base.findPaths(s, g, path_cap=nq * 400)
    gm.sync()
    lanes = 3
    streams = [torch.cuda.Stream() for _ in range(lanes)]
    pls = []
    for st in streams:
        pl = mk()
        pl.setGridMap(gm)
        pl.set_stream(st.cuda_stream)
        pls.append(pl)
    got, errs = [None] * lanes, []

    def lane(i):
        try:
            for _ in range(2):
                got[i] = pls[i].findPaths(s, g, path_cap=nq * 400)
        except Exception as e:
            errs.append(e)

    th = [threading.Thread(target=lane, args=(i,)) for i in range(lanes)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    mode = po.FAITHFUL if sem == "faithful" else po.CANONICAL
    ro, _ = po.OraclePlanner(algo, om, 0.1, 5.0, 200000, mode=mode).find_paths(s, g, nthreads=8)
    for i in range(lanes):
        res, path = got[i]
        for f in ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords"):
            assert np.array_equal(res[f], want[f]), (i, f)
        for q in range(0, nq, 97):  # path blocks land at batch-dependent offsets: compare per query
            assert np.array_equal(b200.path_of(res, path, q), b200.path_of(want, want_path, q)), (i, q)
        assert_results_equal(res, ro)
    pls[0].set_stream(None)  # back to the map's stream
    again, _ = pls[0].findPaths(s, g)
    assert np.array_equal(again["g_final"], want["g_final"])


def test_planner_pool_runs_batches_on_lanes_in_order(b200, po, synth):
    """b200.PlannerPool: five different batches over two lanes, host arrays in — every batch equals its single-handle run."""
    gm, om, c = make_pair(b200, po, synth, "C3")
    batches = []
    for k in range(5):
        s, g = synth.random_queries(20 + k, om.get_inflate_occupancy, c["origin"], c["size"], 300 + 50 * k, min_dist=3.0)
        batches.append((s, g))
    kw = dict(resolution=0.1, lambda_heuristic=5.0, allocate_num=200000, semantics="canonical")
    pool = b200.PlannerPool(gm, "lazythetastar", lanes=2, **kw)
    got = pool.find_paths_many(batches)
    pool.close()
    single = b200.Planner("lazythetastar", **kw)
    single.setGridMap(gm)
    for (s, g), (res, _) in zip(batches, got):
        want, _ = single.findPaths(s, g)
        assert len(res) == len(s)
        for f in ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords"):
            assert np.array_equal(res[f], want[f]), f
