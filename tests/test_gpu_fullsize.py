"""-m gpu: BASELINE.json full-size configurations.  Where the oracle finishes in seconds the comparison is exact;
beyond that, size-independent properties of the results (path validity, cost consistency, determinism)."""
import numpy as np
import pytest

from util import assert_results_equal, make_pair

pytestmark = pytest.mark.gpu


def test_c3_10k_costlazythetastar_queries_exact(b200, po, synth):
    """BASELINE config 3: 10k random start/goal costlazythetastar queries with safety cost on 256x256x64."""
    gm, om, c = make_pair(b200, po, synth, "C3")
    gm.build_edt(d_safe=1.0)
    om.build_edt(d_safe=1.0, nthreads=po.max_threads())
    s, g = synth.random_queries(5, om.get_inflate_occupancy, c["origin"], c["size"], 10000, min_dist=5.0)
    pl = b200.Planner("costlazythetastar", resolution=0.1, lambda_heuristic=5.0, cost_weight=2.0)
    pl.setGridMap(gm)
    rg, pg = pl.findPaths(s, g, path_cap=200000)
    ro, _ = po.OraclePlanner("costlazythetastar", om, 0.1, 5.0, 1000000, mode=po.CANONICAL, cost_weight=2.0).find_paths(
        s, g, nthreads=po.max_threads())
    assert ro["solved"].all()
    assert_results_equal(rg, ro)
    assert (rg["path_offset"] >= 0).all()


def _check_paths(b200, gm, algo, res, path, starts, goals, r=0.1):
    n = len(res)
    assert res["solved"].all()
    seg_a, seg_b, owner = [], [], []
    for i in range(n):
        p = b200.path_of(res, path, i)
        assert p is not None and len(p) == res["n_path"][i] >= 1
        assert np.array_equal(p[0], starts[i])                                    # unsnapped start (A.5)
        assert (np.abs(p[-1] - goals[i]) <= r).all()                              # termination box (B6)
        assert np.array_equal(res["goal_coords"][i], p[-1])
        if len(p) > 1:
            d = np.linalg.norm(np.diff(p, axis=0), axis=1)
            assert abs(d.sum() - res["g_final"][i]) <= 1e-9 * max(1.0, d.sum())   # g is the summed segment length
            assert abs(res["path_length"][i] - d.sum() * gm.resolution) <= 1e-4 * max(1.0, d.sum())   # x res quirk (A.7)
            if algo == "astar":
                assert d.max() <= r * np.sqrt(3) + 1e-9                           # 26-connected lattice steps
            seg_a.append(p[:-1]); seg_b.append(p[1:]); owner.append(np.full(len(p) - 1, i))
    a, b = np.concatenate(seg_a), np.concatenate(seg_b)
    assert (gm.getInflateOccupancy(np.concatenate([a, b])) == 0).all()           # every path node is free space
    if algo != "astar":
        # every any-angle (parent-of-parent) segment was admitted by a fastLOS; lattice steps are never LoS-tested by the
        # reference (ThetaStar.cpp:61-69) and may graze an inflated corner, so they are exempt
        long_seg = np.linalg.norm(a - b, axis=1) > r * np.sqrt(3) + 1e-9
        assert long_seg.sum() > 100
        assert gm.los_batch(a[long_seg], b[long_seg]).all()


@pytest.mark.parametrize("algo", ["astar", "thetastar", "lazythetastar"])
def test_c4_batched_queries_on_the_512x512x128_map(b200, po, synth, algo):
    """BASELINE config 4 (per-GPU shard): batched queries on the replicated 512x512x128 map.  Exact vs the oracle on a
    subsample, properties on the whole batch, and run-to-run determinism."""
    gm, om, c = make_pair(b200, po, synth, "C2", n_points=1_000_000, rng=(60.0, 60.0, 20.0))
    nq = 20000 if algo != "thetastar" else 4000
    s, g = synth.random_queries(6, gm.getInflateOccupancy, c["origin"], c["size"], nq, min_dist=5.0, z_range=(0.3, 3.0))
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, semantics="canonical")
    pl.setGridMap(gm)
    res, path = pl.findPaths(s, g, path_cap=nq * 600)
    _check_paths(b200, gm, algo, res, path, s, g)
    res2, _ = pl.findPaths(s, g)
    for f in ("g_final", "explored_nodes", "iter_num", "los_checks", "n_path"):
        assert np.array_equal(res[f], res2[f])                                    # deterministic despite dynamic scheduling
    sub = slice(0, 1500 if algo != "thetastar" else 300)
    ro, _ = po.OraclePlanner(algo, om, 0.1, 5.0, 1000000, mode=po.CANONICAL).find_paths(s[sub], g[sub], nthreads=po.max_threads())
    assert_results_equal(res[sub], ro)
