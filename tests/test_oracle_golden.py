"""CPU: the oracle's FAITHFUL restatement vs golden vectors produced by the reference's own code
(tests/golden/gen_golden.py -> ref_golden.npz).  This is what pins the oracle."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
SIZE, ORIGIN, RES = tuple(G["size"]), tuple(G["origin"]), float(G["res"])
DIMS = (128, 128, 32)


def unbits(a):
    return np.unpackbits(a)[: np.prod(DIMS)].reshape(DIMS)


@pytest.fixture(scope="module")
def omap(po, synth):
    m = po.OracleMap(ORIGIN, SIZE, RES, 0.099, (1e3, 1e3, 1e3), ORIGIN[2])
    m.update_cloud(synth.pack_cloud(G["cloud"]), tuple(G["cam1"]))
    return m


def test_cloud_update_grid_and_bounds(omap):
    assert omap.dims == DIMS
    assert np.array_equal(omap.inflate(), unbits(G["grid1_bits"]))
    assert np.array_equal(np.array(omap.local_bounds()), G["bounds1"])


def test_second_frame_local_range(po, synth, omap):
    m = po.OracleMap(ORIGIN, SIZE, RES, 0.099, tuple(G["range2"]), ORIGIN[2])
    m.set_inflate(unbits(G["grid1_bits"]))
    m.update_cloud(synth.pack_cloud(G["cloud2"]), tuple(G["cam2"]))
    assert np.array_equal(m.inflate(), unbits(G["grid2_bits"]))
    assert np.array_equal(np.array(m.local_bounds()), G["bounds2"])


def test_inflation_step_two(po, synth):
    m = po.OracleMap(ORIGIN, SIZE, RES, 0.15, (1e3, 1e3, 1e3), ORIGIN[2])
    m.update_cloud(synth.pack_cloud(G["cloud"]), tuple(G["cam1"]))
    assert np.array_equal(m.inflate(), unbits(G["grid3_bits"]))


def test_position_reads(omap):
    assert np.array_equal(omap.get_inflate_occupancy(G["pos"]), G["pos_occ"])
    assert np.array_equal(omap.is_in_map(G["pos"]), G["pos_inmap"])
    assert np.array_equal(omap.pos_to_index(G["pos"]), G["pos_index"])


def test_fast_los(omap):
    assert np.array_equal(omap.los_batch(G["los_a"], G["los_b"]), G["los_vis"])
    assert 0.05 < G["los_vis"].mean() < 0.95


FIELDS = ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords",
          "start_coords")


def _check(res, gold, path=None, gold_path=None):
    for f in FIELDS:
        a, b = np.ascontiguousarray(res[f]), np.ascontiguousarray(gold[f])
        if a.dtype.kind == "f":  # bit-exact, not just ==
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f
        else:
            assert np.array_equal(a, b), f
    if path is not None:
        assert np.array_equal(path.view(np.uint64), gold_path.view(np.uint64))


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
@pytest.mark.parametrize("lam", [5, 1])
def test_planners_faithful_mode_is_the_reference(po, omap, algo, lam):
    gold = G[f"{algo}_lam{lam}_res"]
    n = len(gold)
    pl = po.OraclePlanner(algo, omap, RES, float(lam), 100000, mode=po.FAITHFUL)
    res, paths = pl.find_paths(G["q_start"][:n], G["q_goal"][:n], path_stride=8192, nthreads=4)
    flat = np.concatenate([paths[i * 8192: i * 8192 + res["n_path"][i]] for i in range(n)])
    _check(res, gold, flat, G[f"{algo}_lam{lam}_path"])
    assert gold["solved"].all() or lam == 1   # lambda = 1 exhausts the 100k-node pool on one query (status 2), also pinned


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
def test_pool_exhaustion_and_sealed_goal(po, omap, algo):
    gold = G[f"{algo}_oom_res"]
    pl = po.OraclePlanner(algo, omap, RES, 5.0, 300, mode=po.FAITHFUL)
    res, _ = pl.find_paths(G["q_start"][:16], G["q_goal"][:16], nthreads=2)
    _check(res, gold)
    assert (gold["status"] == 2).any()
    grid = np.zeros(DIMS, np.uint8)
    grid[40:52, 40:52, 5:17] = 1
    grid[42:50, 42:50, 7:15] = 0
    m = po.OracleMap(ORIGIN, SIZE, RES)
    m.set_inflate(grid)
    res, _ = po.OraclePlanner(algo, m, RES, 5.0, 100000, mode=po.FAITHFUL).find_paths(G["sealed_start"], G["sealed_goal"])
    _check(res, G[f"{algo}_sealed_res"])
    assert res["status"][0] == 1


def test_canonical_vs_reference_deviation_is_bounded(po, omap):
    """Documents (not hides) the gap between the clean CANONICAL semantics and the reference's
    rb-tree-with-mutated-keys behaviour: most queries agree exactly, the rest stay close."""
    for algo, min_same, tol in (("astar", 0.6, 0.25), ("thetastar", 0.85, 0.05)):
        gold = G[f"{algo}_lam5_res"]
        pl = po.OraclePlanner(algo, omap, RES, 5.0, 100000, mode=po.CANONICAL)
        res, _ = pl.find_paths(G["q_start"][:64], G["q_goal"][:64], nthreads=4)
        assert res["solved"].all()
        same = (res["g_final"] == gold["g_final"]).mean()
        rel = np.abs(res["g_final"] - gold["g_final"]) / gold["g_final"]
        assert same >= min_same, (algo, same)
        assert rel.max() < tol, (algo, rel.max())


@pytest.mark.parametrize("algo", ["thetastaragr", "thetastaragrpp"])
def test_air_ground_planners_faithful_mode_is_the_reference(po, omap, algo):
    gold = G[f"{algo}_res"]
    n = len(gold)
    for mode in ([po.FAITHFUL] if algo == "thetastaragr" else [po.FAITHFUL, po.CANONICAL]):   # AGRpp: both coincide
        pl = po.OraclePlanner(algo, omap, RES, 5.0, 100000, mode=mode)
        res, paths = pl.find_paths(G["agr_start"], G["agr_goal"], path_stride=8192, nthreads=4)
        flat = np.concatenate([paths[i * 8192: i * 8192 + res["n_path"][i]] for i in range(n)])
        _check(res, gold, flat, G[f"{algo}_path"])
    assert gold["solved"].sum() > 30
