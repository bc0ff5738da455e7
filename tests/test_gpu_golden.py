"""-m gpu: the CUDA path against golden vectors produced by the reference's OWN code (tests/golden/gen_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_golden.npz"))
SIZE, ORIGIN, RES = tuple(G["size"]), tuple(G["origin"]), float(G["res"])
DIMS = (128, 128, 32)


def unbits(a):
    return np.unpackbits(a)[: np.prod(DIMS)].reshape(DIMS)


@pytest.fixture(scope="module")
def gmap(b200, synth):
    m = b200.GridMap(SIZE, RES, obstacles_inflation=0.099, local_update_range=(1e3, 1e3, 1e3), ground_height=ORIGIN[2])
    assert np.allclose(m.origin, ORIGIN)          # origin=None reproduces initMap's (-sx/2, -sy/2, ground_height)
    m.cloudCallback(synth.pack_cloud(G["cloud"]), tuple(G["cam1"]))
    return m


def test_grid_bounds_and_reads(gmap):
    assert gmap.dims == DIMS
    assert np.array_equal(gmap.inflate(), unbits(G["grid1_bits"]))
    assert np.array_equal(np.array(gmap.local_bounds()), G["bounds1"])
    assert np.array_equal(gmap.getInflateOccupancy(G["pos"]), G["pos_occ"])
    assert np.array_equal(gmap.isInMap(G["pos"]), G["pos_inmap"])
    assert np.array_equal(gmap.posToIndex(G["pos"]), G["pos_index"])


def test_second_frame_and_inflation_two(b200, synth):
    m = b200.GridMap(SIZE, RES, local_update_range=tuple(G["range2"]), ground_height=ORIGIN[2])
    m.set_inflate(unbits(G["grid1_bits"]))
    m.cloudCallback(synth.pack_cloud(G["cloud2"]), tuple(G["cam2"]))
    assert np.array_equal(m.inflate(), unbits(G["grid2_bits"]))
    assert np.array_equal(np.array(m.local_bounds()), G["bounds2"])
    m3 = b200.GridMap(SIZE, RES, obstacles_inflation=0.15, local_update_range=(1e3, 1e3, 1e3), ground_height=ORIGIN[2])
    m3.cloudCallback(synth.pack_cloud(G["cloud"]), tuple(G["cam1"]))
    assert np.array_equal(m3.inflate(), unbits(G["grid3_bits"]))


def test_fast_los(gmap):
    assert np.array_equal(gmap.los_batch(G["los_a"], G["los_b"]), G["los_vis"])


FIELDS = ("solved", "status", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords",
          "start_coords")


def _check(res, gold):
    for f in FIELDS:
        a, b = np.ascontiguousarray(res[f]), np.ascontiguousarray(gold[f])
        if a.dtype.kind == "f":
            assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), f
        else:
            assert np.array_equal(a, b), f


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
@pytest.mark.parametrize("lam", [5, 1])
def test_planners_are_bit_identical_to_the_reference(b200, gmap, algo, lam):
    gold = G[f"{algo}_lam{lam}_res"]
    n = len(gold)
    pl = b200.Planner(algo, resolution=RES, lambda_heuristic=float(lam), allocate_num=100000)   # default = faithful
    pl.setGridMap(gmap)
    res, path = pl.findPaths(G["q_start"][:n], G["q_goal"][:n], path_cap=int(gold["n_path"].sum()) + 16)
    _check(res, gold)
    flat = np.concatenate([b200.path_of(res, path, i) for i in range(n) if res["n_path"][i] > 0])
    assert np.array_equal(flat.view(np.uint64), G[f"{algo}_lam{lam}_path"].view(np.uint64))


@pytest.mark.parametrize("algo", ["astar", "thetastar"])
def test_failure_modes_match_the_reference(b200, gmap, algo):
    pl = b200.Planner(algo, resolution=RES, lambda_heuristic=5.0, allocate_num=300)
    pl.setGridMap(gmap)
    res, _ = pl.findPaths(G["q_start"][:16], G["q_goal"][:16])
    _check(res, G[f"{algo}_oom_res"])
    grid = np.zeros(DIMS, np.uint8)
    grid[40:52, 40:52, 5:17] = 1
    grid[42:50, 42:50, 7:15] = 0
    m = b200.GridMap(SIZE, RES, ground_height=ORIGIN[2])
    m.set_inflate(grid)
    pl = b200.Planner(algo, resolution=RES, lambda_heuristic=5.0, allocate_num=100000)
    pl.setGridMap(m)
    res, _ = pl.findPaths(G["sealed_start"], G["sealed_goal"])
    _check(res, G[f"{algo}_sealed_res"])


@pytest.mark.parametrize("algo", ["thetastaragr", "thetastaragrpp"])
def test_air_ground_planners_are_bit_identical_to_the_reference(b200, gmap, algo):
    gold = G[f"{algo}_res"]
    n = len(gold)
    pl = b200.Planner(algo, resolution=RES, lambda_heuristic=5.0, allocate_num=100000)   # launch-file AGR parameters
    pl.setGridMap(gmap)
    res, path = pl.findPaths(G["agr_start"], G["agr_goal"], path_cap=int(gold["n_path"].sum()) + 16)
    _check(res, gold)
    flat = np.concatenate([b200.path_of(res, path, i) for i in range(n) if res["n_path"][i] > 0])
    assert np.array_equal(flat.view(np.uint64), G[f"{algo}_path"].view(np.uint64))
