"""Depth-image fusion: the oracle's restatement against the reference's own code — the committed
golden vectors everywhere, and the live reference build where /root/reference exists."""
import numpy as np
import pytest

from depth_util import GOLDEN, SCENARIOS, check_publishers, replay, scenario


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", SCENARIOS)
def test_oracle_depth_fusion_is_the_reference(po, golden, name):
    sc = scenario(golden, name)
    om = po.OracleMap(tuple(golden["origin"]), tuple(golden["size"]), float(golden["res"]), sc["inflation"], sc["range"],
                      float(golden["origin"][2]))
    om.set_depth_params(**sc["params"])
    replay(sc, om.update_depth, om.occupancy, om.inflate, om.local_bounds, om.dims, om.set_raycast_num)
    assert om.raycast_num() == sc["frames_after"]
    check_publishers(sc, om.export_occupancy_cloud(0, sc["params"]["local_map_margin"] // 2), om.export_occupancy_cloud(1, 0))
    st = om.depth_stats()
    assert st[0] == sc["proj"][-1]
    if name == "count_wrap":
        # more log-odds updates than voxels touched: a `short` counter wrapped past 1 again and re-queued its voxel
        assert int((sc["img"][0] == 0).sum()) > 65536
        fresh = po.OracleMap(tuple(golden["origin"]), tuple(golden["size"]), float(golden["res"]), sc["inflation"], sc["range"],
                             float(golden["origin"][2]))
        fresh.set_depth_params(**sc["params"])
        fresh.update_depth(sc["img"][0], sc["pos"][0], sc["rot"][0])
        occ = fresh.occupancy()
        assert fresh.depth_stats()[3] > int((occ != occ.min()).sum())
    # a camera outside the map drops the frame (grid_map.cpp:688-697)
    before = om.occupancy()
    assert not om.update_depth(sc["img"][0], (0.0, 0.0, float(golden["size"][2]) + 1.0), sc["rot"][0])
    assert np.array_equal(before, om.occupancy())


def test_get_occupancy_matches_threshold(po, golden):
    sc = scenario(golden, "filter")
    om = po.OracleMap(tuple(golden["origin"]), tuple(golden["size"]), float(golden["res"]), sc["inflation"], sc["range"],
                      float(golden["origin"][2]))
    om.set_depth_params(**sc["params"])
    for f in range(len(sc["pos"])):
        om.update_depth(sc["img"][f], sc["pos"][f], sc["rot"][f])
    occ = om.occupancy()
    idx = np.argwhere(occ > np.log(0.8 / 0.2))
    assert len(idx) > 50
    pos = (idx + 0.5) * float(golden["res"]) + golden["origin"]
    assert (om.get_occupancy(pos) == 1).all()
    assert (om.get_occupancy(pos + [0, 0, 100.0]) == -1).all()


def test_oracle_depth_vs_live_reference(po, synth):
    pr = pytest.importorskip("oracle.pyref")
    if not pr.available():
        pytest.skip("reference sources not present")
    size, res, gh = (8.0, 6.0, 2.5), 0.1, -0.01
    origin = (-4.0, -3.0, gh)
    dp = dict(po.DEPTH_DEFAULTS)
    dp.update(fx=60.5, fy=60.5, cx=50.2, cy=38.1, depth_filter_margin=2, skip_pixel=3, local_map_margin=2, virtual_ceil_height=2.2,
              max_ray_length=3.5)
    rm = pr.RefMap(size, res, 0.099, (3.0, 3.0, 2.0), gh, depth=dp)
    om = po.OracleMap(origin, size, res, 0.099, (3.0, 3.0, 2.0), gh)
    om.set_depth_params(**dp)
    boxes = synth.pillar_boxes(5, origin, size, 12)
    rng = np.random.default_rng(9)
    for f in range(7):
        # frames 0-4: a slowly moving camera (surfaces seen repeatedly become occupied); 5-6: somewhere else entirely
        if f < 5:
            pos = (-2.0 + 0.08 * f + rng.uniform(-0.02, 0.02), 1.0 + rng.uniform(-0.02, 0.02), 1.1)
            R, q = synth.look_at(-0.6 + 0.03 * f, 0.2)
        else:
            pos = (rng.uniform(-3.5, 3.5), rng.uniform(-2.5, 2.5), rng.uniform(0.3, 2.0))
            R, q = synth.look_at(rng.uniform(0, 6.28), rng.uniform(-0.4, 0.6))
        img = synth.render_depth(boxes, gh, pos, R, 76, 100, dp["fx"], dp["fy"], dp["cx"], dp["cy"], dropout=0.05, seed=f)
        fused, Rref = rm.update_depth(img, pos, q)
        assert om.update_depth(img, pos, Rref) == fused
        assert np.array_equal(rm.occupancy(), om.occupancy())
        assert np.array_equal(rm.inflate(), om.inflate())
        assert rm.proj_points() == om.depth_stats()[0]
        if rm.proj_points():
            assert rm.local_bounds() == om.local_bounds()
            # the two raw-occupancy publishers (grid_map.cpp:806-849, :902-939)
            assert np.array_equal(rm.publish_occupancy(0), om.export_occupancy_cloud(0, dp["local_map_margin"] // 2))
            assert np.array_equal(rm.publish_occupancy(1), om.export_occupancy_cloud(1, 0))
        if f == 4:
            assert len(om.export_occupancy_cloud(0, 1)) > 20 and len(om.export_occupancy_cloud(1, 0)) > 1000


def _sequential_reach(paths):
    """The reference's loop (grid_map.cpp:397-419): ray i walks its path, counts every voxel it reaches, and stops at
    (and including) the first voxel an earlier ray has flagged; voxels it passes get flagged."""
    flagged, reach = set(), []
    for p in paths:
        n = 0
        for v in p:
            n += 1
            if v in flagged:
                break
            flagged.add(v)
        reach.append(n)
    return reach


def _fixpoint_reach(paths):
    """The order-free formulation the CUDA path uses (csrc/depth_kernels.cu): F(v) = smallest ray id reaching v, iterated
    from +inf until it stops changing; ray i stops at the first voxel with F(v) < i.  Returns (reach, rounds)."""
    inf = len(paths)
    F, rounds = {}, 0
    while True:
        rounds += 1
        Fn = {}
        for i, p in enumerate(paths):
            for v in p:
                if F.get(v, inf) < i:
                    break
                if Fn.get(v, inf) > i:
                    Fn[v] = i
        if Fn == F:
            break
        F = Fn
        assert rounds <= len(paths) + 2
    reach = []
    for i, p in enumerate(paths):
        n = 0
        for v in p:
            n += 1
            if F.get(v, inf) < i:
                break
        reach.append(n)
    return reach, rounds


def test_ray_order_fixed_point_equals_the_sequential_loop(po):
    """The claim behind the parallel depth fusion, checked on the host: for ANY set of ordered paths the fixed point of
    'smallest ray id reaching each voxel' reproduces the sequential first-come flags exactly — on rays to a common
    camera (voxel sequences from the RayCaster restatement) and on arbitrary random paths."""
    r = np.random.default_rng(5)
    cam = np.array([20.3, 17.8, 9.1])
    ends = cam + r.normal(0, 12, (600, 3))
    paths = [[tuple(c) for c in po.ray_cells(e, cam, cap=256)] for e in ends]
    want = _sequential_reach(paths)
    got, rounds = _fixpoint_reach(paths)
    assert got == want and 2 <= rounds <= 60
    assert sum(want) < sum(len(p) for p in paths)  # rays really do stop early
    for trial in range(20):  # no geometry at all: random repeat-free paths over a small voxel set, many collisions
        paths = [list(r.permutation(40)[: r.integers(1, 12)]) for _ in range(60)]
        got, _ = _fixpoint_reach(paths)
        assert got == _sequential_reach(paths), trial
    # a path that meets one of its OWN earlier voxels stops there in the sequential loop (its own flag); the parallel
    # form gets the same by cutting every path after its first repeated voxel (depth_path_count_kernel does that)
    paths = [[1, 2, 2, 3, 4], [5, 2, 6], [7, 8, 7, 9]]
    cut = []
    for p in paths:
        seen, q = set(), []
        for v in p:
            q.append(v)
            if v in seen:
                break
            seen.add(v)
        cut.append(q)
    assert _fixpoint_reach(cut)[0] == _sequential_reach(paths) == [3, 2, 3]
