"""CPU, only where /root/reference exists: randomised diff of the oracle (FAITHFUL) against the reference's own
compiled sources (oracle/_ref/libref.so) on worlds other than the golden one."""
import numpy as np
import pytest

from oracle import pyref as pr
from util import set_positions as _set_positions

pytestmark = pytest.mark.skipif(not pr.available(), reason="/root/reference not present (GPU box): golden vectors cover this")


@pytest.mark.parametrize("seed,size,res,infl", [(101, (25.6, 25.6, 6.4), 0.1, 0.099), (102, (20.0, 16.0, 6.0), 0.2, 0.099),
                                                  (103, (12.8, 12.8, 3.2), 0.1, 0.25)])
def test_map_los_and_planners_match_reference(po, synth, seed, size, res, infl):
    gh = -0.01
    origin = (-size[0] / 2, -size[1] / 2, gh)
    cloud = synth.pack_cloud(np.concatenate([synth.forest(seed, origin, size, res, 30),
                                             synth.uniform_cloud(seed + 1, origin, size, 2000, margin=0.5)]))
    rm = pr.RefMap(size, res, infl, (1e3, 1e3, 1e3), gh)
    om = po.OracleMap(origin, size, res, infl, (1e3, 1e3, 1e3), gh)
    assert rm.dims == om.dims and np.allclose(rm.origin, origin)
    rm.update_cloud(cloud, (0.3, -0.2, 1.0))
    om.update_cloud(cloud, (0.3, -0.2, 1.0))
    assert np.array_equal(rm.inflate(), om.inflate())
    assert rm.local_bounds() == om.local_bounds()
    s, g = synth.random_queries(seed + 2, om.get_inflate_occupancy, origin, size, 40, min_dist=3.0)
    assert np.array_equal(rm.los_batch(s, g), om.los_batch(s, g))
    for algo in ("astar", "thetastar"):
        rr, rpaths = pr.RefPlanner(algo, rm, res, 5.0, 100000).find_paths(s, g)
        ro, pth = po.OraclePlanner(algo, om, res, 5.0, 100000, mode=po.FAITHFUL).find_paths(s, g, path_stride=8192, nthreads=4)
        for f in ("solved", "n_path", "explored_nodes", "iter_num", "los_checks", "g_final", "path_length", "goal_coords"):
            assert np.array_equal(rr[f], ro[f]), (algo, f)
        for i in range(len(s)):
            assert np.array_equal(pth[i * 8192: i * 8192 + ro["n_path"][i]], rpaths[i]), (algo, i)
        # getVisitedNodes (AlgorithmBase.cpp:185-189): every created node but the last, in pool order
        rp, op = pr.RefPlanner(algo, rm, res, 5.0, 100000), po.OraclePlanner(algo, om, res, 5.0, 100000, mode=po.FAITHFUL)
        r1, _ = rp.find_path(s[0], g[0])
        o1, _ = op.find_path(s[0], g[0])
        vr, vo = rp.visited(), op.visited()
        assert len(vr) == int(r1["explored_nodes"]) - 1 and np.array_equal(vr, vo)


def test_clear_box_matches_reference(po, synth):
    size, res = (12.8, 12.8, 3.2), 0.1
    origin = (-6.4, -6.4, -0.01)
    rm = pr.RefMap(size, res)
    om = po.OracleMap(origin, size, res)
    grid = (np.random.default_rng(1).random(om.dims) < 0.2).astype(np.uint8)
    rm.set_inflate(grid)
    om.set_inflate(grid)
    for mn, mx in [((-3.05, -2.0, 0.33), (4.0, 1.0, 2.71)), ((-100, -100, -100), (0, 0, 0)), ((5, 5, 2), (4, 4, 1))]:
        rm.clear_box(mn, mx)
        om.clear_box(mn, mx)
        assert np.array_equal(rm.inflate(), om.inflate())


def test_map_publisher_matches_reference(po, synth):
    size, res = (12.8, 12.8, 3.2), 0.1
    origin = (-6.4, -6.4, -0.01)
    cloud = synth.pack_cloud(synth.forest(21, origin, size, res, 22))
    for rng, cam in [((1e3, 1e3, 1e3), (0, 0, 1)), ((3.0, 2.5, 1.2), (1.7, -2.2, 0.9))]:
        rm = pr.RefMap(size, res, 0.099, rng, -0.01)
        om = po.OracleMap(origin, size, res, 0.099, rng, -0.01)
        rm.update_cloud(cloud, cam)
        om.update_cloud(cloud, cam)
        for all_info in (False, True):
            ref = rm.publish_inflate(all_info)                 # grid_map/visualization_truncate_height 2.4, local_map_margin 1
            mine, n = om.export_inflate_cloud(1 if all_info else 0, 2.4)
            assert n == len(ref) and n > 1000
            assert np.array_equal(ref.view(np.uint32), mine.view(np.uint32))


def test_voxel_traversal_matches_the_reference_raycaster(po):
    """The RayCaster restatement (used by depth fusion and by the DDA line of sight) against the reference's own
    RayCaster::setInput/step (raycast.cpp:253-346): identical cell sequences, including axis-aligned, degenerate and
    negative-coordinate rays."""
    pr = pytest.importorskip("oracle.pyref")
    if not pr.available():
        pytest.skip("reference sources not present")
    r = np.random.default_rng(3)
    a = r.uniform(-40, 40, (400, 3))
    b = r.uniform(-40, 40, (400, 3))
    a[:40, 0] = b[:40, 0]                    # no motion along x
    a[40:60, 1:] = b[40:60, 1:]              # axis-aligned
    a[60:70] = b[60:70]                      # same point
    a[70:100] = np.floor(a[70:100])          # starts exactly on lattice planes
    b[100:130] = np.floor(b[100:130]) + 0.5  # ends at cell centres
    for i in range(len(a)):
        got, want = po.ray_cells(a[i], b[i]), pr.ray_cells(a[i], b[i])
        assert np.array_equal(got, want), (i, a[i], b[i], len(got), len(want))


def test_set_occupied_and_set_occupancy_match_reference(po):
    size, res, gh = (6.4, 4.8, 3.2), 0.1, -0.01
    origin = (-size[0] / 2, -size[1] / 2, gh)
    rm = pr.RefMap(size, res, 0.099, (1e3, 1e3, 1e3), gh)
    om = po.OracleMap(origin, size, res, 0.099, (1e3, 1e3, 1e3), gh)
    om.set_depth_params()
    pos, occ = _set_positions(5, origin, size, 4000)
    rm.set_occupied(pos[:2000])
    om.set_occupied(pos[:2000])
    assert np.array_equal(rm.inflate(), om.inflate()) and rm.inflate().sum() > 1000
    rm.set_occupancy(pos, occ)
    om.set_occupancy(pos, occ)
    a, b = rm.occupancy(), om.occupancy()
    assert np.array_equal(a, b) and (a == 1.0).sum() > 500 and (a == 0.0).sum() > 200


def test_path_distance_metrics_arithmetic_matches_reference():
    """GetPath.srv:20-23: the product's host arithmetic (b200_distance_metrics, pure host code in the C-ABI library) against
    the reference's own calculateDistancesMetrics on random, constant, single-element and empty inputs."""
    import astar_planners_b200 as b200
    rng = np.random.default_rng(4)
    cases = [rng.random(n) * s for n, s in ((1, 3.0), (2, 1.0), (7, 0.3), (300, 5.0), (5000, 12.0))]
    cases += [np.full(40, 0.7), np.array([]), np.sqrt(rng.integers(0, 5000, 900).astype(np.float64)) * 0.1]
    for d in cases:
        m = b200.distance_metrics(d)
        assert (m["mean"], m["std"], m["min"], m["max"]) == pr.distance_metrics(d)


def test_depth_odom_pose_algebra_matches_reference():
    """depthOdomCallback (grid_map.cpp:965-982): body odometry -> camera position and md_.camera_q_ — the product's host
    arithmetic (b200_odom_to_camera_pose, pure host code in the C-ABI library) against the reference's own callback on random
    orientations, rotations near 180 degrees about every axis (the three non-trace branches of matrix -> quaternion),
    axis-aligned ones and unnormalised quaternions; bit for bit."""
    import astar_planners_b200 as b200
    rm = pr.RefMap((4.0, 4.0, 2.0), 0.1)  # scratch map: the callback stores a 1x1 image in it
    rng = np.random.default_rng(12)
    quats = [q / np.linalg.norm(q) for q in rng.normal(size=(300, 4))]
    for ax in range(3):  # near-pi rotations: trace <= 0, largest diagonal element = ax
        for _ in range(40):
            v = rng.normal(size=3) * 0.05
            v[ax] = 1.0
            v /= np.linalg.norm(v)
            ang = np.pi - abs(rng.normal()) * 0.02
            quats.append(np.concatenate([[np.cos(ang / 2)], np.sin(ang / 2) * v]))
    quats += [np.array(q, dtype=np.float64) for q in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (0.5, 0.5, 0.5, 0.5),
                                                      (np.sqrt(0.5), 0, 0, np.sqrt(0.5)), (0.9, 0.1, -0.3, 0.2), (2.0, 0.0, 0.0, 0.1))]
    branches = set()
    for q in quats:
        pos = rng.normal(size=3) * 3.0
        cp, cq = rm.depth_odom_pose(pos, q)
        gp, gq, gr = b200.odom_to_camera_pose(pos, q)
        assert np.array_equal(cp.view(np.uint64), gp.view(np.uint64)), (q, cp, gp)
        assert np.array_equal(cq.view(np.uint64), gq.view(np.uint64)), (q, cq, gq)
        w, x, y, z = gq  # the rotation projectDepthImage derives (:208): Eigen's toRotationMatrix order, restated here
        tx, ty, tz = 2.0 * x, 2.0 * y, 2.0 * z
        twx, twy, twz, txx, txy, txz, tyy, tyz, tzz = tx * w, ty * w, tz * w, tx * x, ty * x, tz * x, ty * y, tz * y, tz * z
        want = np.array([[1.0 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1.0 - (txx + tzz), tyz - twx],
                         [txz - twy, tyz + twx, 1.0 - (txx + tyy)]])
        assert np.array_equal(want.view(np.uint64), gr.view(np.uint64))
        branches.add("trace" if gq[0] >= 0.5 * np.sqrt(1e-30) and abs(gq[0]) == max(abs(gq)) else int(np.argmax(np.abs(gq[1:]))))
    assert {0, 1, 2} <= branches and "trace" in branches
