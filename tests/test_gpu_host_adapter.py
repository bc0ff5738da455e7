"""-m gpu: the C++ adapter classes (3d-astar-path-planners_b200/host) driven like the reference's ROS node, compared
with the same world through the ctypes binding."""
import json
import math
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "3d-astar-path-planners_b200", "host", "host_selftest")


def _cloud():
    pts = []
    for j in range(40, 88):
        for k in range(32):
            pts.append((np.float32(0.05), np.float32(-6.4) + (np.float32(j) + np.float32(0.5)) * np.float32(0.1),
                        np.float32(-0.01) + (np.float32(k) + np.float32(0.5)) * np.float32(0.1), 0.0))
    return np.array(pts, dtype=np.float32)


@pytest.mark.parametrize("algo", ["astar", "thetastar", "lazythetastar", "costlazythetastar"])
def test_cpp_adapter_matches_ctypes_path(b200, algo):
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EXE)])
    out = subprocess.run([EXE, algo], capture_output=True, text=True, timeout=120)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert "error" not in d, d
    gm = b200.GridMap((12.8, 12.8, 3.2), 0.1, local_update_range=(100, 100, 100))
    gm.cloudCallback(_cloud(), (0, 0, 1))
    if algo.startswith("cost"):
        gm.build_edt(1.0)
    pl = b200.Planner(algo, resolution=0.1, lambda_heuristic=5.0, allocate_num=100000, cost_weight=2.0)
    pl.setGridMap(gm)
    r = pl.findPath((-3.0, -1.0, 1.0), (3.0, 1.0, 1.5))
    assert d["solved"] == 1 and r["solved"]
    assert d["algorithm"] == algo
    assert d["n_path"] == len(pl.getPath())
    assert d["g_final_node"] == r["g_final_node"] and d["path_length"] == r["path_length"]
    assert d["explored_nodes"] == r["explored_nodes"] and d["line_of_sight_checks"] == r["line_of_sight_checks"]
    assert np.array_equal(np.array(d["goal_coords"]), r["goal_coords"])
    assert d["collide_wall"] == 1 and d["collide_outside"] == 1 and d["total_cost1"] == 0
    assert d["time_spent_us"] > 0
    # GetPath.srv:20-23 through the adapter == the ctypes path == metrics.cpp arithmetic over the field the GPU built
    gm.build_edt(1.0)
    m, dist = gm.path_distance_metrics(pl.getPath(), return_distances=True)
    assert (d["dist_mean"], d["dist_std"], d["dist_min"], d["dist_max"]) == (m["mean"], m["std"], m["min"], m["max"])
    idx = gm.posToIndex(pl.getPath())
    want = np.sqrt(gm.dist2()[idx[:, 0], idx[:, 1], idx[:, 2]].astype(np.float64)) * 0.1
    assert np.array_equal(dist, want)
    acc = 0.0
    for v in want:
        acc += float(v)
    mean = acc / len(want)
    sd = 0.0
    for v in want:
        sd += (float(v) - mean) ** 2
    assert m["mean"] == mean and m["std"] == math.sqrt(sd / len(want)) and m["min"] == want.min() and m["max"] == want.max()
    assert m["min"] > 0
    assert d["set_occupied"] == [0, 1] and d["edt_stale"] == 1


def test_cpp_adapter_depth_path(b200, po):
    """GridMap::depthPoseCallback of the C++ adapter: a wall seen in 6 identical frames, against the oracle."""
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EXE)])
    out = subprocess.run([EXE, "depth"], capture_output=True, text=True, timeout=120)
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert "error" not in d, d
    om = po.OracleMap((-6.4, -6.4, -0.01), (12.8, 12.8, 3.2), 0.1, 0.099, (100, 100, 100), -0.01)
    om.set_depth_params(fx=96.8, fy=96.8, cx=80.3, cy=60.9)
    qw, qx, qy, qz = 0.5, -0.5, 0.5, -0.5
    R = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)],
                  [2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)],
                  [2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)]])
    img = np.full((120, 160), 2000, dtype=np.uint16)
    fused = sum(int(om.update_depth(img, (0.0, 0.0, 1.0), R)) for _ in range(6))
    st = om.depth_stats()
    want = dict(fused=fused, wall_occupancy=int(om.get_occupancy([(2.02, 0.0, 1.0)])[0]),
                wall_inflated=int(om.get_inflate_occupancy([(2.02, 0.0, 1.0)])[0]),
                free_occupancy=int(om.get_occupancy([(1.0, 0.0, 1.0)])[0]), outside=-1, projected=int(st[0]), rays_cast=int(st[1]),
                ray_steps=int(st[2]), updates=int(st[3]), has_depth=1)
    assert want["wall_occupancy"] == 1 and want["wall_inflated"] == 1 and want["free_occupancy"] == 0
    for k, v in want.items():
        assert d[k] == v, (k, d[k], v)


def test_cpp_adapter_planners_on_their_own_streams(b200):
    """AlgorithmBase::useOwnStream of the C++ adapter: three LazyThetaStar planners on one GridMap, three std::threads, two batches
    each — every batch identical to the single-planner batch (b200_stream_create / b200_planner_set_stream through the adapter)."""
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-s", "-C", os.path.dirname(EXE)])
    out = subprocess.run([EXE, "pool"], capture_output=True, text=True, timeout=120)
    d = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert "error" not in d, d
    assert d["mode"] == "pool" and d["lanes"] == 3 and d["queries"] == 400
    assert d["mismatches"] == 0 and d["solved"] > 300 and out.returncode == 0
