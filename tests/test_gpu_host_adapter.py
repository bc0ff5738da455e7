"""-m gpu: the C++ adapter classes (3d-astar-path-planners_b200/host) driven like the reference's ROS node, compared
with the same world through the ctypes binding."""
import json
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
