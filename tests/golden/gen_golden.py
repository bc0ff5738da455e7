#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz by RUNNING THE REFERENCE'S OWN CODE (oracle/_ref/libref.so: the
unmodified /root/reference sources compiled against oracle/shim/).  Only runnable where /root/reference exists;
the .npz is committed so the GPU box and CI can check the oracle and the CUDA path against it.

    python tests/golden/gen_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import astar_planners_b200  # noqa: F401,E402
from astar_planners_b200 import synth  # noqa: E402
from oracle import pyref as pr  # noqa: E402
from oracle.pyoracle import RESULT_DTYPE  # noqa: E402

SIZE, RES, GH = (12.8, 12.8, 3.2), 0.1, -0.01
ORIGIN = (-6.4, -6.4, GH)  # what initMap derives: (-sx/2, -sy/2, ground_height)  (grid_map.cpp:53)


def main():
    out = {}
    cloud = synth.forest(21, ORIGIN, SIZE, RES, 22)
    extra = synth.uniform_cloud(22, ORIGIN, SIZE, 3000, margin=0.4)  # generic (non-lattice) + out-of-map points
    cloud = np.concatenate([cloud, extra]).astype(np.float32)
    blob = synth.pack_cloud(cloud, 16)
    out["cloud"] = cloud
    out["size"], out["origin"], out["res"] = np.array(SIZE), np.array(ORIGIN), np.array(RES)

    # frame 1: whole-map update range
    rm = pr.RefMap(SIZE, RES, 0.099, (1e3, 1e3, 1e3), GH)
    assert np.allclose(rm.origin, ORIGIN) and rm.dims == (128, 128, 32)
    cam1 = (0.0, 0.0, 1.0)
    rm.update_cloud(blob, cam1)
    g1 = rm.inflate()
    out["cam1"], out["grid1_bits"] = np.array(cam1), np.packbits(g1)
    out["bounds1"] = np.array(rm.local_bounds())

    # frame 2 on a second map: small local range around a moved camera, on top of frame 1's grid
    rng2, cam2 = (3.0, 2.5, 1.2), (1.7, -2.2, 0.9)
    rm2 = pr.RefMap(SIZE, RES, 0.099, rng2, GH)
    rm2.set_inflate(g1)
    cloud2 = synth.uniform_cloud(23, ORIGIN, SIZE, 4000)
    rm2.update_cloud(synth.pack_cloud(cloud2, 16), cam2)
    out["cloud2"], out["range2"], out["cam2"] = cloud2, np.array(rng2), np.array(cam2)
    out["grid2_bits"], out["bounds2"] = np.packbits(rm2.inflate()), np.array(rm2.local_bounds())

    # inflation step 2 (obstacles_inflation 0.15 -> ceil(1.5) = 2)
    rm3 = pr.RefMap(SIZE, RES, 0.15, (1e3, 1e3, 1e3), GH)
    rm3.update_cloud(blob, cam1)
    out["grid3_bits"] = np.packbits(rm3.inflate())

    # per-position reads
    r = np.random.default_rng(31)
    o, s = np.array(ORIGIN), np.array(SIZE)
    pos = o - 0.3 + r.random((4000, 3)) * (s + 0.6)
    edge = np.array([[o[0] + 1e-4, 0, 1], [np.nextafter(o[0] + 1e-4, -1), 0, 1], [o[0] + s[0] - 1e-4, 0, 1],
                     [np.nextafter(o[0] + s[0] - 1e-4, 1e9), 0, 1], [0.6000000000000001, 0.30000000000000004, 1.0]])
    pos = np.concatenate([pos, edge])
    out["pos"], out["pos_occ"] = pos, rm.get_inflate_occupancy(pos)
    out["pos_inmap"], out["pos_index"] = rm.is_in_map(pos), rm.pos_to_index(pos)

    # fastLOS
    a = o + 0.2 + r.random((6000, 3)) * (s - 0.4)
    b = o + 0.2 + r.random((6000, 3)) * (s - 0.4)
    b[:600] = o - 1.0 + r.random((600, 3)) * (s + 2.0)
    b[600:1800] = a[600:1800] + r.normal(0, 0.3, (1200, 3))
    b[1800:1810] = a[1800:1810]
    out["los_a"], out["los_b"], out["los_vis"] = a, b, rm.los_batch(a, b)

    # planners (the reference's AStar / ThetaStar, reset + findPath per query)
    qs, qg = synth.random_queries(32, rm.get_inflate_occupancy, ORIGIN, SIZE, 64, min_dist=3.0)
    out["q_start"], out["q_goal"] = qs, qg
    for algo in ("astar", "thetastar"):
        for lam in (5.0, 1.0):
            n = 64 if lam == 5.0 else 12
            pl = pr.RefPlanner(algo, rm, RES, lam, 100000)
            assert pl.pool_disorder() == 0  # node addresses increase with creation order (fresh-heap tie-break)
            res, paths = pl.find_paths(qs[:n], qg[:n])
            key = f"{algo}_lam{int(lam)}"
            out[key + "_res"] = res
            out[key + "_path"] = np.concatenate(paths) if paths else np.zeros((0, 3))
        # tiny node budget: pool exhaustion branch (AStar.cpp:165-168)
        pl = pr.RefPlanner(algo, rm, RES, 5.0, 300)
        res, paths = pl.find_paths(qs[:16], qg[:16])
        out[f"{algo}_oom_res"] = res
    # the fork's air-ground planners (ThetaStarAGR.cpp, ThetaStarAGRpp.cpp), launch-file parameters; some starts on /
    # below the ground plane exercise the z clamp and the 0/0 -> NaN cost of ThetaStarAGRpp.cpp:63
    qs2, qg2 = qs[:48].copy(), qg[:48].copy()
    qs2[:8, 2] = 0.0
    qs2[8:12, 2] = -0.005
    qg2[12:20, 2] = 0.05
    out["agr_start"], out["agr_goal"] = qs2, qg2
    for algo in ("thetastaragr", "thetastaragrpp"):
        pl = pr.RefPlanner(algo, rm, RES, 5.0, 100000)
        assert pl.pool_disorder() == 0
        res, paths = pl.find_paths(qs2, qg2)
        out[f"{algo}_res"] = res
        out[f"{algo}_path"] = np.concatenate(paths) if paths else np.zeros((0, 3))
    # unreachable goal: sealed hollow box -> open set empties (AStar.cpp:175-178)
    grid = np.zeros((128, 128, 32), np.uint8)
    grid[40:52, 40:52, 5:17] = 1
    grid[42:50, 42:50, 7:15] = 0
    rm4 = pr.RefMap(SIZE, RES, 0.099, (1e3, 1e3, 1e3), GH)
    rm4.set_inflate(grid)
    inside, outside = np.array([[-6.4 + 4.6, -6.4 + 4.6, GH + 1.1]]), np.array([[3.0, 3.0, 2.0]])
    out["sealed_start"], out["sealed_goal"] = inside, outside
    for algo in ("astar", "thetastar"):
        res, _ = pr.RefPlanner(algo, rm4, RES, 5.0, 100000).find_paths(inside, outside)
        out[f"{algo}_sealed_res"] = res
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if k.endswith("_res")})


if __name__ == "__main__":
    main()
