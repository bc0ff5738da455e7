"""Timing driver for the depth-fusion path: 640x480 frames (reference defaults) on the C2 grid, GPU vs the CPU oracle.

usage: python profiles/prof_depth.py [n_frames]
"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
b200 = importlib.import_module("3d-astar-path-planners_b200")
synth = importlib.import_module("3d-astar-path-planners_b200.synth")
from oracle import pyoracle as po  # noqa: E402  (CPU baseline only)


def main():
    n_frames = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    c = synth.CONFIGS["C2"]
    rng_ = (5.5, 5.5, 4.5)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], local_update_range=rng_, ground_height=c["origin"][2])
    om = po.OracleMap(c["origin"], c["size"], c["resolution"], local_update_range=rng_, ground_height=c["origin"][2])
    gm.set_depth_params()
    om.set_depth_params()
    dp = b200.DEPTH_DEFAULTS
    boxes = synth.pillar_boxes(c["seed"], c["origin"], c["size"], c["n_pillars"])
    base = np.array([-3.0, 2.0, 1.2])
    tg, tc = [], []
    for f in range(n_frames):
        pos = base + np.array([0.12 * f, -0.05 * f, 0.0])
        R, _ = synth.look_at(0.7 + 0.03 * f, 0.15)
        img = synth.render_depth(boxes, c["origin"][2], pos, R, 480, 640, dp["fx"], dp["fy"], dp["cx"], dp["cy"], dropout=0.02, seed=f)
        gm.sync()
        t0 = time.perf_counter()
        gm.depthPoseCallback(img, pos, R)
        gm.sync()
        t1 = time.perf_counter()
        om.update_depth(img, pos, R)
        t2 = time.perf_counter()
        tg.append((t1 - t0) * 1e3)
        tc.append((t2 - t1) * 1e3)
        print(f"frame {f}: gpu {tg[-1]:.3f} ms (host image in, results resident)  cpu oracle {tc[-1]:.2f} ms  {gm.depth_stats()}")
    same = np.array_equal(gm.occupancy(), om.occupancy()) and np.array_equal(gm.inflate(), om.inflate())
    print(f"median gpu {np.median(tg[2:]):.3f} ms  cpu {np.median(tc[2:]):.2f} ms  identical grids: {same}")


if __name__ == "__main__":
    main()
