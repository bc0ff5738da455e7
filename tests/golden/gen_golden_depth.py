#!/usr/bin/env python
"""Generates tests/golden/ref_depth_golden.npz by RUNNING THE REFERENCE'S OWN depth-fusion code
(GridMap::depthPoseCallback + updateOccupancyCallback of the unmodified /root/reference sources, compiled against
oracle/shim/ into oracle/_ref/libref.so).  Only runnable where /root/reference exists; the .npz is committed so the GPU
box can check the oracle and the CUDA path against it.

    python tests/golden/gen_golden_depth.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import astar_planners_b200  # noqa: F401,E402
from astar_planners_b200 import synth  # noqa: E402
from oracle import pyref as pr  # noqa: E402
from oracle.pyoracle import DEPTH_DEFAULTS  # noqa: E402

SIZE, RES, GH = (10.0, 10.0, 3.0), 0.1, -0.01
ORIGIN = (-5.0, -5.0, GH)
ROWS, COLS = 120, 160
INTR = dict(fx=96.8, fy=96.8, cx=80.3, cy=60.9)

# name: (depth params, obstacles_inflation, local_update_range, first raycast_num, n_frames, camera box, image shape)
# name: (depth params, obstacles_inflation, local_update_range, first raycast_num_, n_frames, camera box, image shape,
#        rewind: frame index before which raycast_num_ is set back to `first`, as if 256 frames had gone by)
SCENARIOS = {
    "filter": (dict(), 0.099, (5.5, 5.5, 4.5), 0, 8, 3.5, (ROWS, COLS), None),
    "nofilter_infl2_edge": (dict(use_depth_filter=0, local_map_margin=3), 0.15, (3.0, 3.0, 2.0), 0, 6, 4.8, (ROWS, COLS), None),
    # `char raycast_num_` (grid_map.h:124): 125, 126, 127, -128, -127 ...
    "counter_wrap": (dict(skip_pixel=1, depth_filter_margin=2), 0.099, (5.5, 5.5, 4.5), 124, 6, 3.5, (ROWS, COLS), None),
    # ... -2, -1: in round -1 every flag still holding its initial -1 looks set (no ray is cast into untouched space), then 0, 1
    "initial_flag_alias": (dict(use_depth_filter=0), 0.099, (5.5, 5.5, 4.5), -4, 6, 3.5, (ROWS, COLS), None),
    # flags left by frames 0-2 are met again by frames 3-5, which reuse the same round numbers
    "stale_flag_alias": (dict(use_depth_filter=0), 0.099, (5.5, 5.5, 4.5), 40, 6, 3.5, (ROWS, COLS), 3),
    # > 65536 zero-depth pixels all project onto the camera's own voxel: the `short` hit/miss counters wrap and the voxel
    # is queued a second time (setCacheOccupancy, grid_map.cpp:182-187)
    "count_wrap": (dict(use_depth_filter=0, skip_pixel=1, fx=387.2, fy=387.2, cx=321.0, cy=243.4, max_ray_length=3.0), 0.099,
                   (5.5, 5.5, 4.5), 7, 2, 1.0, (480, 640), None),
}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frames_of(name):
    dp_over, infl, rng, first, n_frames, box, shape, rewind = SCENARIOS[name]
    dp = dict(DEPTH_DEFAULTS)
    dp.update(INTR)
    dp.update(dp_over)
    r = np.random.default_rng(sum(map(ord, name)))
    boxes = synth.pillar_boxes(11, ORIGIN, SIZE, 25)
    out = []
    base = np.array([r.uniform(-box, box), r.uniform(-box, box), r.uniform(0.6, 1.8)])
    yaw, pitch = r.uniform(0, 2 * np.pi), r.uniform(-0.1, 0.4)
    for f in range(n_frames):
        # a slowly moving camera, so that surfaces are seen several times and cross the occupancy threshold
        pos = tuple(base + r.uniform(-0.15, 0.15, 3) + np.array([0.05 * f, 0.0, 0.0]))
        R, q = synth.look_at(yaw + r.uniform(-0.1, 0.1), pitch + r.uniform(-0.05, 0.05))
        img = synth.render_depth(boxes, GH, pos, R, shape[0], shape[1], dp["fx"], dp["fy"], dp["cx"], dp["cy"], dropout=0.3 if name == "count_wrap" else 0.03, seed=f)
        out.append((np.array(pos), q, img))
    return dp, infl, rng, first, out, rewind


def main():
    out = {"size": np.array(SIZE), "origin": np.array(ORIGIN), "res": np.array(RES)}
    for name in SCENARIOS:
        dp, infl, rng, first, frames, rewind = frames_of(name)
        rm = pr.RefMap(SIZE, RES, infl, rng, GH, depth=dp)
        assert np.allclose(rm.origin, ORIGIN)
        if first:
            rm.set_raycast_num(first)
        out[f"{name}/params"] = np.array([dp[k] for k in DEPTH_DEFAULTS], dtype=np.float64)
        out[f"{name}/inflation"], out[f"{name}/range"], out[f"{name}/first"] = np.array(infl), np.array(rng), np.array(first)
        occ_sha, inf_bits, bounds, rots, proj = [], [], [], [], []
        out[f"{name}/rewind"] = np.array(-1 if rewind is None else rewind)
        for fi, (pos, q, img) in enumerate(frames):
            if rewind is not None and fi == rewind:
                rm.set_raycast_num(first)
            fused, R = rm.update_depth(img, pos, q)
            assert fused
            occ_sha.append(sha(rm.occupancy()))
            inf_bits.append(np.packbits(rm.inflate()))
            bounds.append(np.array(rm.local_bounds()))
            rots.append(R)
            proj.append(rm.proj_points())
        out[f"{name}/pos"] = np.array([f[0] for f in frames])
        out[f"{name}/rot"] = np.array(rots)
        out[f"{name}/img"] = np.array([f[2] for f in frames])
        out[f"{name}/occ_sha"] = np.array(occ_sha)
        out[f"{name}/inflate_bits"] = np.array(inf_bits)
        out[f"{name}/bounds"] = np.array(bounds)
        out[f"{name}/proj"] = np.array(proj)
        out[f"{name}/frames_after"] = np.array(rm.raycast_num())
        pm, pu = rm.publish_occupancy(0), rm.publish_occupancy(1)  # publishMap / publishUnknown (grid_map.cpp:806-849, :902-939)
        out[f"{name}/pub_map_sha"], out[f"{name}/pub_unknown_sha"] = np.array(sha(pm)), np.array(sha(pu))
        out[f"{name}/pub_counts"] = np.array([len(pm), len(pu)])
        occ = rm.occupancy()
        out[f"{name}/occ_final_values"], inv = np.unique(occ, return_inverse=True)
        assert len(out[f"{name}/occ_final_values"]) < 65536
        out[f"{name}/occ_final_index"] = inv.astype(np.uint16).reshape(occ.shape)
        # a camera outside the map: the frame is dropped (grid_map.cpp:688-697)
        fused, _ = rm.update_depth(frames[0][2], (0.0, 0.0, SIZE[2] + 1.0), frames[0][1])
        assert not fused and sha(rm.occupancy()) == occ_sha[-1]
        print(name, "frames", len(frames), "proj", proj, "occupied", int((occ > np.log(0.8 / 0.2)).sum()), "inflated",
              int(np.unpackbits(inf_bits[-1]).sum()), "distinct log-odds", len(out[f"{name}/occ_final_values"]))
    path = os.path.join(ROOT, "tests", "golden", "ref_depth_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
