"""Replays the committed depth-fusion golden scenarios (tests/golden/ref_depth_golden.npz, produced by the reference's
own code via tests/golden/gen_golden_depth.py) through any implementation with the GridMap-like depth interface."""
import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_depth_golden.npz")
SCENARIOS = ("filter", "nofilter_infl2_edge", "counter_wrap", "initial_flag_alias", "stale_flag_alias", "count_wrap")
PARAM_NAMES = ("fx", "fy", "cx", "cy", "k_depth_scaling_factor", "depth_filter_mindist", "depth_filter_maxdist", "max_ray_length",
               "p_hit", "p_miss", "p_min", "p_max", "p_occ", "virtual_ceil_height", "use_depth_filter", "depth_filter_margin",
               "skip_pixel", "local_map_margin")
INT_PARAMS = ("use_depth_filter", "depth_filter_margin", "skip_pixel", "local_map_margin")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scenario(g, name):
    dp = {k: (int(v) if k in INT_PARAMS else float(v)) for k, v in zip(PARAM_NAMES, g[f"{name}/params"])}
    return dict(params=dp, inflation=float(g[f"{name}/inflation"]), range=tuple(g[f"{name}/range"]), first=int(g[f"{name}/first"]),
                pos=g[f"{name}/pos"], rot=g[f"{name}/rot"], img=g[f"{name}/img"], occ_sha=g[f"{name}/occ_sha"],
                inflate_bits=g[f"{name}/inflate_bits"], bounds=g[f"{name}/bounds"], proj=g[f"{name}/proj"],
                frames_after=int(g[f"{name}/frames_after"]), rewind=int(g[f"{name}/rewind"]),
                pub_map_sha=str(g[f"{name}/pub_map_sha"]), pub_unknown_sha=str(g[f"{name}/pub_unknown_sha"]),
                pub_counts=tuple(int(x) for x in g[f"{name}/pub_counts"]),
                occ_final=g[f"{name}/occ_final_values"][g[f"{name}/occ_final_index"]])


def check_publishers(sc, pub_map, pub_unknown):
    """pub_map / pub_unknown: (n,4) float32 PointXYZ records of publishMap / publishUnknown after the last frame."""
    assert (len(pub_map), len(pub_unknown)) == sc["pub_counts"]
    assert sha(pub_map) == sc["pub_map_sha"] and sha(pub_unknown) == sc["pub_unknown_sha"]


def replay(sc, update, occupancy, inflate, bounds, dims, set_frame_count):
    """update(img, pos, rot) -> fused; asserts every frame against the reference's outputs."""
    set_frame_count(sc["first"])
    for f in range(len(sc["pos"])):
        if f == sc["rewind"]:
            set_frame_count(sc["first"])
        assert update(sc["img"][f], sc["pos"][f], sc["rot"][f])
        occ = occupancy()
        if f == len(sc["pos"]) - 1:
            bad = np.argwhere(occ != sc["occ_final"])
            assert len(bad) == 0, f"log-odds differ at {len(bad)} voxels, first {bad[:3].tolist()}"
        assert sha(occ) == str(sc["occ_sha"][f]), f"frame {f}: log-odds grid differs from the reference"
        want = np.unpackbits(sc["inflate_bits"][f])[: int(np.prod(dims))].reshape(dims)
        got = inflate()
        bad = np.argwhere(got != want)
        assert len(bad) == 0, f"frame {f}: inflated grid differs at {len(bad)} voxels, first {bad[:3].tolist()}"
        if sc["proj"][f] > 0:
            assert np.array_equal(np.array(bounds()), sc["bounds"][f]), f"frame {f}: local bounds"
