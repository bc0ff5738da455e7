"""Shared builders for the parity tests: the same seeded world is given to the CUDA path and to the oracle."""
import numpy as np


def make_pair(b200, po, synth, name="C3", n_points=None, cam=(0.0, 0.0, 1.0), rng=(1e3, 1e3, 1e3), inflation=0.099,
              cloud=None, point_step=16):
    c = synth.CONFIGS[name]
    if cloud is None:
        cloud = synth.config_cloud(name, n_points=n_points)
    blob = synth.pack_cloud(cloud, point_step)
    gm = b200.GridMap(c["size"], c["resolution"], origin=c["origin"], obstacles_inflation=inflation,
                      local_update_range=rng, ground_height=c["origin"][2])
    om = po.OracleMap(c["origin"], c["size"], c["resolution"], obstacles_inflation=inflation, local_update_range=rng,
                      ground_height=c["origin"][2])
    gm.cloudCallback(blob, cam, point_step=point_step)
    om.update_cloud(blob, cam, point_step=point_step)
    return gm, om, c


def assert_results_equal(res_gpu, res_cpu, fields=("solved", "status", "n_path", "open_peak", "explored_nodes", "iter_num", "los_checks",
                                                   "los_samples")):
    for f in fields:
        bad = np.nonzero(res_gpu[f] != res_cpu[f])[0]
        assert len(bad) == 0, f"{f}: {len(bad)} mismatches, first {bad[:5]} gpu={res_gpu[f][bad[:5]]} cpu={res_cpu[f][bad[:5]]}"
    for f in ("g_final", "path_length", "goal_coords", "start_coords"):
        a, b = res_gpu[f].view(np.uint64), res_cpu[f].view(np.uint64)
        bad = np.nonzero((a != b).reshape(len(res_gpu), -1).any(axis=1))[0]
        assert len(bad) == 0, f"{f}: {len(bad)} bit mismatches, first {bad[:5]} gpu={res_gpu[f][bad[:5]]} cpu={res_cpu[f][bad[:5]]}"


def set_positions(seed, origin, size, n):
    """positions inside, on the +-1e-4 rim, outside, and repeated voxels; occ values incl. invalid ones"""
    rng = np.random.default_rng(seed)
    o, s = np.asarray(origin), np.asarray(size)
    pos = o - 0.3 + rng.random((n, 3)) * (s + 0.6)
    pos[: n // 8] = pos[n // 8: n // 4]                       # repeats: the last call on a voxel wins
    pos[-3] = o + 1e-4
    pos[-2] = np.nextafter(o + 1e-4, 10.0)
    pos[-1] = o + s - 1e-4
    occ = rng.choice([0.0, 1.0, 1.0, 0.5, -1.0, 2.0], n)
    return pos, occ
