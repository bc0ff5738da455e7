"""ctypes binding of oracle/liboracle.so — TEST INFRASTRUCTURE ONLY (see oracle/oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

ALGOS = {"astar": 0, "thetastar": 1, "lazythetastar": 2, "costastar": 3, "costlazythetastar": 4, "thetastaragr": 5,
         "thetastaragrpp": 6}
# thetastaragr/{barrier, flying_cost, flying_cost_default, ground_judge, epsilon}: launch/planner.launch:71-77 defaults
AGR_DEFAULTS = dict(barrier=2.0, flying_cost=10.0, flying_cost_default=0.0, ground_judge=0.1, epsilon=0.1)
FAITHFUL, CANONICAL = 0, 1


class Result(C.Structure):
    _fields_ = [("solved", C.c_int32), ("status", C.c_int32), ("n_path", C.c_int32), ("open_peak", C.c_int32),
                ("path_offset", C.c_int64), ("g_final", C.c_double), ("path_length", C.c_double),
                ("explored_nodes", C.c_int64), ("iter_num", C.c_int64), ("los_checks", C.c_int64),
                ("los_samples", C.c_int64), ("goal_coords", C.c_double * 3), ("start_coords", C.c_double * 3),
                ("time_us", C.c_double)]


RESULT_DTYPE = np.dtype([("solved", "i4"), ("status", "i4"), ("n_path", "i4"), ("open_peak", "i4"), ("path_offset", "i8"),
                         ("g_final", "f8"), ("path_length", "f8"), ("explored_nodes", "i8"), ("iter_num", "i8"),
                         ("los_checks", "i8"), ("los_samples", "i8"), ("goal_coords", "f8", 3),
                         ("start_coords", "f8", 3), ("time_us", "f8")])
assert RESULT_DTYPE.itemsize == C.sizeof(Result)


class DepthParams(C.Structure):
    """Same layout as b200_depth_params_t (include/b200_planner.h); defaults = grid_map.cpp:22-49."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k_depth_scaling_factor", C.c_double), ("depth_filter_mindist", C.c_double),
                ("depth_filter_maxdist", C.c_double), ("max_ray_length", C.c_double),
                ("p_hit", C.c_double), ("p_miss", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("p_occ", C.c_double), ("virtual_ceil_height", C.c_double),
                ("use_depth_filter", C.c_int32), ("depth_filter_margin", C.c_int32), ("skip_pixel", C.c_int32),
                ("local_map_margin", C.c_int32)]


DEPTH_DEFAULTS = dict(fx=387.229248046875, fy=387.229248046875, cx=321.04638671875, cy=243.44969177246094,
                      k_depth_scaling_factor=1000.0, depth_filter_mindist=0.2, depth_filter_maxdist=5.0, max_ray_length=4.5,
                      p_hit=0.70, p_miss=0.35, p_min=0.12, p_max=0.97, p_occ=0.80, virtual_ceil_height=2.5,
                      use_depth_filter=1, depth_filter_margin=1, skip_pixel=2, local_map_margin=1)


def depth_params(**kw):
    d = dict(DEPTH_DEFAULTS)
    d.update(kw)
    p = DepthParams()
    for k, v in d.items():
        setattr(p, k, v)
    return p


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "CXX=g++"])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
        L.orc_map_create.restype = vp
        L.orc_map_create.argtypes = [dp, dp, C.c_double, C.c_double, dp, C.c_double]
        L.orc_map_destroy.argtypes = [vp]
        L.orc_map_dims.argtypes = [vp, C.POINTER(C.c_int)]
        L.orc_map_update_cloud.argtypes = [vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, dp, C.c_int]
        L.orc_map_clear_box.argtypes = [vp, dp, dp]
        L.orc_map_download_inflate.argtypes = [vp, vp]
        L.orc_map_upload_inflate.argtypes = [vp, vp]
        L.orc_map_local_bounds.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_map_get_inflate_occupancy.argtypes = [vp, vp, i64, vp]
        L.orc_map_export_inflate_cloud.argtypes = [vp, C.c_int, C.c_double, vp, i64]
        L.orc_map_export_inflate_cloud.restype = i64
        L.orc_map_is_in_map.argtypes = [vp, vp, i64, vp]
        L.orc_pos_to_index.argtypes = [vp, vp, i64, vp]
        L.orc_los_batch.argtypes = [vp, vp, vp, i64, vp, vp, C.c_int]
        L.orc_los_dk.argtypes = [vp, i64]
        L.orc_edt_brute.argtypes = [vp, vp]
        L.orc_map_build_edt.argtypes = [vp, C.c_double, vp, vp, vp, C.c_int]
        L.orc_planner_create.restype = vp
        L.orc_planner_create.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, dp]
        L.orc_planner_destroy.argtypes = [vp]
        L.orc_planner_set_map.argtypes = [vp, vp]
        L.orc_planner_find_path.argtypes = [vp, dp, dp, C.POINTER(Result), vp, i64]
        L.orc_planner_find_paths.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_double, C.c_double, dp,
                                             vp, vp, vp, i64, vp, vp, i64, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_planner_visited.argtypes = [vp, vp, i64]
        L.orc_planner_visited.restype = i64
        L.orc_map_set_depth_params.argtypes = [vp, vp]
        L.orc_map_update_depth.argtypes = [vp, vp, C.c_int, C.c_int, dp, dp]
        L.orc_map_update_depth.restype = C.c_int
        L.orc_map_download_occupancy.argtypes = [vp, vp]
        L.orc_map_set_raycast_num.argtypes = [vp, C.c_int]
        L.orc_map_raycast_num.argtypes = [vp]
        L.orc_map_raycast_num.restype = C.c_int
        L.orc_map_depth_stats.argtypes = [vp, vp]
        L.orc_map_get_occupancy.argtypes = [vp, vp, i64, vp]
        L.orc_map_set_occupied.argtypes = [vp, vp, i64]
        L.orc_map_set_occupancy.argtypes = [vp, vp, vp, i64]
        L.orc_map_export_occupancy_cloud.argtypes = [vp, C.c_int, C.c_int, C.c_double, vp, i64]
        L.orc_map_export_occupancy_cloud.restype = i64
        L.orc_los_batch_dda.argtypes = [vp, vp, vp, i64, vp, vp]
        L.orc_ray_cells.argtypes = [dp, dp, vp, i64]
        L.orc_ray_cells.restype = i64
        _LIB = L
    return _LIB


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleMap:
    """Restatement of GridMap's cloud path (grid_map.cpp:52-80,155-171,711-804; grid_map.h:257-404)."""

    def __init__(self, origin, size, resolution, obstacles_inflation=0.099, local_update_range=(1e9, 1e9, 1e9),
                 ground_height=None):
        self.origin, self.size, self.res = tuple(origin), tuple(size), float(resolution)
        gh = origin[2] if ground_height is None else ground_height
        self.h = lib().orc_map_create(_d3(origin), _d3(size), resolution, obstacles_inflation, _d3(local_update_range), gh)
        d = (C.c_int * 3)()
        lib().orc_map_dims(self.h, d)
        self.dims = tuple(d)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_map_destroy(self.h)
            self.h = None

    def update_cloud(self, cloud, cam, point_step=None, offsets=(0, 4, 8), nthreads=1):
        cloud = np.ascontiguousarray(cloud)
        if point_step is None:
            point_step = cloud.shape[1] * cloud.itemsize if cloud.ndim == 2 else 16
        n = cloud.nbytes // point_step
        lib().orc_map_update_cloud(self.h, _ptr(cloud), n, point_step, *offsets, _d3(cam), nthreads)

    def clear_box(self, mn, mx):
        lib().orc_map_clear_box(self.h, _d3(mn), _d3(mx))

    def inflate(self):
        out = np.empty(self.dims, dtype=np.uint8)
        lib().orc_map_download_inflate(self.h, _ptr(out))
        return out

    def set_inflate(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        assert grid.shape == self.dims
        lib().orc_map_upload_inflate(self.h, _ptr(grid))

    def local_bounds(self):
        a, b = (C.c_int * 3)(), (C.c_int * 3)()
        lib().orc_map_local_bounds(self.h, a, b)
        return tuple(a), tuple(b)

    def export_inflate_cloud(self, margin=0, truncate_height=2.4, cap=None):
        cap = int(np.prod(self.dims)) if cap is None else cap
        out = np.empty((cap, 4), dtype=np.float32)
        n = lib().orc_map_export_inflate_cloud(self.h, margin, truncate_height, _ptr(out), cap)
        return out[: min(n, cap)].copy(), n

    # ---- depth-image fusion (grid_map.cpp:173-697) ----
    def set_depth_params(self, **kw):
        self._dp = depth_params(**kw)
        lib().orc_map_set_depth_params(self.h, C.byref(self._dp))

    def update_depth(self, depth, cam_pos, cam_rot):
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        rot = np.ascontiguousarray(cam_rot, dtype=np.float64).reshape(9)
        r = lib().orc_map_update_depth(self.h, _ptr(depth), depth.shape[0], depth.shape[1], _d3(cam_pos),
                                       rot.ctypes.data_as(C.POINTER(C.c_double)))
        if r < 0:
            raise RuntimeError("set_depth_params first")
        return bool(r)

    def occupancy(self):
        out = np.empty(self.dims, dtype=np.float64)
        lib().orc_map_download_occupancy(self.h, _ptr(out))
        return out

    def set_raycast_num(self, v):
        lib().orc_map_set_raycast_num(self.h, int(v))

    def raycast_num(self):
        return lib().orc_map_raycast_num(self.h)

    def depth_stats(self):
        out = np.zeros(8, dtype=np.int64)
        lib().orc_map_depth_stats(self.h, _ptr(out))
        return out

    def export_occupancy_cloud(self, which, margin=0, truncate_height=2.4):
        """publishMap (which=0, margin = local_map_margin // 2) / publishUnknown (which=1) -> (n,4) float32."""
        cap = int(np.prod(self.dims))
        out = np.empty((cap, 4), dtype=np.float32)
        n = lib().orc_map_export_occupancy_cloud(self.h, which, margin, truncate_height, _ptr(out), cap)
        return out[:n].copy()

    def get_occupancy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        lib().orc_map_get_occupancy(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def set_occupied(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        lib().orc_map_set_occupied(self.h, _ptr(pos), len(pos))

    def set_occupancy(self, pos, occ):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        occ = np.ascontiguousarray(np.broadcast_to(np.asarray(occ, dtype=np.float64), (len(pos),)))
        lib().orc_map_set_occupancy(self.h, _ptr(pos), _ptr(occ), len(pos))

    def get_inflate_occupancy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        lib().orc_map_get_inflate_occupancy(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def is_in_map(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.uint8)
        lib().orc_map_is_in_map(self.h, _ptr(pos), len(pos), _ptr(out))
        return out.astype(bool)

    def pos_to_index(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(pos), 3), dtype=np.int32)
        lib().orc_pos_to_index(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def los_batch(self, starts, ends, nthreads=1, return_samples=False):
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(ends, dtype=np.float64).reshape(-1, 3)
        vis = np.empty(len(s), dtype=np.uint8)
        ns = np.empty(len(s), dtype=np.int64)
        lib().orc_los_batch(self.h, _ptr(s), _ptr(e), len(s), _ptr(vis), _ptr(ns), nthreads)
        return (vis, ns) if return_samples else vis

    def los_batch_dda(self, starts, ends, return_cells=False):
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(ends, dtype=np.float64).reshape(-1, 3)
        vis = np.empty(len(s), dtype=np.uint8)
        nc = np.empty(len(s), dtype=np.int32)
        lib().orc_los_batch_dda(self.h, _ptr(s), _ptr(e), len(s), _ptr(vis), _ptr(nc))
        return (vis, nc) if return_cells else vis

    def edt_brute(self):
        out = np.empty(self.dims, dtype=np.int32)
        lib().orc_edt_brute(self.h, _ptr(out))
        return out

    def build_edt(self, d_safe=0.0, nthreads=1, want_cost=False):
        d2 = np.empty(self.dims, dtype=np.int32)
        cq = np.empty(self.dims, dtype=np.uint16) if d_safe > 0 else None
        cf = np.empty(self.dims, dtype=np.float32) if (d_safe > 0 and want_cost) else None
        lib().orc_map_build_edt(self.h, d_safe, _ptr(d2), _ptr(cq) if cq is not None else None,
                                _ptr(cf) if cf is not None else None, nthreads)
        return (d2, cq, cf) if want_cost else (d2, cq)


def ray_cells(a, b, cap=4096):
    """Voxel sequence of the RayCaster restatement between two lattice points (raycast.cpp:253-346)."""
    out = np.empty((cap, 3), dtype=np.int32)
    n = lib().orc_ray_cells(_d3(a), _d3(b), _ptr(out), cap)
    return out[: min(n, cap)].copy()


def los_dk(n):
    out = np.empty(n, dtype=np.float64)
    lib().orc_los_dk(_ptr(out), n)
    return out


class OraclePlanner:
    """Restatement of AStar/ThetaStar::findPath (AStar.cpp:80-179, ThetaStar.cpp:37-89) + our lazy/cost variants."""

    def __init__(self, algo, omap, resolution, lambda_heuristic=5.0, allocate_num=1000000, mode=CANONICAL,
                 cost_weight=0.0, max_los=0.0, **agr):
        self.algo, self.mode, self.map = ALGOS[algo], mode, omap
        a = dict(AGR_DEFAULTS, **agr)
        self._agr = (C.c_double * 5)(a["barrier"], a["flying_cost"], a["flying_cost_default"], a["ground_judge"], a["epsilon"])
        self.args = (float(resolution), float(lambda_heuristic), int(allocate_num), float(cost_weight), float(max_los), self._agr)
        self.h = lib().orc_planner_create(self.algo, mode, *self.args)
        lib().orc_planner_set_map(self.h, omap.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_planner_destroy(self.h)
            self.h = None

    def find_path(self, start, goal, path_cap=100000):
        r = Result()
        path = np.empty((path_cap, 3), dtype=np.float64)
        lib().orc_planner_find_path(self.h, _d3(start), _d3(goal), C.byref(r), _ptr(path), path_cap)
        res = np.frombuffer(bytes(r), dtype=RESULT_DTYPE)[0]
        return res, path[: min(r.n_path, path_cap)].copy()

    def visited(self):
        """getVisitedNodes of the last find_path: (n,3) positions in creation order."""
        n = lib().orc_planner_visited(self.h, None, 0)
        out = np.empty((n, 3), dtype=np.float64)
        lib().orc_planner_visited(self.h, _ptr(out), n)
        return out

    def find_paths(self, starts, goals, path_stride=0, nthreads=1):
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        g = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, 3)
        nq = len(s)
        res = np.zeros(nq, dtype=RESULT_DTYPE)
        path = np.empty((nq * path_stride, 3), dtype=np.float64) if path_stride else None
        lib().orc_planner_find_paths(self.algo, self.mode, *self.args, self.map.h, _ptr(s), _ptr(g), nq, _ptr(res),
                                     _ptr(path) if path is not None else None, path_stride, nthreads)
        return res, path


def max_threads():
    return lib().orc_max_threads()
