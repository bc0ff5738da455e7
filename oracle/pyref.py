"""ctypes binding of oracle/_ref/libref.so — the reference's OWN unmodified sources (AStar.cpp, ThetaStar.cpp,
AlgorithmBase.cpp, LineOfSight.cpp, geometry_utils.cpp, grid_map.cpp, ...) compiled against the header shim in
oracle/shim/.  TEST INFRASTRUCTURE ONLY.  Exists only where /root/reference does (this container); tests that
need it skip elsewhere and fall back to the golden vectors generated from it (tests/golden/)."""
import ctypes as C
import os
import subprocess

import numpy as np

from .pyoracle import RESULT_DTYPE, Result, _d3, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None


def available():
    return os.path.exists(_PATH) or os.path.isdir("/root/reference/src")


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(_PATH):
            subprocess.check_call(["make", "-s", "-C", _HERE, "ref", "CXX=g++"])
        L = C.CDLL(_PATH)
        vp, dp, i64 = C.c_void_p, C.POINTER(C.c_double), C.c_int64
        L.ref_map_create.restype = vp
        L.ref_map_create.argtypes = [dp, C.c_double, C.c_double, dp, C.c_double]
        L.ref_map_destroy.argtypes = [vp]
        L.ref_map_dims.argtypes = [vp, C.POINTER(C.c_int)]
        L.ref_map_region.argtypes = [vp, dp, dp]
        L.ref_map_update_cloud.argtypes = [vp, vp, i64, C.c_int, C.c_int, C.c_int, C.c_int, dp]
        L.ref_map_clear_box.argtypes = [vp, dp, dp]
        L.ref_map_download_inflate.argtypes = [vp, vp]
        L.ref_map_upload_inflate.argtypes = [vp, vp]
        L.ref_map_local_bounds.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.ref_map_get_inflate_occupancy.argtypes = [vp, vp, i64, vp]
        L.ref_map_publish_inflate.argtypes = [vp, C.c_int, vp, i64]
        L.ref_map_publish_inflate.restype = i64
        L.ref_map_is_in_map.argtypes = [vp, vp, i64, vp]
        L.ref_pos_to_index.argtypes = [vp, vp, i64, vp]
        L.ref_los_batch.argtypes = [vp, vp, vp, i64, vp]
        L.ref_set_param.argtypes = [C.c_char_p, C.c_double]
        L.ref_map_update_depth.argtypes = [vp, vp, C.c_int, C.c_int, dp, dp, dp]
        L.ref_map_update_depth.restype = C.c_int
        L.ref_map_depth_odom_pose.argtypes = [vp, dp, dp, dp, dp]
        L.ref_map_download_occupancy.argtypes = [vp, vp]
        L.ref_map_raycast_num.argtypes = [vp]
        L.ref_map_raycast_num.restype = C.c_int
        L.ref_map_proj_points.argtypes = [vp]
        L.ref_map_proj_points.restype = C.c_int
        L.ref_map_set_raycast_num.argtypes = [vp, C.c_int]
        L.ref_map_get_occupancy.argtypes = [vp, vp, i64, vp]
        L.ref_map_set_occupied.argtypes = [vp, vp, i64]
        L.ref_distance_metrics.argtypes = [vp, i64, vp]
        L.ref_map_set_occupancy.argtypes = [vp, vp, vp, i64]
        L.ref_map_publish_occupancy.argtypes = [vp, C.c_int, vp, i64]
        L.ref_map_publish_occupancy.restype = i64
        L.ref_ray_cells.argtypes = [dp, dp, vp, i64]
        L.ref_ray_cells.restype = i64
        L.ref_planner_create.restype = vp
        L.ref_planner_create.argtypes = [C.c_int, C.c_double, C.c_double, C.c_int, vp]
        L.ref_planner_destroy.argtypes = [vp]
        L.ref_planner_pool_disorder.argtypes = [vp]
        L.ref_planner_pool_disorder.restype = C.c_int
        L.ref_planner_find_path.argtypes = [vp, dp, dp, C.POINTER(Result), vp, i64]
        L.ref_planner_visited.argtypes = [vp, vp, i64]
        L.ref_planner_visited.restype = i64
        _LIB = L
    return _LIB


def distance_metrics(distances):
    """The reference's own calculateDistancesMetrics (src/utils/metrics.cpp:62-78) -> (mean, std_dev, min, max)."""
    d = np.ascontiguousarray(distances, dtype=np.float64)
    out = np.zeros(4)
    lib().ref_distance_metrics(_ptr(d), len(d), _ptr(out))
    return tuple(out)


def ray_cells(a, b, cap=4096):
    """Cells the reference's own RayCaster::step returns between two lattice points (raycast.cpp:253-346)."""
    out = np.empty((cap, 3), dtype=np.int32)
    n = lib().ref_ray_cells(_d3(a), _d3(b), _ptr(out), cap)
    return out[: min(n, cap)].copy()


class RefMap:
    """The reference's GridMap (initMap + odomCallback + cloudCallback); origin is always (-sx/2,-sy/2,ground_height)."""

    def __init__(self, size, resolution, obstacles_inflation=0.099, local_update_range=(1e9, 1e9, 1e9), ground_height=-0.01,
                 depth=None):
        """depth: dict of grid_map/* depth-fusion params (names as in oracle.pyoracle.DEPTH_DEFAULTS), set before initMap."""
        for k, v in (depth or {}).items():
            lib().ref_set_param(("grid_map/" + k).encode(), float(v))
        self.h = lib().ref_map_create(_d3(size), resolution, obstacles_inflation, _d3(local_update_range), ground_height)
        d = (C.c_int * 3)()
        lib().ref_map_dims(self.h, d)
        self.dims = tuple(d)
        o, s = (C.c_double * 3)(), (C.c_double * 3)()
        lib().ref_map_region(self.h, o, s)
        self.origin, self.size = tuple(o), tuple(s)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_map_destroy(self.h)
            self.h = None

    def update_cloud(self, cloud, cam, point_step=None, offsets=(0, 4, 8)):
        cloud = np.ascontiguousarray(cloud)
        if point_step is None:
            point_step = cloud.shape[1] * cloud.itemsize if cloud.ndim == 2 else 16
        lib().ref_map_update_cloud(self.h, _ptr(cloud), cloud.nbytes // point_step, point_step, *offsets, _d3(cam))

    def clear_box(self, mn, mx):
        lib().ref_map_clear_box(self.h, _d3(mn), _d3(mx))

    def inflate(self):
        out = np.empty(self.dims, dtype=np.uint8)
        lib().ref_map_download_inflate(self.h, _ptr(out))
        return out

    def set_inflate(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        assert grid.shape == self.dims
        lib().ref_map_upload_inflate(self.h, _ptr(grid))

    def local_bounds(self):
        a, b = (C.c_int * 3)(), (C.c_int * 3)()
        lib().ref_map_local_bounds(self.h, a, b)
        return tuple(a), tuple(b)

    def publish_inflate(self, all_info=False):
        """publishMapInflate (grid_map.cpp:851-900): the PointXYZ records (x, y, z, 1.0) it would publish; uses the
        reference defaults grid_map/visualization_truncate_height = 2.4 and local_map_margin = 1."""
        cap = int(np.prod(self.dims))
        out = np.empty((cap, 4), dtype=np.float32)
        n = lib().ref_map_publish_inflate(self.h, 1 if all_info else 0, _ptr(out), cap)
        return out[:n].copy()

    def update_depth(self, depth, cam_pos, quat_wxyz):
        """depthPoseCallback + updateOccupancyCallback. Returns (fused, rotation 3x3 the reference derived)."""
        depth = np.ascontiguousarray(depth, dtype=np.uint16)
        rot = np.zeros(9)
        q = (C.c_double * 4)(*[float(x) for x in quat_wxyz])
        r = lib().ref_map_update_depth(self.h, _ptr(depth), depth.shape[0], depth.shape[1], _d3(cam_pos), q,
                                       rot.ctypes.data_as(C.POINTER(C.c_double)))
        return bool(r), rot.reshape(3, 3)

    def depth_odom_pose(self, body_pos, body_q_wxyz):
        """depthOdomCallback's pose algebra (grid_map.cpp:965-982): (camera position, camera quaternion w x y z)."""
        q = (C.c_double * 4)(*[float(x) for x in body_q_wxyz])
        cp, cq = (C.c_double * 3)(), (C.c_double * 4)()
        lib().ref_map_depth_odom_pose(self.h, _d3(body_pos), q, cp, cq)
        return np.array(cp[:]), np.array(cq[:])

    def occupancy(self):
        out = np.empty(self.dims, dtype=np.float64)
        lib().ref_map_download_occupancy(self.h, _ptr(out))
        return out

    def publish_occupancy(self, which):
        """publishMap (which=0) / publishUnknown (which=1): the PointXYZ records the reference published, (n,4) float32."""
        cap = int(np.prod(self.dims))
        out = np.empty((cap, 4), dtype=np.float32)
        n = lib().ref_map_publish_occupancy(self.h, which, _ptr(out), cap)
        return out[:n].copy()

    def raycast_num(self):
        return lib().ref_map_raycast_num(self.h)

    def set_raycast_num(self, v):
        lib().ref_map_set_raycast_num(self.h, int(v))

    def proj_points(self):
        return lib().ref_map_proj_points(self.h)

    def get_occupancy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        lib().ref_map_get_occupancy(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def set_occupied(self, pos):
        """GridMap::setOccupied (grid_map.h:310-321), one call per position in order."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        lib().ref_map_set_occupied(self.h, _ptr(pos), len(pos))

    def set_occupancy(self, pos, occ):
        """GridMap::setOccupancy (grid_map.h:323-337), one call per position in order."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        occ = np.ascontiguousarray(np.broadcast_to(np.asarray(occ, dtype=np.float64), (len(pos),)))
        lib().ref_map_set_occupancy(self.h, _ptr(pos), _ptr(occ), len(pos))

    def get_inflate_occupancy(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        lib().ref_map_get_inflate_occupancy(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def is_in_map(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.uint8)
        lib().ref_map_is_in_map(self.h, _ptr(pos), len(pos), _ptr(out))
        return out.astype(bool)

    def pos_to_index(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(pos), 3), dtype=np.int32)
        lib().ref_pos_to_index(self.h, _ptr(pos), len(pos), _ptr(out))
        return out

    def los_batch(self, starts, ends):
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(ends, dtype=np.float64).reshape(-1, 3)
        vis = np.empty(len(s), dtype=np.uint8)
        lib().ref_los_batch(self.h, _ptr(s), _ptr(e), len(s), _ptr(vis))
        return vis


class RefPlanner:
    """The reference's AStar / ThetaStar (setParam + setGridMap + init; reset + findPath per query)."""

    def __init__(self, algo, rmap, resolution, lambda_heuristic=5.0, allocate_num=100000, **agr):
        self.map = rmap
        from .pyoracle import AGR_DEFAULTS
        for k, v in dict(AGR_DEFAULTS, **agr).items():      # thetastaragr/* ROS params (launch/planner.launch:71-77)
            lib().ref_set_param(("thetastaragr/" + k).encode(), float(v))
        self.h = lib().ref_planner_create({"astar": 0, "thetastar": 1, "thetastaragr": 2, "thetastaragrpp": 3}[algo], resolution,
                                          lambda_heuristic, allocate_num, rmap.h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ref_planner_destroy(self.h)
            self.h = None

    def pool_disorder(self):
        """Adjacent pool nodes not in increasing address order (0 on a fresh heap). The reference's tie-break
        is the node ADDRESS, so its results equal 'creation order' semantics only when this is 0."""
        return lib().ref_planner_pool_disorder(self.h)

    def find_path(self, start, goal, path_cap=65536):
        r = Result()
        path = np.empty((path_cap, 3), dtype=np.float64)
        lib().ref_planner_find_path(self.h, _d3(start), _d3(goal), C.byref(r), _ptr(path), path_cap)
        res = np.frombuffer(bytes(r), dtype=RESULT_DTYPE)[0]
        return res, path[: min(r.n_path, path_cap)].copy()

    def visited(self):
        """AlgorithmBase::getVisitedNodes() after the last find_path: (n,3) positions."""
        n = lib().ref_planner_visited(self.h, None, 0)
        out = np.empty((n, 3), dtype=np.float64)
        lib().ref_planner_visited(self.h, _ptr(out), n)
        return out

    def find_paths(self, starts, goals, path_cap=65536):
        res = np.zeros(len(starts), dtype=RESULT_DTYPE)
        paths = []
        for i, (s, g) in enumerate(zip(starts, goals)):
            r, p = self.find_path(s, g, path_cap)
            res[i] = r
            paths.append(p)
        return res, paths
