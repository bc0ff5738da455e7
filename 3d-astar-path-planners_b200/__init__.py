"""B200-native data-parallel core of ChanJoon/3D-Astar-Path-Planners — Python face of the C-ABI.

Everything here is a thin ctypes binding of `libb200planner.so` (include/b200_planner.h).  The classes mirror
the reference's two C++ seams — `GridMap` (include/plan_env/grid_map.h:139-196) and the
`AlgorithmBase` family (include/Planners/AlgorithmBase.hpp:28-53) — with the reference's own method names
where a Python spelling exists.  There is NO CPU fallback: importing works without a GPU (so symbol/ABI
tests can run), but every compute call raises `B200Error` unless an sm_100 device executes it, and loading
fails loudly when the shared library is missing.

The directory name is not a valid Python identifier; import it through the repo-root shim
`astar_planners_b200.py`.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200planner.so")

ALGORITHMS = ("astar", "thetastar", "thetastaragr", "thetastaragrpp", "lazythetastar", "costastar", "costlazythetastar")
# thetastaragr/* defaults of launch/planner.launch:71-77
AGR_DEFAULTS = dict(barrier=2.0, flying_cost=10.0, flying_cost_default=0.0, ground_judge=0.1, epsilon=0.1)
STATUS_NAMES = {0: "solved", 1: "open_set_empty", 2: "pool_exhausted", 3: "tier_overflow", 4: "open_list_full"}
SEMANTICS = {"default": 0, "canonical": 1, "faithful": 2}


class B200Error(RuntimeError):
    pass


class MapParams(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("size", C.c_double * 3), ("resolution", C.c_double),
                ("obstacles_inflation", C.c_double), ("local_update_range", C.c_double * 3),
                ("ground_height", C.c_double), ("device", C.c_int32), ("reserved", C.c_int32)]


class PlannerParams(C.Structure):
    _fields_ = [("resolution", C.c_double), ("lambda_heuristic", C.c_double), ("allocate_num", C.c_int32),
                ("semantics", C.c_int32), ("cost_weight", C.c_double), ("max_los", C.c_double),
                ("agr_barrier", C.c_double), ("agr_flying_cost", C.c_double), ("agr_flying_cost_default", C.c_double),
                ("agr_ground_judge", C.c_double), ("agr_epsilon", C.c_double)]


class DepthParams(C.Structure):
    """b200_depth_params_t: the grid_map/* depth-fusion parameters (grid_map.cpp:21-43)."""
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("k_depth_scaling_factor", C.c_double), ("depth_filter_mindist", C.c_double),
                ("depth_filter_maxdist", C.c_double), ("max_ray_length", C.c_double),
                ("p_hit", C.c_double), ("p_miss", C.c_double), ("p_min", C.c_double), ("p_max", C.c_double),
                ("p_occ", C.c_double), ("virtual_ceil_height", C.c_double),
                ("use_depth_filter", C.c_int32), ("depth_filter_margin", C.c_int32), ("skip_pixel", C.c_int32),
                ("local_map_margin", C.c_int32)]


# the reference's defaults (grid_map.cpp:21-43)
DEPTH_DEFAULTS = dict(fx=387.229248046875, fy=387.229248046875, cx=321.04638671875, cy=243.44969177246094,
                      k_depth_scaling_factor=1000.0, depth_filter_mindist=0.2, depth_filter_maxdist=5.0, max_ray_length=4.5,
                      p_hit=0.70, p_miss=0.35, p_min=0.12, p_max=0.97, p_occ=0.80, virtual_ceil_height=2.5,
                      use_depth_filter=1, depth_filter_margin=1, skip_pixel=2, local_map_margin=1)


RESULT_DTYPE = np.dtype([("solved", "i4"), ("status", "i4"), ("n_path", "i4"), ("open_peak", "i4"), ("path_offset", "i8"),
                         ("g_final", "f8"), ("path_length", "f8"), ("explored_nodes", "i8"), ("iter_num", "i8"),
                         ("los_checks", "i8"), ("los_samples", "i8"), ("goal_coords", "f8", 3),
                         ("start_coords", "f8", 3), ("time_us", "f8")])
assert RESULT_DTYPE.itemsize == 128

# every symbol include/b200_planner.h declares: (name, restype, argtypes)
_vp, _i32, _i64, _dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
_pp = C.POINTER(C.c_void_p)
ABI = [
    ("b200_map_create", C.c_int, [C.POINTER(MapParams), _pp]),
    ("b200_map_destroy", C.c_int, [_vp]),
    ("b200_map_set_stream", C.c_int, [_vp, _vp]),
    ("b200_map_sync", C.c_int, [_vp]),
    ("b200_map_dims", C.c_int, [_vp, C.POINTER(_i32)]),
    ("b200_map_get_region", C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    ("b200_map_get_resolution", _dbl, [_vp]),
    ("b200_map_update_cloud", C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _i32, C.POINTER(_dbl), _i32]),
    ("b200_map_clear_box", C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_dbl)]),
    ("b200_map_reset", C.c_int, [_vp]),
    ("b200_map_local_bounds", C.c_int, [_vp, C.POINTER(_i32), C.POINTER(_i32)]),
    ("b200_map_get_inflate_occupancy", C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    ("b200_map_is_in_map", C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    ("b200_map_pos_to_index", C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    ("b200_map_download_inflate", C.c_int, [_vp, _vp]),
    ("b200_map_upload_inflate", C.c_int, [_vp, _vp]),
    ("b200_map_device_bits", C.c_int, [_vp, _pp, C.POINTER(_i64)]),
    ("b200_map_export_inflate_cloud", C.c_int, [_vp, _i32, _dbl, _vp, _i64, C.POINTER(_i64), _i32]),
    ("b200_map_set_depth_params", C.c_int, [_vp, C.POINTER(DepthParams)]),
    ("b200_map_update_depth", C.c_int, [_vp, _vp, _i32, _i32, C.POINTER(_dbl), C.POINTER(_dbl), _i32, C.POINTER(_i32)]),
    ("b200_map_get_occupancy", C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    ("b200_map_get_log_odds", C.c_int, [_vp, _vp, _i64, _vp, _i32]),
    ("b200_map_download_occupancy", C.c_int, [_vp, _vp]),
    ("b200_map_export_occupancy_cloud", C.c_int, [_vp, _i32, _i32, _dbl, _vp, _i64, C.POINTER(_i64), _i32]),
    ("b200_map_depth_stats", C.c_int, [_vp, C.POINTER(_i64)]),
    ("b200_map_set_depth_frame_count", C.c_int, [_vp, _i32]),
    ("b200_map_build_edt", C.c_int, [_vp, _dbl]),
    ("b200_map_set_edt_pipeline", C.c_int, [_vp, _i32]),
    ("b200_map_has_distance_field", C.c_int, [_vp, _dbl]),
    ("b200_map_set_occupied", C.c_int, [_vp, _vp, _i64, _i32]),
    ("b200_map_set_occupancy", C.c_int, [_vp, _vp, _vp, _i64, _i32]),
    ("b200_map_path_distance_metrics", C.c_int, [_vp, _vp, _i64, C.POINTER(_dbl), _vp]),
    ("b200_distance_metrics", C.c_int, [_vp, _i64, C.POINTER(_dbl)]),
    ("b200_map_edt_pipeline_used", C.c_int, [_vp]),
    ("b200_map_column_extents", C.c_int, [_vp, _vp, _vp, _i32]),
    ("b200_map_build_edt_slab", C.c_int, [_vp, _dbl, _vp, _vp, _i32]),
    ("b200_map_set_slab_extent", C.c_int, [_vp, _i32]),
    ("b200_map_device_fields", C.c_int, [_vp, C.POINTER(_vp), C.POINTER(_vp)]),
    ("b200_map_column_extents16", C.c_int, [_vp, _vp, _i32]),
    ("b200_map_slab_gaps", C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp]),
    ("b200_map_download_dist2", C.c_int, [_vp, _vp]),
    ("b200_map_download_costq", C.c_int, [_vp, _vp]),
    ("b200_map_download_cost", C.c_int, [_vp, _vp]),
    ("b200_los_batch", C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i32]),
    ("b200_los_batch_dda", C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i32]),
    ("b200_map_set_profiling", C.c_int, [_vp, _i32]),
    ("b200_map_launch_count", C.c_int, [_vp, C.POINTER(_i64)]),
    ("b200_map_get_stage_ms", C.c_int, [_vp, C.POINTER(_dbl)]),
    ("b200_planner_create", C.c_int, [C.c_char_p, C.POINTER(PlannerParams), _pp]),
    ("b200_planner_destroy", C.c_int, [_vp]),
    ("b200_planner_set_map", C.c_int, [_vp, _vp]),
    ("b200_odom_to_camera_pose", C.c_int, [C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl), C.POINTER(_dbl)]),
    ("b200_planner_set_stream", C.c_int, [_vp, _vp]),
    ("b200_stream_create", C.c_int, [_i32, _pp]),
    ("b200_stream_destroy", C.c_int, [_vp]),
    ("b200_planner_set_cost_factor", C.c_int, [_vp, _dbl]),
    ("b200_planner_detect_collision", C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_i32)]),
    ("b200_planner_find_path", C.c_int, [_vp, C.POINTER(_dbl), C.POINTER(_dbl), _vp, _vp, _i64]),
    ("b200_planner_find_paths", C.c_int, [_vp, _vp, _vp, _i64, _vp, _vp, _i64, _i32]),
    ("b200_planner_get_visited_nodes", C.c_int, [_vp, _vp, _i64, C.POINTER(_i64)]),
    ("b200_planner_last_launches", C.c_int, [_vp, C.POINTER(_i32)]),
    ("b200_last_error", C.c_char_p, []),
    ("b200_version", C.c_char_p, []),
    ("b200_host_alloc", C.c_int, [_pp, _i64]),
    ("b200_host_free", C.c_int, [_vp]),
]

_LIB = None


def build(verbose=False):
    """Compile libb200planner.so for sm_100a (nvcc cross-compiles without a GPU)."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-s", "-C", _HERE, "-j4"], stdout=out)
    return LIB_PATH


def lib():
    """Load the C-ABI library; raises (never falls back) when it is missing."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise B200Error(f"{LIB_PATH} is missing: run `make -C {_HERE}` (or __graft_entry__.build()). "
                            "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, res, args in ABI:
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = L
    return _LIB


def _ck(rc):
    if rc != 0:
        raise B200Error(f"b200 error {rc}: {lib().b200_last_error().decode()}")


def _d3(v):
    return (C.c_double * 3)(*[float(x) for x in v])


def _ptr(a):
    """host numpy array or device pointer carrier (anything with data_ptr(), e.g. a torch CUDA tensor)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return a.ctypes.data_as(C.c_void_p)


def _is_dev(a):
    return hasattr(a, "data_ptr") and getattr(a, "is_cuda", False)


class GridMap:
    """Mirror of the reference `GridMap`'s cloud path (src/plan_env/grid_map.cpp:6-171,711-804).

    `origin=None` reproduces initMap's origin (-sx/2, -sy/2, ground_height) (grid_map.cpp:53).
    """

    def __init__(self, size, resolution, origin=None, obstacles_inflation=0.099, local_update_range=(5.0, 5.0, 5.0),
                 ground_height=-0.01, device=0):
        if origin is None:
            origin = (-size[0] / 2.0, -size[1] / 2.0, ground_height)
        p = MapParams()
        p.origin[:] = [float(x) for x in origin]
        p.size[:] = [float(x) for x in size]
        p.resolution = float(resolution)
        p.obstacles_inflation = float(obstacles_inflation)
        p.local_update_range[:] = [float(x) for x in local_update_range]
        p.ground_height = float(ground_height)
        p.device = int(device)
        self.params = p
        self.h = C.c_void_p()
        _ck(lib().b200_map_create(C.byref(p), C.byref(self.h)))
        d = (C.c_int32 * 3)()
        _ck(lib().b200_map_dims(self.h, d))
        self.dims = tuple(d)
        self.origin, self.size, self.resolution = tuple(p.origin), tuple(p.size), p.resolution

    def close(self):
        if getattr(self, "h", None) and lib is not None:  # (module globals are gone when this runs at interpreter exit)
            lib().b200_map_destroy(self.h)
            self.h = None

    __del__ = close

    # --- GridMap API (reference names in comments) -------------------------------------------------
    def cloudCallback(self, cloud, camera_pos, point_step=None, offsets=(0, 4, 8)):
        """cloudCallback (grid_map.cpp:711-804). `cloud`: float32 array (n, step/4) or a CUDA tensor."""
        if _is_dev(cloud):
            step = point_step or cloud.stride(0) * cloud.element_size()
            n = cloud.shape[0]
            dev = 1
        else:
            cloud = np.ascontiguousarray(cloud)
            step = point_step or (cloud.shape[1] * cloud.itemsize if cloud.ndim == 2 else 16)
            n = cloud.nbytes // step
            dev = 0
        self._keep = cloud
        _ck(lib().b200_map_update_cloud(self.h, _ptr(cloud), n, step, *offsets, _d3(camera_pos), dev))

    update_cloud = cloudCallback

    def resetBuffer(self, min_pos=None, max_pos=None):
        """resetBuffer() / resetBuffer(min,max) (grid_map.cpp:144-171)."""
        if min_pos is None:
            _ck(lib().b200_map_reset(self.h))
        else:
            _ck(lib().b200_map_clear_box(self.h, _d3(min_pos), _d3(max_pos)))

    # ---- depth-image fusion (grid_map.cpp:173-697) ----
    def set_depth_params(self, **kw):
        """grid_map/* depth parameters by the reference's names; unspecified ones take its defaults."""
        d = dict(DEPTH_DEFAULTS)
        d.update(kw)
        p = DepthParams()
        for k, v in d.items():
            setattr(p, k, v)
        _ck(lib().b200_map_set_depth_params(self.h, C.byref(p)))
        self.depth_params = d

    def depthPoseCallback(self, depth, cam_pos, cam_rot):
        """depthPoseCallback + updateOccupancyCallback (grid_map.cpp:635-697) for one CV_16UC1 frame (numpy uint16
        (rows, cols) or a CUDA tensor); cam_rot = camera-to-world rotation (3x3). Returns False if the frame was dropped."""
        rot = (C.c_double * 9)(*[float(x) for x in np.asarray(cam_rot, dtype=np.float64).reshape(9)])
        fused = C.c_int32(0)
        if _is_dev(depth):
            rows, cols, dev = depth.shape[0], depth.shape[1], 1
        else:
            depth = np.ascontiguousarray(depth, dtype=np.uint16)
            rows, cols, dev = depth.shape[0], depth.shape[1], 0
        _ck(lib().b200_map_update_depth(self.h, _ptr(depth), rows, cols, _d3(cam_pos), rot, dev, C.byref(fused)))
        return bool(fused.value)

    def depthOdomCallback(self, depth, body_pos, body_q_wxyz):
        """depthOdomCallback (grid_map.cpp:965-994) + the occupancy timer body: the camera pose is the body odometry composed with
        cam2body_ and passed through the reference's matrix -> quaternion -> matrix round trip (`odom_to_camera_pose`)."""
        cam_pos, _, cam_rot = odom_to_camera_pose(body_pos, body_q_wxyz)
        return self.depthPoseCallback(depth, cam_pos, cam_rot)

    def getOccupancy(self, pos):
        """getOccupancy(Vector3d) (grid_map.h:339-348) for an (n,3) batch -> int8 {-1,0,1}."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        _ck(lib().b200_map_get_occupancy(self.h, _ptr(pos), len(pos), _ptr(out), 0))
        return out

    def setOccupied(self, pos):
        """setOccupied(Vector3d) (grid_map.h:310-321) for an (n,3) batch: inflated grid <- 1 where isInMap."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        _ck(lib().b200_map_set_occupied(self.h, _ptr(pos), len(pos), 0))

    def setOccupancy(self, pos, occ):
        """setOccupancy(Vector3d, double) (grid_map.h:323-337) for an (n,3) batch: log-odds grid <- occ (0 or 1 only)."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        occ = np.ascontiguousarray(np.broadcast_to(np.asarray(occ, dtype=np.float64), (len(pos),)))
        _ck(lib().b200_map_set_occupancy(self.h, _ptr(pos), _ptr(occ), len(pos), 0))

    def path_distance_metrics(self, path, return_distances=False):
        """GetPath.srv:20-23 -> dict(mean, std, min, max) of the distance to the nearest obstacle (m) over the path points
        (metrics.cpp:62-97 arithmetic over sqrt(dist2) * resolution)."""
        path = np.ascontiguousarray(path, dtype=np.float64).reshape(-1, 3)
        out = (C.c_double * 4)()
        d = np.empty(len(path), dtype=np.float64)
        _ck(lib().b200_map_path_distance_metrics(self.h, _ptr(path), len(path), out, _ptr(d)))
        r = dict(mean=out[0], std=out[1], min=out[2], max=out[3])
        return (r, d) if return_distances else r

    def log_odds(self, pos):
        """occupancy_buffer_ at the (clamped) voxel of each position."""
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.float64)
        _ck(lib().b200_map_get_log_odds(self.h, _ptr(pos), len(pos), _ptr(out), 0))
        return out

    def isUnknown(self, pos):
        """isUnknown(Vector3d) (grid_map.h:276-289)."""
        p = self.depth_params
        return self.log_odds(pos) < np.log(p["p_min"] / (1 - p["p_min"])) - 1e-3

    def occupancy(self):
        """occupancy_buffer_ (log-odds) as a (Nx,Ny,Nz) float64 array."""
        out = np.empty(self.dims, dtype=np.float64)
        _ck(lib().b200_map_download_occupancy(self.h, _ptr(out)))
        return out

    def _publish_occupancy(self, which, margin, truncate_height):
        n = C.c_int64(0)
        _ck(lib().b200_map_export_occupancy_cloud(self.h, which, margin, truncate_height, None, 0, C.byref(n), 0))
        out = np.empty((n.value, 4), dtype=np.float32)
        if n.value:
            _ck(lib().b200_map_export_occupancy_cloud(self.h, which, margin, truncate_height, _ptr(out), n.value, C.byref(n), 0))
        return out

    def publishMap(self, visualization_truncate_height=2.4):
        """publishMap (grid_map.cpp:806-849): occupied voxels (log-odds) of the local box as PointXYZ records (n,4)."""
        return self._publish_occupancy(0, int(self.depth_params["local_map_margin"]) // 2, visualization_truncate_height)

    def publishUnknown(self, visualization_truncate_height=2.4):
        """publishUnknown (grid_map.cpp:902-939)."""
        return self._publish_occupancy(1, 0, visualization_truncate_height)

    def depth_stats(self):
        out = (C.c_int64 * 8)()
        _ck(lib().b200_map_depth_stats(self.h, out))
        return dict(zip(("projected", "rays_cast", "ray_steps", "updates", "order_rounds", "frames"), list(out)[:6]))

    def set_depth_frame_count(self, n):
        _ck(lib().b200_map_set_depth_frame_count(self.h, int(n)))

    def getInflateOccupancy(self, pos, out=None):
        """getInflateOccupancy (grid_map.h:350-359) for an (n,3) batch -> int8 {-1,0,1}."""
        if _is_dev(pos):
            _ck(lib().b200_map_get_inflate_occupancy(self.h, _ptr(pos), pos.shape[0], _ptr(out), 1))
            return out
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.int8)
        _ck(lib().b200_map_get_inflate_occupancy(self.h, _ptr(pos), len(pos), _ptr(out), 0))
        return out

    def isInMap(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty(len(pos), dtype=np.uint8)
        _ck(lib().b200_map_is_in_map(self.h, _ptr(pos), len(pos), _ptr(out), 0))
        return out.astype(bool)

    def posToIndex(self, pos):
        pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
        out = np.empty((len(pos), 3), dtype=np.int32)
        _ck(lib().b200_map_pos_to_index(self.h, _ptr(pos), len(pos), _ptr(out), 0))
        return out

    def getRegion(self):
        o, s = (C.c_double * 3)(), (C.c_double * 3)()
        _ck(lib().b200_map_get_region(self.h, o, s))
        return tuple(o), tuple(s)

    def getResolution(self):
        return lib().b200_map_get_resolution(self.h)

    def local_bounds(self):
        a, b = (C.c_int32 * 3)(), (C.c_int32 * 3)()
        _ck(lib().b200_map_local_bounds(self.h, a, b))
        return tuple(a), tuple(b)

    def inflate(self):
        """occupancy_buffer_inflate_ as uint8 [Nx,Ny,Nz] (reference byte layout)."""
        out = np.empty(self.dims, dtype=np.uint8)
        _ck(lib().b200_map_download_inflate(self.h, _ptr(out)))
        return out

    def set_inflate(self, grid):
        grid = np.ascontiguousarray(grid, dtype=np.uint8)
        assert grid.shape == self.dims
        _ck(lib().b200_map_upload_inflate(self.h, _ptr(grid)))

    def publishMapInflate(self, all_info=False, local_map_margin=1, visualization_truncate_height=2.4, cap=None):
        """publishMapInflate (grid_map.cpp:851-900): the PointXYZ records (x, y, z, 1.0) of the PointCloud2 it publishes."""
        cap = int(np.prod(self.dims)) if cap is None else int(cap)
        out = np.empty((cap, 4), dtype=np.float32)
        n = C.c_int64()
        _ck(lib().b200_map_export_inflate_cloud(self.h, local_map_margin if all_info else 0, float(visualization_truncate_height),
                                                _ptr(out), cap, C.byref(n), 0))
        return out[: min(n.value, cap)].copy()

    # --- new functionality ------------------------------------------------------------------------
    def build_edt(self, d_safe=0.0):
        _ck(lib().b200_map_build_edt(self.h, float(d_safe)))

    def has_distance_field(self, d_safe=0.0):
        return bool(lib().b200_map_has_distance_field(self.h, float(d_safe)))

    def set_edt_pipeline(self, mode):
        """0: automatic (fused shared-memory kernels when they fit), 1: always the seven-kernel HBM-stack pipeline."""
        _ck(lib().b200_map_set_edt_pipeline(self.h, int(mode)))

    def edt_pipeline_used(self):
        return "fused" if lib().b200_map_edt_pipeline_used(self.h) else "hbm-stack"

    def column_extents(self, lowest=None, highest=None):
        """Lowest / highest occupied z per (x,y) column (-1 = empty): what a z-slab publishes to the others."""
        if _is_dev(lowest):
            _ck(lib().b200_map_column_extents(self.h, _ptr(lowest), _ptr(highest), 1))
            return lowest, highest
        lo = np.empty(self.dims[:2], dtype=np.int32)
        hi = np.empty(self.dims[:2], dtype=np.int32)
        _ck(lib().b200_map_column_extents(self.h, _ptr(lo), _ptr(hi), 0))
        return lo, hi

    def column_extents16(self, lo_hi=None):
        """Packed (lowest, highest) int16 pair per column; lo_hi: optional device tensor [Nx, Ny, 2] int16 to fill."""
        if lo_hi is not None and _is_dev(lo_hi):
            _ck(lib().b200_map_column_extents16(self.h, _ptr(lo_hi), 1))
            return lo_hi
        out = np.empty(self.dims[:2] + (2,), dtype=np.int16)
        _ck(lib().b200_map_column_extents16(self.h, _ptr(out), 0))
        return out

    def slab_gaps(self, gathered, slab_nz, rank, below, above):
        """Device-side halo derivation: gathered = device tensor [P, Nx, Ny, 2] int16 (all-gathered column_extents16),
        below / above = device int32 [Nx, Ny] outputs for b200_map_build_edt_slab."""
        nz = np.ascontiguousarray(slab_nz, dtype=np.int32)
        _ck(lib().b200_map_slab_gaps(self.h, _ptr(gathered), _ptr(nz), len(nz), int(rank), _ptr(below), _ptr(above)))
        return below, above

    def set_slab_extent(self, total_nz):
        """z-slab sharding: height (voxels) of the whole stacked grid — bounds the halo distances (narrower hull entries)."""
        _ck(lib().b200_map_set_slab_extent(self.h, int(total_nz)))

    def build_edt_slab(self, d_safe=0.0, below_gap=None, above_gap=None):
        """EDT of one z-slab given the per-column gaps to the nearest occupied voxel in the slabs below / above."""
        dev = 1 if (_is_dev(below_gap) or _is_dev(above_gap)) else 0
        if not dev:
            below_gap = None if below_gap is None else np.ascontiguousarray(below_gap, dtype=np.int32)
            above_gap = None if above_gap is None else np.ascontiguousarray(above_gap, dtype=np.int32)
        self._keep_gaps = (below_gap, above_gap)
        _ck(lib().b200_map_build_edt_slab(self.h, float(d_safe), _ptr(below_gap), _ptr(above_gap), dev))

    def device_fields(self):
        """(dist2, costq) as zero-copy torch CUDA tensors [Nx, Ny, Nz] (int32 / int16 view of the uint16 cost); None if absent."""
        import torch
        d, c = C.c_void_p(), C.c_void_p()
        _ck(lib().b200_map_device_fields(self.h, C.byref(d), C.byref(c)))

        class _Arr:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = dict(shape=shape, typestr=typestr, data=(ptr, False), version=3)

        dev = torch.device("cuda", int(self.params.device))
        td = torch.as_tensor(_Arr(d.value, tuple(self.dims), "<i4"), device=dev) if d.value else None
        tc = torch.as_tensor(_Arr(c.value, tuple(self.dims), "<i2"), device=dev) if c.value else None
        return td, tc

    def dist2(self):
        out = np.empty(self.dims, dtype=np.int32)
        _ck(lib().b200_map_download_dist2(self.h, _ptr(out)))
        return out

    def costq(self):
        out = np.empty(self.dims, dtype=np.uint16)
        _ck(lib().b200_map_download_costq(self.h, _ptr(out)))
        return out

    def cost(self):
        out = np.empty(self.dims, dtype=np.float32)
        _ck(lib().b200_map_download_cost(self.h, _ptr(out)))
        return out

    def los_batch(self, starts, ends, return_samples=False, vis_out=None, samples_out=None):
        """LineOfSight::fastLOS (LineOfSight.cpp:9-27) for n segments."""
        if _is_dev(starts):
            _ck(lib().b200_los_batch(self.h, _ptr(starts), _ptr(ends), starts.shape[0], _ptr(vis_out), _ptr(samples_out), 1))
            return vis_out
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(ends, dtype=np.float64).reshape(-1, 3)
        vis = np.empty(len(s), dtype=np.uint8)
        ns = np.empty(len(s), dtype=np.int32)
        _ck(lib().b200_los_batch(self.h, _ptr(s), _ptr(e), len(s), _ptr(vis), _ptr(ns), 0))
        return (vis, ns) if return_samples else vis

    def los_batch_dda(self, starts, ends, return_cells=False, vis_out=None, cells_out=None):
        """Voxel-traversal (Amanatides-Woo, raycast.cpp:253-346) line of sight for n segments."""
        if _is_dev(starts):
            _ck(lib().b200_los_batch_dda(self.h, _ptr(starts), _ptr(ends), starts.shape[0], _ptr(vis_out), _ptr(cells_out), 1))
            return vis_out
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        e = np.ascontiguousarray(ends, dtype=np.float64).reshape(-1, 3)
        vis = np.empty(len(s), dtype=np.uint8)
        nc = np.empty(len(s), dtype=np.int32)
        _ck(lib().b200_los_batch_dda(self.h, _ptr(s), _ptr(e), len(s), _ptr(vis), _ptr(nc), 0))
        return (vis, nc) if return_cells else vis

    def set_stream(self, cuda_stream_ptr):
        _ck(lib().b200_map_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        _ck(lib().b200_map_sync(self.h))

    def launch_count(self):
        n = C.c_int64()
        _ck(lib().b200_map_launch_count(self.h, C.byref(n)))
        return int(n.value)

    def set_profiling(self, on=True):
        _ck(lib().b200_map_set_profiling(self.h, 1 if on else 0))

    def stage_ms(self):
        out = (C.c_double * 16)()
        _ck(lib().b200_map_get_stage_ms(self.h, out))
        return list(out)


class Planner:
    """Mirror of `Planners::AlgorithmBase` (AlgorithmBase.hpp:28-53): setGridMap / setCostFactor /
    detectCollision / findPath / getPath, plus the batched `findPaths`."""

    def __init__(self, algorithm="astar", resolution=0.1, lambda_heuristic=5.0, allocate_num=1000000, cost_weight=0.0,
                 max_los=0.0, semantics="default", **agr):
        """semantics: "faithful" (bit-identical to the reference's std::set behaviour; astar/thetastar only),
        "canonical" (exact priority queue, faster), "default" (faithful where the reference has the algorithm)."""
        a = dict(AGR_DEFAULTS, **agr)
        p = PlannerParams(float(resolution), float(lambda_heuristic), int(allocate_num), SEMANTICS[semantics],
                          float(cost_weight), float(max_los), a["barrier"], a["flying_cost"], a["flying_cost_default"],
                          a["ground_judge"], a["epsilon"])
        self.h = C.c_void_p()
        _ck(lib().b200_planner_create(algorithm.encode(), C.byref(p), C.byref(self.h)))
        self.algorithm = algorithm if algorithm in ALGORITHMS else "astar"
        self._path = np.zeros((0, 3))
        self.map = None

    def close(self):
        if getattr(self, "h", None) and lib is not None:
            lib().b200_planner_destroy(self.h)
            self.h = None

    __del__ = close

    def setGridMap(self, grid_map):
        self.map = grid_map
        _ck(lib().b200_planner_set_map(self.h, grid_map.h))

    def setCostFactor(self, w):
        _ck(lib().b200_planner_set_cost_factor(self.h, float(w)))

    def set_stream(self, cuda_stream_ptr):
        """Own CUDA stream for this planner (None/0 = the map's stream): batches of planners on different streams overlap."""
        _ck(lib().b200_planner_set_stream(self.h, C.c_void_p(cuda_stream_ptr or 0)))

    def detectCollision(self, pos):
        out = C.c_int32()
        _ck(lib().b200_planner_detect_collision(self.h, _d3(pos), C.byref(out)))
        return bool(out.value)

    def findPath(self, source, target, path_cap=65536):
        """reset() + findPath(src,tgt) (AStar.cpp:80-179). Returns the PathData-like dict of
        createResultDataObject (AlgorithmBase.cpp:123-175)."""
        res = np.zeros(1, dtype=RESULT_DTYPE)
        path = np.empty((path_cap, 3), dtype=np.float64)
        _ck(lib().b200_planner_find_path(self.h, _d3(source), _d3(target), _ptr(res), _ptr(path), path_cap))
        r = res[0]
        self._path = path[: min(int(r["n_path"]), path_cap)].copy()
        return {
            "solved": bool(r["solved"]), "goal_coords": r["goal_coords"].copy(), "g_final_node": float(r["g_final"]),
            "algorithm": self.algorithm, "path": self._path, "time_spent": float(r["time_us"]),
            "explored_nodes": int(r["explored_nodes"]), "start_coords": r["start_coords"].copy(),
            "path_length": float(r["path_length"]), "total_cost1": 0, "total_cost2": 0, "h_cost": 0, "g_cost1": 0,
            "g_cost2": 0, "c_cost": 0, "grid_cost1": 0, "grid_cost2": 0, "line_of_sight_checks": int(r["los_checks"]),
            "status": STATUS_NAMES.get(int(r["status"]), "?"), "iter_num": int(r["iter_num"]),
            "los_samples": int(r["los_samples"]),
        }

    def getPath(self):
        return self._path

    def getVisitedNodes(self):
        """getVisitedNodes (AlgorithmBase.cpp:185-189) of the last findPath: (n,3) positions in creation order."""
        n = C.c_int64(0)
        _ck(lib().b200_planner_get_visited_nodes(self.h, None, 0, C.byref(n)))
        out = np.empty((n.value, 3), dtype=np.float64)
        if n.value:
            _ck(lib().b200_planner_get_visited_nodes(self.h, _ptr(out), n.value, C.byref(n)))
        return out

    def findPaths(self, starts, goals, path_cap=None, results=None, path_buf=None):
        """nq independent reset()+findPath() pairs. Host numpy in/out, or CUDA tensors for all of
        starts/goals/results(/path_buf) (results: uint8 tensor of nq*128 bytes)."""
        if _is_dev(starts):
            nq = starts.shape[0]
            cap = 0 if path_buf is None else path_buf.shape[0]
            _ck(lib().b200_planner_find_paths(self.h, _ptr(starts), _ptr(goals), nq, _ptr(results), _ptr(path_buf), cap, 1))
            return results, path_buf
        s = np.ascontiguousarray(starts, dtype=np.float64).reshape(-1, 3)
        g = np.ascontiguousarray(goals, dtype=np.float64).reshape(-1, 3)
        nq = len(s)
        res = np.zeros(nq, dtype=RESULT_DTYPE) if results is None else results
        cap = int(path_cap) if path_cap is not None else 0
        path = (np.empty((cap, 3), dtype=np.float64) if path_buf is None else path_buf) if cap else None
        _ck(lib().b200_planner_find_paths(self.h, _ptr(s), _ptr(g), nq, _ptr(res), _ptr(path), cap, 0))
        return res, path

    def last_launches(self):
        out = C.c_int32()
        _ck(lib().b200_planner_last_launches(self.h, C.byref(out)))
        return out.value


class PlannerPool:
    """Several `Planner` handles on ONE map, each with its own CUDA stream and host thread: query batches submitted to
    the pool overlap on the device, so the straggling queries of one batch run beside the first queries of the next
    instead of leaving the SMs idle (a batch is as long as its slowest query; measured on one B200, 131 072-query
    batches: Lazy Theta* 0.87 -> 1.02 M queries/s with two lanes, faithful A* 12 k -> 43 k with four).  New: the reference
    plans one query at a time on the caller's thread.  Results are bit-identical to the single-handle run."""

    def __init__(self, grid_map, algorithm="astar", lanes=2, **planner_kw):
        self.lanes = int(lanes)
        self.streams, self.planners = [], []  # streams: raw cudaStream_t values (b200_stream_create: no CUDA runtime binding needed)
        for _ in range(self.lanes):
            st = C.c_void_p()
            _ck(lib().b200_stream_create(int(grid_map.params.device), C.byref(st)))
            self.streams.append(st.value)
            pl = Planner(algorithm, **planner_kw)
            pl.setGridMap(grid_map)
            pl.set_stream(st.value)
            self.planners.append(pl)
        grid_map.sync()  # the lanes' streams start after whatever built the map

    def close(self):
        for pl in self.planners:
            pl.close()
        self.planners = []
        if lib is not None:
            for st in self.streams:
                lib().b200_stream_destroy(C.c_void_p(st))
        self.streams = []

    __del__ = close

    def find_paths_many(self, batches, **kw):
        """batches: list of (starts, goals[, results]) — host arrays, or CUDA tensors with a preallocated results tensor.
        Batch i runs on lane i % lanes, the batches of a lane in order; returns the per-batch `findPaths` results in order."""
        import threading
        out, errs = [None] * len(batches), []

        def lane_main(l):
            try:
                for i in range(l, len(batches), self.lanes):
                    b = batches[i]
                    out[i] = self.planners[l].findPaths(b[0], b[1], results=b[2] if len(b) > 2 else None, **kw)
            except Exception as e:  # re-raised by the caller's thread
                errs.append(e)

        th = [threading.Thread(target=lane_main, args=(l,)) for l in range(min(self.lanes, len(batches)))]
        for t in th:
            t.start()
        for t in th:
            t.join()
        if errs:
            raise errs[0]
        return out


def odom_to_camera_pose(body_pos, body_q_wxyz):
    """(camera position, camera quaternion w x y z, rotation 3x3) exactly as GridMap::depthOdomCallback + projectDepthImage derive
    them from a body odometry (grid_map.cpp:965-982, :208) — host only, no GPU."""
    q = (C.c_double * 4)(*[float(x) for x in body_q_wxyz])
    cp, cq, cr = (C.c_double * 3)(), (C.c_double * 4)(), (C.c_double * 9)()
    _ck(lib().b200_odom_to_camera_pose(_d3(body_pos), q, cp, cq, cr))
    return np.array(cp[:]), np.array(cq[:]), np.array(cr[:]).reshape(3, 3)


def distance_metrics(distances):
    """{mean, std, min, max} with the arithmetic of the reference's calculateDistancesMetrics (host only, no GPU)."""
    d = np.ascontiguousarray(distances, dtype=np.float64)
    out = (C.c_double * 4)()
    _ck(lib().b200_distance_metrics(_ptr(d), len(d), out))
    return dict(mean=out[0], std=out[1], min=out[2], max=out[3])


def path_of(results, path, i):
    r = results[i]
    if r["path_offset"] < 0:
        return None
    return path[r["path_offset"]: r["path_offset"] + r["n_path"]]
