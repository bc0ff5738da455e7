/*
 * b200_planner.h — C-ABI of libb200planner.so: the B200 (sm_100a) replacement for the data-parallel hot
 * path of ChanJoon/3D-Astar-Path-Planners (ROS package heuristic_planners).
 *
 * The reference has no FFI layer: its seam is two C++ classes, `GridMap` (include/plan_env/grid_map.h:139-196)
 * and `Planners::AlgorithmBase` (include/Planners/AlgorithmBase.hpp:28-53), consumed by
 * `HeuristicPlannerROS` (src/planner_ros_node.cpp).  Every entry point below names the reference member it
 * replaces (file:line relative to the reference root).  The C++ adapter classes that put the reference's
 * own method names back on top of this ABI live in 3d-astar-path-planners_b200/host/; INTEGRATION.md shows
 * the binding a maintainer would add to the ROS node.
 *
 * Conventions: plain pointers and sizes, opaque handles, `int` status (0 = OK, <0 = error, see
 * b200_status_t), no exceptions across the boundary, caller-owned output buffers.  A handle is
 * thread-compatible (not thread-safe).  All work of a handle is issued on one CUDA stream (default: the
 * legacy default stream, so it orders with a framework's default stream; see b200_map_set_stream); calls
 * that return data to host memory synchronise that stream before returning.  Calls taking `on_device != 0`
 * expect device pointers on the handle's device and do not synchronise.
 *
 * There is no CPU fallback: every call fails with B200_ERR_CUDA if no sm_100 device is usable.
 */
#ifndef B200_PLANNER_H
#define B200_PLANNER_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200_OK = 0,
  B200_ERR_INVALID = -1,     /* bad argument / NULL handle */
  B200_ERR_CUDA = -2,        /* CUDA runtime error; text via b200_last_error() */
  B200_ERR_NOMEM = -3,       /* device or host allocation failed */
  B200_ERR_UNSUPPORTED = -4, /* parameter combination outside the supported envelope */
  B200_ERR_NO_MAP = -5       /* planner used before b200_planner_set_map */
} b200_status_t;

typedef struct b200_map b200_map_t;
typedef struct b200_planner b200_planner_t;

/* ---------------------------------------------------------------------------------------------------- */
/* GridMap                                                                                              */
/* ---------------------------------------------------------------------------------------------------- */

/* MappingParameters subset used by the cloud path (include/plan_env/grid_map.h:50-58).  `origin` is explicit;
 * the reference derives it as (-size_x/2, -size_y/2, ground_height) (src/plan_env/grid_map.cpp:53). */
typedef struct {
  double origin[3];             /* map_origin_            grid_map.cpp:53  */
  double size[3];               /* map_size_              grid_map.cpp:54  */
  double resolution;            /* grid_map/resolution    grid_map.cpp:12  */
  double obstacles_inflation;   /* grid_map/obstacles_inflation grid_map.cpp:19 */
  double local_update_range[3]; /* grid_map/local_update_range_{x,y,z} grid_map.cpp:16-18 */
  double ground_height;         /* grid_map/ground_height grid_map.cpp:50  */
  int32_t device;               /* CUDA device ordinal */
  int32_t reserved;
} b200_map_params_t;

/* GridMap::initMap — geometry + buffer allocation (grid_map.cpp:52-54,69-80).  Grid starts all-free. */
int b200_map_create(const b200_map_params_t* params, b200_map_t** out);
int b200_map_destroy(b200_map_t* map);
/* Use an existing cudaStream_t (passed as void*) for all work of this map and of planners bound to it. */
int b200_map_set_stream(b200_map_t* map, void* cuda_stream);
int b200_map_sync(b200_map_t* map);

/* map_voxel_num_ (grid_map.cpp:69-70) */
int b200_map_dims(const b200_map_t* map, int32_t n[3]);
/* GridMap::getRegion (grid_map.cpp:960-963) */
int b200_map_get_region(const b200_map_t* map, double origin[3], double size[3]);
/* GridMap::getResolution (grid_map.h:443-446) */
double b200_map_get_resolution(const b200_map_t* map);

/* GridMap::cloudCallback (grid_map.cpp:711-804): clear the camera-centred box, then voxelise + inflate every
 * point (float32 x,y,z at byte offsets off_* inside records of `point_step` bytes — the PointCloud2 layout
 * pcl::fromROSMsg reads, grid_map.cpp:713-714), and update the local bounds.  n == 0 or a NaN camera is a
 * no-op exactly like :724-728.  `pts` is a host pointer (copied inside) unless on_device != 0. */
int b200_map_update_cloud(b200_map_t* map, const void* pts, int64_t n, int32_t point_step, int32_t off_x,
                          int32_t off_y, int32_t off_z, const double cam[3], int32_t on_device);
/* GridMap::resetBuffer(min,max) (grid_map.cpp:155-171) and resetBuffer() (:144-153). */
int b200_map_clear_box(b200_map_t* map, const double min_pos[3], const double max_pos[3]);
int b200_map_reset(b200_map_t* map);
/* md_.local_bound_min_/max_ after the last cloud update (grid_map.cpp:799-803). */
int b200_map_local_bounds(b200_map_t* map, int32_t min_idx[3], int32_t max_idx[3]);

/* GridMap::getInflateOccupancy (grid_map.h:350-359): out[i] in {-1 outside map, 0 free, 1 occupied}. */
int b200_map_get_inflate_occupancy(b200_map_t* map, const double* pos_xyz, int64_t n, int8_t* out,
                                   int32_t on_device);
/* GridMap::isInMap(Vector3d) (grid_map.h:370-385): out[i] in {0,1}. */
int b200_map_is_in_map(b200_map_t* map, const double* pos_xyz, int64_t n, uint8_t* out, int32_t on_device);
/* GridMap::posToIndex (grid_map.h:400-404): out[3*i..] = floor((pos-origin)*res_inv). */
int b200_map_pos_to_index(b200_map_t* map, const double* pos_xyz, int64_t n, int32_t* out_idx, int32_t on_device);

/* occupancy_buffer_inflate_ in the reference's own layout: one byte per voxel, address x*Ny*Nz + y*Nz + z
 * (grid_map.h:257-260).  Download is for parity / publishers; upload replaces the grid (≡ setOccupied loops). */
int b200_map_download_inflate(b200_map_t* map, uint8_t* dst);
int b200_map_upload_inflate(b200_map_t* map, const uint8_t* src);
/* Device pointer + word count of the bit-packed grid (uint32 words, [x][y][ceil(Nz/32)], bit z&31). */
int b200_map_device_bits(b200_map_t* map, void** dev_ptr, int64_t* n_words);

/* GridMap::publishMapInflate (grid_map.cpp:851-900): the occupied inflated voxels of the local-bounds box (extended by
 * `margin` voxels = local_map_margin when all_info, 0 otherwise), x -> y -> z order, voxel centres not above
 * `truncate_height` (grid_map/visualization_truncate_height), as 16-byte PointXYZ records (float x, y, z, 1.0f) — the
 * PointCloud2 data blob the reference publishes.  *n_out = number of points (may exceed cap; only cap are written). */
int b200_map_export_inflate_cloud(b200_map_t* map, int32_t margin, double truncate_height, float* xyz1, int64_t cap,
                                  int64_t* n_out, int32_t on_device);

/* ---- depth-image fusion: the other producer of occupancy_buffer_inflate_ -------------------------
 * grid_map/... parameters of the depth path (grid_map.cpp:21-43, same names and defaults).  depth_filter_tolerance and
 * min_ray_length are read by the reference but never used by live code, so they are not here. */
typedef struct b200_depth_params {
  double fx, fy, cx, cy;              /* 387.229248046875, 387.229248046875, 321.04638671875, 243.44969177246094 */
  double k_depth_scaling_factor;      /* 1000.0 */
  double depth_filter_mindist;        /* 0.2 */
  double depth_filter_maxdist;        /* 5.0 */
  double max_ray_length;              /* 4.5 */
  double p_hit, p_miss, p_min, p_max, p_occ; /* 0.70 0.35 0.12 0.97 0.80 */
  double virtual_ceil_height;         /* 2.5 (a ceiling plane is stamped when > -0.5) */
  int32_t use_depth_filter;           /* 1 */
  int32_t depth_filter_margin;        /* 1 */
  int32_t skip_pixel;                 /* 2 */
  int32_t local_map_margin;           /* 1 */
} b200_depth_params_t;
/* Stores the parameters and derives the log-odds constants (logit macro, grid_map.h:27; grid_map.cpp:57-62). */
int b200_map_set_depth_params(b200_map_t* map, const b200_depth_params_t* p);
/* One depth frame: GridMap::depthPoseCallback (grid_map.cpp:668-697: image, pose, the camera-inside-the-map gate)
 * followed by the occupancy timer body GridMap::updateOccupancyCallback (grid_map.cpp:635-666): projectDepthImage ->
 * raycastProcess -> clearAndInflateLocalMap.  `depth`: rows*cols uint16 (CV_16UC1), host or device memory.
 * `cam_rot`: camera-to-world rotation, row-major 3x3 — what Eigen's camera_q_.toRotationMatrix() holds on the caller's
 * side.  *fused = 0 when the camera is outside the map (the frame is dropped, as in the reference).
 * Results are bit-identical to the reference's sequential loop, including its ray-order dependence (see
 * csrc/depth_kernels.cu).  Allocates the log-odds grid, counters and flags on first use (38 bytes per voxel). */
int b200_map_update_depth(b200_map_t* map, const uint16_t* depth, int32_t rows, int32_t cols, const double cam_pos[3],
                          const double cam_rot[9], int32_t on_device, int32_t* fused);
/* GridMap::getOccupancy(Vector3d) (grid_map.h:339-348): -1 outside, else occupancy_buffer_ > logit(p_occ). */
int b200_map_get_occupancy(b200_map_t* map, const double* pos_xyz, int64_t n, int8_t* out, int32_t on_device);
/* GridMap::setOccupied(Vector3d) (grid_map.h:310-321), batched: the inflated grid is set to 1 at the voxel of every
 * position that isInMap; positions outside are ignored.  Invalidates the distance field. */
int b200_map_set_occupied(b200_map_t* map, const double* pos_xyz, int64_t n, int32_t on_device);
/* GridMap::setOccupancy(Vector3d, double) (grid_map.h:323-337), batched: occupancy_buffer_ (the log-odds grid the depth
 * path fuses into) is set to occ[i] at the voxel of pos i.  Like the reference, values other than 0 and 1 and
 * positions outside the map are ignored; positions apply in index order (the last one on a voxel wins).  Needs
 * b200_map_set_depth_params (it owns that grid). */
int b200_map_set_occupancy(b200_map_t* map, const double* pos_xyz, const double* occ, int64_t n, int32_t on_device);
/* The raw log-odds of the voxel holding each position, index clamped into the map (posToIndex + boundIndex): the read
 * behind GridMap::isUnknown (value < logit(p_min) - 1e-3) and isKnownFree (value >= logit(p_min) and the inflated voxel
 * is free), grid_map.h:276-300. */
int b200_map_get_log_odds(b200_map_t* map, const double* pos_xyz, int64_t n, double* out, int32_t on_device);
/* occupancy_buffer_ (log-odds, double per voxel, address x*Ny*Nz + y*Nz + z) to host memory. */
int b200_map_download_occupancy(b200_map_t* map, double* dst);
/* GridMap::publishMap (grid_map.cpp:806-849; which = 0: voxels whose log-odds are not below logit(p_occ), box = local
 * bounds +- margin with margin = local_map_margin / 2) and GridMap::publishUnknown (:902-939; which = 1: log-odds below
 * clamp_min - 1e-3, margin 0): voxel centres in x -> y -> z order, not above `truncate_height`, as 16-byte PointXYZ
 * records (float x, y, z, 1.0f).  *n_out = number of points (may exceed cap; only cap are written). */
int b200_map_export_occupancy_cloud(b200_map_t* map, int32_t which, int32_t margin, double truncate_height, float* xyz1,
                                    int64_t cap, int64_t* n_out, int32_t on_device);
/* Counters of the last fused frame: [0] projected points, [1] rays cast, [2] voxels stepped through, [3] log-odds
 * updates applied, [4] rounds of the ray-order fixed-point iteration, [5] raycast_num_ (int8, wraps). */
int b200_map_depth_stats(b200_map_t* map, int64_t out[8]);
/* Sets raycast_num_ (a `char`, grid_map.h:124; the value is reduced to int8).  It matters beyond bookkeeping: the
 * reference never clears its per-voxel ray flags, they hold the frame number that last set them, and the number wraps
 * every 256 fused frames — stale flags then suppress rays (see csrc/depth_kernels.cu).  Kept bit-exactly, so a
 * restored map must restore the counter as well. */
int b200_map_set_depth_frame_count(b200_map_t* map, int32_t raycast_num);

/* NEW (not in the reference; replaces the commented-out evaluateCoarseEDT hook, AlgorithmBase.cpp:40-42):
 * exact squared Euclidean distance (voxel units, int32; 1<<29 where no obstacle exists) of every voxel to
 * the nearest inflated-occupied voxel, plus — when d_safe > 0 — the safety-cost field
 * c = max(0, 1 - sqrt(d2)*res/d_safe) (float32) and its uint16 quantisation the cost-aware planners read. */
int b200_map_build_edt(b200_map_t* map, double d_safe);
/* The distance / cost field describes the inflated grid at the time of the build: every call that rewrites the grid
 * (update_cloud, update_depth, clear_box, reset, upload_inflate, set_occupancy) invalidates it, and the field readers and
 * cost-aware planners then fail with B200_ERR_INVALID until build_edt runs again.
 * Pipeline selection (diagnostic; outputs are bit-identical): 0 = automatic (two shared-memory-staged kernels when a
 * line's hull stack fits in shared memory, else the seven-kernel pipeline with stacks in HBM), 1 = always the latter.
 * b200_map_edt_pipeline_used reports what the last build ran (1 = fused, 0 = seven-kernel). */
int b200_map_has_distance_field(const b200_map_t* map, double d_safe); /* 1: built for the current grid with this d_safe */
int b200_map_set_edt_pipeline(b200_map_t* map, int32_t mode);
int b200_map_edt_pipeline_used(const b200_map_t* map);
/* z-slab sharding of the distance transform for grids too large for one GPU (north star).  The map is
 * one z-slab (full x,y extent).  Because the z pass runs first and on bits, the only information an EXACT,
 * untruncated transform needs from the other slabs is, per (x,y) column (index x*Ny + y), the distance in voxels
 * from this slab's first / last z-plane to the nearest occupied voxel below / above it (>= 1; < 0 = none).
 * b200_map_column_extents publishes a slab's lowest / highest occupied z per column (-1 = empty column); the host
 * all-gathers those (NCCL/NVLink) and derives the gaps (shard.py: slab_gaps). NULL gap arrays = no neighbour. */
int b200_map_column_extents(b200_map_t* map, int32_t* lowest_z, int32_t* highest_z, int32_t on_device);
/* The exchange in 4 bytes per column: lo_hi = [Nx*Ny][2] int16 (lowest, highest occupied z; -1 = empty column), the
 * buffer the slabs all-gather; b200_map_slab_gaps then derives THIS slab's two gap arrays on the device from the
 * gathered [n_slabs][Nx*Ny][2] tensor (device pointers; slab_nz = the slabs' heights in voxels, host array). */
int b200_map_column_extents16(b200_map_t* map, int16_t* lo_hi, int32_t on_device);
int b200_map_slab_gaps(b200_map_t* map, const int16_t* gathered_lo_hi, const int32_t* slab_nz, int32_t n_slabs, int32_t rank,
                       int32_t* below_gap, int32_t* above_gap);
/* Optional hint for the slab build: the height (voxels) of the whole stacked grid, i.e. a bound on every gap + slab
 * height.  It only selects narrower hull entries (32 instead of 64 bits) in the shared-memory kernels; results do not
 * depend on it as long as the bound holds.  0 = unknown (default). */
int b200_map_set_slab_extent(b200_map_t* map, int32_t total_nz);
int b200_map_build_edt_slab(b200_map_t* map, double d_safe, const int32_t* below_gap, const int32_t* above_gap,
                            int32_t on_device);
int b200_map_download_dist2(b200_map_t* map, int32_t* dst);
/* Device pointers of the resident fields ([Nx][Ny][Nz] int32 squared distance, uint16 quantised cost; NULL when not
 * built) for consumers that stay on the GPU (the planners read them this way; bench.py compares slabs with them). */
int b200_map_device_fields(b200_map_t* map, void** dist2, void** costq);
/* The distance-to-obstacle statistics GetPath.srv:20-23 carries (min/max/mean/std_dev; the reference leaves them 0
 * because its EDT hook is commented out): d_i = sqrt(dist2[voxel(p_i)]) * resolution at every path point (index clamped
 * into the map, grid_map.h:267-274), then out = {mean, std_dev, min, max} with the element order and arithmetic of
 * calculateDistancesMetrics / getAverage / getStdDev (src/utils/metrics.cpp:62-97; population std-dev).  n = 0 gives
 * zeros (:77).  distances (optional, n doubles) receives the d_i.  path_xyz is a host pointer. */
int b200_map_path_distance_metrics(b200_map_t* map, const double* path_xyz, int64_t n, double out_mean_std_min_max[4],
                                   double* distances);
/* The host half of it on its own (no GPU): {mean, std_dev, min, max} of n distances exactly as
 * calculateDistancesMetrics (src/utils/metrics.cpp:62-97) computes them. */
int b200_distance_metrics(const double* distances, int64_t n, double out_mean_std_min_max[4]);
/* Host-only (no device work): the pose algebra of GridMap::depthOdomCallback (grid_map.cpp:965-982).  body odometry
 * (position, quaternion w x y z) -> camera position, md_.camera_q_ (w x y z) and the row-major rotation projectDepthImage
 * derives from it (:208) — body quaternion -> matrix, body2world * cam2body_ (:91), matrix -> quaternion -> matrix, in Eigen's
 * published operation order.  Feed cam_pos / cam_rot to b200_map_update_depth. */
int b200_odom_to_camera_pose(const double body_pos[3], const double body_q_wxyz[4], double cam_pos[3], double cam_q_wxyz[4],
                             double cam_rot[9]);
int b200_map_download_costq(b200_map_t* map, uint16_t* dst);
int b200_map_download_cost(b200_map_t* map, float* dst);

/* LineOfSight::fastLOS (src/utils/LineOfSight.cpp:9-27) for n independent segments:
 * visible[i] = 1 iff no 0.05 m sample of start→end is occupied or outside the map.
 * n_samples (optional, may be NULL) = samples the reference loop evaluates. */
int b200_los_batch(b200_map_t* map, const double* seg_start_xyz, const double* seg_end_xyz, int64_t n,
                   uint8_t* visible, int32_t* n_samples, int32_t on_device);

/* NEW (SURVEY Appendix D "3-D DDA LoS"; not what the reference's planners call): exact voxel-traversal line of sight.
 * visible[i] = 1 iff both end points are inside the map and no voxel visited by the reference's Amanatides-Woo
 * RayCaster (src/plan_env/raycast.cpp:253-346; start and end voxel included) between them is occupied in the inflated
 * grid.  The walk runs on map-relative voxel coordinates (pos - origin) * res_inv.  Its booleans differ from fastLOS on
 * grazing rays (a sampled march skips corner voxels), so the planners keep fastLOS.  n_cells (optional) = voxels
 * examined. */
int b200_los_batch_dda(b200_map_t* map, const double* seg_start_xyz, const double* seg_end_xyz, int64_t n,
                       uint8_t* visible, int32_t* n_cells, int32_t on_device);

/* Per-stage device times (ms, CUDA events on the map's stream) of the last update_cloud / build_edt when
 * profiling is enabled: [0] clear, [1] scatter; seven-kernel EDT: [2] edt z, [3] scan y, [4] band crossovers y,
 * [5] fill y, [6] scan x, [7] band crossovers x, [8] fill x (+cost); fused EDT: [9] pass z+y, [10] pass x (+cost);
 * [11..15] reserved. */
int b200_map_set_profiling(b200_map_t* map, int32_t enable);
/* Kernel launches this map has issued since it was created (for bench accounting; memsets and copies not counted). */
int b200_map_launch_count(const b200_map_t* map, int64_t* n_launches);
int b200_map_get_stage_ms(b200_map_t* map, double out_ms[16]);

/* ---------------------------------------------------------------------------------------------------- */
/* Planners                                                                                             */
/* ---------------------------------------------------------------------------------------------------- */

/* ROS params of AStar::setParam / ThetaStar::setParam (src/Planners/AStar.cpp:31-41, ThetaStar.cpp:25-35)
 * + AlgorithmBase::setCostFactor (AlgorithmBase.hpp:39) + GetPath.srv:26 max_los. */
/* Open-list semantics (DESIGN.md "Search semantics").
 *  FAITHFUL : bit-identical to the reference's own code — the open list is the red-black tree std::set builds,
 *             with keys mutated in place, A*'s unconditional overwrite and Theta*'s erase-by-mutated-key
 *             (AStar.cpp:56-78, ThetaStar.cpp:72-89, utils.hpp:52-61).  astar / thetastar only.
 *  CANONICAL: exact priority queue on (f, creation index), true decrease-key, guarded relaxations — the clean
 *             semantics of SURVEY Appendix B; several times faster; the only one for the extension algorithms.
 *  DEFAULT  : FAITHFUL for astar / thetastar / thetastaragr, CANONICAL otherwise.  thetastaragrpp relaxes every node
 *             exactly once (ThetaStarAGRpp.cpp:208-210), so its keys never change inside the set and the two semantics
 *             coincide: it always runs the (faster) canonical open list and is still identical to the reference. */
enum { B200_SEMANTICS_DEFAULT = 0, B200_SEMANTICS_CANONICAL = 1, B200_SEMANTICS_FAITHFUL = 2 };

typedef struct {
  double resolution;        /* <algo>/resolution */
  double lambda_heuristic;  /* <algo>/lambda_heuristic */
  int32_t allocate_num;     /* <algo>/allocate_num: node budget per query */
  int32_t semantics;        /* B200_SEMANTICS_* */
  double cost_weight;       /* cost-aware variants only */
  double max_los;           /* <= 0: unlimited */
  /* thetastaragr/{barrier, flying_cost, flying_cost_default, ground_judge, epsilon} — read by BOTH air-ground
   * variants (src/Planners/ThetaStarAGR.cpp:33-36, ThetaStarAGRpp.cpp:33-37; launch/planner.launch:71-77) */
  double agr_barrier, agr_flying_cost, agr_flying_cost_default, agr_ground_judge, agr_epsilon;
} b200_planner_params_t;

/* AlgorithmBase::createResultDataObject (src/Planners/AlgorithmBase.cpp:123-175) as a plain struct. */
typedef struct {
  int32_t solved;           /* "solved" */
  int32_t status;           /* 0 solved, 1 open set empty (AStar.cpp:175-178), 2 pool exhausted (:165-168),
                               4 open-list workspace exhausted (faithful mode only; never seen in practice) */
  int32_t n_path;           /* path.size() */
  int32_t open_peak;        /* largest open-list size at a pop (extra counter) */
  int64_t path_offset;      /* first point of this path inside the caller's path buffer (in points) */
  double g_final;           /* "g_final_node" */
  double path_length;       /* "path_length" (float accumulation, geometry_utils.cpp:10-20) */
  int64_t explored_nodes;   /* "explored_nodes" = use_node_num_ */
  int64_t iter_num;         /* iter_num_ (expansions) */
  int64_t los_checks;       /* "line_of_sight_checks" */
  int64_t los_samples;      /* fastLOS samples evaluated (extra counter) */
  double goal_coords[3];    /* "goal_coords" = position of the LAST node (AlgorithmBase.cpp:130) */
  double start_coords[3];   /* "start_coords" */
  double time_us;           /* "time_spent": device time of this query (the reference reports 0) */
} b200_path_result_t;

/* algorithm: "astar", "thetastar", "thetastaragr", "thetastaragrpp" (the four the reference dispatches,
 * src/planner_ros_node.cpp:261-285), and the north-star extensions "lazythetastar", "costastar", "costlazythetastar".  Unknown names fall back to "astar" exactly
 * like planner_ros_node.cpp:280-284.  ≡ new AStar()/ThetaStar() + setParam() + init(). */
int b200_planner_create(const char* algorithm, const b200_planner_params_t* params, b200_planner_t** out);
int b200_planner_destroy(b200_planner_t* planner);
/* AlgorithmBase::setGridMap (AlgorithmBase.cpp:10-13) */
int b200_planner_set_map(b200_planner_t* planner, b200_map_t* map);
/* NEW (no reference counterpart: the reference plans one query at a time on the caller's thread).  Give this planner its
 * own cudaStream_t (void*; NULL = back to the map's stream).  Planners bound to the same map but to different streams may be
 * driven from different host threads, and their batches overlap on the device: the straggling queries of one batch run
 * beside the first queries of the next instead of leaving the SMs idle.  The caller orders the stream after the map's
 * last write (b200_map_sync, or an event on the map's stream). */
int b200_planner_set_stream(b200_planner_t* planner, void* cuda_stream);
/* NEW: a non-blocking CUDA stream on `device` for callers that do not link the CUDA runtime (the reference's node does not);
 * pass it to b200_planner_set_stream / b200_map_set_stream, destroy it after the handles that use it. */
int b200_stream_create(int32_t device, void** cuda_stream_out);
int b200_stream_destroy(void* cuda_stream);
/* AlgorithmBase::setCostFactor (AlgorithmBase.hpp:39) */
int b200_planner_set_cost_factor(b200_planner_t* planner, double cost_weight);
/* AlgorithmBase::detectCollision (AlgorithmBase.cpp:38-44): *collides = bool(getInflateOccupancy(pos)). */
int b200_planner_detect_collision(b200_planner_t* planner, const double pos[3], int32_t* collides);

/* AlgorithmBase::reset + findPath(src, tgt) + getPath (AStar.cpp:80-179; AlgorithmBase.cpp:17-32,177-183).
 * path_xyz receives up to path_cap points (3 doubles each); result->n_path is the full length. */
int b200_planner_find_path(b200_planner_t* planner, const double start[3], const double goal[3],
                           b200_path_result_t* result, double* path_xyz, int64_t path_cap);
/* nq independent queries (one reset+findPath each, one query per thread block).  Paths are packed into
 * path_xyz (capacity path_cap points) at results[i].path_offset; a path that does not fit keeps n_path but
 * gets path_offset = -1.  Pointers are host pointers unless on_device != 0. */
int b200_planner_find_paths(b200_planner_t* planner, const double* starts_xyz, const double* goals_xyz, int64_t nq,
                            b200_path_result_t* results, double* path_xyz, int64_t path_cap, int32_t on_device);
/* AlgorithmBase::getVisitedNodes (AlgorithmBase.cpp:185-189) for the last b200_planner_find_path: the positions of the
 * nodes created by that query, in creation order, all but the last one (the reference returns
 * path_node_pool_[0 .. use_node_num_ - 1)).  *n_out = their number (may exceed cap; only cap are written).
 * B200_ERR_INVALID after a batched call. */
int b200_planner_get_visited_nodes(b200_planner_t* planner, double* xyz, int64_t cap, int64_t* n_out);
/* Kernel launches issued by the last find_paths call (for bench accounting). */
int b200_planner_last_launches(b200_planner_t* planner, int32_t* n_launches);

/* ---------------------------------------------------------------------------------------------------- */
const char* b200_last_error(void);
const char* b200_version(void);
/* Pinned host memory helpers for callers that want the fast H2D/D2H path. */
int b200_host_alloc(void** ptr, int64_t bytes);
int b200_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif /* B200_PLANNER_H */
