// kernels.h — launchers shared between the .cu translation units and the C-ABI layer (api.cu).
#pragma once
#include "common.cuh"
#include "../../include/b200_planner.h"

namespace b200 {

// ---- map_kernels.cu
void launch_clear_box(const MapDev& m, const int a[3], const int b[3], int sm_count, cudaStream_t st);
int launch_scatter(const MapDev& m, const uint8_t* pts, long long n, int step, int offx, int offy, int offz, const double cam[3],
                   const double range[3], int s, long long* bounds, cudaStream_t st);
void launch_unpack(const MapDev& m, uint8_t* dst, int sm_count, cudaStream_t st);
void launch_pack(const MapDev& m, const uint8_t* src, int sm_count, cudaStream_t st);
void launch_occupancy(const MapDev& m, const double* pos, long long n, int8_t* out, cudaStream_t st);
void launch_in_map(const MapDev& m, const double* pos, long long n, uint8_t* out, cudaStream_t st);
void launch_pos_to_index(const MapDev& m, const double* pos, long long n, int32_t* out, cudaStream_t st);
void launch_set_occupied(const MapDev& m, const double* pos, long long n, cudaStream_t st);
void launch_set_occupancy(const MapDev& m, double* logodds, const double* pos, const double* occ, long long n, uint32_t* winner,
                          cudaStream_t st);
void launch_distance_at(const MapDev& m, const double* pos, long long n, double* out, cudaStream_t st);
void launch_los_batch(const MapDev& m, const double* s, const double* e, long long n, uint8_t* vis, int32_t* nsamples, int sm_count,
                      cudaStream_t st);

size_t export_scan_temp_bytes(long long cols);
void launch_export(const MapDev& m, const int a[3], const int b[3], double truncate, int* counts, int* offsets, void* temp,
                   size_t temp_bytes, float4* out, long long cap, cudaStream_t st);

// ---- edt_kernels.cu
int edt_lut_size(double res, double d_safe);
// ev (optional): 8 events recorded before z, scan_y, cross_y, fill_y, scan_x, cross_x, fill_x and after fill_x
bool edt_dz_is_wide(const MapDev& m, bool halo);
void launch_edt(const MapDev& m, void* dzbuf, int32_t* tmp, uint2* st, int* counts, int4* cross, int32_t* dist2, uint16_t* costq,
                uint16_t* lut, int lut_n, double d_safe, const int32_t* below_gap, const int32_t* above_gap, cudaStream_t stream,
                cudaEvent_t* ev);
void launch_cost_lut(uint16_t* lut, int lut_n, double res, double d_safe, cudaStream_t stream);
// ---- edt_fused.cu: two shared-memory-staged kernels; false when no configuration fits the grid
struct EdtOverlap {  // optional: two auxiliary streams + events so the z range can be processed as two overlapping halves
  cudaStream_t s[2];
  cudaEvent_t fork, y_done[2], x_done[2];
  int parts_used;  // set by launch_edt_fused: z parts the last build was cut into (launches = 2 * parts_used)
};
size_t edt_fused_plane_flags_bytes(const MapDev& m);
bool launch_edt_fused(const MapDev& m, void* tmp, uint8_t* plane_empty, int32_t* dist2, uint16_t* costq, const uint16_t* lut, int lut_n,
                      const int32_t* below_gap, const int32_t* above_gap, int dz_bound, cudaStream_t stream, cudaEvent_t* ev, EdtOverlap* ov);
void launch_column_extents(const MapDev& m, int32_t* lowest, int32_t* highest, cudaStream_t stream);
constexpr int kMaxSlabs = 64;
void launch_column_extents16(const MapDev& m, short2* lo_hi, cudaStream_t stream);
void launch_slab_gaps(const short2* gathered, const int* slab_nz, int P, int rank, long long cols, int32_t* below, int32_t* above,
                      cudaStream_t stream);

// ---- search_kernels.cu
enum Algo { ALGO_ASTAR = 0, ALGO_THETASTAR = 1, ALGO_LAZYTHETASTAR = 2, ALGO_COSTASTAR = 3, ALGO_COSTLAZYTHETASTAR = 4,
            ALGO_THETASTARAGR = 5, ALGO_THETASTARAGRPP = 6 };
enum { ST_SOLVED = 0, ST_OPEN_EMPTY = 1, ST_POOL_EXHAUSTED = 2, ST_TIER_OVERFLOW = 3, ST_OPEN_LIST_FULL = 4 };

// Per-slot search workspace in HBM (one slot per resident thread block; SoA of slots, AoS inside):
//   nodes  NodeRec [cap]      64 B records, index = creation order (the reference's pool index / tie-break)
//   hash   uint32  [hcap]     open addressing on the exact 3xFP64 key -> node index + 1
//   heap   HeapEnt [2cap+32]  canonical: 32-ary min-heap on (f, creation index), children of k are one aligned 512 B row;
//                             faithful: the red-black tree nodes of rbtree.cuh (16 B each, duplicates possible)
struct __align__(16) NodeRec {
  // first 32-byte sector: what the neighbour probe reads (key compare, closed test, open-list position)
  double px, py, pz;
  int state;      // low byte: NODE_*; next byte: via direction code (0..26); bit 16: AGR motion_state (flying)
  int hpos;       // heap position, -1 when not in the open list
  // second sector: what UpdateVertex reads and writes
  double g, f;
  int parent;     // node index or -1
  uint32_t slot;  // hash slot (for O(used) clearing)
  union {
    struct { int via; int pad; };  // lazy variants: expanded node that last relaxed this one
    double penalty;                // thetastaragr: penalty_g_score carried in g (utils.hpp:28)
  };
};
static_assert(sizeof(NodeRec) == 64, "NodeRec must be one 64-byte record");
struct __align__(16) HeapEnt {
  unsigned long long key;  // order-preserving image of f
  uint32_t idx;
  uint32_t pad;
};

struct SearchParams {
  MapDev map;
  double r, lambda, tie, cost_weight, max_los;
  double agr_barrier, agr_flying_cost, agr_flying_cost_default, agr_ground_judge, agr_epsilon;  // thetastaragr/*
  double dv[3];
  int allocate_num;
  int algo;
  int faithful;  // 1: reference-identical open list (astar / thetastar only)
  const double* starts;
  const double* goals;
  const int* qlist;  // optional indirection (second-tier pass)
  long long nq;
  b200_path_result_t* results;
  double* path_pool;
  long long path_cap;
  unsigned long long* path_cursor;
  unsigned long long* q_counter;
  int* overflow_list;
  int* overflow_count;
  NodeRec* nodes;
  uint32_t* hash;
  HeapEnt* heap;
  double* fkey;   // faithful mode: compact copy of the nodes' f (the tree comparator's only input), cap per slot
  int cap;        // physical nodes per slot
  uint32_t open_entries;  // 16-byte open-list entries per slot (heap / red-black tree nodes)
  uint32_t hcap;  // hash entries per slot (power of two >= 2*cap)
  // promotion slots (canonical 32-thread kernels): full-size workspaces a query moves into when it outgrows `cap`
  NodeRec* promo_nodes;
  uint32_t* promo_hash;
  HeapEnt* promo_heap;
  int promo_cap, promo_slots;
  uint32_t promo_open_entries, promo_hcap;
  int* promo_count;  // slots handed out by this launch
};

// Open-list region of a slot, in 16-byte entries.  The canonical heap needs cap + 32.  The faithful red-black
// tree holds DUPLICATES of a node (AStar.cpp:71-72 inserts without erasing): measured peak |open| / nodes is
// <= 1.0 for thetastar and up to 4.4 for astar at lambda = 1 (oracle, C3 map), hence 8x for faithful astar.
inline int open_region_factor(int algo, int faithful) { return (faithful && algo == 0) ? 8 : 2; }
inline size_t open_region_entries(int cap, int factor) { return (size_t)factor * (size_t)cap + 32; }
// ---- depth-image fusion (depth_kernels.cu) ----
struct DepthCtl {                 // device-side counters of one frame
  unsigned long long bbox[6];     // order-preserving keys: min x,y,z then max x,y,z of the ray end points
  unsigned long long n_cast, n_steps, n_updated, n_path_total;
  unsigned n_valid, n_touched, n_walked, rounds, n_cast_rays;
  unsigned changed[3];
};
struct DepthDev {
  // grid_map.h:60-78 (MappingParameters), logits precomputed on the host with libm like the reference
  double fx, fy, cx, cy, scale, inv_scale, mindist, maxdist, max_ray_length;
  double prob_hit_log, prob_miss_log, clamp_min_log, clamp_max_log, min_occ_log, unknown;
  int use_filter, margin, skip;
  // this frame
  int rows, cols, cols_s, n_rays, max_steps;
  int8_t round;                   // raycast_num_ of this frame (`char`: wraps 127 -> -128)
  double cam[3], R[9];            // R row-major, camera -> world
  double minb[3], maxb[3];        // map_min_boundary_ / map_max_boundary_
  int loc_min[3], loc_max[3];     // index box of camera +- local_update_range
  // state in HBM (per voxel unless noted)
  double* occ;                    // occupancy_buffer_ (log-odds)
  uint32_t *cnt_all, *cnt_hit;    // count_hit_and_miss_, count_hit_ (exact totals; the reference's shorts are their low 16 bits)
  uint32_t *F, *Fnew, *E;         // this frame: E = smallest ray id ending in the voxel; F / Fnew = (round stamp | smallest ray
                                  // id reaching the voxel), written alternately by the ordering rounds (0xffffffff = none)
  int8_t *flagT, *flagE;          // flag_traverse_, flag_rayend_: raycast_num_ of the last frame that set them (never cleared)
  uint32_t *touched, *walked;     // voxel lists: counted this frame / reached in the first round
  double* ray_pt;                 // [n_rays][3] clamped ray end points
  uint32_t* ray_addr;             // [n_rays] end voxel address
  uint32_t *cast_id, *path_off, *path_len;  // [n_cast_rays] the rays that are cast: id, and their slice of `paths`
  uint32_t* stop;                 // [n_cast_rays] path voxels the ray visited in the last ordering round
  int id_bits;                    // bits of a ray id inside an F / Fnew entry (the rest is the round stamp)
  uint32_t* paths;                // voxel address sequences of the cast rays (0xffffffff = cell outside the buffer)
  DepthCtl* ctl;
};
void launch_depth_init(const DepthDev& D, double unknown, long long nvox, cudaStream_t st);
void launch_depth_project(const DepthDev& D, const MapDev& m, const uint16_t* img, cudaStream_t st);
void launch_depth_path_count(const DepthDev& D, const MapDev& m, cudaStream_t st);
void launch_depth_path_fill(const DepthDev& D, const MapDev& m, unsigned n_cast, cudaStream_t st);
int launch_depth_order(const DepthDev& D, unsigned n_cast, int sm_count, cudaStream_t st);
void launch_depth_count(const DepthDev& D, unsigned n_cast, cudaStream_t st);
void launch_depth_update(const DepthDev& D, const MapDev& m, int sm_count, cudaStream_t st);
void launch_depth_local_map(const DepthDev& D, const MapDev& m, const int lb_min[3], const int lb_max[3], const int cut_min[3],
                            const int cut_max[3], const int cutm_min[3], const int cutm_max[3], int inf_step, int ceil_on, int ceil_id,
                            int sm_count, cudaStream_t st);
void launch_exclusive_scan(void* temp, size_t temp_bytes, const int* in, int* out, long long n, cudaStream_t st);  // map_kernels.cu (CUB)
void launch_depth_publish(const DepthDev& D, const MapDev& m, int which, const int a[3], const int b[3], double truncate, int* counts,
                          int* offsets, void* temp, size_t temp_bytes, float4* out, long long cap, cudaStream_t st);
void launch_los_dda(const MapDev& m, const double* s, const double* e, long long n, uint8_t* vis, int32_t* ncells, cudaStream_t st);
void launch_depth_get_log_odds(const MapDev& m, const double* occ, const double* pos, long long n, double* out, cudaStream_t st);
void launch_depth_get_occupancy(const MapDev& m, const double* occ, double min_occ_log, const double* pos, long long n, int8_t* out,
                                cudaStream_t st);

size_t search_slot_bytes(int cap, int open_factor, uint32_t* hcap_out);
int search_threads_per_block(int algo);
int search_slots_per_sm(int algo, int faithful);
void launch_search(const SearchParams& p, int slots, cudaStream_t st);
size_t query_order_bytes(int nq);
const int* launch_query_order(const MapDev& m, const double* starts, const double* goals, int nq, void* buf, cudaStream_t st);

}  // namespace b200
