// depth_kernels.cu — depth-image fusion into the log-odds grid and the inflated occupancy grid.
//
// Behaviour restated from the reference (not ported): grid_map.cpp:195-315 projectDepthImage, :317-482 raycastProcess,
// :484-507 closetPointInMap, :509-627 clearAndInflateLocalMap, :173-193 setCacheOccupancy, raycast.cpp:32-49,253-346
// (RayCaster::setInput / step).  All position arithmetic is FP64 with separately rounded operations.
//
// The reference processes rays ONE AFTER THE OTHER and its result depends on that order in two places:
//   (1) only the first ray (in pixel order) that ends in a voxel is cast at all (flag_rayend_, :385-395);
//   (2) a ray stops at the first voxel an EARLIER ray of the same frame has crossed (flag_traverse_, :408-418).
// Both are reproduced exactly without serialising the rays.  Every ray carries its position in the reference's
// order as an id.  (1) is an atomicMin of ids per end voxel.  For (2) let F(v) be the smallest id of a ray that
// REACHES voxel v; ray i stops at the first voxel with F(v) < i.  F is the unique fixed point of
//   F(v) = min { i : v on ray i's path before i's first voxel w with F(w) < i }
// (induction over i: ray 0 is never stopped, and what ray i reaches depends only on rays < i), so iterating
// F_{k+1} = T(F_k) from F_0 = +inf until it stops changing yields exactly the sequential result; the iterates bracket
// it (F_1 <= F_3 <= ... <= F <= ... <= F_4 <= F_2).  Rays to one camera centre almost form a tree, so a few rounds
// suffice (the count is reported in the stats).  Hit/miss counters are then accumulated with plain atomics by one
// last walk under the converged F, and the per-voxel log-odds update is order-free.
//
// Reference quirks kept on purpose:
//   * the zero-depth test of the filtered projection reads the NEXT sampled pixel (:260);
//   * the two flag arrays are never cleared: they hold the `char` frame counter raycast_num_ (grid_map.h:123-124) of
//     the last frame that set them, and the counter wraps 127 -> -128.  A voxel last flagged exactly 256 fused frames
//     ago — or never, in frame 255, whose number -1 is the arrays' initial value — therefore looks already taken at
//     the start of the frame: no ray is cast to it, every ray stops at it.  Kept as persistent int8 arrays here;
//   * the counters are `short` and a voxel is queued again each time its counter wraps round to 1;
//   * inflation range-tests only the LINEAR address (:608), so a cube sticking out in y or z wraps around.
#include <cooperative_groups.h>
#include <cstdint>

#include <cstdlib>
#include "kernels.h"

namespace cg = cooperative_groups;

namespace b200 {

namespace {
constexpr uint32_t kNone = 0xffffffffu;      // F / E: no ray; ray_addr: pixel produced no ray
constexpr uint32_t kOutside = 0xfffffffeu;   // ray_addr: end point outside the buffer (never deduplicated)

__device__ __forceinline__ unsigned long long dkey(double v) {
  unsigned long long u = (unsigned long long)__double_as_longlong(v);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
// grid_map.h:400-404 + :257-265 on a world position; -1 when the linear address falls outside the buffer
__device__ __forceinline__ long long pos_address(const MapDev& m, double x, double y, double z) {
  const long long ix = pos_to_idx1(x, m.ox, m.inv_res), iy = pos_to_idx1(y, m.oy, m.inv_res), iz = pos_to_idx1(z, m.oz, m.inv_res);
  const long long a = ix * m.Ny * m.Nz + iy * m.Nz + iz;
  const long long total = (long long)m.Nx * m.Ny * m.Nz;
  return (a < 0 || a >= total) ? -1 : a;
}
__device__ __forceinline__ void append(uint32_t* list, unsigned* n, uint32_t a) { list[atomicAdd(n, 1u)] = a; }

// raycast.cpp:32-49
__device__ __forceinline__ double ray_mod1(double v) { return fmod(dadd(fmod(v, 1.0), 1.0), 1.0); }
__device__ __forceinline__ double ray_intbound(double s, double ds) {
  if (ds < 0) { s = -s; ds = -ds; }
  return __ddiv_rn(dsub(1.0, ray_mod1(s)), ds);
}
struct RayWalk {  // raycast.cpp:253-346
  int x, y, z, ex, ey, ez, sx, sy, sz;
  double tmx, tmy, tmz, tdx, tdy, tdz;
  __device__ __forceinline__ void set(double ax, double ay, double az, double bx, double by, double bz) {
    x = (int)floor(ax); y = (int)floor(ay); z = (int)floor(az);
    ex = (int)floor(bx); ey = (int)floor(by); ez = (int)floor(bz);
    const double dx = ex - x, dy = ey - y, dz = ez - z;
    sx = (dx > 0) - (dx < 0); sy = (dy > 0) - (dy < 0); sz = (dz > 0) - (dz < 0);
    tmx = ray_intbound(ax, dx); tmy = ray_intbound(ay, dy); tmz = ray_intbound(az, dz);
    tdx = __ddiv_rn((double)sx, dx); tdy = __ddiv_rn((double)sy, dy); tdz = __ddiv_rn((double)sz, dz);
  }
  __device__ __forceinline__ bool step(int& ox, int& oy, int& oz) {
    ox = x; oy = y; oz = z;
    if (x == ex && y == ey && z == ez) return false;
    if (tmx < tmy) {
      if (tmx < tmz) { x += sx; tmx = dadd(tmx, tdx); } else { z += sz; tmz = dadd(tmz, tdz); }
    } else {
      if (tmy < tmz) { y += sy; tmy = dadd(tmy, tdy); } else { z += sz; tmz = dadd(tmz, tdz); }
    }
    return true;
  }
};
}  // namespace

// --------------------------------------------------------------------------------------------------
// projection + ray end points (projectDepthImage, first half of raycastProcess' per-point body)
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_project_kernel(DepthDev D, MapDev m, const uint16_t* __restrict__ img) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  double px = 0, py = 0, pz = 0;
  bool valid = id < D.n_rays;
  if (valid) {
    int u, v;
    double depth;
    if (D.use_filter) {
      v = D.margin + (id / D.cols_s) * D.skip;
      u = D.margin + (id % D.cols_s) * D.skip;
      const size_t at = (size_t)v * D.cols + u;
      depth = dmul((double)img[at], D.inv_scale);
      if (img[at + D.skip] == 0) depth = dadd(D.max_ray_length, 0.1);      // :260 (the pointer was advanced first)
      else if (depth < D.mindist) valid = false;                          // :264
      else if (depth > D.maxdist) depth = dadd(D.max_ray_length, 0.1);    // :268
    } else {
      v = id / D.cols; u = id % D.cols;
      depth = __ddiv_rn((double)img[id], D.scale);                         // :221
    }
    if (valid) {
      const double qx = __ddiv_rn(dmul(dsub((double)u, D.cx), depth), D.fx);
      const double qy = __ddiv_rn(dmul(dsub((double)v, D.cy), depth), D.fy);
      const double qz = depth;
      px = dadd(dadd(dadd(dmul(D.R[0], qx), dmul(D.R[1], qy)), dmul(D.R[2], qz)), D.cam[0]);
      py = dadd(dadd(dadd(dmul(D.R[3], qx), dmul(D.R[4], qy)), dmul(D.R[5], qz)), D.cam[1]);
      pz = dadd(dadd(dadd(dmul(D.R[6], qx), dmul(D.R[7], qy)), dmul(D.R[8], qz)), D.cam[2]);
    }
  }
  int occ = 0;
  if (valid) {
    const double cx = D.cam[0], cy = D.cam[1], cz = D.cam[2];
    bool inside = is_in_map(m, px, py, pz);
    if (!inside) {  // closetPointInMap, :484-507
      const double diff[3] = {dsub(px, cx), dsub(py, cy), dsub(pz, cz)};
      double min_t = 1000000;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (fabs(diff[i]) > 0) {
          const double t1 = __ddiv_rn(dsub(D.maxb[i], D.cam[i]), diff[i]);
          if (t1 > 0 && t1 < min_t) min_t = t1;
          const double t2 = __ddiv_rn(dsub(D.minb[i], D.cam[i]), diff[i]);
          if (t2 > 0 && t2 < min_t) min_t = t2;
        }
      const double k = dsub(min_t, 1e-3);
      px = dadd(cx, dmul(k, diff[0])); py = dadd(cy, dmul(k, diff[1])); pz = dadd(cz, dmul(k, diff[2]));
    }
    const double dx = dsub(px, cx), dy = dsub(py, cy), dz = dsub(pz, cz);
    const double len = norm3(dx, dy, dz);
    if (len > D.max_ray_length) {  // :353-357 / :364-367
      px = dadd(dmul(__ddiv_rn(dx, len), D.max_ray_length), cx);
      py = dadd(dmul(__ddiv_rn(dy, len), D.max_ray_length), cy);
      pz = dadd(dmul(__ddiv_rn(dz, len), D.max_ray_length), cz);
    } else if (inside) {
      occ = 1;  // :371
    }
  }
  // bounding box of the (clamped) end points, :375-381
  unsigned long long k[6];
  k[0] = valid ? dkey(px) : ~0ULL; k[1] = valid ? dkey(py) : ~0ULL; k[2] = valid ? dkey(pz) : ~0ULL;
  k[3] = valid ? dkey(px) : 0ULL;  k[4] = valid ? dkey(py) : 0ULL;  k[5] = valid ? dkey(pz) : 0ULL;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      k[i] = min(k[i], __shfl_xor_sync(0xffffffffu, k[i], o));
      k[i + 3] = max(k[i + 3], __shfl_xor_sync(0xffffffffu, k[i + 3], o));
    }
  }
  const unsigned any = __ballot_sync(0xffffffffu, valid);
  if ((threadIdx.x & 31) == 0 && any) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { atomicMin(&D.ctl->bbox[i], k[i]); atomicMax(&D.ctl->bbox[i + 3], k[i + 3]); }
    atomicAdd(&D.ctl->n_valid, (unsigned)__popc(any));
  }
  if (id >= D.n_rays) return;
  uint32_t addr = kNone;
  if (valid) {
    const long long a = pos_address(m, px, py, pz);
    if (a < 0) {
      addr = kOutside;
    } else {  // setCacheOccupancy, :173-193
      addr = (uint32_t)a;
      if (atomicAdd(&D.cnt_all[a], 1u) == 0u) append(D.touched, &D.ctl->n_touched, addr);
      if (occ) atomicAdd(&D.cnt_hit[a], 1u);
      atomicMin(&D.E[a], (uint32_t)id);
    }
    D.ray_pt[3 * (size_t)id] = px; D.ray_pt[3 * (size_t)id + 1] = py; D.ray_pt[3 * (size_t)id + 2] = pz;
  }
  D.ray_addr[id] = addr;
}

// --------------------------------------------------------------------------------------------------
// ray paths.  The voxel sequence of a ray (RayCaster, raycast.cpp:253-346) depends only on its two end points, so it
// is traced ONCE per frame into a compact list; every round of the ordering iteration and the counting pass then read
// addresses instead of re-running the FP64 traversal (measured: 35-40 rounds per 640x480 frame, and a round of
// dependent DDA steps costs ~50 us against ~5 us for a warp reading 32 cached voxels at a time).
// --------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ray_is_cast(const DepthDev& D, int id) {
  const uint32_t ea = D.ray_addr[id];
  if (ea == kNone) return false;
  // :385-395: the end voxel was already taken — by an earlier ray of this frame, or its flag still holds this round's number
  return ea == kOutside || (D.flagE[ea] != D.round && D.E[ea] == (uint32_t)id);
}
__device__ __forceinline__ void ray_setup(const DepthDev& D, const MapDev& m, int id, RayWalk& w) {
  const double px = D.ray_pt[3 * (size_t)id], py = D.ray_pt[3 * (size_t)id + 1], pz = D.ray_pt[3 * (size_t)id + 2];
  w.set(__ddiv_rn(px, m.res), __ddiv_rn(py, m.res), __ddiv_rn(pz, m.res), __ddiv_rn(D.cam[0], m.res), __ddiv_rn(D.cam[1], m.res),
        __ddiv_rn(D.cam[2], m.res));  // :397
}
// pass 1: which rays are cast, and how long their paths are
__global__ void __launch_bounds__(128) depth_path_count_kernel(DepthDev D, MapDev m) {
  const int id = blockIdx.x * blockDim.x + threadIdx.x;
  if (id >= D.n_rays || !ray_is_cast(D, id)) return;
  RayWalk w;
  ray_setup(D, m, id, w);
  int ix, iy, iz;
  unsigned steps = 0;
  // Lattice cells map to voxels monotonically along every axis, so a ray can only meet one of its OWN voxels again as
  // an immediate repeat (two consecutive cells in one voxel — possible when origin / resolution has fractional part
  // 1/2).  The reference then stops the ray at its own flag (:410): the cached path ends with that repeated voxel.
  long long prev = -1;
  while (steps < (unsigned)D.max_steps && w.step(ix, iy, iz)) {
    ++steps;
    const long long a = pos_address(m, dmul(dadd((double)ix, 0.5), m.res), dmul(dadd((double)iy, 0.5), m.res),
                                    dmul(dadd((double)iz, 0.5), m.res));
    if (a >= 0) {
      if (a == prev) break;
      prev = a;
    }
  }
  const unsigned slot = atomicAdd(&D.ctl->n_cast_rays, 1u);
  D.cast_id[slot] = (uint32_t)id;
  D.path_len[slot] = steps;
  D.path_off[slot] = atomicAdd(&D.ctl->n_path_total, (unsigned long long)steps);
}
// pass 2: the voxel addresses (kNone where the cell centre falls outside the buffer: such a cell is stepped over, :401-404)
__global__ void __launch_bounds__(128) depth_path_fill_kernel(DepthDev D, MapDev m) {
  const unsigned slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= D.ctl->n_cast_rays) return;
  RayWalk w;
  ray_setup(D, m, (int)D.cast_id[slot], w);
  uint32_t* out = D.paths + D.path_off[slot];
  const unsigned len = D.path_len[slot];
  int ix, iy, iz;
  for (unsigned k = 0; k < len && w.step(ix, iy, iz); ++k) {
    const long long a = pos_address(m, dmul(dadd((double)ix, 0.5), m.res), dmul(dadd((double)iy, 0.5), m.res),
                                    dmul(dadd((double)iz, 0.5), m.res));  // :401
    out[k] = a < 0 ? kNone : (uint32_t)a;
  }
}
// One cast ray handled by one warp, 32 path voxels per step.  MODE 0: a round of the ordering iteration; MODE 1: the
// counting walk under the converged state.  Returns the number of path voxels the ray visits (warp-uniform).
//
// The two per-voxel arrays F / Fnew alternate as "written this round" / "written last round" WITHOUT ever being cleared
// inside a frame: an entry is (stamp << id_bits) | ray id with stamp = kMaxStamp - round, so under atomicMin an entry of
// the current round always beats whatever an older round left in the same array, and a reader only accepts entries that
// carry the previous round's stamp.  0xffffffff (the between-frames value) carries a stamp no round ever has.
struct StampCode {
  int idb;
  uint32_t idmask, max_stamp;
  __device__ __forceinline__ explicit StampCode(int id_bits) : idb(id_bits), idmask((1u << id_bits) - 1u), max_stamp((1u << (32 - id_bits)) - 2u) {}
  __device__ __forceinline__ uint32_t key(unsigned round, uint32_t id) const { return ((max_stamp - round) << idb) | id; }
  // ray id stored for `round` in v, kNone when v is from another round (or empty)
  __device__ __forceinline__ uint32_t id_of(uint32_t v, unsigned round) const { return (v >> idb) == max_stamp - round ? (v & idmask) : kNone; }
};
constexpr int kOrderThreads = 1024;  // few, large blocks: the grid barrier is what a round costs, and its price grows with the block count
template <int MODE>
__device__ __forceinline__ unsigned walk_path(const DepthDev& D, unsigned slot, int lane, unsigned round, unsigned base0 = 0) {
  const StampCode sc(D.id_bits);
  const uint32_t id = D.cast_id[slot];
  const uint32_t* path = D.paths + D.path_off[slot];
  const unsigned len = D.path_len[slot];
  // MODE 0 reads what round - 1 wrote (nothing in round 0) and writes the other array; MODE 1 reads what `round` wrote
  const unsigned rd_round = MODE == 0 ? round - 1u : round;
  const uint32_t* rd = (rd_round & 1u) ? D.Fnew : D.F;
  uint32_t* wr = (round & 1u) ? D.Fnew : D.F;
  const bool have_prev = MODE == 1 || round > 0;
  unsigned steps = len;
  for (unsigned base = base0; base < len; base += 32) {
    const unsigned k = base + lane;
    const uint32_t a = k < len ? path[k] : kNone;
    // :410 flag_traverse_[v] == raycast_num_ at the moment ray `id` arrives
    bool taken = false;
    if (a != kNone) taken = D.flagT[a] == D.round || (have_prev && sc.id_of(rd[a], rd_round) < id);
    const unsigned tm = __ballot_sync(0xffffffffu, taken);
    const int first = tm ? __ffs(tm) - 1 : 32;
    if (MODE == 0) {
      if (a != kNone && lane < first) {
        // only round 0 needs the old value (first visit of the frame -> walked list); later rounds use the fire-and-forget form
        if (round == 0) { if (atomicMin(&wr[a], sc.key(round, id)) == kNone) append(D.walked, &D.ctl->n_walked, a); }
        else atomicMin(&wr[a], sc.key(round, id));
      }
    } else {
      // :406 the voxel the ray stops at is counted too (the count comes before the stop test)
      if (a != kNone && lane <= first && atomicAdd(&D.cnt_all[a], 1u) == 0u) append(D.touched, &D.ctl->n_touched, a);
    }
    if (tm) { steps = base + (unsigned)first + 1u; break; }
  }
  if (MODE == 1 && lane == 0) {
    atomicAdd(&D.ctl->n_cast, 1ull);
    atomicAdd(&D.ctl->n_steps, (unsigned long long)steps);
  }
  return steps;
}
__global__ void __launch_bounds__(128) depth_count_kernel(DepthDev D) {
  const unsigned slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot < D.ctl->n_cast_rays) walk_path<1>(D, slot, threadIdx.x & 31, D.ctl->rounds - 1u);
}
// The whole fixed-point iteration in one cooperative launch, ONE grid barrier per round: every cast ray walks under what
// the previous round wrote and writes this round's entries into the other array.  The state of a round is determined by
// where each ray stops, so "no ray stopped anywhere else than in the previous round" is the fixed point — no per-voxel
// comparison or copy pass (the first version had both, and two barriers per round: 12 us per round instead of 6).
__global__ void __launch_bounds__(kOrderThreads) depth_order_kernel(DepthDev D) {
  cg::grid_group grid = cg::this_grid();
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthreads = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31;
  const unsigned n_cast = D.ctl->n_cast_rays;
  const StampCode sc(D.id_bits);
  const unsigned round_limit = min(n_cast + 2u, sc.max_stamp - 1u);
  // Usual case (a 640x480 frame casts ~3 000 rays of ~8 voxels): every ray has a warp of its own and its path fits one
  // 32-voxel step, so everything that does not change between rounds — ray id, path voxels, the persistent
  // flag_traverse_ test — is read ONCE into registers and a round is one load, one ballot, one atomicMin and the barrier
  // (a round cost 12 us when the ray record, path and flags were re-read through L2 every time).
  const unsigned slot0 = tid >> 5;
  const bool one_ray_per_warp = n_cast <= (nthreads >> 5);
  // (a longer path keeps its first 32 voxels in registers too: rays almost always stop within them, and only a ray that
  // does not goes on through the cached path in global memory)
  uint32_t c_id = 0, c_len = 0, c_a = kNone;
  bool c_flag = false;
  if (one_ray_per_warp && slot0 < n_cast) {
    c_id = D.cast_id[slot0];
    c_len = D.path_len[slot0];
    c_a = (unsigned)lane < c_len ? D.paths[D.path_off[slot0] + lane] : kNone;
    c_flag = c_a != kNone && D.flagT[c_a] == D.round;
  }
  unsigned my_stop = 0xffffffffu;
  unsigned round = 0;
  for (;; ++round) {
    bool changed = false;
    if (one_ray_per_warp) {
      if (slot0 < n_cast) {
        unsigned steps;
        {
          const uint32_t* rd = ((round - 1u) & 1u) ? D.Fnew : D.F;
          uint32_t* wr = (round & 1u) ? D.Fnew : D.F;
          bool taken = c_flag;
          if (!taken && round > 0 && c_a != kNone) taken = sc.id_of(rd[c_a], round - 1u) < c_id;
          const unsigned tm = __ballot_sync(0xffffffffu, taken);
          const int first = tm ? __ffs(tm) - 1 : 32;
          if (c_a != kNone && lane < first) {
            if (round == 0) { if (atomicMin(&wr[c_a], sc.key(round, c_id)) == kNone) append(D.walked, &D.ctl->n_walked, c_a); }
            else atomicMin(&wr[c_a], sc.key(round, c_id));  // result unused: fire-and-forget reduction
          }
          steps = tm ? (unsigned)first + 1u : (c_len <= 32u ? c_len : walk_path<0>(D, slot0, lane, round, 32u));
        }
        changed = steps != my_stop;
        my_stop = steps;
      }
    } else {
      for (unsigned slot = slot0; slot < n_cast; slot += nthreads >> 5) {  // a slot is always handled by the same warp
        const unsigned steps = walk_path<0>(D, slot, lane, round);
        if (round == 0 || D.stop[slot] != steps) changed = true;
        if (lane == 0) D.stop[slot] = steps;
      }
    }
    if (changed && lane == 0) D.ctl->changed[round % 3] = 1u;
    if (tid == 0) D.ctl->changed[(round + 1) % 3] = 0u;  // three slots: the one read at the end of the previous round is left alone
    grid.sync();
    if (D.ctl->changed[round % 3] == 0u || round >= round_limit) break;
  }
  // the last round's array holds the converged state; 0xffffffff = no fixed point within the stamp range (the host fails the call)
  if (tid == 0) D.ctl->rounds = D.ctl->changed[round % 3] == 0u ? round + 1u : 0xffffffffu;
}
// end of frame: the persistent flags take this frame's number, the per-frame ray ids go back to "none"
__global__ void __launch_bounds__(256) depth_reset_kernel(DepthDev D) {
  const unsigned n = D.ctl->n_walked, stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  const StampCode sc(D.id_bits);
  const unsigned last = D.ctl->rounds - 1u;
  const uint32_t* fin = (last & 1u) ? D.Fnew : D.F;
  for (unsigned i = t0; i < n; i += stride) {
    const uint32_t a = D.walked[i];
    // :416: every voxel some ray reached (in the converged state = the last round's entries) keeps this frame's number
    if (sc.id_of(fin[a], last) != kNone) D.flagT[a] = D.round;
    D.F[a] = kNone; D.Fnew[a] = kNone;
  }
  for (unsigned i = t0; i < (unsigned)D.n_rays; i += stride) {
    const uint32_t a = D.ray_addr[i];
    if (a < kOutside) { D.flagE[a] = D.round; D.E[a] = kNone; }  // :393
  }
}

// --------------------------------------------------------------------------------------------------
// log-odds update of every voxel touched this frame (raycastProcess :448-481)
// --------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_update_kernel(DepthDev D, MapDev m) {
  const unsigned n = D.ctl->n_touched;
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
  const uint32_t a = D.touched[i];
  const uint32_t n_all = D.cnt_all[a], n_hit = D.cnt_hit[a];
  D.cnt_all[a] = 0u; D.cnt_hit[a] = 0u;
  // the reference's counters are `short`: they hold the totals modulo 2^16, and the voxel was queued once per wrap to 1
  const int all_s = (int)(short)(uint16_t)n_all, hit_s = (int)(short)(uint16_t)n_hit;
  const unsigned pushes = (n_all - 1u) / 65536u + 1u;
  const int nyz = m.Ny * m.Nz;
  const int ix = (int)(a / (uint32_t)nyz), iy = (int)((a % (uint32_t)nyz) / (uint32_t)m.Nz), iz = (int)(a % (uint32_t)m.Nz);
  const bool in_local = ix >= D.loc_min[0] && ix <= D.loc_max[0] && iy >= D.loc_min[1] && iy <= D.loc_max[1] &&
                        iz >= D.loc_min[2] && iz <= D.loc_max[2];
  double o = D.occ[a];
  for (unsigned p = 0; p < pushes; ++p) {
    const double upd = (p == 0 ? hit_s >= all_s - hit_s : true) ? D.prob_hit_log : D.prob_miss_log;
    if (upd >= 0 && o >= D.clamp_max_log) continue;
    if (upd <= 0 && o <= D.clamp_min_log) { o = D.clamp_min_log; continue; }
    if (!in_local) o = D.clamp_min_log;
    o = fmin(fmax(dadd(o, upd), D.clamp_min_log), D.clamp_max_log);
  }
  D.occ[a] = o;
  atomicAdd(&D.ctl->n_updated, (unsigned long long)pushes);
  }
}

// --------------------------------------------------------------------------------------------------
// clearAndInflateLocalMap (:509-627)
// --------------------------------------------------------------------------------------------------
struct Box { int lo[3], hi[3]; };
__device__ __forceinline__ bool box_cell(const Box& b, long long t, int& x, int& y, int& z) {
  const int nx = b.hi[0] - b.lo[0] + 1, ny = b.hi[1] - b.lo[1] + 1, nz = b.hi[2] - b.lo[2] + 1;
  if (nx <= 0 || ny <= 0 || nz <= 0 || t >= (long long)nx * ny * nz) return false;
  z = b.lo[2] + (int)(t % nz);
  y = b.lo[1] + (int)((t / nz) % ny);
  x = b.lo[0] + (int)(t / ((long long)nz * ny));
  return true;
}
// every voxel of box M outside box C becomes "unknown" (:529-579: three slab loops whose union is M \ C)
__global__ void __launch_bounds__(256) depth_shell_kernel(DepthDev D, MapDev m, Box M, Box C) {
  int x, y, z;
  if (!box_cell(M, blockIdx.x * (long long)blockDim.x + threadIdx.x, x, y, z)) return;
  const bool inside = x >= C.lo[0] && x <= C.hi[0] && y >= C.lo[1] && y <= C.hi[1] && z >= C.lo[2] && z <= C.hi[2];
  if (!inside) D.occ[vox_addr(m, x, y, z)] = D.unknown;
}
__device__ __forceinline__ void set_inflate_linear(const MapDev& m, long long a) {
  const long long total = (long long)m.Nx * m.Ny * m.Nz;
  if (a < 0 || a >= total) return;  // :608 — the only range test there is
  const int nyz = m.Ny * m.Nz;
  const int x = (int)(a / nyz), y = (int)((a % nyz) / m.Nz), z = (int)(a % m.Nz);
  atomicOr(m.occ + ((size_t)x * m.Ny + y) * m.wz + (z >> 5), 1u << (z & 31));
}
// occupied voxels of the local box stamp the full (2*step+1)^3 cube (:595-616, grid_map.h:412-441)
__global__ void __launch_bounds__(256) depth_inflate_kernel(DepthDev D, MapDev m, Box L, int step) {
  int x, y, z;
  if (!box_cell(L, blockIdx.x * (long long)blockDim.x + threadIdx.x, x, y, z)) return;
  if (!(D.occ[vox_addr(m, x, y, z)] > D.min_occ_log)) return;
  for (int dx = -step; dx <= step; ++dx)
    for (int dy = -step; dy <= step; ++dy)
      for (int dz = -step; dz <= step; ++dz)
        set_inflate_linear(m, (long long)(x + dx) * m.Ny * m.Nz + (long long)(y + dy) * m.Nz + (z + dz));
}
// virtual ceiling (:618-627)
__global__ void __launch_bounds__(256) depth_ceiling_kernel(MapDev m, Box L, int ceil_id) {
  const int nx = L.hi[0] - L.lo[0] + 1, ny = L.hi[1] - L.lo[1] + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (nx <= 0 || ny <= 0 || t >= (long long)nx * ny) return;
  const int x = L.lo[0] + (int)(t / ny), y = L.lo[1] + (int)(t % ny);
  set_inflate_linear(m, (long long)x * m.Ny * m.Nz + (long long)y * m.Nz + ceil_id);
}

// publishMap (grid_map.cpp:806-849, WHICH 0: occupancy_buffer_ >= logit(p_occ)) and publishUnknown (:902-939, WHICH 1:
// occupancy_buffer_ < clamp_min - 1e-3): voxel centres of the box, x -> y -> z order, not above the truncate height.
// Thread per (x,y) column; `out == nullptr` counts, otherwise writes from the column's scanned offset.
template <int WHICH>
__device__ __forceinline__ bool publish_keep(const DepthDev& D, double o) {
  return WHICH == 0 ? !(o < D.min_occ_log) : (o < dsub(D.clamp_min_log, 1e-3));
}
template <int WHICH>
__global__ void __launch_bounds__(256) depth_publish_kernel(DepthDev D, MapDev m, Box B, double truncate, int* __restrict__ counts,
                                                            const int* __restrict__ offsets, float4* __restrict__ out, long long cap) {
  const int ny = B.hi[1] - B.lo[1] + 1;
  const long long cols = (long long)(B.hi[0] - B.lo[0] + 1) * ny;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cols) return;
  const int x = B.lo[0] + (int)(i / ny), y = B.lo[1] + (int)(i % ny);
  const double* col = D.occ + vox_addr(m, x, y, 0);
  const float px = __double2float_rn(dadd(dmul(dadd((double)x, 0.5), m.res), m.ox));  // indexToPos, grid_map.h:406-410
  const float py = __double2float_rn(dadd(dmul(dadd((double)y, 0.5), m.res), m.oy));
  long long o = out ? offsets[i] : 0;
  int n = 0;
  for (int z = B.lo[2]; z <= B.hi[2]; ++z) {
    if (!publish_keep<WHICH>(D, col[z])) continue;
    const double pz = dadd(dmul(dadd((double)z, 0.5), m.res), m.oz);
    if (pz > truncate) continue;
    if (out) {
      if (o < cap) out[o] = make_float4(px, py, __double2float_rn(pz), 1.0f);  // pcl::PointXYZ: data[3] = 1
      ++o;
    }
    ++n;
  }
  if (!out) counts[i] = n;
}

// ---- 3-D DDA line of sight (ours — the reference's planners use the sampled fastLOS) ----------
// visible iff both end points are inside the map and no voxel the reference's RayCaster (raycast.cpp:253-346) visits
// between them — start and end voxel included — is occupied in the inflated grid.  The traversal runs on
// map-relative voxel coordinates (pos - origin) * res_inv.  It is a serial FP64 recurrence, so one THREAD walks one
// segment (a warp takes 32 segments); splitting a segment over lanes would restart the recurrence at interior
// points and could visit different voxels on grazing rays.
__global__ void __launch_bounds__(128) los_dda_kernel(MapDev m, const double* __restrict__ s, const double* __restrict__ e, long long n,
                                                      int max_steps, uint8_t* __restrict__ vis, int32_t* __restrict__ ncells) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sx = s[3 * i], sy = s[3 * i + 1], sz = s[3 * i + 2], ex = e[3 * i], ey = e[3 * i + 1], ez = e[3 * i + 2];
  int cells = 0;
  bool ok = is_in_map(m, sx, sy, sz) && is_in_map(m, ex, ey, ez);
  if (ok) {
    RayWalk w;
    w.set(dmul(dsub(sx, m.ox), m.inv_res), dmul(dsub(sy, m.oy), m.inv_res), dmul(dsub(sz, m.oz), m.inv_res),
          dmul(dsub(ex, m.ox), m.inv_res), dmul(dsub(ey, m.oy), m.inv_res), dmul(dsub(ez, m.oz), m.inv_res));
    for (;;) {
      int ix, iy, iz;
      const bool more = w.step(ix, iy, iz);
      ++cells;
      if (ix < 0 || iy < 0 || iz < 0 || ix >= m.Nx || iy >= m.Ny || iz >= m.Nz || occ_bit(m, ix, iy, iz)) { ok = false; break; }
      if (!more) break;
      if (cells > max_steps) { ok = false; break; }  // a walk that misses its end cell counts as blocked
    }
  }
  vis[i] = ok ? 1 : 0;
  if (ncells) ncells[i] = cells;
}

// getOccupancy(Vector3d), grid_map.h:339-348
__global__ void depth_get_occupancy_kernel(MapDev m, const double* __restrict__ occ, double min_occ_log, const double* __restrict__ pos,
                                           long long n, int8_t* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  if (!is_in_map(m, x, y, z)) { out[i] = -1; return; }
  const int ix = pos_to_idx1(x, m.ox, m.inv_res), iy = pos_to_idx1(y, m.oy, m.inv_res), iz = pos_to_idx1(z, m.oz, m.inv_res);
  out[i] = occ[vox_addr(m, ix, iy, iz)] > min_occ_log ? 1 : 0;
}
// occupancy_buffer_[toAddress(boundIndex(posToIndex(pos)))] — the read behind isUnknown / isKnownFree (grid_map.h:276-300)
__global__ void depth_get_log_odds_kernel(MapDev m, const double* __restrict__ occ, const double* __restrict__ pos, long long n,
                                          double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ix = min(max(pos_to_idx1(pos[3 * i], m.ox, m.inv_res), 0), m.Nx - 1);
  const int iy = min(max(pos_to_idx1(pos[3 * i + 1], m.oy, m.inv_res), 0), m.Ny - 1);
  const int iz = min(max(pos_to_idx1(pos[3 * i + 2], m.oz, m.inv_res), 0), m.Nz - 1);
  out[i] = occ[vox_addr(m, ix, iy, iz)];
}
__global__ void depth_fill_kernel(double* __restrict__ p, double v, uint32_t* __restrict__ f0, uint32_t* __restrict__ f1,
                                  uint32_t* __restrict__ f2, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  p[i] = v; f0[i] = kNone; f1[i] = kNone; f2[i] = kNone;
}

// --------------------------------------------------------------------------------------------------
// launchers
// --------------------------------------------------------------------------------------------------
static inline unsigned blocks_for(long long n, int per) { return (unsigned)((n + per - 1) / per); }

void launch_depth_init(const DepthDev& D, double unknown, long long nvox, cudaStream_t st) {
  depth_fill_kernel<<<blocks_for(nvox, 256), 256, 0, st>>>(D.occ, unknown, D.F, D.Fnew, D.E, nvox);
}
void launch_depth_project(const DepthDev& D, const MapDev& m, const uint16_t* img, cudaStream_t st) {
  if (D.n_rays > 0) depth_project_kernel<<<blocks_for(D.n_rays, 256), 256, 0, st>>>(D, m, img);
}
void launch_depth_path_count(const DepthDev& D, const MapDev& m, cudaStream_t st) {
  depth_path_count_kernel<<<blocks_for(D.n_rays, 128), 128, 0, st>>>(D, m);
}
void launch_depth_path_fill(const DepthDev& D, const MapDev& m, unsigned n_cast, cudaStream_t st) {
  if (n_cast) depth_path_fill_kernel<<<blocks_for(n_cast, 128), 128, 0, st>>>(D, m);
}
// the ray-order fixed point; D.ctl->changed[0] must be zero.  Returns a CUDA error code.
int launch_depth_order(const DepthDev& D, unsigned n_cast, int sm_count, cudaStream_t st) {
  if (n_cast == 0) return 0;
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, depth_order_kernel, kOrderThreads, 0);
  if (e != cudaSuccess) return (int)e;
  const unsigned want = blocks_for((long long)n_cast * 32, kOrderThreads);  // a warp per cast ray
  static const int cap_env = getenv("B200_DEPTH_ORDER_BLOCKS") ? atoi(getenv("B200_DEPTH_ORDER_BLOCKS")) : 0;
  unsigned grid = cap_env > 0 ? (unsigned)cap_env : (unsigned)sm_count;  // one block per SM at most
  if (per_sm < 1) return (int)cudaErrorLaunchOutOfResources;
  if (grid > want) grid = want;
  DepthDev d = D;
  void* args[] = {&d};
  return (int)cudaLaunchCooperativeKernel((void*)depth_order_kernel, dim3(grid), dim3(kOrderThreads), args, 0, st);
}
void launch_depth_count(const DepthDev& D, unsigned n_cast, cudaStream_t st) {
  if (n_cast) depth_count_kernel<<<blocks_for((long long)n_cast * 32, 128), 128, 0, st>>>(D);
}
void launch_depth_update(const DepthDev& D, const MapDev& m, int sm_count, cudaStream_t st) {
  depth_update_kernel<<<sm_count * 8, 256, 0, st>>>(D, m);
  depth_reset_kernel<<<sm_count * 8, 256, 0, st>>>(D);
}
void launch_depth_local_map(const DepthDev& D, const MapDev& m, const int lb_min[3], const int lb_max[3], const int cut_min[3],
                            const int cut_max[3], const int cutm_min[3], const int cutm_max[3], int inf_step, int ceil_on, int ceil_id,
                            int sm_count, cudaStream_t st) {
  Box L, C, M;
  for (int i = 0; i < 3; ++i) {
    L.lo[i] = lb_min[i]; L.hi[i] = lb_max[i];
    C.lo[i] = cut_min[i]; C.hi[i] = cut_max[i];
    M.lo[i] = cutm_min[i]; M.hi[i] = cutm_max[i];
  }
  auto cells = [](const Box& b) {
    long long n = 1;
    for (int i = 0; i < 3; ++i) n *= (b.hi[i] >= b.lo[i]) ? (b.hi[i] - b.lo[i] + 1) : 0;
    return n;
  };
  if (cells(M) > 0) depth_shell_kernel<<<blocks_for(cells(M), 256), 256, 0, st>>>(D, m, M, C);
  if (cells(L) > 0) {
    launch_clear_box(m, lb_min, lb_max, sm_count, st);  // :587-593
    depth_inflate_kernel<<<blocks_for(cells(L), 256), 256, 0, st>>>(D, m, L, inf_step);
    if (ceil_on) {
      const long long cols = (long long)(L.hi[0] - L.lo[0] + 1) * (L.hi[1] - L.lo[1] + 1);
      depth_ceiling_kernel<<<blocks_for(cols, 256), 256, 0, st>>>(m, L, ceil_id);
    }
  }
}
// counts/offsets: cols + 1 ints each; temp: export_scan_temp_bytes(cols + 1) bytes
void launch_depth_publish(const DepthDev& D, const MapDev& m, int which, const int a[3], const int b[3], double truncate, int* counts,
                          int* offsets, void* temp, size_t temp_bytes, float4* out, long long cap, cudaStream_t st) {
  Box B;
  for (int i = 0; i < 3; ++i) { B.lo[i] = a[i]; B.hi[i] = b[i]; }
  const long long cols = (long long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1);
  const unsigned g = blocks_for(cols, 256);
  cudaMemsetAsync(counts + cols, 0, sizeof(int), st);
  if (which == 0) depth_publish_kernel<0><<<g, 256, 0, st>>>(D, m, B, truncate, counts, nullptr, nullptr, 0);
  else depth_publish_kernel<1><<<g, 256, 0, st>>>(D, m, B, truncate, counts, nullptr, nullptr, 0);
  launch_exclusive_scan(temp, temp_bytes, counts, offsets, cols + 1, st);
  if (!out) return;  // count only
  if (which == 0) depth_publish_kernel<0><<<g, 256, 0, st>>>(D, m, B, truncate, nullptr, offsets, out, cap);
  else depth_publish_kernel<1><<<g, 256, 0, st>>>(D, m, B, truncate, nullptr, offsets, out, cap);
}
void launch_los_dda(const MapDev& m, const double* s, const double* e, long long n, uint8_t* vis, int32_t* ncells, cudaStream_t st) {
  if (n > 0) los_dda_kernel<<<blocks_for(n, 128), 128, 0, st>>>(m, s, e, n, m.Nx + m.Ny + m.Nz + 4, vis, ncells);
}
void launch_depth_get_log_odds(const MapDev& m, const double* occ, const double* pos, long long n, double* out, cudaStream_t st) {
  if (n > 0) depth_get_log_odds_kernel<<<blocks_for(n, 256), 256, 0, st>>>(m, occ, pos, n, out);
}
void launch_depth_get_occupancy(const MapDev& m, const double* occ, double min_occ_log, const double* pos, long long n, int8_t* out,
                                cudaStream_t st) {
  if (n > 0) depth_get_occupancy_kernel<<<blocks_for(n, 256), 256, 0, st>>>(m, occ, min_occ_log, pos, n, out);
}

}  // namespace b200
