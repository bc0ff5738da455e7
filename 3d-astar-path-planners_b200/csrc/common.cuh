// common.cuh — device-side view of the map and the FP64 helpers every kernel shares.
//
// FP64 discipline (DESIGN.md "Floating-point pins"): the reference is built -O2 without -march or
// fast-math (CMakeLists.txt:4-7), i.e. every double op is a separately rounded IEEE op.  All position
// arithmetic below therefore uses the explicit round-to-nearest intrinsics (never contracted to FMA), and
// the library is additionally compiled with -fmad=false.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace b200 {

constexpr int32_t kDistInf = 1 << 29;
constexpr double kLosStep = 0.1 / 2.0;  // LineOfSight.cpp:13-14

// HBM layout of one map:
//   occ    uint32 [Nx][Ny][wz]   bit-packed occupancy_buffer_inflate_, bit (z & 31) of word (z >> 5);
//                                z is the contiguous axis exactly like the reference's toAddress (grid_map.h:257)
//   dist2  int32  [Nx][Ny][Nz]   exact squared EDT in voxel units
//   costq  uint16 [Nx][Ny][Nz]   quantised safety cost
//   dk     double [dk_len]       the d_k = d_{k-1} + 0.05 sequence of fastLOS (LineOfSight.cpp:19)
struct MapDev {
  uint32_t* occ;
  int32_t* dist2;
  uint16_t* costq;
  const double* dk;
  int dk_len;
  int Nx, Ny, Nz, wz;
  double ox, oy, oz;  // map_origin_
  double res, inv_res;
  double lo[3], hi[3];  // map_min_boundary_ + 1e-4, map_max_boundary_ - 1e-4 (grid_map.h:372-379)
};

__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }

// Eigen 3.3 squaredNorm(): (x*x + y*y) + z*z
__device__ __forceinline__ double sqnorm3(double x, double y, double z) {
  return dadd(dadd(dmul(x, x), dmul(y, y)), dmul(z, z));
}
__device__ __forceinline__ double norm3(double x, double y, double z) { return __dsqrt_rn(sqnorm3(x, y, z)); }

// grid_map.h:370-385
__device__ __forceinline__ bool is_in_map(const MapDev& m, double x, double y, double z) {
  if (x < m.lo[0] || y < m.lo[1] || z < m.lo[2]) return false;
  if (x > m.hi[0] || y > m.hi[1] || z > m.hi[2]) return false;
  return true;
}
// grid_map.h:400-404 (floor then int conversion == round-down conversion for every in-range value)
__device__ __forceinline__ int pos_to_idx1(double p, double o, double inv) { return __double2int_rd(dmul(dsub(p, o), inv)); }

__device__ __forceinline__ uint32_t occ_bit(const MapDev& m, int ix, int iy, int iz) {
  ix = min(max(ix, 0), m.Nx - 1);
  iy = min(max(iy, 0), m.Ny - 1);
  iz = min(max(iz, 0), m.Nz - 1);
  uint32_t w = __ldg(m.occ + ((size_t)ix * m.Ny + iy) * m.wz + (iz >> 5));
  return (w >> (iz & 31)) & 1u;
}
// grid_map.h:350-359
__device__ __forceinline__ int inflate_occ(const MapDev& m, double x, double y, double z) {
  if (!is_in_map(m, x, y, z)) return -1;
  return (int)occ_bit(m, pos_to_idx1(x, m.ox, m.inv_res), pos_to_idx1(y, m.oy, m.inv_res), pos_to_idx1(z, m.oz, m.inv_res));
}
__device__ __forceinline__ size_t vox_addr(const MapDev& m, int ix, int iy, int iz) {
  return ((size_t)ix * m.Ny + iy) * m.Nz + iz;
}

// One fastLOS (LineOfSight.cpp:9-27) evaluated by a full warp: lane k of each batch evaluates sample
// base+k; __ballot_sync gives the early-out.  Returns visibility (warp-uniform).  *samples receives the
// number of samples the sequential reference loop would have evaluated.  When SUMQ, *sumq receives the
// sum of the quantised cost over the sampled voxels (only meaningful when visible).
template <bool SUMQ>
__device__ __forceinline__ bool los_warp(const MapDev& m, double sx, double sy, double sz, double ex, double ey, double ez,
                                         int lane, int* samples, unsigned long long* sumq) {
  const double vx = dsub(ex, sx), vy = dsub(ey, sy), vz = dsub(ez, sz);
  const double sq = sqnorm3(vx, vy, vz);
  const double dist = __dsqrt_rn(sq);
  double dx = vx, dy = vy, dz = vz;  // Eigen normalized(): v / sqrt(sq) when sq > 0 else v
  if (sq > 0.0) {
    dx = __ddiv_rn(vx, dist);
    dy = __ddiv_rn(vy, dist);
    dz = __ddiv_rn(vz, dist);
  }
  unsigned long long acc = 0;
  for (int base = 0;; base += 32) {
    const int k = base + lane;
    bool active = false, blocked = false;
    if (k < m.dk_len) {
      const double d = __ldg(m.dk + k);
      active = d <= dist;
      if (active) {
        const double px = dadd(sx, dmul(dx, d)), py = dadd(sy, dmul(dy, d)), pz = dadd(sz, dmul(dz, d));
        if (!is_in_map(m, px, py, pz)) blocked = true;
        else {
          const int ix = pos_to_idx1(px, m.ox, m.inv_res), iy = pos_to_idx1(py, m.oy, m.inv_res),
                    iz = pos_to_idx1(pz, m.oz, m.inv_res);
          blocked = occ_bit(m, ix, iy, iz) != 0;
          if (SUMQ && !blocked)
            acc += m.costq[vox_addr(m, min(max(ix, 0), m.Nx - 1), min(max(iy, 0), m.Ny - 1), min(max(iz, 0), m.Nz - 1))];
        }
      }
    }
    const unsigned bm = __ballot_sync(0xffffffffu, blocked);
    const unsigned am = __ballot_sync(0xffffffffu, active);
    if (bm) {
      *samples = base + __ffs(bm);
      return false;
    }
    if (am != 0xffffffffu) {
      *samples = base + __popc(am);
      break;
    }
  }
  if (SUMQ) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    *sumq = acc;
  }
  return true;
}

}  // namespace b200
