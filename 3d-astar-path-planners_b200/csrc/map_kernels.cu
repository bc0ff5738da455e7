// map_kernels.cu — GridMap cloud path on B200: clear box, voxelise+inflate scatter, occupancy reads,
// byte<->bit (un)packing for parity, and the batched fastLOS kernel.
//
// Reference behaviour restated (not ported): src/plan_env/grid_map.cpp:155-171 (resetBuffer),
// :711-804 (cloudCallback), include/plan_env/grid_map.h:350-404 (occupancy inlines),
// src/utils/LineOfSight.cpp:9-27 (fastLOS).
#include <cub/device/device_scan.cuh>

#include "kernels.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// clear box: one thread per 32-bit word of the box's columns.  Whole words get a plain store; edge words
// a masked read-modify-write (each word is owned by exactly one thread, so no atomics).
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) clear_box_kernel(MapDev m, int ax, int ay, int az, int bx, int by, int bz) {
  const int w0 = az >> 5, w1 = bz >> 5, nw = w1 - w0 + 1;
  const int ny = by - ay + 1;
  const long long total = (long long)(bx - ax + 1) * ny * nw;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = w0 + (int)(i % nw);
    const long long c = i / nw;
    const int y = ay + (int)(c % ny), x = ax + (int)(c / ny);
    const int zlo = max(az, w << 5), zhi = min(bz, (w << 5) + 31);
    const int nb = zhi - zlo + 1;
    const uint32_t mask = (nb == 32 ? 0xffffffffu : ((1u << nb) - 1u)) << (zlo & 31);
    uint32_t* p = m.occ + ((size_t)x * m.Ny + y) * m.wz + w;
    if (mask == 0xffffffffu) *p = 0u;
    else *p &= ~mask;
  }
}

// ---------------------------------------------------------------------------------------------------
// cloud scatter: one thread per point.  Each point produces (2s+1)^2 columns x 3 z-cells in POSITION
// space (pt + k*res, then floor — grid_map.cpp:765-777), which is not the same as voxelise-then-dilate
// for lattice-aligned points.  The three z cells of a column almost always share a
// 32-bit word, so a column costs one OR.  A word that already holds the bits is skipped after a plain
// load; the remaining ORs go through or_aggregated below (one atomicOr for the whole warp when every voting
// lane targets the same word, otherwise one per lane).
// bounds[0..2] = running min, bounds[3..5] = running max of every generated position (:769-775), kept
// as order-preserving int64 keys so atomicMin/atomicMax apply.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long dkey(double v) {
  long long b = __double_as_longlong(v);
  return b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
}

// Warp-aggregated atomicOr, cheap form.  Each thread has already merged its three z bits into one mask and skipped
// words that are fully set, so what is left to guard against is the whole warp hammering ONE word (a dense surface
// patch, or a degenerate cloud): if every voting lane targets the same word the leader issues a single atomicOr of
// the OR-ed masks, otherwise the lanes issue their own.  A general per-address grouping (__match_any_sync) was measured
// slower than it saves on the C2 frame: scatter 82 us with it, 60 us with plain atomics (its cost grows with the
// number of distinct addresses in the warp, which is the common case).
__device__ __forceinline__ void or_aggregated(uint32_t* addr, uint32_t mask, bool need) {
  // every collective runs with the full mask in warp-uniform control flow: sub-warp masks inside divergent code cost a
  // WARPSYNC / collective bracket per call (96 of them in the first version of this kernel)
  const unsigned todo = __ballot_sync(0xffffffffu, need);
  if (todo == 0u) return;
  const int leader = __ffs(todo) - 1;
  const unsigned long long a0 = __shfl_sync(0xffffffffu, (unsigned long long)addr, leader);
  if (__all_sync(0xffffffffu, !need || (unsigned long long)addr == a0)) {
    const uint32_t merged = __reduce_or_sync(0xffffffffu, need ? mask : 0u);
    if ((int)(threadIdx.x & 31) == leader) atomicOr(addr, merged);
  } else if (need) {
    atomicOr(addr, mask);
  }
}

template <int MAXS>
__global__ void __launch_bounds__(256)
scatter_kernel(MapDev m, const uint8_t* __restrict__ pts, long long n, int step, int offx, int offy, int offz, double camx,
               double camy, double camz, double rx, double ry, double rz, int s, long long* __restrict__ bounds) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  bool accept = false;
  float fx = 0.f, fy = 0.f, fz = 0.f;
  if (i < n) {
    const uint8_t* p = pts + i * step;
    if (step == 16 && offx == 0 && offy == 4 && offz == 8) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p));  // coalesced 128-bit loads (PointXYZ stride)
      fx = v.x; fy = v.y; fz = v.z;
    } else {
      fx = __ldg(reinterpret_cast<const float*>(p + offx));
      fy = __ldg(reinterpret_cast<const float*>(p + offy));
      fz = __ldg(reinterpret_cast<const float*>(p + offz));
    }
    // :754-758  |pt - cam| < range per axis (NaN/Inf fail the compare)
    accept = fabs(dsub((double)fx, camx)) < rx && fabs(dsub((double)fy, camy)) < ry && fabs(dsub((double)fz, camz)) < rz;
  }
  constexpr int SIDE = 2 * MAXS + 1;
  int ixs[SIDE], iys[SIDE];
  uint32_t zmask[2] = {0u, 0u};
  int zw[2] = {-1, -1};
  if (accept) {
#pragma unroll
    for (int k = 0; k < SIDE; ++k) {  // :765-766  pt + k*res (float promoted to double), :777 floor
      const double off = dmul((double)(k - s), m.res);
      ixs[k] = pos_to_idx1(dadd((double)fx, off), m.ox, m.inv_res);
      iys[k] = pos_to_idx1(dadd((double)fy, off), m.oy, m.inv_res);
    }
    for (int k = -1; k <= 1; ++k) {  // inf_step_z = 1 (:736)
      const int iz = pos_to_idx1(dadd((double)fz, dmul((double)k, m.res)), m.oz, m.inv_res);
      if (iz < 0 || iz > m.Nz - 1) continue;  // isInMap(Vector3i) :779
      const int w = iz >> 5;
      const int slot = (zw[0] == w || zw[0] < 0) ? 0 : 1;
      zw[slot] = w;
      zmask[slot] |= 1u << (iz & 31);
    }
  }
  const int side = 2 * s + 1;
  // the three z cells straddle a word boundary for 2 points in 32: the second word's pass runs only for warps that need it
  const int nslots = __any_sync(0xffffffffu, zw[1] >= 0) ? 2 : 1;
#pragma unroll
  for (int a = 0; a < SIDE; ++a) {
    if (a >= side) break;  // warp-uniform
    // a row's words are all requested first (independent loads in flight together), then voted on: with a vote right
    // after each load a warp waits for SIDE * nslots L2 round trips one after the other
    uint32_t* addr[SIDE][2];
    uint32_t cur[SIDE][2];
#pragma unroll
    for (int b = 0; b < SIDE; ++b) {
      bool in = false;
      size_t col = 0;
      if (accept && b < side) {
        const int ix = ixs[a], iy = iys[b];
        in = ix >= 0 && ix <= m.Nx - 1 && iy >= 0 && iy <= m.Ny - 1;
        col = ((size_t)ix * m.Ny + iy) * m.wz;
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const bool live = in && q < nslots && zw[q] >= 0;
        addr[b][q] = live ? m.occ + col + zw[q] : m.occ;
        cur[b][q] = live ? *addr[b][q] : 0xffffffffu;  // "everything already set" for cells that do not exist
      }
    }
#pragma unroll
    for (int b = 0; b < SIDE; ++b) {
      if (b >= side) break;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        if (q >= nslots) break;  // warp-uniform
        // a word that already holds the bits (set by an earlier point) needs nothing
        or_aggregated(addr[b][q], zmask[q], (cur[b][q] & zmask[q]) != zmask[q]);
      }
    }
  }
  // running min/max of the generated positions; monotone rounding => extremes are at k = -s / +s (z: -1 / +1) of the
  // extreme coordinates, so the warp reduces the float coordinates themselves (one redux.sync per key on
  // order-preserving uint32 images) and only lane 0 forms the FP64 positions
  const unsigned am = __ballot_sync(0xffffffffu, accept);
  const float f3[3] = {fx, fy, fz};
  float fmin3[3], fmax3[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const uint32_t b = __float_as_uint(f3[a]);
    const uint32_t k = (b >> 31) ? ~b : (b | 0x80000000u);
    const uint32_t kmin = __reduce_min_sync(0xffffffffu, accept ? k : 0xffffffffu);
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, accept ? k : 0u);
    fmin3[a] = __uint_as_float((kmin >> 31) ? (kmin & 0x7fffffffu) : ~kmin);
    fmax3[a] = __uint_as_float((kmax >> 31) ? (kmax & 0x7fffffffu) : ~kmax);
  }
  if (am && (threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double k = a == 2 ? 1.0 : (double)s;
      const long long lo = dkey(dadd((double)fmin3[a], dmul(-k, m.res)));
      const long long hi = dkey(dadd((double)fmax3[a], dmul(k, m.res)));
      if (lo < bounds[a]) atomicMin(bounds + a, lo);
      if (hi > bounds[3 + a]) atomicMax(bounds + 3 + a, hi);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// parity helpers: bit grid <-> the reference's byte-per-voxel layout
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) unpack_kernel(MapDev m, uint8_t* __restrict__ dst) {
  const long long total = (long long)m.Nx * m.Ny * m.Nz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % m.Nz);
    const long long c = i / m.Nz;
    dst[i] = (uint8_t)((m.occ[c * m.wz + (z >> 5)] >> (z & 31)) & 1u);
  }
}
__global__ void __launch_bounds__(256) pack_kernel(MapDev m, const uint8_t* __restrict__ src) {
  const long long total = (long long)m.Nx * m.Ny * m.wz;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % m.wz);
    const long long c = i / m.wz;
    uint32_t v = 0;
    for (int b = 0; b < 32; ++b) {
      const int z = (w << 5) + b;
      if (z < m.Nz && src[c * m.Nz + z]) v |= 1u << b;
    }
    m.occ[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// per-position reads
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) occupancy_kernel(MapDev m, const double* __restrict__ pos, long long n, int8_t* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int8_t)inflate_occ(m, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
}
__global__ void __launch_bounds__(256) in_map_kernel(MapDev m, const double* __restrict__ pos, long long n, uint8_t* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) out[i] = is_in_map(m, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]) ? 1 : 0;
}
__global__ void __launch_bounds__(256) pos_to_index_kernel(MapDev m, const double* __restrict__ pos, long long n, int32_t* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) {
    out[3 * i] = pos_to_idx1(pos[3 * i], m.ox, m.inv_res);
    out[3 * i + 1] = pos_to_idx1(pos[3 * i + 1], m.oy, m.inv_res);
    out[3 * i + 2] = pos_to_idx1(pos[3 * i + 2], m.oz, m.inv_res);
  }
}

// GridMap::setOccupied (grid_map.h:310-321): inflated buffer <- 1 at the voxel of every in-map position
__global__ void __launch_bounds__(256) set_occupied_kernel(MapDev m, const double* __restrict__ pos, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  if (!is_in_map(m, x, y, z)) return;
  const int ix = pos_to_idx1(x, m.ox, m.inv_res), iy = pos_to_idx1(y, m.oy, m.inv_res), iz = pos_to_idx1(z, m.oz, m.inv_res);
  if (ix < 0 || iy < 0 || iz < 0 || ix >= m.Nx || iy >= m.Ny || iz >= m.Nz) return;  // the reference would write outside its buffer
  atomicOr(m.occ + ((size_t)ix * m.Ny + iy) * m.wz + (iz >> 5), 1u << (iz & 31));
}
// GridMap::setOccupancy (grid_map.h:323-337): occupancy_buffer_ (log-odds grid) <- occ for occ in {0, 1}.  Where several
// positions of one call hit the same voxel the LAST one wins, like the sequential calls: pass 1 leaves the largest
// index per voxel in `winner` (a per-voxel uint32 scratch holding 0xffffffff; keys are n-1-i under atomicMin), pass 2
// lets that position write and restores the scratch.
template <bool APPLY>
__global__ void __launch_bounds__(256) set_occupancy_kernel(MapDev m, double* __restrict__ logodds, const double* __restrict__ pos,
                                                            const double* __restrict__ occ, long long n, uint32_t* __restrict__ winner) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double o = occ[i];
  if (o != 1.0 && o != 0.0) return;  // "occ value error!"
  const double x = pos[3 * i], y = pos[3 * i + 1], z = pos[3 * i + 2];
  if (!is_in_map(m, x, y, z)) return;
  const int ix = pos_to_idx1(x, m.ox, m.inv_res), iy = pos_to_idx1(y, m.oy, m.inv_res), iz = pos_to_idx1(z, m.oz, m.inv_res);
  if (ix < 0 || iy < 0 || iz < 0 || ix >= m.Nx || iy >= m.Ny || iz >= m.Nz) return;
  const size_t a = vox_addr(m, ix, iy, iz);
  const uint32_t key = (uint32_t)(n - 1 - i);
  if (!APPLY) atomicMin(winner + a, key);
  else if (winner[a] == key) { logodds[a] = o; winner[a] = 0xffffffffu; }
}
// distance to the nearest obstacle (metres) at the voxel of every position: sqrt(dist2) * resolution, index clamped
// into the map like GridMap::boundIndex (grid_map.h:267-274)
__global__ void __launch_bounds__(256) distance_at_kernel(MapDev m, const double* __restrict__ pos, long long n, double* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int ix = min(max(pos_to_idx1(pos[3 * i], m.ox, m.inv_res), 0), m.Nx - 1);
  const int iy = min(max(pos_to_idx1(pos[3 * i + 1], m.oy, m.inv_res), 0), m.Ny - 1);
  const int iz = min(max(pos_to_idx1(pos[3 * i + 2], m.oz, m.inv_res), 0), m.Nz - 1);
  out[i] = dmul(__dsqrt_rn((double)m.dist2[vox_addr(m, ix, iy, iz)]), m.res);
}

// ---------------------------------------------------------------------------------------------------
// batched fastLOS: one warp per segment, grid-stride over segments
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) los_batch_kernel(MapDev m, const double* __restrict__ s, const double* __restrict__ e,
                                                        long long n, uint8_t* __restrict__ vis, int32_t* __restrict__ nsamples) {
  const int lane = threadIdx.x & 31;
  const long long warp0 = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long i = warp0; i < n; i += nwarps) {
    int samples = 0;
    const bool v = los_warp<false>(m, s[3 * i], s[3 * i + 1], s[3 * i + 2], e[3 * i], e[3 * i + 1], e[3 * i + 2], lane, &samples, nullptr);
    if (lane == 0) {
      vis[i] = v ? 1 : 0;
      if (nsamples) nsamples[i] = samples;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// map publisher (grid_map.cpp:851-900 publishMapInflate): ordered stream compaction of the occupied voxels of an
// index box into PointXYZ records (x, y, z, 1.0f), x -> y -> z order like the reference's triple loop.
// count per column -> exclusive scan (CUB, not a hot path) -> ordered write.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool export_keep(const MapDev& m, int z, double truncate) {
  return !(dadd(dmul(dadd((double)z, 0.5), m.res), m.oz) > truncate);  // indexToPos (grid_map.h:406-410), :875-876
}
__global__ void __launch_bounds__(256) export_count_kernel(MapDev m, int ax, int ay, int az, int bx, int by, int bz, double truncate,
                                                           int* __restrict__ counts) {
  const int ny = by - ay + 1;
  const long long cols = (long long)(bx - ax + 1) * ny;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cols) return;
  const int x = ax + (int)(i / ny), y = ay + (int)(i % ny);
  const uint32_t* c = m.occ + ((size_t)x * m.Ny + y) * m.wz;
  int n = 0;
  for (int w = az >> 5; w <= bz >> 5; ++w) {
    uint32_t v = __ldg(c + w);
    while (v) {
      const int z = (w << 5) + __ffs(v) - 1;
      v &= v - 1;
      if (z >= az && z <= bz && export_keep(m, z, truncate)) ++n;
    }
  }
  counts[i] = n;
}
__global__ void __launch_bounds__(256) export_write_kernel(MapDev m, int ax, int ay, int az, int bx, int by, int bz, double truncate,
                                                           const int* __restrict__ offsets, float4* __restrict__ out, long long cap) {
  const int ny = by - ay + 1;
  const long long cols = (long long)(bx - ax + 1) * ny;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cols) return;
  const int x = ax + (int)(i / ny), y = ay + (int)(i % ny);
  const uint32_t* c = m.occ + ((size_t)x * m.Ny + y) * m.wz;
  long long o = offsets[i];
  const float px = __double2float_rn(dadd(dmul(dadd((double)x, 0.5), m.res), m.ox));
  const float py = __double2float_rn(dadd(dmul(dadd((double)y, 0.5), m.res), m.oy));
  for (int w = az >> 5; w <= bz >> 5; ++w) {
    uint32_t v = __ldg(c + w);
    while (v) {
      const int z = (w << 5) + __ffs(v) - 1;
      v &= v - 1;
      if (z >= az && z <= bz && export_keep(m, z, truncate)) {
        if (o < cap) out[o] = make_float4(px, py, __double2float_rn(dadd(dmul(dadd((double)z, 0.5), m.res), m.oz)), 1.0f);
        ++o;
      }
    }
  }
}
size_t export_scan_temp_bytes(long long cols) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int*)nullptr, (int*)nullptr, (int)cols);
  return bytes;
}
void launch_exclusive_scan(void* temp, size_t temp_bytes, const int* in, int* out, long long n, cudaStream_t st) {
  cub::DeviceScan::ExclusiveSum(temp, temp_bytes, in, out, (int)n, st);
}
// counts/offsets: cols + 1 ints each (the extra element receives the total); temp: export_scan_temp_bytes(cols + 1)
void launch_export(const MapDev& m, const int a[3], const int b[3], double truncate, int* counts, int* offsets, void* temp,
                   size_t temp_bytes, float4* out, long long cap, cudaStream_t st) {
  const long long cols = (long long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1);
  const unsigned g = (unsigned)((cols + 255) / 256);
  cudaMemsetAsync(counts + cols, 0, sizeof(int), st);
  export_count_kernel<<<g, 256, 0, st>>>(m, a[0], a[1], a[2], b[0], b[1], b[2], truncate, counts);
  cub::DeviceScan::ExclusiveSum(temp, temp_bytes, counts, offsets, (int)(cols + 1), st);
  export_write_kernel<<<g, 256, 0, st>>>(m, a[0], a[1], a[2], b[0], b[1], b[2], truncate, offsets, out, cap);
}

// ---------------------------------------------------------------------------------------------------
// host-side launchers
// ---------------------------------------------------------------------------------------------------
static inline int grid_for(long long n, int block, int cap) {
  long long g = (n + block - 1) / block;
  if (g < 1) g = 1;
  if (g > cap) g = cap;
  return (int)g;
}

void launch_clear_box(const MapDev& m, const int a[3], const int b[3], int sm_count, cudaStream_t st) {
  if (a[0] > b[0] || a[1] > b[1] || a[2] > b[2]) return;
  if (a[0] == 0 && a[1] == 0 && a[2] == 0 && b[0] == m.Nx - 1 && b[1] == m.Ny - 1 && b[2] == m.Nz - 1) {
    cudaMemsetAsync(m.occ, 0, (size_t)m.Nx * m.Ny * m.wz * 4, st);
    return;
  }
  const long long total = (long long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1) * ((b[2] >> 5) - (a[2] >> 5) + 1);
  clear_box_kernel<<<grid_for(total, 256, sm_count * 8), 256, 0, st>>>(m, a[0], a[1], a[2], b[0], b[1], b[2]);
}

int launch_scatter(const MapDev& m, const uint8_t* pts, long long n, int step, int offx, int offy, int offz, const double cam[3],
                   const double range[3], int s, long long* bounds, cudaStream_t st) {
  if (n <= 0) return 0;
  const long long g = (n + 255) / 256;
  if (s <= 1)
    scatter_kernel<1><<<(unsigned)g, 256, 0, st>>>(m, pts, n, step, offx, offy, offz, cam[0], cam[1], cam[2], range[0], range[1], range[2], s, bounds);
  else if (s <= 4)
    scatter_kernel<4><<<(unsigned)g, 256, 0, st>>>(m, pts, n, step, offx, offy, offz, cam[0], cam[1], cam[2], range[0], range[1], range[2], s, bounds);
  else
    return -1;
  return 0;
}

void launch_unpack(const MapDev& m, uint8_t* dst, int sm_count, cudaStream_t st) {
  unpack_kernel<<<grid_for((long long)m.Nx * m.Ny * m.Nz, 256, sm_count * 16), 256, 0, st>>>(m, dst);
}
void launch_pack(const MapDev& m, const uint8_t* src, int sm_count, cudaStream_t st) {
  pack_kernel<<<grid_for((long long)m.Nx * m.Ny * m.wz, 256, sm_count * 16), 256, 0, st>>>(m, src);
}
void launch_occupancy(const MapDev& m, const double* pos, long long n, int8_t* out, cudaStream_t st) {
  if (n > 0) occupancy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, pos, n, out);
}
void launch_in_map(const MapDev& m, const double* pos, long long n, uint8_t* out, cudaStream_t st) {
  if (n > 0) in_map_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, pos, n, out);
}
void launch_pos_to_index(const MapDev& m, const double* pos, long long n, int32_t* out, cudaStream_t st) {
  if (n > 0) pos_to_index_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, pos, n, out);
}
void launch_set_occupied(const MapDev& m, const double* pos, long long n, cudaStream_t st) {
  if (n > 0) set_occupied_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, pos, n);
}
// winner: per-voxel uint32 scratch holding 0xffffffff (restored on return); n < 2^32 - 1
void launch_set_occupancy(const MapDev& m, double* logodds, const double* pos, const double* occ, long long n, uint32_t* winner,
                          cudaStream_t st) {
  if (n <= 0) return;
  const unsigned g = (unsigned)((n + 255) / 256);
  set_occupancy_kernel<false><<<g, 256, 0, st>>>(m, logodds, pos, occ, n, winner);
  set_occupancy_kernel<true><<<g, 256, 0, st>>>(m, logodds, pos, occ, n, winner);
}
void launch_distance_at(const MapDev& m, const double* pos, long long n, double* out, cudaStream_t st) {
  if (n > 0) distance_at_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(m, pos, n, out);
}
void launch_los_batch(const MapDev& m, const double* s, const double* e, long long n, uint8_t* vis, int32_t* nsamples, int sm_count,
                      cudaStream_t st) {
  if (n <= 0) return;
  const long long warps_needed = n;
  const int blocks = grid_for(warps_needed * 32, 256, sm_count * 8);
  los_batch_kernel<<<blocks, 256, 0, st>>>(m, s, e, n, vis, nsamples);
}

}  // namespace b200
