// edt_fused.cu — the exact squared EDT + cost field as TWO shared-memory-staged kernels (pass y fused with the z
// pass, pass x fused with the cost table), replacing the seven-kernel pipeline of edt_kernels.cu wherever a line's
// hull stack fits in shared memory.  NEW functionality like edt_kernels.cu (the reference only has commented-out
// EDT hooks: AlgorithmBase.cpp:40-42, ThetaStarAGR.cpp:221-229); same definition, same outputs, bit for bit.
//
// Block = tile of one (x) [pass 1] or one (y) [pass 2], W consecutive z, the WHOLE line in the third axis:
//   pass 1 (y):  the plane's bit words arrive in shared memory by one bulk async copy (cp.async.bulk + mbarrier);
//                thread per column finds the nearest occupied z outside the tile's own word; thread (band, lane) derives
//                each site's z distance from the bit word on the fly (two masks, clz / ffs) and builds the band's lower
//                hull in the shared stack; log2(NB) zipper levels; prefix of the surviving
//                intervals; thread (band, lane) walks the hull over its 1/NB of the line and stores, per voxel, the
//                PAIR (|y - y*|, dz(y*)) packed in 16 bits instead of the 19-bit squared distance: the intermediate
//                volume is 2 B/voxel (67 MB at C2: it stays in the 126 MB L2 for pass 2).
//   pass 2 (x):  thread (band, lane) streams its sites from that volume (register prefetch, 64 B per warp row),
//                g = dy^2 + dz^2, hull in shared memory, zipper, walk; writes dist2 (int32) and costq (uint16, table).
// HBM traffic per voxel: 1/8 B in, 2 B + 2 B through L2, 4 B + 2 B out; no stack, z-distance or crossover arrays.
// Per-thread logic: edt_fused.cuh (also compiled for the host by tests/cpu/edt_fused_host_test.cpp).
#include <algorithm>
#include <cstdlib>

#include "kernels.h"
#include "edt_fused.cuh"

namespace b200 {
using namespace edtf;

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
  }
}

// zipper levels + prefix; every thread of the block calls it (it contains the barriers)
template <int NB, int W, class Tr>
__device__ __forceinline__ int merge_bands(const typename Tr::E* st, uint32_t* rng, uint16_t* off, int* total, int SB, int BL, int band,
                                           int lane) {
  LineHull<Tr> L{st + lane, rng + lane, W, SB, BL};
#pragma unroll
  for (int gs = 2; gs <= NB; gs <<= 1) {
    __syncthreads();
    if ((band & (gs - 1)) == 0) zipper<Tr>(L, band, band + gs / 2, band + gs);
  }
  // band 0's thread made the last merge itself: it can go straight on to the prefix of the surviving intervals
  if (band == 0) total[lane] = hull_offsets(rng + lane, off + lane, W, NB);
  __syncthreads();
  return total[lane];
}

__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// pass 1: bits -> z distance (on the fly, per site) -> hull along y -> packed (dy, dz) or int32 squared distance
// ---------------------------------------------------------------------------------------------------
// BULK: how a block learns its columns' bits — true: one bulk async copy of the plane into shared memory (mbarrier), every
// thread then prepares columns of the whole line from it, block barrier; false (planes that do not fit the stack region,
// or B200_EDT_Y_SETUP=2): the lanes of a band prepare only the band's OWN columns straight from global memory and a
// warp-level sync follows.  Measured equal at C2 (169 vs 168 us): resident blocks hide the set-up either way.
template <int NB, int W, class Tr, class TmpT, bool BULK>
__global__ void __launch_bounds__(NB* W, 5)
    edt_fused_y_kernel(MapDev m, const int32_t* __restrict__ below_gap, const int32_t* __restrict__ above_gap, TmpT* __restrict__ tmp,
                       uint8_t* __restrict__ plane_empty, int SB, int BL, int DYB, int zt0, int nzt) {
  using E = typename Tr::E;
  constexpr bool PACKED = sizeof(TmpT) == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ny = m.Ny, Nz = m.Nz, wz = m.wz;
  E* st = reinterpret_cast<E*>(smem_raw);
  // BULK: the plane's bit words [Ny][wz] land in the (not yet used) stack region; what the scan needs of them — the
  // tile's own word of every column and the nearest occupied z outside it — is copied out before the stack is built
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem_raw);
  uint32_t* own = reinterpret_cast<uint32_t*>(smem_raw + align16((size_t)NB * BL * W * sizeof(E)));
  int2* outside = reinterpret_cast<int2*>(own + (((size_t)Ny + 3) & ~(size_t)3));  // per column: nearest z below / above the own word
  uint32_t* rng = reinterpret_cast<uint32_t*>(outside + Ny);
  uint16_t* off = reinterpret_cast<uint16_t*>(rng + NB * W);
  int* total = reinterpret_cast<int*>(off + NB * W);
  __shared__ __align__(8) uint64_t bar;

  const int x = blockIdx.x / nzt, z0 = (zt0 + blockIdx.x % nzt) * W;  // this launch covers the z-tiles [zt0, zt0 + nzt)
  const int tid = threadIdx.x, band = tid / W, lane = tid % W;
  const int wi = z0 >> 5;
  const uint32_t* plane_bits = m.occ + (size_t)x * Ny * wz;
  const int z = z0 + lane, bi = z & 31;
  const bool active = z < Nz;
  const int u0 = band * BL, u1 = min(Ny, u0 + BL);
  if (BULK) {  // one bulk async copy of the plane's words (TMA engine), completion on an mbarrier
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)((size_t)Ny * wz * 4);
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(bits, plane_bits, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  }
  auto column = [&](int y) {
    const size_t col = (size_t)x * Ny + y;
    const int bg = below_gap ? __ldg(below_gap + col) : 0, ag = above_gap ? __ldg(above_gap + col) : 0;
    int lo, hi;
    if (BULK) {
      column_outside(bits + (size_t)y * wz, wz, Nz, wi, bg, ag, &lo, &hi);
      own[y] = bits[(size_t)y * wz + wi];
    } else {  // the other words of the column come straight from L2 / L1
      column_outside(plane_bits + (size_t)y * wz, wz, Nz, wi, bg, ag, &lo, &hi);
      own[y] = __ldg(plane_bits + (size_t)y * wz + wi);
    }
    outside[y] = make_int2(lo, hi);
  };
  if (!BULK) {
    for (int y = u0 + lane; y < u1; y += W) column(y);
    __syncwarp();  // a band's columns are written and read by the band's own lanes only (a warp holds whole bands)
  } else {
    for (int y = tid; y < Ny; y += NB * W) column(y);
    __syncthreads();
  }
  {
    HullBuilder<Tr> hb;
    hb.init(st + (size_t)band * BL * W + lane, W, SB, u0);
    if (active) {
      for (int u = u0; u < u1; ++u) {
        const int2 o = outside[u];
        const int d = site_dz(own[u], bi, z, o.x, o.y);
        if (d < kDzNone) hb.push(u, (uint32_t)d);
      }
    }
    rng[band * W + lane] = hb.range();
  }
  const int T = merge_bands<NB, W, Tr>(st, rng, off, total, SB, BL, band, lane);
  if (!active) return;
  if (PACKED && blockIdx.x % nzt == 0 && tid == 0) plane_empty[x] = T == 0;  // the same value from every z-tile
  TmpT* o = tmp + ((size_t)x * Ny + u0) * Nz + z;
  if (T == 0) {
    if (!PACKED)
      for (int u = u0; u < u1; ++u, o += Nz) *o = (TmpT)kDistInf;
    return;
  }
  LineHull<Tr> L{st + lane, rng + lane, W, SB, BL};
  hull_fill<Tr>(L, off + lane, NB, T, u0, u1, (int)off[band * W + lane], [&](uint32_t p) { return p << DYB; },  // shifted once per vertex
                [&](int d, uint32_t ps, int32_t v) {
                  if (PACKED) *o = (TmpT)((uint32_t)abs(d) | ps);
                  else *o = (TmpT)v;
                  o += Nz;
                });
}

// ---------------------------------------------------------------------------------------------------
// pass 2: hull along x over pass 1's volume -> dist2 (+ costq)
// ---------------------------------------------------------------------------------------------------
constexpr int kSmemLut = 2048;  // cost-table entries kept in shared memory (d_safe / resolution up to 43 voxels)

template <int NB, int W, class Tr, class TmpT>
__global__ void __launch_bounds__(NB* W, 3) edt_fused_x_kernel(MapDev m, const TmpT* __restrict__ tmp, const uint8_t* __restrict__ plane_empty,
                                                           int32_t* __restrict__ dist2, uint16_t* __restrict__ costq,
                                                           const uint16_t* __restrict__ lut, int lut_n, int SB, int BL, int DYB, int zt0, int nzt) {
  using E = typename Tr::E;
  constexpr bool PACKED = sizeof(TmpT) == 2;
  constexpr int PF = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Nx = m.Nx, Ny = m.Ny, Nz = m.Nz;
  E* st = reinterpret_cast<E*>(smem_raw);
  uint32_t* rng = reinterpret_cast<uint32_t*>(smem_raw + align16((size_t)NB * BL * W * sizeof(E)));
  uint16_t* off = reinterpret_cast<uint16_t*>(rng + NB * W);
  int* total = reinterpret_cast<int*>(off + NB * W);
  uint16_t* slut = reinterpret_cast<uint16_t*>(total + W);  // [min(lut_n, kSmemLut) + 1], last entry 0

  const int y = blockIdx.x / nzt, z0 = (zt0 + blockIdx.x % nzt) * W;
  const int tid = threadIdx.x, band = tid / W, lane = tid % W;
  const bool smem_lut = costq != nullptr && lut_n <= kSmemLut;
  const int z = z0 + lane;
  const bool active = z < Nz;
  const int u0 = band * BL, u1 = min(Nx, u0 + BL);
  const size_t plane = (size_t)Ny * Nz, base = (size_t)y * Nz + z;
  // No block-wide barrier in front of the scan (loading the cost table and the empty-plane flags into shared memory behind
  // a __syncthreads put two serial memory latencies at the start of every block): the cost table is only needed by the
  // walk, so it is fetched after the scan and published by the merge's first barrier; the flags of a band are read by the
  // band's own lanes straight from global memory and combined with one vote.
  int any_empty = 0;
  if (PACKED) {
    if (active && u0 + PF <= u1) {  // the first rows of the scan start travelling before the flags are known
#pragma unroll
      for (int j = 0; j < PF; ++j) asm volatile("prefetch.global.L1 [%0];" ::"l"(tmp + base + (size_t)(u0 + j) * plane));
    }
    for (int u = u0 + lane; u < u1; u += W) any_empty |= __ldg(plane_empty + u);
    const unsigned grp = W >= 32 ? 0xffffffffu : (((1u << (W & 31)) - 1u) << (((tid & 31) / W) * W));  // the lanes of this band
    any_empty = (__ballot_sync(0xffffffffu, any_empty != 0) & grp) != 0u;
  }
  {
    HullBuilder<Tr> hb;
    hb.init(st + (size_t)band * BL * W + lane, W, SB, u0);
    if (active) {
      const TmpT* in = tmp + base + (size_t)u0 * plane;
      const uint32_t dmask = (1u << DYB) - 1u;
      auto site = [&](int u, TmpT raw) {
        if (PACKED) {
          const uint32_t c = (uint32_t)raw, dy = c & dmask, dz = c >> DYB;
          hb.push(u, dy * dy + dz * dz);
        } else if ((int32_t)raw < kDistInf) hb.push(u, (uint32_t)raw);
      };
      int ub = u0;
      if (!(PACKED && any_empty)) {
        for (; ub + PF <= u1; ub += PF) {  // PF sites requested ahead of the serial hull update
          TmpT buf[PF];
#pragma unroll
          for (int j = 0; j < PF; ++j, in += plane) buf[j] = __ldg(in);
#pragma unroll
          for (int j = 0; j < PF; ++j) site(ub + j, buf[j]);
        }
      }
      for (; ub < u1; ++ub, in += plane)
        if (!(PACKED && __ldg(plane_empty + ub))) site(ub, __ldg(in));
    }
    rng[band * W + lane] = hb.range();
  }
  if (smem_lut)
    for (int i = tid; i <= lut_n; i += NB * W) slut[i] = i < lut_n ? __ldg(lut + i) : (uint16_t)0;
  const int T = merge_bands<NB, W, Tr>(st, rng, off, total, SB, BL, band, lane);
  if (!active) return;
  int32_t* o = dist2 + base + (size_t)u0 * plane;
  uint16_t* cq = costq ? costq + base + (size_t)u0 * plane : nullptr;
  if (T == 0) {
    for (int u = u0; u < u1; ++u, o += plane) {
      *o = kDistInf;
      if (cq) { *cq = 0; cq += plane; }
    }
    return;
  }
  LineHull<Tr> L{st + lane, rng + lane, W, SB, BL};
  if (!cq) {
    hull_fill<Tr>(L, off + lane, NB, T, u0, u1, (int)off[band * W + lane], [](uint32_t) { return 0; }, [&](int, int, int32_t v) { *o = (int32_t)v; o += plane; });
  } else if (smem_lut) {
    hull_fill<Tr>(L, off + lane, NB, T, u0, u1, (int)off[band * W + lane], [](uint32_t) { return 0; }, [&](int, int, int32_t v) {
      const int32_t d = (int32_t)v;
      *o = d;
      *cq = slut[min(d, lut_n)];
      o += plane; cq += plane;
    });
  } else {
    hull_fill<Tr>(L, off + lane, NB, T, u0, u1, (int)off[band * W + lane], [](uint32_t) { return 0; }, [&](int, int, int32_t v) {
      const int32_t d = (int32_t)v;
      *o = d;
      *cq = d < lut_n ? __ldg(lut + d) : (uint16_t)0;
      o += plane; cq += plane;
    });
  }
}

// ---------------------------------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------------------------------
namespace {
int bits_of(unsigned long long v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }
constexpr size_t kMaxSmem = 227 * 1024;

struct Cfg { size_t smem; int blocks_per_sm; };
// blocks of `smem` dynamic bytes (+1 KB reserved each) an SM holds
Cfg cfg_of(size_t smem, int threads) {
  Cfg c{smem, 0};
  if (smem <= kMaxSmem) c.blocks_per_sm = (int)std::min<size_t>((228 * 1024) / (smem + 1024), 2048 / threads);
  return c;
}
template <int NB, int W, class E>
bool bulk_ok(const MapDev& m) {  // 16-byte granularity of the bulk copy, and the plane has to fit in the stack region it lands in
  const int BL = (m.Ny + NB - 1) / NB;
  return ((size_t)m.Ny * m.wz) % 4 == 0 && (size_t)m.Ny * m.wz * 4 <= (size_t)NB * BL * W * sizeof(E);
}
template <int NB, int W, class E>
size_t smem_y(const MapDev& m) {
  const int BL = (m.Ny + NB - 1) / NB;
  return align16((size_t)NB * BL * W * sizeof(E)) + ((((size_t)m.Ny) + 3) & ~(size_t)3) * 4 + (size_t)m.Ny * 8 + (size_t)NB * W * 6 + (size_t)W * 4 + 16;
}
template <int NB, int W, class E>
size_t smem_x(const MapDev& m) {
  const int BL = (m.Nx + NB - 1) / NB;
  return align16((size_t)NB * BL * W * sizeof(E)) + (size_t)NB * W * 6 + (size_t)W * 4 + (size_t)(kSmemLut + 2) * 2 + 32;
}

template <int NB, int W, class Tr, class TmpT>
bool launch_y(const MapDev& m, void* tmp, uint8_t* plane_empty, const int32_t* below_gap, const int32_t* above_gap, int zt0, int nzt,
              cudaStream_t stream) {
  const size_t sy = smem_y<NB, W, typename Tr::E>(m);
  static const bool no_bulk = getenv("B200_EDT_Y_SETUP") && atoi(getenv("B200_EDT_Y_SETUP")) == 2;  // tuning knob
  const bool bulk = !no_bulk && bulk_ok<NB, W, typename Tr::E>(m);
  auto k = bulk ? edt_fused_y_kernel<NB, W, Tr, TmpT, true> : edt_fused_y_kernel<NB, W, Tr, TmpT, false>;
  static size_t configured[64][2] = {};  // per instantiation and device: the opt-in is a driver call, pay it once per size
  int dev = 0;
  cudaGetDevice(&dev);
  if (sy > configured[dev & 63][bulk]) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sy) != cudaSuccess) { cudaGetLastError(); return false; }
    configured[dev & 63][bulk] = sy;
  }
  const int BL = (m.Ny + NB - 1) / NB;
  k<<<(unsigned)(m.Nx * nzt), NB * W, sy, stream>>>(m, below_gap, above_gap, (TmpT*)tmp, plane_empty, bits_of((unsigned)(BL - 1)), BL,
                                                     bits_of((unsigned)(m.Ny - 1)), zt0, nzt);
  return true;
}
template <int NB, int W, class Tr, class TmpT>
bool launch_x(const MapDev& m, const void* tmp, const uint8_t* plane_empty, int32_t* dist2, uint16_t* costq, const uint16_t* lut, int lut_n,
              int zt0, int nzt, cudaStream_t stream) {
  const size_t sx = smem_x<NB, W, typename Tr::E>(m);
  auto k = edt_fused_x_kernel<NB, W, Tr, TmpT>;
  static size_t configured[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (sx > configured[dev & 63]) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sx) != cudaSuccess) { cudaGetLastError(); return false; }
    configured[dev & 63] = sx;
  }
  const int BL = (m.Nx + NB - 1) / NB;
  k<<<(unsigned)(m.Ny * nzt), NB * W, sx, stream>>>(m, (const TmpT*)tmp, plane_empty, dist2, costq, lut, lut_n, bits_of((unsigned)(BL - 1)), BL,
                                                     bits_of((unsigned)(m.Ny - 1)), zt0, nzt);
  return true;
}
// The three tile shapes (bands x lanes, 256 threads): the one that keeps the most blocks resident per SM (the kernels are
// latency-bound chains: resident warps matter more than tile width), the wider one on ties.  0 = none fits.
// B200_EDT_SHAPE_Y / _X = 1, 2, 3 force a shape (tuning knob; ignored when that shape does not fit).
template <class E, bool Y>
int pick_shape(const MapDev& m) {
  const size_t s[3] = {Y ? smem_y<8, 32, E>(m) : smem_x<8, 32, E>(m), Y ? smem_y<16, 16, E>(m) : smem_x<16, 16, E>(m),
                       Y ? smem_y<32, 8, E>(m) : smem_x<32, 8, E>(m)};
  static const int forced_y = getenv("B200_EDT_SHAPE_Y") ? atoi(getenv("B200_EDT_SHAPE_Y")) : 0;
  static const int forced_x = getenv("B200_EDT_SHAPE_X") ? atoi(getenv("B200_EDT_SHAPE_X")) : 0;
  const int forced = Y ? forced_y : forced_x;
  if (forced >= 1 && forced <= 3 && cfg_of(s[forced - 1], 256).blocks_per_sm >= 1) return forced;
  int best = 0, best_blocks = 0;
  for (int i = 0; i < 3; ++i) {
    const int b = cfg_of(s[i], 256).blocks_per_sm;
    if (b > best_blocks) { best = i + 1; best_blocks = b; }
  }
  return best;
}
}  // namespace

size_t edt_fused_plane_flags_bytes(const MapDev& m) { return (size_t)m.Nx + 16; }

// Returns false when no fused configuration fits this grid (the caller then runs the seven-kernel pipeline).
// tmp: 4 B/voxel scratch (2 B/voxel used in the packed case); ev (optional): 3 events before y, between, after x.
// dz_bound: with a halo, an upper bound on any z distance (the stacked grid's total height), 0 when unknown.
bool launch_edt_fused(const MapDev& m, void* tmp, uint8_t* plane_empty, int32_t* dist2, uint16_t* costq, const uint16_t* lut, int lut_n,
                      const int32_t* below_gap, const int32_t* above_gap, int dz_bound, cudaStream_t stream, cudaEvent_t* ev, EdtOverlap* ov) {
  if (m.Nx < 2 || m.Ny < 2 || m.Nx > 32768 || m.Ny > 32768) return false;
  if (ov) ov->parts_used = 1;
  const bool halo = below_gap || above_gap;
  const unsigned long long ny1 = m.Ny - 1, nx1 = m.Nx - 1, nz1 = m.Nz - 1;
  const unsigned long long g1 = nz1 * nz1 + ny1 * ny1;  // largest finite value after pass 1 (no halo)
  // "small" family (C1-C3 shapes): 16-bit (s, dz) / 32-bit (s, g) hull entries, 16-bit (dy, dz) intermediate volume,
  // 32-bit cross products; 8 bands x 32 lanes.  Band-relative sites need at most bits(ceil(N/8) - 1) bits.
  const int sby8 = bits_of((unsigned)((m.Ny + 7) / 8 - 1)), sbx8 = bits_of((unsigned)((m.Nx + 7) / 8 - 1));
  const bool small = !halo && bits_of(ny1) + bits_of(nz1) <= 16 && sby8 + bits_of(nz1) <= 16 && sbx8 + bits_of(g1) <= 32 &&
                     (ny1 * ny1 + nz1 * nz1) * ny1 < (1ull << 31) && (nx1 * nx1 + g1) * nx1 < (1ull << 31) &&
                     cfg_of(smem_y<8, 32, uint16_t>(m), 256).blocks_per_sm >= 2 && cfg_of(smem_x<8, 32, uint32_t>(m), 256).blocks_per_sm >= 2;
  // The z range is cut into parts that alternate between two streams: pass 2 of a z-tile only needs pass 1 of the same
  // z-tile, so the last, partly filled wave of one kernel overlaps the first wave of the next instead of idling the SMs
  // (each kernel's tail hidden but the last one's).  B200_EDT_PARTS = number of parts (default 2; 1 = one stream, two
  // launches).  ov == nullptr or a single z-tile: one stream.
  static const int want_parts = getenv("B200_EDT_PARTS") ? std::max(1, atoi(getenv("B200_EDT_PARTS"))) : 2;
  auto run = [&](auto&& launch_y_fn, auto&& launch_x_fn, int W) -> bool {
    const int ztiles = (m.Nz + W - 1) / W;
    const int parts = std::min(ztiles, want_parts);
    if (ov) ov->parts_used = 1;
    if (!ov || parts < 2) {
      bool ok = launch_y_fn(0, ztiles, stream);
      if (ev) cudaEventRecord(ev[1], stream);
      return ok && launch_x_fn(0, ztiles, stream);
    }
    ov->parts_used = parts;
    auto lo = [&](int i) { return (int)((long long)ztiles * i / parts); };
    cudaEventRecord(ov->fork, stream);
    for (int i = 0; i < 2; ++i) cudaStreamWaitEvent(ov->s[i], ov->fork, 0);
    bool ok = true;
    for (int r = 0; r < parts; r += 2) {  // rounds of two parts: Y_r (s0), Y_r+1 (s1), X_r (s0), X_r+1 (s1)
      for (int i = r; i < std::min(parts, r + 2); ++i) {
        ok = ok && launch_y_fn(lo(i), lo(i + 1) - lo(i), ov->s[i & 1]);
        if (i + 2 >= parts) cudaEventRecord(ov->y_done[i & 1], ov->s[i & 1]);  // the stream's last pass-1 launch
      }
      for (int i = r; i < std::min(parts, r + 2); ++i) ok = ok && launch_x_fn(lo(i), lo(i + 1) - lo(i), ov->s[i & 1]);
    }
    for (int i = 0; i < 2; ++i) {
      cudaEventRecord(ov->x_done[i], ov->s[i]);
      cudaStreamWaitEvent(stream, ov->x_done[i], 0);
    }
    return ok;
  };
  if (ev) cudaEventRecord(ev[0], stream);
  bool ok;
  if (small) {
    using TY = PassY<uint16_t, int32_t>;
    using TX = PassX<uint32_t, int32_t>;
    ok = run([&](int z0, int nz, cudaStream_t st) { return launch_y<8, 32, TY, uint16_t>(m, tmp, plane_empty, nullptr, nullptr, z0, nz, st); },
             [&](int z0, int nz, cudaStream_t st) { return launch_x<8, 32, TX, uint16_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st); }, 32);
  } else {
    // "large" family: int32 intermediate, 64-bit cross products, 32-bit (s, dz) entries (z distances < 65535, as in the
    // seven-kernel pipeline's wide mode), (s, g) entries of 32 bits when 6..8 bits of band-relative site + g fit, else 64;
    // tile shape chosen per pass
    using GY = PassY<uint32_t, long long>;
    using GY2 = PassY<uint16_t, long long>;
    using GX4 = PassX<uint32_t, long long>;
    using GX8 = PassX<uint64_t, long long>;
    // pass 1 entries: 16 bits when band-relative site + z distance fit (with a halo: only under the caller's bound)
    const unsigned long long dzmax = !halo ? nz1 : (dz_bound > 0 ? (unsigned long long)std::max(dz_bound, m.Nz) : 65535ull);
    auto sby = [&](int nb) { return bits_of((unsigned)((m.Ny + nb - 1) / nb - 1)); };
    int shy = pick_shape<uint16_t, true>(m);
    const bool y16 = shy != 0 && sby(shy == 1 ? 8 : shy == 2 ? 16 : 32) + bits_of(dzmax) <= 16;
    if (!y16) shy = pick_shape<uint32_t, true>(m);
    // g after pass 1 is < 2^29 by construction (kDistInf); 32-bit entries need the real bound
    // (with a halo the z distances reach into the neighbouring slabs: dz_bound = the caller's bound on them, 0 = unknown)
    const unsigned long long dzb = (unsigned long long)std::max(dz_bound, m.Nz);
    const unsigned long long gmax = !halo ? g1 : (dz_bound > 0 ? std::min<unsigned long long>(dzb * dzb + ny1 * ny1, (1ull << 29) - 1) : (1ull << 29) - 1);
    auto sbx = [&](int nb) { return bits_of((unsigned)((m.Nx + nb - 1) / nb - 1)); };
    int shx = pick_shape<uint32_t, false>(m);
    const bool x32 = shx != 0 && sbx(shx == 1 ? 8 : shx == 2 ? 16 : 32) + bits_of(gmax) <= 32;
    if (!x32) shx = pick_shape<uint64_t, false>(m);
    if (!shy || !shx) return false;
    auto ly = [&](int z0, int nz, cudaStream_t st) -> bool {
      if (y16)
        return shy == 1   ? launch_y<8, 32, GY2, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st)
               : shy == 2 ? launch_y<16, 16, GY2, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st)
                          : launch_y<32, 8, GY2, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st);
      return shy == 1   ? launch_y<8, 32, GY, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st)
             : shy == 2 ? launch_y<16, 16, GY, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st)
                        : launch_y<32, 8, GY, int32_t>(m, tmp, plane_empty, below_gap, above_gap, z0, nz, st);
    };
    auto lx = [&](int z0, int nz, cudaStream_t st) -> bool {
      if (x32)
        return shx == 1   ? launch_x<8, 32, GX4, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st)
               : shx == 2 ? launch_x<16, 16, GX4, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st)
                          : launch_x<32, 8, GX4, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st);
      return shx == 1   ? launch_x<8, 32, GX8, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st)
             : shx == 2 ? launch_x<16, 16, GX8, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st)
                        : launch_x<32, 8, GX8, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, z0, nz, st);
    };
    // the two passes may use different lane widths: split on the coarser tile so both launches cover the same z range
    const int wy = shy == 1 ? 32 : shy == 2 ? 16 : 8, wx = shx == 1 ? 32 : shx == 2 ? 16 : 8;
    if (wy == wx) ok = run(ly, lx, wy);
    else {
      EdtOverlap* keep = ov;
      ov = nullptr;  // different tilings: keep it simple, one stream
      ok = ly(0, (m.Nz + wy - 1) / wy, stream);
      if (ev) cudaEventRecord(ev[1], stream);
      ok = ok && lx(0, (m.Nz + wx - 1) / wx, stream);
      ov = keep;
    }
  }
  if (ev) cudaEventRecord(ev[2], stream);
  return ok;
}

}  // namespace b200
