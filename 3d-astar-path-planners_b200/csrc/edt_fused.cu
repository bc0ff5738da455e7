// edt_fused.cu — the exact squared EDT + cost field as TWO shared-memory-staged kernels (pass y fused with the z
// pass, pass x fused with the cost table), replacing the seven-kernel pipeline of edt_kernels.cu wherever a line's
// hull stack fits in shared memory.  NEW functionality like edt_kernels.cu (the reference only has commented-out
// EDT hooks: AlgorithmBase.cpp:40-42, ThetaStarAGR.cpp:221-229); same definition, same outputs, bit for bit.
//
// Block = tile of one (x) [pass 1] or one (y) [pass 2], W consecutive z, the WHOLE line in the third axis:
//   pass 1 (y):  the plane's bit words arrive in shared memory by one bulk async copy (cp.async.bulk + mbarrier);
//                thread per column turns them into W z distances (clz/ffs sweep) in a shared tile; thread (band, lane)
//                builds the band's lower hull in the shared stack; log2(NB) zipper levels; prefix of the surviving
//                intervals; thread (band, lane) walks the hull over its 1/NB of the line and stores, per voxel, the
//                PAIR (|y - y*|, dz(y*)) packed in 16 bits instead of the 19-bit squared distance: the intermediate
//                volume is 2 B/voxel (67 MB at C2: it stays in the 126 MB L2 for pass 2).
//   pass 2 (x):  thread (band, lane) streams its sites from that volume (register prefetch, 64 B per warp row),
//                g = dy^2 + dz^2, hull in shared memory, zipper, walk; writes dist2 (int32) and costq (uint16, table).
// HBM traffic per voxel: 1/8 B in, 2 B + 2 B through L2, 4 B + 2 B out; no stack, z-distance or crossover arrays.
// Per-thread logic: edt_fused.cuh (also compiled for the host by tests/cpu/edt_fused_host_test.cpp).
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
  __syncthreads();
  if (band == 0) total[lane] = hull_offsets(rng + lane, off + lane, W, NB);
  __syncthreads();
  return total[lane];
}

__host__ __device__ constexpr size_t align16(size_t v) { return (v + 15) & ~(size_t)15; }

}  // namespace

// ---------------------------------------------------------------------------------------------------
// pass 1: bits -> z distances -> hull along y -> packed (dy, dz) or int32 squared distance
// ---------------------------------------------------------------------------------------------------
template <int NB, int W, class Tr, class DZ, class TmpT, bool BULK>
__global__ void __launch_bounds__(NB* W) edt_fused_y_kernel(MapDev m, const int32_t* __restrict__ below_gap, const int32_t* __restrict__ above_gap,
                                                           TmpT* __restrict__ tmp, uint8_t* __restrict__ plane_empty, int SB, int BL) {
  using E = typename Tr::E;
  using X = typename Tr::X;
  constexpr bool PACKED = sizeof(TmpT) == 2;
  constexpr int kNone = sizeof(DZ) == 1 ? 255 : 65535;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Ny = m.Ny, Nz = m.Nz, wz = m.wz;
  E* st = reinterpret_cast<E*>(smem_raw);
  DZ* dzt = reinterpret_cast<DZ*>(smem_raw + align16((size_t)NB * BL * W * sizeof(E)));
  uint32_t* bits = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(dzt) + align16((size_t)Ny * W * sizeof(DZ)));
  uint32_t* rng = bits + (BULK ? (((size_t)Ny * wz + 3) & ~(size_t)3) : 0);
  uint16_t* off = reinterpret_cast<uint16_t*>(rng + NB * W);
  int* total = reinterpret_cast<int*>(off + NB * W);
  __shared__ __align__(8) uint64_t bar;

  const int ztiles = (Nz + W - 1) / W;
  const int x = blockIdx.x / ztiles, z0 = (blockIdx.x % ztiles) * W;
  const int tid = threadIdx.x, band = tid / W, lane = tid % W;
  const uint32_t* plane_bits = m.occ + (size_t)x * Ny * wz;
  if (BULK) {
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (tid == 0) {
      const uint32_t bytes = (uint32_t)((size_t)Ny * wz * 4);
      mbar_expect_tx(&bar, bytes);
      bulk_g2s(bits, plane_bits, bytes, &bar);
    }
    mbar_wait(&bar, 0);
  }
  // z distances of the tile: thread per column
  for (int y = tid; y < Ny; y += NB * W) {
    const size_t col = (size_t)x * Ny + y;
    const int bg = below_gap ? __ldg(below_gap + col) : 0, ag = above_gap ? __ldg(above_gap + col) : 0;
    __align__(16) DZ v[W];
    uint32_t cw[8];
    const uint32_t* c = BULK ? bits + (size_t)y * wz : plane_bits + (size_t)y * wz;
    if (!BULK && wz <= 8) {  // read the column once through the read-only path
      for (int j = 0; j < wz; ++j) cw[j] = __ldg(c + j);
      c = cw;
    }
    column_dz<W, DZ>(c, wz, Nz, z0, bg, ag, v, 1);
    constexpr int kBytes = W * (int)sizeof(DZ);
    if (kBytes >= 16) {
#pragma unroll
      for (int k = 0; k < kBytes / 16; ++k) reinterpret_cast<uint4*>(dzt + (size_t)y * W)[k] = reinterpret_cast<const uint4*>(v)[k];
    } else {
      reinterpret_cast<uint2*>(dzt + (size_t)y * W)[0] = reinterpret_cast<const uint2*>(v)[0];
    }
  }
  __syncthreads();
  const int z = z0 + lane;
  const bool active = z < Nz;
  const int u0 = band * BL, u1 = min(Ny, u0 + BL);
  {
    HullBuilder<Tr> hb;
    hb.init(st + (size_t)band * BL * W + lane, W, SB);
    if (active)
      for (int u = u0; u < u1; ++u) {
        const int d = dzt[(size_t)u * W + lane];
        if (d != kNone) hb.push(u, (uint32_t)d);
      }
    rng[band * W + lane] = hb.range();
  }
  const int T = merge_bands<NB, W, Tr>(st, rng, off, total, SB, BL, band, lane);
  if (!active) return;
  if (PACKED && blockIdx.x % ztiles == 0 && tid == 0) plane_empty[x] = T == 0;
  TmpT* o = tmp + (size_t)x * Ny * Nz + z;
  if (T == 0) {
    if (!PACKED)
      for (int u = u0; u < u1; ++u) o[(size_t)u * Nz] = (TmpT)kDistInf;
    return;
  }
  LineHull<Tr> L{st + lane, rng + lane, W, SB, BL};
  hull_fill<Tr>(L, off + lane, NB, T, u0, u1, [&](int u, int s, uint32_t p, X v) {
    if (PACKED) o[(size_t)u * Nz] = (TmpT)((uint32_t)abs(u - s) | (p << SB));
    else o[(size_t)u * Nz] = (TmpT)v;
  });
}

// ---------------------------------------------------------------------------------------------------
// pass 2: hull along x over pass 1's volume -> dist2 (+ costq)
// ---------------------------------------------------------------------------------------------------
template <int NB, int W, class Tr, class TmpT>
__global__ void __launch_bounds__(NB* W) edt_fused_x_kernel(MapDev m, const TmpT* __restrict__ tmp, const uint8_t* __restrict__ plane_empty,
                                                           int32_t* __restrict__ dist2, uint16_t* __restrict__ costq,
                                                           const uint16_t* __restrict__ lut, int lut_n, int SB, int BL, int DYB) {
  using E = typename Tr::E;
  using X = typename Tr::X;
  constexpr bool PACKED = sizeof(TmpT) == 2;
  constexpr int PF = 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int Nx = m.Nx, Ny = m.Ny, Nz = m.Nz;
  E* st = reinterpret_cast<E*>(smem_raw);
  uint32_t* rng = reinterpret_cast<uint32_t*>(smem_raw + align16((size_t)NB * BL * W * sizeof(E)));
  uint16_t* off = reinterpret_cast<uint16_t*>(rng + NB * W);
  int* total = reinterpret_cast<int*>(off + NB * W);
  uint8_t* pe = reinterpret_cast<uint8_t*>(total + W);

  const int ztiles = (Nz + W - 1) / W;
  const int y = blockIdx.x / ztiles, z0 = (blockIdx.x % ztiles) * W;
  const int tid = threadIdx.x, band = tid / W, lane = tid % W;
  if (PACKED) {
    for (int i = tid; i < Nx; i += NB * W) pe[i] = __ldg(plane_empty + i);
    __syncthreads();
  }
  const int z = z0 + lane;
  const bool active = z < Nz;
  const int u0 = band * BL, u1 = min(Nx, u0 + BL);
  const size_t plane = (size_t)Ny * Nz, base = (size_t)y * Nz + z;
  {
    HullBuilder<Tr> hb;
    hb.init(st + (size_t)band * BL * W + lane, W, SB);
    if (active) {
      const TmpT* in = tmp + base;
      const uint32_t dmask = (1u << DYB) - 1u;
      for (int ub = u0; ub < u1; ub += PF) {
        TmpT buf[PF];
#pragma unroll
        for (int j = 0; j < PF; ++j) buf[j] = (ub + j < u1) ? __ldg(in + (size_t)(ub + j) * plane) : (TmpT)0;
#pragma unroll
        for (int j = 0; j < PF; ++j) {
          const int u = ub + j;
          if (u < u1) {
            if (PACKED) {
              if (!pe[u]) {
                const uint32_t c = (uint32_t)buf[j], dy = c & dmask, dz = c >> DYB;
                hb.push(u, dy * dy + dz * dz);
              }
            } else if ((int32_t)buf[j] < kDistInf) hb.push(u, (uint32_t)buf[j]);
          }
        }
      }
    }
    rng[band * W + lane] = hb.range();
  }
  const int T = merge_bands<NB, W, Tr>(st, rng, off, total, SB, BL, band, lane);
  if (!active) return;
  int32_t* o = dist2 + base;
  uint16_t* cq = costq ? costq + base : nullptr;
  if (T == 0) {
    for (int u = u0; u < u1; ++u) {
      o[(size_t)u * plane] = kDistInf;
      if (cq) cq[(size_t)u * plane] = 0;
    }
    return;
  }
  LineHull<Tr> L{st + lane, rng + lane, W, SB, BL};
  hull_fill<Tr>(L, off + lane, NB, T, u0, u1, [&](int u, int, uint32_t, X v) {
    const int32_t d = (int32_t)v;
    o[(size_t)u * plane] = d;
    if (cq) cq[(size_t)u * plane] = d < lut_n ? __ldg(lut + d) : (uint16_t)0;
  });
}

// ---------------------------------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------------------------------
namespace {
int bits_of(unsigned long long v) { int b = 0; while (v) { ++b; v >>= 1; } return b; }
constexpr int kMaxSmem = 227 * 1024;

template <int NB, int W, class E, class DZ>
size_t smem_y(int Ny, int wz, bool bulk) {
  const int BL = (Ny + NB - 1) / NB;
  return align16((size_t)NB * BL * W * sizeof(E)) + align16((size_t)Ny * W * sizeof(DZ)) + (bulk ? (((size_t)Ny * wz + 3) & ~(size_t)3) * 4 : 0) +
         (size_t)NB * W * 6 + (size_t)W * 4 + 16;
}
template <int NB, int W, class E>
size_t smem_x(int Nx) {
  const int BL = (Nx + NB - 1) / NB;
  return align16((size_t)NB * BL * W * sizeof(E)) + (size_t)NB * W * 6 + (size_t)W * 4 + (size_t)Nx + 32;
}

template <int NB, int W, class TrY, class TrX, class DZ, class TmpT>
bool launch_family(const MapDev& m, void* tmp, uint8_t* plane_empty, int32_t* dist2, uint16_t* costq, const uint16_t* lut, int lut_n,
                   const int32_t* below_gap, const int32_t* above_gap, cudaStream_t stream, cudaEvent_t* ev) {
  const bool bulk = ((size_t)m.Ny * m.wz) % 4 == 0 && (((size_t)m.Ny * m.wz * 4) < (1u << 20));
  const size_t sy = smem_y<NB, W, typename TrY::E, DZ>(m.Ny, m.wz, bulk), sx = smem_x<NB, W, typename TrX::E>(m.Nx);
  if (sy > (size_t)kMaxSmem || sx > (size_t)kMaxSmem) return false;
  const int ztiles = (m.Nz + W - 1) / W;
  const int SBy = bits_of((unsigned)(m.Ny - 1)), SBx = bits_of((unsigned)(m.Nx - 1));
  const int BLy = (m.Ny + NB - 1) / NB, BLx = (m.Nx + NB - 1) / NB;
  auto ky_bulk = edt_fused_y_kernel<NB, W, TrY, DZ, TmpT, true>;
  auto ky_plain = edt_fused_y_kernel<NB, W, TrY, DZ, TmpT, false>;
  auto kx = edt_fused_x_kernel<NB, W, TrX, TmpT>;
  auto ky = bulk ? ky_bulk : ky_plain;
  if (cudaFuncSetAttribute(ky, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sy) != cudaSuccess ||
      cudaFuncSetAttribute(kx, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sx) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (ev) cudaEventRecord(ev[0], stream);
  ky<<<(unsigned)(m.Nx * ztiles), NB * W, sy, stream>>>(m, below_gap, above_gap, (TmpT*)tmp, plane_empty, SBy, BLy);
  if (ev) cudaEventRecord(ev[1], stream);
  kx<<<(unsigned)(m.Ny * ztiles), NB * W, sx, stream>>>(m, (const TmpT*)tmp, plane_empty, dist2, costq, lut, lut_n, SBx, BLx, SBy);
  if (ev) cudaEventRecord(ev[2], stream);
  return true;
}
}  // namespace

size_t edt_fused_plane_flags_bytes(const MapDev& m) { return (size_t)m.Nx + 16; }

// Returns false when no fused configuration fits this grid (the caller then runs the seven-kernel pipeline).
// tmp: 4 B/voxel scratch (2 B/voxel used in the packed case); ev (optional): 3 events before y, between, after x.
bool launch_edt_fused(const MapDev& m, void* tmp, uint8_t* plane_empty, int32_t* dist2, uint16_t* costq, const uint16_t* lut, int lut_n,
                      const int32_t* below_gap, const int32_t* above_gap, cudaStream_t stream, cudaEvent_t* ev) {
  if (m.Nx < 2 || m.Ny < 2 || m.Nx > 32768 || m.Ny > 32768) return false;
  const bool halo = below_gap || above_gap;
  const unsigned long long ny1 = m.Ny - 1, nx1 = m.Nx - 1, nz1 = m.Nz - 1;
  const unsigned long long g1 = nz1 * nz1 + ny1 * ny1;  // largest finite value after pass 1
  // "small" family: 16-bit (s, dz) hull entries and 16-bit (dy, dz) intermediate, 32-bit cross products
  const bool small = !halo && m.Nz <= 254 && bits_of(ny1) + bits_of(nz1) <= 16 && bits_of(nx1) + bits_of(g1) <= 32 &&
                     (ny1 * ny1 + nz1 * nz1) * ny1 < (1ull << 31) && (nx1 * nx1 + g1) * nx1 < (1ull << 31);
  if (small) {
    using TY = PassY<uint16_t, int32_t>;
    using TX = PassX<uint32_t, int32_t>;
    if (launch_family<8, 32, TY, TX, uint8_t, uint16_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, nullptr, nullptr, stream, ev)) return true;
  }
  // generic family: 32-bit (s, dz) / 64-bit (s, g) entries, int32 intermediate, 64-bit cross products; z distances < 65535
  if (bits_of(ny1) > 16 || bits_of(nx1) > 16) return false;
  using GY = PassY<uint32_t, long long>;
  using GX = PassX<uint64_t, long long>;
  if (launch_family<8, 32, GY, GX, uint16_t, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, below_gap, above_gap, stream, ev)) return true;
  if (launch_family<16, 16, GY, GX, uint16_t, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, below_gap, above_gap, stream, ev)) return true;
  if (launch_family<32, 8, GY, GX, uint16_t, int32_t>(m, tmp, plane_empty, dist2, costq, lut, lut_n, below_gap, above_gap, stream, ev)) return true;
  return false;
}

}  // namespace b200
