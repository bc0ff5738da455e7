// edt_kernels.cu — exact separable 3-D squared Euclidean distance transform + safety-cost field.
//
// NEW functionality (the reference only has commented-out EDT hooks: AlgorithmBase.cpp:40-42,
// ThetaStarAGR.cpp:221-229).  Definition: dist2[v] = min over inflated-occupied u of |v-u|^2 in voxel
// units, int32, exact; kDistInf where the map holds no obstacle.  cost = max(0, 1 - sqrt(d2)*res/d_safe)
// (float32), costq = (uint16)(cost*65535 + 0.5).
//
// z is the contiguous axis, so a thread that owns line (x,*,z) or (*,y,z) is coalesced with its
// neighbours in z without any transpose.  Five kernels:
//   z     (parallel, one thread per 32-bit word of a column): distance along z to the nearest set bit,
//         from clz/ffs on the bit-packed column, stored as one byte (uint16 when Nz > 254) per voxel.
//   scan  (one thread per line; twice: along y on the z distances, along x on pass 1's output):
//         Meijster's integer lower envelope of parabolas.  Builds the line's stack of
//         (site s, first covered index t, height g) entries — 8 bytes each — into `st`.  The top of stack
//         and the entry below it live in registers; a pop is one early-issued 8-byte load; the separator's
//         division is a shift when the previous site is the top of stack, else a float reciprocal with an
//         integer fix-up (exact for every quotient that matters).
//   fill  (parallel, one thread per 8 consecutive sites of a line; twice): entry k owns sites
//         [t_k, t_{k+1}); the thread binary-searches the entry covering its first site, walks on from there
//         and stores (u-s)^2 + g.  Pass 2's fill also emits the quantised cost through a d2 -> costq table.
// History (profiles/README.md): v1 fused everything in one in-place sweep per line (0.85 + 1.0 ms, ~144
// warp-instructions per site, IPC 1.2); ncu showed ~0.5 pops per site on the C2 map, each recomputing g(s).
#include "kernels.h"

namespace b200 {

// ---------------------------------------------------------------------------------------------------
// z pass
// ---------------------------------------------------------------------------------------------------
// halo (optional, z-slab sharding): below_gap[c] / above_gap[c] = distance in voxels from the slab's first / last
// z-plane to the nearest occupied voxel in the slabs below / above (>= 1), or < 0 when there is none.
template <class T>
__global__ void __launch_bounds__(256) edt_z_kernel(MapDev m, T* __restrict__ dz, const int32_t* __restrict__ below_gap,
                                                    const int32_t* __restrict__ above_gap) {
  constexpr int kNone = sizeof(T) == 1 ? 255 : 65535;
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // (column, word)
  const long long total = (long long)m.Nx * m.Ny * m.wz;
  if (i >= total) return;
  const int wi = (int)(i % m.wz);
  const long long col = i / m.wz;
  const uint32_t* c = m.occ + col * m.wz;
  const uint32_t own = __ldg(c + wi);
  // nearest set bit outside the own word (absolute z), far away when none
  int below = -(1 << 20), above = 1 << 20;
  if (below_gap) { const int g = __ldg(below_gap + col); if (g > 0) below = -g; }
  if (above_gap) { const int g = __ldg(above_gap + col); if (g > 0) above = m.Nz - 1 + g; }
  for (int j = wi - 1; j >= 0; --j) {
    const uint32_t v = __ldg(c + j);
    if (v) { below = (j << 5) + 31 - __clz(v); break; }
  }
  for (int j = wi + 1; j < m.wz; ++j) {
    const uint32_t v = __ldg(c + j);
    if (v) { above = (j << 5) + __ffs(v) - 1; break; }
  }
  const int z0 = wi << 5;
  int up[32];
  int last = below;
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    if ((own >> b) & 1u) last = z0 + b;
    up[b] = z0 + b - last;
  }
  int next = above;
  T* o = dz + col * m.Nz + z0;
#pragma unroll
  for (int b = 31; b >= 0; --b) {
    if ((own >> b) & 1u) next = z0 + b;
    up[b] = min(min(up[b], next - (z0 + b)), kNone);
  }
  if ((m.Nz & 31) == 0) {
    uint4* o4 = reinterpret_cast<uint4*>(o);
    if (sizeof(T) == 1) {
      uint32_t pk[8];
#pragma unroll
      for (int k = 0; k < 8; ++k)
        pk[k] = (uint32_t)up[4 * k] | ((uint32_t)up[4 * k + 1] << 8) | ((uint32_t)up[4 * k + 2] << 16) | ((uint32_t)up[4 * k + 3] << 24);
      o4[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      o4[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        o4[k] = make_uint4((uint32_t)up[8 * k] | ((uint32_t)up[8 * k + 1] << 16), (uint32_t)up[8 * k + 2] | ((uint32_t)up[8 * k + 3] << 16),
                           (uint32_t)up[8 * k + 4] | ((uint32_t)up[8 * k + 5] << 16), (uint32_t)up[8 * k + 6] | ((uint32_t)up[8 * k + 7] << 16));
    }
  } else {
#pragma unroll
    for (int b = 0; b < 32; ++b)
      if (z0 + b < m.Nz) o[b] = (T)up[b];
  }
}

// per (x,y) column: lowest and highest occupied z of this map (-1 when the column is empty) — what a z-slab
// publishes to the other slabs
__global__ void __launch_bounds__(256) column_extents_kernel(MapDev m, int32_t* __restrict__ lowest, int32_t* __restrict__ highest) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)m.Nx * m.Ny) return;
  const uint32_t* c = m.occ + col * m.wz;
  int lo = -1, hi = -1;
  for (int j = 0; j < m.wz; ++j) {
    const uint32_t v = __ldg(c + j);
    if (v) { if (lo < 0) lo = (j << 5) + __ffs(v) - 1; hi = (j << 5) + 31 - __clz(v); }
  }
  lowest[col] = lo;
  highest[col] = hi;
}
void launch_column_extents(const MapDev& m, int32_t* lowest, int32_t* highest, cudaStream_t stream) {
  const long long cols = (long long)m.Nx * m.Ny;
  column_extents_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>(m, lowest, highest);
}

// ---------------------------------------------------------------------------------------------------
// scan
// ---------------------------------------------------------------------------------------------------
template <class T>
struct DzInput {  // 1-D z distance -> squared, streamed along the line
  const T* cur;
  size_t stride;
  __device__ __forceinline__ int32_t first() const { return conv(__ldg(cur)); }
  __device__ __forceinline__ int32_t next() { cur += stride; return conv(__ldg(cur)); }
  static __device__ __forceinline__ int32_t conv(T v) {
    constexpr int kNone = sizeof(T) == 1 ? 255 : 65535;
    return (int)v == kNone ? kDistInf : (int)v * (int)v;
  }
};
struct ArrayInput {
  const int32_t* cur;
  size_t stride;
  __device__ __forceinline__ int32_t first() const { return __ldg(cur); }
  __device__ __forceinline__ int32_t next() { cur += stride; return __ldg(cur); }
};

__device__ __forceinline__ uint32_t pack_st(int s, int t) { return ((uint32_t)t << 16) | (uint32_t)s; }

// floor(num / den) for 0 <= num < 2^31, 2 <= den <= 2^15, exact whenever the quotient is < 2^15; larger
// quotients only need to be reported as "large" (the caller compares against n <= 16384).
__device__ __forceinline__ int floor_div_small(int num, int den) {
  if (den == 2) return num >> 1;  // previous site is the top of stack: the common case
  const float qf = __fmul_rn((float)num, __frcp_rn((float)den));
  if (qf >= 32768.0f) return 32768;
  int q = (int)qf;
  const int rem = num - q * den;
  if (rem < 0) --q;
  else if (rem >= den) ++q;
  return q;
}

// Stack construction for one line of n sites (Meijster et al. 2000, integer separator):
// st[k*stride] = (s_k | t_k << 16, g_k) for k = 0..q; returns q.
constexpr int PF = 4;
template <class In>
__device__ __forceinline__ int envelope_scan(In& in, uint2* __restrict__ st, size_t stride, int n) {
  int q = 0;
  int sq = 0, tq = 0, sb = 0, tb = 0;  // top entry, entry below it
  int32_t gq = in.first(), gb = 0;
  uint2* sp = st;                      // slot of the top entry
  *sp = make_uint2(pack_st(0, 0), (uint32_t)gq);
  for (int u0 = 1; u0 < n; u0 += PF) {
    int32_t gbuf[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) gbuf[j] = (u0 + j < n) ? in.next() : 0;
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int u = u0 + j;
      if (u < n) {
        const int32_t gu = gbuf[j];
        while (q >= 0) {
          const int a = tq - sq, b = tq - u;
          if (a * a + gq > b * b + gu) {
            --q;  // pop: the cached entry below becomes the top; refill the cache with one early load
            sp -= stride;
            sq = sb; tq = tb; gq = gb;
            if (q >= 1) {
              const uint2 e = *(sp - stride);
              sb = (int)(e.x & 0xffffu);
              tb = (int)(e.x >> 16);
              gb = (int32_t)e.y;
            }
          } else break;
        }
        if (q < 0) {
          q = 0; sq = u; tq = 0; gq = gu;
          sp = st;
          *sp = make_uint2(pack_st(u, 0), (uint32_t)gu);
        } else {
          // after the loop parabola sq is still no worse than u at tq >= 0, so the numerator is >= 0
          const int w = 1 + floor_div_small(u * u - sq * sq + gu - gq, 2 * (u - sq));
          if (w < n) {
            ++q;
            sp += stride;
            sb = sq; tb = tq; gb = gq;
            sq = u; tq = w; gq = gu;
            *sp = make_uint2(pack_st(u, w), (uint32_t)gu);
          }
        }
      }
    }
  }
  return q;
}

// pass 1 scan: one thread per (x, z) line along y, input = z distances
template <class T>
__global__ void __launch_bounds__(128) edt_scan_y_kernel(MapDev m, const T* __restrict__ dz, uint2* __restrict__ st, int* __restrict__ counts) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= (long long)m.Nx * m.Nz) return;
  const int x = (int)(l / m.Nz), z = (int)(l % m.Nz);
  const size_t base = (size_t)x * m.Ny * m.Nz + z;
  DzInput<T> in{dz + base, (size_t)m.Nz};
  counts[l] = envelope_scan(in, st + base, (size_t)m.Nz, m.Ny);
}
// pass 2 scan: one thread per (y, z) line along x
__global__ void __launch_bounds__(128) edt_scan_x_kernel(MapDev m, const int32_t* __restrict__ in_arr, uint2* __restrict__ st,
                                                         int* __restrict__ counts) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long plane = (long long)m.Ny * m.Nz;
  if (l >= plane) return;
  ArrayInput in{in_arr + l, (size_t)plane};
  counts[l] = envelope_scan(in, st + l, (size_t)plane, m.Nx);
}

// ---------------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------------
constexpr int SPT = 8;
template <class Emit>
__device__ __forceinline__ void envelope_fill(const uint2* __restrict__ st, size_t stride, int n, int q, int u0, const Emit& emit) {
  if (u0 >= n) return;
  int lo = 0, hi = q;  // invariant: t_lo <= u0 (t_0 == 0)
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int tm = (int)(__ldg(&st[(size_t)mid * stride].x) >> 16);
    if (tm <= u0) lo = mid; else hi = mid - 1;
  }
  int k = lo;
  const uint2* ep = st + (size_t)k * stride;
  uint2 e = __ldg(ep);
  uint2 en = make_uint2(0u, 0u);
  int tnext = n;
  if (k < q) { en = __ldg(ep + stride); tnext = (int)(en.x >> 16); }
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    const int u = u0 + j;
    if (u < n) {
      while (u >= tnext) {  // t is strictly increasing: at most one step per site on average
        ++k;
        ep += stride;
        e = en;
        tnext = n;
        if (k < q) { en = __ldg(ep + stride); tnext = (int)(en.x >> 16); }
      }
      const int d = u - (int)(e.x & 0xffffu);
      const int32_t v = d * d + (int32_t)e.y;
      emit(j, v >= kDistInf ? kDistInf : v);
    }
  }
}

__global__ void __launch_bounds__(128) edt_fill_y_kernel(MapDev m, const uint2* __restrict__ st, const int* __restrict__ counts,
                                                         int32_t* __restrict__ out) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= (long long)m.Nx * m.Nz) return;
  const int x = (int)(l / m.Nz), z = (int)(l % m.Nz);
  const size_t base = (size_t)x * m.Ny * m.Nz + z, stride = (size_t)m.Nz;
  const int u0 = blockIdx.y * SPT;
  int32_t* o = out + base + (size_t)u0 * stride;
  envelope_fill(st + base, stride, m.Ny, counts[l], u0, [&](int j, int32_t v) { o[(size_t)j * stride] = v; });
}

__global__ void __launch_bounds__(128)
edt_fill_x_kernel(MapDev m, const uint2* __restrict__ st, const int* __restrict__ counts, int32_t* __restrict__ out,
                  uint16_t* __restrict__ costq, const uint16_t* __restrict__ lut, int lut_n) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long plane = (long long)m.Ny * m.Nz;
  if (l >= plane) return;
  const size_t stride = (size_t)plane;
  const int u0 = blockIdx.y * SPT;
  int32_t* o = out + l + (size_t)u0 * stride;
  uint16_t* c = costq ? costq + l + (size_t)u0 * stride : nullptr;
  envelope_fill(st + l, stride, m.Nx, counts[l], u0, [&](int j, int32_t v) {
    o[(size_t)j * stride] = v;
    if (c) c[(size_t)j * stride] = v < lut_n ? __ldg(lut + v) : (uint16_t)0;
  });
}

// d2 -> quantised cost table: costq = (uint16)(max(0, 1 - sqrt(d2)*res/d_safe) * 65535 + 0.5), float32 ops
// individually rounded (same formula as the oracle's cost_from_d2 / costq_from_cost)
__global__ void cost_lut_kernel(uint16_t* __restrict__ lut, int n, float res, float d_safe) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float c = 0.0f;
  const float d = __fmul_rn(__fsqrt_rn((float)i), res);
  if (d < d_safe) c = __fsub_rn(1.0f, __fdiv_rn(d, d_safe));
  lut[i] = (uint16_t)__fadd_rn(__fmul_rn(c, 65535.0f), 0.5f);
}

int edt_lut_size(double res, double d_safe) {
  if (!(d_safe > 0)) return 0;
  const double r = d_safe / res + 2.0;  // every d2 with sqrt(d2)*res < d_safe is below r*r
  double n = r * r + 2.0;
  if (n > 64.0 * 1024 * 1024) n = 64.0 * 1024 * 1024;
  return (int)n;
}

// ev (optional): 6 events recorded before z, scan_y, fill_y, scan_x, fill_x and after fill_x
bool edt_dz_is_wide(const MapDev& m, bool halo) { return m.Nz > 254 || halo; }  // halo: z distances are not bounded by the slab

void launch_edt(const MapDev& m, void* dzbuf, int32_t* tmp, uint2* st, int* counts, int32_t* dist2, uint16_t* costq, uint16_t* lut,
                int lut_n, double d_safe, const int32_t* below_gap, const int32_t* above_gap, cudaStream_t stream, cudaEvent_t* ev) {
  const long long lines1 = (long long)m.Nx * m.Nz, lines2 = (long long)m.Ny * m.Nz;
  const unsigned g1 = (unsigned)((lines1 + 127) / 128), g2 = (unsigned)((lines2 + 127) / 128);
  const dim3 f1(g1, (unsigned)((m.Ny + SPT - 1) / SPT)), f2(g2, (unsigned)((m.Nx + SPT - 1) / SPT));
  const long long words = (long long)m.Nx * m.Ny * m.wz;
  const unsigned gz = (unsigned)((words + 255) / 256);
  const bool wide = edt_dz_is_wide(m, below_gap || above_gap);
  if (d_safe > 0 && lut_n > 0) cost_lut_kernel<<<(lut_n + 255) / 256, 256, 0, stream>>>(lut, lut_n, (float)m.res, (float)d_safe);
  if (ev) cudaEventRecord(ev[0], stream);
  if (wide) edt_z_kernel<uint16_t><<<gz, 256, 0, stream>>>(m, (uint16_t*)dzbuf, below_gap, above_gap);
  else edt_z_kernel<uint8_t><<<gz, 256, 0, stream>>>(m, (uint8_t*)dzbuf, below_gap, above_gap);
  if (ev) cudaEventRecord(ev[1], stream);
  if (wide) edt_scan_y_kernel<uint16_t><<<g1, 128, 0, stream>>>(m, (const uint16_t*)dzbuf, st, counts);
  else edt_scan_y_kernel<uint8_t><<<g1, 128, 0, stream>>>(m, (const uint8_t*)dzbuf, st, counts);
  if (ev) cudaEventRecord(ev[2], stream);
  edt_fill_y_kernel<<<f1, 128, 0, stream>>>(m, st, counts, tmp);
  if (ev) cudaEventRecord(ev[3], stream);
  edt_scan_x_kernel<<<g2, 128, 0, stream>>>(m, tmp, st, counts);
  if (ev) cudaEventRecord(ev[4], stream);
  edt_fill_x_kernel<<<f2, 128, 0, stream>>>(m, st, counts, dist2, d_safe > 0 ? costq : nullptr, lut, lut_n);
  if (ev) cudaEventRecord(ev[5], stream);
}

}  // namespace b200
