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
__global__ void __launch_bounds__(256, 4) edt_z_kernel(  // 64 registers: 28.7 -> 26.9 us; 48 or fewer spill the 32 distances (51+ us)
MapDev m, T* __restrict__ dz, const int32_t* __restrict__ below_gap,
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
// the same extents as one packed (lowest, highest) int16 pair per column: what the slabs exchange (4 B per column)
__global__ void __launch_bounds__(256) column_extents16_kernel(MapDev m, short2* __restrict__ lo_hi) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= (long long)m.Nx * m.Ny) return;
  const uint32_t* c = m.occ + col * m.wz;
  int lo = -1, hi = -1;
  for (int j = 0; j < m.wz; ++j) {
    const uint32_t v = __ldg(c + j);
    if (v) { if (lo < 0) lo = (j << 5) + __ffs(v) - 1; hi = (j << 5) + 31 - __clz(v); }
  }
  lo_hi[col] = make_short2((short)lo, (short)hi);
}
void launch_column_extents16(const MapDev& m, short2* lo_hi, cudaStream_t stream) {
  const long long cols = (long long)m.Nx * m.Ny;
  column_extents16_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>(m, lo_hi);
}
// z-slab halo from the all-gathered extents: per column, the distance from this slab's first / last z-plane to the
// nearest occupied voxel in the slabs below / above (>= 1), -1 where there is none.  gathered = [P][cols] pairs.
struct SlabHeights { int nz[kMaxSlabs]; };
__global__ void __launch_bounds__(256) slab_gaps_kernel(const short2* __restrict__ gathered, SlabHeights h, int P, int rank, long long cols,
                                                        int32_t* __restrict__ below, int32_t* __restrict__ above) {
  const long long col = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (col >= cols) return;
  int b = -1, gap = 0;
  for (int r = rank - 1; r >= 0; --r) {
    const int hi = __ldg(gathered + (size_t)r * cols + col).y;
    if (hi >= 0) { b = gap + (h.nz[r] - hi); break; }  // local z = hi of slab r sits nz[r] - hi voxels below our z = 0, plus the slabs in between
    gap += h.nz[r];
  }
  int a = -1;
  gap = 0;
  for (int r = rank + 1; r < P; ++r) {
    const int lo = __ldg(gathered + (size_t)r * cols + col).x;
    if (lo >= 0) { a = gap + lo + 1; break; }
    gap += h.nz[r];
  }
  below[col] = b;
  above[col] = a;
}
void launch_slab_gaps(const short2* gathered, const int* slab_nz, int P, int rank, long long cols, int32_t* below, int32_t* above,
                      cudaStream_t stream) {
  SlabHeights h;
  for (int i = 0; i < kMaxSlabs; ++i) h.nz[i] = i < P ? slab_nz[i] : 0;
  slab_gaps_kernel<<<(unsigned)((cols + 255) / 256), 256, 0, stream>>>(gathered, h, P, rank, cols, below, above);
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
  // One straight-line path for every lane.  Measured at C2 (scan_y / scan_x us): __frcp_rn + a `den == 2` shortcut
  // 141 / 154; approximate reciprocal (1 ulp — the remainder test below absorbs it — and no slow-path call in the
  // loop) 123 / 140; clamp instead of an early return 121 / 139; no shortcut (lanes no longer split on it) 112 / 131.
  float rc;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rc) : "f"((float)den));
  int q = (int)fminf(__fmul_rn((float)num, rc), 32768.0f);
  const int rem = num - q * den;
  if (rem < 0) --q;
  else if (rem >= den) ++q;
  return q;
}

// Stack construction (Meijster et al. 2000, integer separator) for the sites [u_begin, u_end) of a line whose
// domain is [0, n): st[k*stride] = (s_k | t_k << 16, g_k) for k = 0..q; returns q.  `st` points at the band's own
// stack region and `in` at site u_begin.
//
// Bands: a line is cut into B bands scanned by B independent threads (B x the warps: the scan is latency-bound,
// one line per thread gives only ~14 warps/SM at 512x512x128).  Band b's stack is the lower envelope of ITS sites
// over the whole domain.  Envelopes of site sets that are ordered left-to-right cross exactly once (the right
// one's parabolas all have larger abscissae, so E_right - E_left is non-increasing), so a tiny per-line kernel
// finds the crossover positions and the fill kernel still does one search per site, in the band that owns it.
// PF = sites whose input is requested ahead of the serial stack update.  Measured at C2 (us): scan_y 128 / 123 / 113 / 121
// and scan_x 147 / 136 / 131 / 127 for PF = 1 / 2 / 4 / 8, hence 4 for the u8 input of pass 1 and 8 for the int32 input of pass 2.
template <int PF, class In>
__device__ __forceinline__ int envelope_scan(In& in, uint2* __restrict__ st, size_t stride, int n, int u_begin, int u_end) {
  int q = 0;
  int sq = u_begin, tq = 0;  // top entry (site, start)
  int32_t gq0 = in.first();
  int cq = sq * sq + gq0;    // ... and c = s^2 + g: with it the pop test and the separator share one numerator,
                             //   f(t; s) > f(t; u)  <=>  c_u - c_s < 2 t (u - s),   sep(s, u) = floor((c_u - c_s) / (2 (u - s)))
  uint2* sp = st;            // slot of the top entry
  *sp = make_uint2(pack_st(u_begin, 0), (uint32_t)gq0);
  for (int u0 = u_begin + 1; u0 < u_end; u0 += PF) {
    int32_t gbuf[PF];
#pragma unroll
    for (int j = 0; j < PF; ++j) gbuf[j] = (u0 + j < u_end) ? in.next() : 0;
#pragma unroll
    for (int j = 0; j < PF; ++j) {
      const int u = u0 + j;
      if (u < u_end) {
        const int32_t gu = gbuf[j];
        const int cu = u * u + gu;
        int num = cu - cq, den = 2 * (u - sq);
        while (q >= 0 && num < tq * den) {
          // pop: reload the new top (an L1 hit).  Keeping the entry below the top in registers paid while the scan was
          // latency-bound; now that it is issue-bound the register shuffling costs more than the load (112 -> 107 us)
          --q;
          sp -= stride;
          if (q >= 0) {
            const uint2 e = *sp;
            sq = (int)(e.x & 0xffffu);
            tq = (int)(e.x >> 16);
            cq = sq * sq + (int32_t)e.y;
            num = cu - cq; den = 2 * (u - sq);
          }
        }
        if (q < 0) {
          q = 0; sq = u; tq = 0; cq = cu;
          sp = st;
          *sp = make_uint2(pack_st(u, 0), (uint32_t)gu);
        } else {
          // after the loop parabola sq is still no worse than u at tq >= 0, so the numerator is >= 0
          const int w = 1 + floor_div_small(num, den);
          if (w < n) {
            ++q;
            sp += stride;
            sq = u; tq = w; cq = cu;
            *sp = make_uint2(pack_st(u, w), (uint32_t)gu);
          }
        }
      }
    }
  }
  return q;
}

__host__ __device__ inline int band_len(int n, int nb) { return (n + nb - 1) / nb; }

// pass 1 scan: thread (band, line) with line = (x, z) along y; input = z distances
template <class T>
__global__ void __launch_bounds__(128) edt_scan_y_kernel(MapDev m, const T* __restrict__ dz, uint2* __restrict__ st, int* __restrict__ counts,
                                                         int nb) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= (long long)m.Nx * m.Nz) return;
  const int band = blockIdx.y, bl = band_len(m.Ny, nb);
  const int u0 = band * bl, u1 = min(m.Ny, u0 + bl);
  if (u0 >= u1) { counts[l * nb + band] = -1; return; }
  const int x = (int)(l / m.Nz), z = (int)(l % m.Nz);
  const size_t base = (size_t)x * m.Ny * m.Nz + z + (size_t)u0 * m.Nz;
  DzInput<T> in{dz + base, (size_t)m.Nz};
  counts[l * nb + band] = envelope_scan<4>(in, st + base, (size_t)m.Nz, m.Ny, u0, u1);
}
// pass 2 scan: thread (band, line) with line = (y, z) along x
__global__ void __launch_bounds__(128) edt_scan_x_kernel(MapDev m, const int32_t* __restrict__ in_arr, uint2* __restrict__ st,
                                                         int* __restrict__ counts, int nb) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long plane = (long long)m.Ny * m.Nz;
  if (l >= plane) return;
  const int band = blockIdx.y, bl = band_len(m.Nx, nb);
  const int u0 = band * bl, u1 = min(m.Nx, u0 + bl);
  if (u0 >= u1) { counts[l * nb + band] = -1; return; }
  ArrayInput in{in_arr + l + (size_t)u0 * plane, (size_t)plane};
  counts[l * nb + band] = envelope_scan<8>(in, st + l + (size_t)u0 * plane, (size_t)plane, m.Nx, u0, u1);
}

// ---------------------------------------------------------------------------------------------------
// envelope evaluation and band crossovers
// ---------------------------------------------------------------------------------------------------
// index of the stack entry covering site u: largest k in [0, q] with t_k <= u (t_0 == 0)
__device__ __forceinline__ int covering_entry(const uint2* __restrict__ st, size_t stride, int q, int u) {
  int lo = 0, hi = q;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    const int tm = (int)(__ldg(&st[(size_t)mid * stride].x) >> 16);
    if (tm <= u) lo = mid; else hi = mid - 1;
  }
  return lo;
}
struct LineBands {  // one line's band stacks + crossovers (nb <= 4)
  const uint2* st;  // slot of site 0 of the line
  size_t stride;
  int n, nb, bl;
  int q[4];
  int c01, cmid, c23;  // first site where band 1 beats band 0 / bands {2,3} beat {0,1} / band 3 beats band 2
  __device__ __forceinline__ int band_of(int u) const {
    if (nb == 1) return 0;
    if (nb == 2) return u < c01 ? 0 : 1;
    return u < cmid ? (u < c01 ? 0 : 1) : (u < c23 ? 2 : 3);
  }
  __device__ __forceinline__ const uint2* band_stack(int b) const { return st + (size_t)b * bl * stride; }
};
// A band envelope evaluated at a moving site: the covering entry is remembered, so the next evaluation searches
// outward from it (no load at all while the site stays inside the entry's interval, a gallop of log(distance) probes
// otherwise) instead of restarting from an end of the stack.  The crossover search probes sites that move away from
// the band boundary and then close in on the answer, so the fingers only ever move a little.
struct Finger {
  const uint2* st;
  size_t stride;
  int q;       // last entry; < 0: the band does not exist
  int k;       // covering entry of the last evaluation
  int t, tn;   // its start and the start of the next entry (INT_MAX at the top)
  int s, g;
  __device__ __forceinline__ int T(int i) const { return (int)(__ldg(&st[(size_t)i * stride].x) >> 16); }
  __device__ __forceinline__ void load(int i) {
    const uint2 e = __ldg(st + (size_t)i * stride);
    k = i; s = (int)(e.x & 0xffffu); t = (int)(e.x >> 16); g = (int)e.y;
    tn = i < q ? T(i + 1) : 0x7fffffff;
  }
  __device__ __forceinline__ void init(const uint2* stack, size_t str, int last, bool from_top) {
    st = stack; stride = str; q = last;
    if (q >= 0) load(from_top ? q : 0);
  }
  __device__ __forceinline__ long long at(int u) {
    if (q < 0) return 1LL << 60;  // a band that does not exist never wins
    if (u >= tn) {         // covering entry is above k: tn = t_{k+1} <= u
      int lo = k + 1, hi;  // invariant: t_lo <= u, and (hi > q or t_hi > u)
      for (int step = 1;; step <<= 1) {
        hi = lo + step;
        if (hi > q) { hi = q + 1; break; }
        if (T(hi) > u) break;
        lo = hi;
      }
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T(mid) <= u) lo = mid; else hi = mid;
      }
      load(lo);
    } else if (u < t) {    // below k: t_k > u
      int hi = k, lo;
      for (int step = 1;; step <<= 1) {
        lo = hi - step;
        if (lo <= 0) { lo = 0; break; }
        if (T(lo) <= u) break;
        hi = lo;
      }
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (T(mid) <= u) lo = mid; else hi = mid;
      }
      load(lo);
    }
    const long long d = u - s;
    return d * d + (long long)g;
  }
};
// first u in [0, n] with F_right(u) < F_left(u) (monotone predicate: false ... false true ... true).
// The crossover of two neighbouring band envelopes is almost always at (or a few sites from) their common
// boundary b0, so gallop outward from b0 and bisect only the last bracket.
template <class FL, class FR>
__device__ __forceinline__ int first_right_wins(int n, int b0, const FL& left, const FR& right) {
  int lo, hi;  // predicate false at lo (or lo == -1), true at hi (or hi == n)
  if (b0 < n && right(b0) < left(b0)) {
    hi = b0; lo = b0 - 1;
    int step = 1;
    while (lo >= 0 && right(lo) < left(lo)) { hi = lo; lo -= step; step <<= 1; }
    if (lo < -1) lo = -1;
  } else {
    lo = min(b0, n - 1); hi = lo + 1;
    int step = 1;
    while (hi < n && !(right(hi) < left(hi))) { lo = hi; hi += step; step <<= 1; }
    if (hi > n) hi = n;
  }
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (right(mid) < left(mid)) hi = mid; else lo = mid;
  }
  return hi;
}
// thread (line, which): crossover positions c[which] of a line — which = 0: band 1 over band 0; 1: bands {2,3} over
// bands {0,1}; 2: band 3 over band 2.  The three are independent (the pair envelopes are evaluated as minima).
// A band that does not exist (counts < 0) never wins.
__global__ void __launch_bounds__(128) edt_cross_kernel(const uint2* __restrict__ st, const int* __restrict__ counts, int* __restrict__ cross,
                                                        long long lines, int n, int nb, size_t stride, int line_div, size_t line_mul) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= lines) return;
  const int which = blockIdx.y;
  // slot of site 0: pass 1 lines are (x,z): x*line_mul + z with line_div = Nz; pass 2 lines are flat (line_div = 0)
  const size_t base = line_div ? (size_t)(l / line_div) * line_mul + (size_t)(l % line_div) : (size_t)l;
  const int bl = band_len(n, nb);
  const int* q = counts + l * nb;
  // left set = bands [l_lo, l_hi] probed from their tops, right set = [r_lo, r_hi] probed from their bottoms
  const int l_hi = nb == 2 || which == 0 ? 0 : (which == 2 ? 2 : 1);
  const int l_lo = (nb == 4 && which == 1) ? 0 : l_hi;
  const int r_lo = l_hi + 1;
  const int r_hi = (nb == 4 && which == 1) ? 3 : r_lo;
  Finger fl0, fl1, fr0, fr1;  // (fl1 / fr1 only exist for the middle crossover)
  fl0.init(st + base + (size_t)l_hi * bl * stride, stride, __ldg(q + l_hi), true);
  fr0.init(st + base + (size_t)r_lo * bl * stride, stride, __ldg(q + r_lo), false);
  const bool two = l_lo != l_hi;
  fl1.init(st + base + (size_t)l_lo * bl * stride, stride, two ? __ldg(q + l_lo) : -1, true);
  fr1.init(st + base + (size_t)r_hi * bl * stride, stride, two ? __ldg(q + r_hi) : -1, false);
  const int b0 = r_lo * bl;
  const int c = first_right_wins(n, b0, [&](int u) { return min(fl0.at(u), fl1.at(u)); }, [&](int u) { return min(fr0.at(u), fr1.at(u)); });
  cross[l * 4 + which] = c;
}

// ---------------------------------------------------------------------------------------------------
// fill
// ---------------------------------------------------------------------------------------------------
// sites per fill thread.  The binary search for a thread's first site is what loads L2 (its probes hit a different
// 32-byte sector per lane), so more sites per search help until too few threads are left: with the kernels capped at
// 32 registers, 8 / 12 / 16 sites give fill_y 108 / 101 / 108 us and fill_x 121 / 115 / 114 us at C2.
constexpr int SPT = 12;
template <class Emit>
__device__ __forceinline__ void envelope_fill(const LineBands& L, int u0, const Emit& emit) {
  if (u0 >= L.n) return;
  int band = -1, k = 0, q = 0, tnext = 0;
  const uint2* ep = nullptr;
  uint2 e = make_uint2(0u, 0u), en = make_uint2(0u, 0u);
#pragma unroll
  for (int j = 0; j < SPT; ++j) {
    const int u = u0 + j;
    if (u < L.n) {
      const int b = L.band_of(u);
      if (b != band) {  // (re)locate: binary search in the owning band's stack
        band = b;
        q = L.q[b];
        const uint2* bs = L.band_stack(b);
        k = covering_entry(bs, L.stride, q, u);
        ep = bs + (size_t)k * L.stride;
        e = __ldg(ep);
        tnext = L.n;
        if (k < q) { en = __ldg(ep + L.stride); tnext = (int)(en.x >> 16); }
      }
      if (u >= tnext) {  // t is strictly increasing and sites are consecutive integers: at most one new entry per site
        ++k;
        ep += L.stride;
        e = en;
        tnext = L.n;
        if (k < q) { en = __ldg(ep + L.stride); tnext = (int)(en.x >> 16); }
      }
      const int d = u - (int)(e.x & 0xffffu);
      const int32_t v = d * d + (int32_t)e.y;
      emit(j, v >= kDistInf ? kDistInf : v);
    }
  }
}
__device__ __forceinline__ LineBands load_bands(const uint2* st_line, size_t stride, int n, int nb, const int* __restrict__ counts,
                                                const int4* __restrict__ cross, long long l) {
  LineBands L;
  L.st = st_line; L.stride = stride; L.n = n; L.nb = nb; L.bl = band_len(n, nb);
#pragma unroll
  for (int b = 0; b < 4; ++b) L.q[b] = b < nb ? __ldg(counts + l * nb + b) : -1;
  const int4 c = nb > 1 ? __ldg(cross + l) : make_int4(n, n, n, 0);
  L.c01 = c.x; L.cmid = nb == 4 ? c.y : n; L.c23 = nb == 4 ? c.z : n;
  return L;
}

__global__ void __launch_bounds__(128, 16) edt_fill_y_kernel(MapDev m, const uint2* __restrict__ st, const int* __restrict__ counts,
                                                         const int4* __restrict__ cross, int nb, int32_t* __restrict__ out) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= (long long)m.Nx * m.Nz) return;
  const int x = (int)(l / m.Nz), z = (int)(l % m.Nz);
  const size_t base = (size_t)x * m.Ny * m.Nz + z, stride = (size_t)m.Nz;
  const int u0 = blockIdx.y * SPT;
  const LineBands L = load_bands(st + base, stride, m.Ny, nb, counts, cross, l);
  int32_t* o = out + base + (size_t)u0 * stride;
  envelope_fill(L, u0, [&](int j, int32_t v) { o[(size_t)j * stride] = v; });
}

__global__ void __launch_bounds__(128, 16)  // 32 registers: the fills are latency-bound and want all 64 warps (36 registers cost 12 %)
edt_fill_x_kernel(MapDev m, const uint2* __restrict__ st, const int* __restrict__ counts, const int4* __restrict__ cross, int nb,
                  int32_t* __restrict__ out, uint16_t* __restrict__ costq, const uint16_t* __restrict__ lut, int lut_n) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long plane = (long long)m.Ny * m.Nz;
  if (l >= plane) return;
  const size_t stride = (size_t)plane;
  const int u0 = blockIdx.y * SPT;
  const LineBands L = load_bands(st + l, stride, m.Nx, nb, counts, cross, l);
  int32_t* o = out + l + (size_t)u0 * stride;
  uint16_t* c = costq ? costq + l + (size_t)u0 * stride : nullptr;
  envelope_fill(L, u0, [&](int j, int32_t v) {
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

void launch_cost_lut(uint16_t* lut, int lut_n, double res, double d_safe, cudaStream_t stream) {
  if (d_safe > 0 && lut_n > 0) cost_lut_kernel<<<(lut_n + 255) / 256, 256, 0, stream>>>(lut, lut_n, (float)res, (float)d_safe);
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
int edt_bands(int n) { return n >= 256 ? 4 : (n >= 64 ? 2 : 1); }

// ev (optional): 8 events recorded before z, scan_y, cross_y, fill_y, scan_x, cross_x, fill_x and after fill_x
void launch_edt(const MapDev& m, void* dzbuf, int32_t* tmp, uint2* st, int* counts, int4* cross, int32_t* dist2, uint16_t* costq,
                uint16_t* lut, int lut_n, double d_safe, const int32_t* below_gap, const int32_t* above_gap, cudaStream_t stream,
                cudaEvent_t* ev) {
  const long long lines1 = (long long)m.Nx * m.Nz, lines2 = (long long)m.Ny * m.Nz;
  const unsigned g1 = (unsigned)((lines1 + 127) / 128), g2 = (unsigned)((lines2 + 127) / 128);
  const int nb1 = edt_bands(m.Ny), nb2 = edt_bands(m.Nx);
  const dim3 s1(g1, nb1), s2(g2, nb2);
  const dim3 f1(g1, (unsigned)((m.Ny + SPT - 1) / SPT)), f2(g2, (unsigned)((m.Nx + SPT - 1) / SPT));
  const long long words = (long long)m.Nx * m.Ny * m.wz;
  const unsigned gz = (unsigned)((words + 255) / 256);
  const bool wide = edt_dz_is_wide(m, below_gap || above_gap);
  if (ev) cudaEventRecord(ev[0], stream);
  if (wide) edt_z_kernel<uint16_t><<<gz, 256, 0, stream>>>(m, (uint16_t*)dzbuf, below_gap, above_gap);
  else edt_z_kernel<uint8_t><<<gz, 256, 0, stream>>>(m, (uint8_t*)dzbuf, below_gap, above_gap);
  if (ev) cudaEventRecord(ev[1], stream);
  if (wide) edt_scan_y_kernel<uint16_t><<<s1, 128, 0, stream>>>(m, (const uint16_t*)dzbuf, st, counts, nb1);
  else edt_scan_y_kernel<uint8_t><<<s1, 128, 0, stream>>>(m, (const uint8_t*)dzbuf, st, counts, nb1);
  if (ev) cudaEventRecord(ev[2], stream);
  if (nb1 > 1) edt_cross_kernel<<<dim3(g1, nb1 == 4 ? 3 : 1), 128, 0, stream>>>(st, counts, (int*)cross, lines1, m.Ny, nb1, (size_t)m.Nz, m.Nz, (size_t)m.Ny * m.Nz);
  if (ev) cudaEventRecord(ev[3], stream);
  edt_fill_y_kernel<<<f1, 128, 0, stream>>>(m, st, counts, cross, nb1, tmp);
  if (ev) cudaEventRecord(ev[4], stream);
  edt_scan_x_kernel<<<s2, 128, 0, stream>>>(m, tmp, st, counts, nb2);
  if (ev) cudaEventRecord(ev[5], stream);
  if (nb2 > 1) edt_cross_kernel<<<dim3(g2, nb2 == 4 ? 3 : 1), 128, 0, stream>>>(st, counts, (int*)cross, lines2, m.Nx, nb2, (size_t)m.Ny * m.Nz, 0, 0);
  if (ev) cudaEventRecord(ev[6], stream);
  edt_fill_x_kernel<<<f2, 128, 0, stream>>>(m, st, counts, cross, nb2, dist2, d_safe > 0 ? costq : nullptr, lut, lut_n);
  if (ev) cudaEventRecord(ev[7], stream);
}

}  // namespace b200
