// edt_kernels.cu — exact separable 3-D squared Euclidean distance transform + safety-cost field.
//
// NEW functionality (the reference only has commented-out EDT hooks: AlgorithmBase.cpp:40-42,
// ThetaStarAGR.cpp:221-229).  Definition: dist2[v] = min over inflated-occupied u of |v-u|^2 in voxel
// units, int32, exact; kDistInf where the map holds no obstacle.  cost = max(0, 1 - sqrt(d2)*res/d_safe)
// (float32), costq = (uint16)(cost*65535 + 0.5).
//
// Passes (z is the contiguous axis, so a thread that owns line (x,*,z) or (*,y,z) is coalesced with its
// neighbours in z without any transpose):
//   pass 1 (fused z+y): the 1-D distance along z is computed on the fly from the bit-packed column
//           (clz/ffs on <= Nz/32 words), and fed straight into the exact 1-D lower-envelope scan along y.
//   pass 2 (x, + cost): lower-envelope scan along x on pass 1's output; the final write also emits
//           costq / cost.
// The lower envelope is Meijster's integer formulation (integer separator, no floats).  Its stack
// (site s, start t; 16 bits each) lives IN the line's own output slots: the stack never holds more than
// u+1 entries at step u and entry q is only needed while the backward sweep is still at u >= t[q] >= q, so
// the sweep overwrites a slot only after the entry is in registers.  The top of stack stays in registers.
#include "kernels.h"

namespace b200 {

// squared distance along z from voxel z to the nearest set bit of a bit-packed column
__device__ __forceinline__ int32_t zdist2(const uint32_t* __restrict__ col, int wz, int z) {
  int best = 1 << 14;
  const int w0 = z >> 5, b0 = z & 31;
  // below (<= z)
  for (int w = w0; w >= 0; --w) {
    uint32_t v = __ldg(col + w);
    if (w == w0) v &= (b0 == 31 ? 0xffffffffu : ((2u << b0) - 1u));
    if (v) { best = z - ((w << 5) + 31 - __clz(v)); break; }
  }
  // above (>= z)
  for (int w = w0; w < wz; ++w) {
    uint32_t v = __ldg(col + w);
    if (w == w0) v &= ~((1u << b0) - 1u);
    if (v) { best = min(best, (w << 5) + __ffs(v) - 1 - z); break; }
    if (((w + 1) << 5) - z >= best) break;
  }
  return best >= (1 << 14) ? kDistInf : best * best;
}

struct BitsInput {
  const uint32_t* occ;  // column base of y = 0 for this line's x
  int wz, z;
  size_t col_stride;    // words between consecutive sites
  __device__ __forceinline__ int32_t operator()(int i) const { return zdist2(occ + (size_t)i * col_stride, wz, z); }
};
struct ArrayInput {
  const int32_t* base;
  size_t stride;
  __device__ __forceinline__ int32_t operator()(int i) const { return base[(size_t)i * stride]; }
};

__device__ __forceinline__ uint32_t pack_st(int s, int t) { return ((uint32_t)t << 16) | (uint32_t)s; }

// out[u*stride] = min_i in(i) + (u-i)^2 for one line of n sites.  `emit(u, value)` performs the store.
template <class In, class Emit>
__device__ __forceinline__ void envelope_line(const In& in, int32_t* __restrict__ out, size_t stride, int n, const Emit& emit) {
  uint32_t* stack = reinterpret_cast<uint32_t*>(out);
  int q = 0;
  int sq = 0, tq = 0;
  int32_t gq = in(0);
  stack[0] = pack_st(0, 0);
  for (int u = 1; u < n; ++u) {
    const int32_t gu = in(u);
    while (q >= 0) {
      const int a = tq - sq, b = tq - u;
      if (a * a + gq > b * b + gu) {
        --q;
        if (q >= 0) {
          const uint32_t e = stack[(size_t)q * stride];
          sq = (int)(e & 0xffffu);
          tq = (int)(e >> 16);
          gq = in(sq);
        }
      } else break;
    }
    if (q < 0) {
      q = 0; sq = u; tq = 0; gq = gu;
      stack[0] = pack_st(u, 0);
    } else {
      const int w = 1 + (u * u - sq * sq + gu - gq) / (2 * (u - sq));
      if (w < n) {
        ++q; sq = u; tq = w; gq = gu;
        stack[(size_t)q * stride] = pack_st(u, w);
      }
    }
  }
  for (int u = n - 1; u >= 0; --u) {
    const int d = u - sq;
    const int32_t v = d * d + gq;
    const bool popnext = (u == tq);
    emit(u, v >= kDistInf ? kDistInf : v);
    if (popnext && u > 0) {
      --q;
      const uint32_t e = stack[(size_t)q * stride];
      sq = (int)(e & 0xffffu);
      tq = (int)(e >> 16);
      gq = in(sq);
    }
  }
}

// pass 1: one thread per (x, z) line, scanning y.  Output dist (int32) holds d_z^2 (+) d_y^2.
__global__ void __launch_bounds__(128) edt_zy_kernel(MapDev m, int32_t* __restrict__ out) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (l >= (long long)m.Nx * m.Nz) return;
  const int x = (int)(l / m.Nz), z = (int)(l % m.Nz);
  BitsInput in{m.occ + (size_t)x * m.Ny * m.wz, m.wz, z, (size_t)m.wz};
  int32_t* line = out + (size_t)x * m.Ny * m.Nz + z;
  const size_t stride = (size_t)m.Nz;
  envelope_line(in, line, stride, m.Ny, [&](int u, int32_t v) { line[(size_t)u * stride] = v; });
}

__device__ __forceinline__ float cost_from_d2(int32_t d2, float res, float d_safe) {
  if (d2 >= kDistInf) return 0.0f;
  const float d = __fmul_rn(__fsqrt_rn((float)d2), res);
  if (d >= d_safe) return 0.0f;
  return __fsub_rn(1.0f, __fdiv_rn(d, d_safe));
}

// pass 2: one thread per (y, z) line, scanning x; emits dist2 and (optionally) costq / cost.
__global__ void __launch_bounds__(128)
edt_x_kernel(MapDev m, const int32_t* __restrict__ in_arr, int32_t* __restrict__ out, uint16_t* __restrict__ costq,
             float* __restrict__ cost, float resf, float d_safe) {
  const long long l = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long plane = (long long)m.Ny * m.Nz;
  if (l >= plane) return;
  ArrayInput in{in_arr + l, (size_t)plane};
  int32_t* line = out + l;
  const size_t stride = (size_t)plane;
  envelope_line(in, line, stride, m.Nx, [&](int u, int32_t v) {
    line[(size_t)u * stride] = v;
    if (costq) {
      const float c = cost_from_d2(v, resf, d_safe);
      costq[l + (size_t)u * stride] = (uint16_t)__fadd_rn(__fmul_rn(c, 65535.0f), 0.5f);
      if (cost) cost[l + (size_t)u * stride] = c;
    }
  });
}

void launch_edt(const MapDev& m, int32_t* tmp, int32_t* dist2, uint16_t* costq, float* cost, double d_safe, cudaStream_t st,
                cudaEvent_t mid) {
  const long long lines1 = (long long)m.Nx * m.Nz;
  edt_zy_kernel<<<(unsigned)((lines1 + 127) / 128), 128, 0, st>>>(m, tmp);
  if (mid) cudaEventRecord(mid, st);
  const long long lines2 = (long long)m.Ny * m.Nz;
  edt_x_kernel<<<(unsigned)((lines2 + 127) / 128), 128, 0, st>>>(m, tmp, dist2, d_safe > 0 ? costq : nullptr,
                                                                d_safe > 0 ? cost : nullptr, (float)m.res, (float)d_safe);
}

}  // namespace b200
