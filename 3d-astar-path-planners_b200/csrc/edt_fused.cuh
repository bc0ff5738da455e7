// edt_fused.cuh — per-thread building blocks of the shared-memory-staged EDT passes (edt_fused.cu).
//
// One pass of the separable squared EDT is, per line, out[u] = min_s (u - s)^2 + g[s].  Writing c_s = s^2 + g[s]
// this is u^2 + min_s (c_s - 2 u s): the minimum of a linear functional over the points (s, c_s), which is attained
// on their LOWER CONVEX HULL.  The sites arrive sorted by s, so the hull is a monotone-chain stack — no separator
// division, and an entry is just (s, payload) where the payload (z distance in pass 1, g in pass 2) gives c_s back.
//
// A line is cut into NB bands; thread (band, lane) builds the hull of its band in the block's shared-memory stack
// (slot [site][lane]: bank = lane for 4-byte entries, so lanes never collide whatever their stack depths).  Hulls of
// neighbouring bands are then zipped pairwise (tree of log2 NB levels): starting from the junction, the last left /
// first right vertex is dropped while it lies on or above the segment joining its two neighbours — exactly the pops
// a sequential scan would have made, just executed later.  Survivors stay where they are; a band only remembers the
// interval [lo, hi] of its stack that is still on the line's hull, and a prefix sum over the bands gives a virtual
// index into the concatenated hull.  The fill then walks the hull once per 1/NB of the line, starting from a binary
// search for the tangent vertex of its first site.
//
// Everything here is plain C++ over pointers so that tests/cpu/edt_fused_host_test.cpp can run the same code on the
// host (phases executed thread after thread) against a brute-force transform.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define EDT_HD __host__ __device__ __forceinline__
#else
#define EDT_HD inline
#endif

namespace b200 {
namespace edtf {

// Pass traits: E = stack entry type (s in the low SB bits, payload above), X = type wide enough for the cross
// products (max c difference times max site difference), g_of = payload -> g.
template <class E_, class X_>
struct PassY {  // payload = z distance
  using E = E_;
  using X = X_;
  static EDT_HD X g_of(uint32_t p) { return (X)p * (X)p; }
};
template <class E_, class X_>
struct PassX {  // payload = g itself
  using E = E_;
  using X = X_;
  static EDT_HD X g_of(uint32_t p) { return (X)p; }
};

template <class Tr>
struct Vtx {
  int s;
  uint32_t p;
  typename Tr::X c;
};
template <class Tr>
EDT_HD Vtx<Tr> unpack(typename Tr::E e, int SB) {
  Vtx<Tr> v;
  v.s = (int)((uint32_t)e & ((1u << SB) - 1u));
  v.p = (uint32_t)(e >> SB);
  v.c = (typename Tr::X)v.s * v.s + Tr::g_of(v.p);
  return v;
}
// b is on or above the segment a--c (s_a < s_b < s_c): b can never be the unique minimiser, drop it
template <class X>
EDT_HD bool drops(int sa, X ca, int sb, X cb, int sc, X cc) {
  return (cb - ca) * (X)(sc - sb) >= (cc - cb) * (X)(sb - sa);
}

// interval of a band's stack that is still on the line's hull: lo | hi << 16; empty when lo > hi
EDT_HD uint32_t rng_pack(int lo, int hi) { return (uint32_t)lo | ((uint32_t)hi << 16); }
EDT_HD int rng_lo(uint32_t r) { return (int)(r & 0xffffu); }
EDT_HD int rng_hi(uint32_t r) { return (int)(r >> 16); }
constexpr uint32_t kRngEmpty = 1u;  // lo = 1, hi = 0

// ---------------------------------------------------------------------------------------------------
// band scan: monotone-chain lower hull of the band's sites.  st = slot 0 of the band for this lane, stride W.
// ---------------------------------------------------------------------------------------------------
template <class Tr>
struct HullBuilder {
  using E = typename Tr::E;
  using X = typename Tr::X;
  E* st;
  int W, SB, q;
  int sA, sB;
  X cA, cB;
  EDT_HD void init(E* st_, int W_, int SB_) { st = st_; W = W_; SB = SB_; q = 0; sA = sB = 0; cA = cB = 0; }
  EDT_HD void push(int u, uint32_t p) {
    const X cu = (X)u * u + Tr::g_of(p);
    while (q >= 2 && drops<X>(sA, cA, sB, cB, u, cu)) {
      --q;
      sB = sA; cB = cA;
      if (q >= 2) {
        const Vtx<Tr> a = unpack<Tr>(st[(q - 2) * W], SB);
        sA = a.s; cA = a.c;
      }
    }
    st[q * W] = (E)((E)(uint32_t)u | ((E)p << SB));
    sA = sB; cA = cB;
    sB = u; cB = cu;
    ++q;
  }
  EDT_HD uint32_t range() const { return q > 0 ? rng_pack(0, q - 1) : kRngEmpty; }
};

// ---------------------------------------------------------------------------------------------------
// zipper: merge the hulls of band groups [b0, bm) and [bm, b1) of one line.
//   st  = slot of (site 0, this lane); entry i of band b is st[(b * BL + i) * W]
//   rng = this lane's interval table, rng[b * W]
// ---------------------------------------------------------------------------------------------------
template <class Tr>
struct LineHull {
  using E = typename Tr::E;
  const E* st;
  uint32_t* rng;
  int W, SB, BL;
  EDT_HD Vtx<Tr> at(int b, int i) const { return unpack<Tr>(st[(b * BL + i) * W], SB); }
  EDT_HD uint32_t r(int b) const { return rng[b * W]; }
};

template <class Tr>
EDT_HD void zipper(const LineHull<Tr>& L, int b0, int bm, int b1) {
  using X = typename Tr::X;
  int ba = bm - 1, bb = bm;
  uint32_t ra = kRngEmpty, rb = kRngEmpty;
  for (; ba >= b0; --ba) { ra = L.r(ba); if (rng_lo(ra) <= rng_hi(ra)) break; }
  if (ba < b0) return;  // nothing on the left: the right group's hull stands
  for (; bb < b1; ++bb) { rb = L.r(bb); if (rng_lo(rb) <= rng_hi(rb)) break; }
  if (bb >= b1) return;
  int ia = rng_hi(ra), la = rng_lo(ra);  // a = (ba, ia); la = first surviving entry of band ba
  int ib = rng_lo(rb), hb = rng_hi(rb);
  Vtx<Tr> a = L.at(ba, ia), b = L.at(bb, ib);
  // a' = predecessor of a, b' = successor of b inside their groups
  int bap = ba, iap = ia, lap = la;
  bool has_ap;
  Vtx<Tr> ap = a, bp = b;
  auto step_pred = [&]() {  // (bap, iap) <- predecessor of it; false when there is none
    if (iap > lap) { --iap; return true; }
    for (int t = bap - 1; t >= b0; --t) {
      const uint32_t rr = L.r(t);
      if (rng_lo(rr) <= rng_hi(rr)) { bap = t; iap = rng_hi(rr); lap = rng_lo(rr); return true; }
    }
    return false;
  };
  int bbp = bb, ibp = ib, hbp = hb;
  bool has_bp;
  auto step_succ = [&]() {
    if (ibp < hbp) { ++ibp; return true; }
    for (int t = bbp + 1; t < b1; ++t) {
      const uint32_t rr = L.r(t);
      if (rng_lo(rr) <= rng_hi(rr)) { bbp = t; ibp = rng_lo(rr); hbp = rng_hi(rr); return true; }
    }
    return false;
  };
  has_ap = step_pred();
  if (has_ap) ap = L.at(bap, iap);
  has_bp = step_succ();
  if (has_bp) bp = L.at(bbp, ibp);
  for (;;) {
    if (has_ap && drops<X>(ap.s, ap.c, a.s, a.c, b.s, b.c)) {
      a = ap; ba = bap; ia = iap; la = lap;
      has_ap = step_pred();
      if (has_ap) ap = L.at(bap, iap);
      continue;
    }
    if (has_bp && drops<X>(a.s, a.c, b.s, b.c, bp.s, bp.c)) {
      b = bp; bb = bbp; ib = ibp; hb = hbp;
      has_bp = step_succ();
      if (has_bp) bp = L.at(bbp, ibp);
      continue;
    }
    break;
  }
  // survivors: left group up to (ba, ia), right group from (bb, ib)
  L.rng[ba * L.W] = rng_pack(la, ia);
  for (int t = ba + 1; t < bm; ++t) L.rng[t * L.W] = kRngEmpty;
  for (int t = bm; t < bb; ++t) L.rng[t * L.W] = kRngEmpty;
  L.rng[bb * L.W] = rng_pack(ib, hb);
}

// prefix of the surviving interval sizes: off[b * W] = number of hull vertices in bands < b; returns the total
EDT_HD int hull_offsets(const uint32_t* rng, uint16_t* off, int W, int NB) {
  int acc = 0;
  for (int b = 0; b < NB; ++b) {
    off[b * W] = (uint16_t)acc;
    const uint32_t r = rng[b * W];
    const int lo = rng_lo(r), hi = rng_hi(r);
    if (lo <= hi) acc += hi - lo + 1;
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------------
// fill: evaluate the line's hull at the sites [u0, u1).  emit(u, s, payload, value) with
// value = (u - s)^2 + g_of(payload) for a minimising site s.  T = number of hull vertices (> 0).
// ---------------------------------------------------------------------------------------------------
template <class Tr>
struct HullCursor {
  using X = typename Tr::X;
  const LineHull<Tr>* L;
  const uint16_t* off;
  int NB, T;
  int v;        // virtual index of the current vertex
  int b, i, hi; // its band, entry, and the band's last surviving entry
  // band of virtual index x: largest b with off[b] <= x (lands on a non-empty band for x < T)
  EDT_HD void seek(int x) {
    int lo_b = 0, hi_b = NB - 1;
    while (lo_b < hi_b) {
      const int mid = (lo_b + hi_b + 1) >> 1;
      if ((int)off[mid * L->W] <= x) lo_b = mid; else hi_b = mid - 1;
    }
    const uint32_t r = L->r(lo_b);
    v = x; b = lo_b; i = rng_lo(r) + (x - (int)off[lo_b * L->W]); hi = rng_hi(r);
  }
  EDT_HD Vtx<Tr> cur() const { return L->at(b, i); }
  EDT_HD void next() {  // v + 1 < T
    if (i < hi) { ++i; ++v; }
    else seek(v + 1);
  }
};

template <class Tr, class Emit>
EDT_HD void hull_fill(const LineHull<Tr>& L, const uint16_t* off, int NB, int T, int u0, int u1, const Emit& emit) {
  using X = typename Tr::X;
  HullCursor<Tr> cu;
  cu.L = &L; cu.off = off; cu.NB = NB; cu.T = T;
  // tangent vertex of u0: first k whose successor is not strictly better
  int lo = 0, hi = T - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    cu.seek(mid);
    const Vtx<Tr> a = cu.cur();
    cu.next();
    const Vtx<Tr> b = cu.cur();
    if (b.c - a.c < (X)(2 * u0) * (X)(b.s - a.s)) lo = mid + 1; else hi = mid;
  }
  cu.seek(lo);
  Vtx<Tr> k = cu.cur(), n = k;
  bool has_n = cu.v + 1 < T;
  if (has_n) { cu.next(); n = cu.cur(); }
  for (int u = u0; u < u1; ++u) {
    while (has_n && n.c - k.c <= (X)(2 * u) * (X)(n.s - k.s)) {
      k = n;
      has_n = cu.v + 1 < T;
      if (has_n) { cu.next(); n = cu.cur(); }
    }
    const int d = u - k.s;
    emit(u, k.s, k.p, (X)d * d + Tr::g_of(k.p));
  }
}

// ---------------------------------------------------------------------------------------------------
// z distances of W consecutive voxels [z0, z0 + W) of one column from its bit words (W divides 32).
// below_gap / above_gap: z-slab halo (distance from the slab's first / last plane to the nearest occupied voxel in
// the slabs below / above, <= 0 when none).  out[j] = kNone when the column (and the halo) holds no obstacle.
// ---------------------------------------------------------------------------------------------------
template <int W, class DZ>
EDT_HD void column_dz(const uint32_t* col, int wz, int Nz, int z0, int below_gap, int above_gap, DZ* out, int out_stride) {
  constexpr int kNone = sizeof(DZ) == 1 ? 255 : 65535;
  const int wi = z0 >> 5, o = z0 & 31;
  const uint32_t own = col[wi];
  int below = -(1 << 20), above = 1 << 20;
  if (below_gap > 0) below = -below_gap;
  if (above_gap > 0) above = Nz - 1 + above_gap;
  {
    const uint32_t lowbits = o ? (own & ((1u << o) - 1u)) : 0u;
    if (lowbits) {
#ifdef __CUDA_ARCH__
      below = (wi << 5) + 31 - __clz(lowbits);
#else
      below = (wi << 5) + 31 - __builtin_clz(lowbits);
#endif
    } else {
      for (int j = wi - 1; j >= 0; --j) {
        const uint32_t v = col[j];
        if (v) {
#ifdef __CUDA_ARCH__
          below = (j << 5) + 31 - __clz(v);
#else
          below = (j << 5) + 31 - __builtin_clz(v);
#endif
          break;
        }
      }
    }
    const uint32_t highbits = (o + W < 32) ? (own >> (o + W)) : 0u;
    if (highbits) {
#ifdef __CUDA_ARCH__
      above = z0 + W + __ffs(highbits) - 1;
#else
      above = z0 + W + __builtin_ffs(highbits) - 1;
#endif
    } else {
      for (int j = wi + 1; j < wz; ++j) {
        const uint32_t v = col[j];
        if (v) {
#ifdef __CUDA_ARCH__
          above = (j << 5) + __ffs(v) - 1;
#else
          above = (j << 5) + __builtin_ffs(v) - 1;
#endif
          break;
        }
      }
    }
  }
  const uint32_t bits = own >> o;
  int up[W];
  int last = below;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if ((bits >> j) & 1u) last = z0 + j;
    up[j] = z0 + j - last;
  }
  int nxt = above;
#pragma unroll
  for (int j = W - 1; j >= 0; --j) {
    if ((bits >> j) & 1u) nxt = z0 + j;
    int d = up[j] < nxt - (z0 + j) ? up[j] : nxt - (z0 + j);
    if (d > kNone) d = kNone;
    out[j * out_stride] = (DZ)d;
  }
}

}  // namespace edtf
}  // namespace b200
