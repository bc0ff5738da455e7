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

// Pass traits: E = stack entry type (s in the low SB bits, payload above), X = type of the cross PRODUCTS (c difference
// times site difference: int32 when they fit, else int64 — one IMAD.WIDE of two 32-bit operands), g_of = payload -> g.
// c = s^2 + g itself always fits 32 bits (s < 2^15, g < 2^29).
template <class E_, class X_>
struct PassY {  // payload = z distance
  using E = E_;
  using X = X_;
  static EDT_HD int32_t g_of(uint32_t p) { return (int32_t)(p * p); }
};
template <class E_, class X_>
struct PassX {  // payload = g itself
  using E = E_;
  using X = X_;
  static EDT_HD int32_t g_of(uint32_t p) { return (int32_t)p; }
};

template <class Tr>
struct Vtx {
  int s;
  uint32_t p;
  int32_t c;
};
// entries hold the site RELATIVE to their band's first site (SB = bits of band length - 1): 6 bits for 64-site bands,
// which is what lets a 16-bit (pass 1) / 32-bit (pass 2) entry carry the payload of grids as large as 2048^2 x 512
template <class Tr>
EDT_HD Vtx<Tr> unpack(typename Tr::E e, int SB, int sbase) {
  Vtx<Tr> v;
  v.s = (int)((uint32_t)e & ((1u << SB) - 1u)) + sbase;
  v.p = (uint32_t)(e >> SB);
  v.c = v.s * v.s + Tr::g_of(v.p);
  return v;
}
// b is on or above the segment a--c (s_a < s_b < s_c): b can never be the unique minimiser, drop it
template <class X>
EDT_HD bool drops(int sa, int32_t ca, int sb, int32_t cb, int sc, int32_t cc) {
  return (X)(cb - ca) * (X)(sc - sb) >= (X)(cc - cb) * (X)(sb - sa);
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
  E* top;      // slot of the top entry B (one stride below the band's first slot while the stack is empty)
  int W, SB, sbase, q;
  int sB, dS;  // top entry's site, and its distance to the entry A below it (valid while q >= 2)
  int32_t cB, dC;  // likewise for c: A is recovered as (sB - dS, cB - dC), so a pop needs no copy of A
  // With fewer than two entries there is nothing to pop: (dS, dC) = (0, -1) makes the pop test  dC (u - sB) >= (cu - cB) dS
  // read  -(u - sB) >= 0, false for every later site, so the loop needs no separate test of the count.
  EDT_HD void init(E* st_, int W_, int SB_, int sbase_) { top = st_ - W_; W = W_; SB = SB_; sbase = sbase_; q = 0; sB = -1; dS = 0; cB = 0; dC = -1; }
  EDT_HD void push(int u, uint32_t p) {
    const int32_t cu = u * u + Tr::g_of(p);
    // B is on or above the segment A--C  <=>  (cB - cA) (u - sB) >= (cu - cB) (sB - sA)
    while ((X)dC * (X)(u - sB) >= (X)(cu - cB) * (X)dS) {
      --q;
      top -= W;
      sB -= dS; cB -= dC;
      if (q >= 2) {
        const Vtx<Tr> a = unpack<Tr>(top[-W], SB, sbase);
        dS = sB - a.s; dC = cB - a.c;
      } else {
        dS = 0; dC = -1;
      }
    }
    top += W;
    *top = (E)((E)(uint32_t)(u - sbase) | ((E)p << SB));
    if (q >= 1) { dS = u - sB; dC = cu - cB; }
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
  EDT_HD Vtx<Tr> at(int b, int i) const { return unpack<Tr>(st[(b * BL + i) * W], SB, b * BL); }
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
  using E = typename Tr::E;
  const LineHull<Tr>* L;
  const uint16_t* off;
  int NB, T;
  int v;            // virtual index of the current vertex
  const E* ep;      // its stack slot
  int before, after;  // surviving entries of its band below / above it
  int sbase;          // first site of its band
  int band;           // that band
  // band of virtual index x: largest b with off[b] <= x (lands on a non-empty band for x < T)
  EDT_HD void seek(int x) {
    int lo_b = 0, hi_b = NB - 1;
    while (lo_b < hi_b) {
      const int mid = (lo_b + hi_b + 1) >> 1;
      if ((int)off[mid * L->W] <= x) lo_b = mid; else hi_b = mid - 1;
    }
    const uint32_t r = L->r(lo_b);
    before = x - (int)off[lo_b * L->W];
    const int i = rng_lo(r) + before;
    after = rng_hi(r) - i;
    sbase = lo_b * L->BL;
    band = lo_b;
    ep = L->st + (sbase + i) * L->W;
    v = x;
  }
  // entering a neighbouring band: the next / previous NON-EMPTY one (it exists: the callers check v against T / 0),
  // almost always the adjacent band — a short linear scan instead of seek()'s bisection over the prefix table
  EDT_HD void enter(uint32_t r, bool first) {
    const int lo = rng_lo(r), hi = rng_hi(r);
    sbase = band * L->BL;
    const int i = first ? lo : hi;
    before = i - lo;
    after = hi - i;
    ep = L->st + (sbase + i) * L->W;
  }
  EDT_HD Vtx<Tr> cur() const { return unpack<Tr>(*ep, L->SB, sbase); }
  EDT_HD void next() {  // v + 1 < T
    if (after > 0) { ep += L->W; --after; ++before; ++v; return; }
    uint32_t r;
    do { ++band; r = L->r(band); } while (rng_lo(r) > rng_hi(r));
    enter(r, true);
    ++v;
  }
  EDT_HD void prev() {  // v > 0
    if (before > 0) { ep -= L->W; --before; ++after; --v; return; }
    uint32_t r;
    do { --band; r = L->r(band); } while (rng_lo(r) > rng_hi(r));
    enter(r, false);
    --v;
  }
};

// emit(d, prep(payload), value) is called once per site in increasing order, d = u - s for a minimising site s and
// value = d^2 + g_of(payload) (int32); the caller keeps its own output cursor.  hint = virtual index to start the search for the
// tangent vertex of u0 from (the junction of the caller's own band: the nearest obstacle is usually close).
// prep(payload) is evaluated once per vertex (not per site); its result is what emit receives as second argument.
template <class Tr, class Prep, class Emit>
EDT_HD void hull_fill(const LineHull<Tr>& L, const uint16_t* off, int NB, int T, int u0, int u1, int hint, Prep&& prep, Emit&& emit) {
  using X = typename Tr::X;
  HullCursor<Tr> cu;
  cu.L = &L; cu.off = off; cu.NB = NB; cu.T = T;
  const X two_u0 = (X)(2 * u0);
  // tangent vertex of u0 = first k whose successor is not strictly better:  val(k+1) < val(k)  <=>  dc < 2 u0 ds.
  // A few steps from the hint, else bisection over the side that is left.
  int lo = 0, hi = T - 1;
  cu.seek(hint < T ? hint : T - 1);
  Vtx<Tr> k = cu.cur();
  {
    int steps = 0;
    bool moved = false, resolved = false;
    while (cu.v + 1 < T) {  // right while the successor is strictly better
      HullCursor<Tr> nx = cu;
      nx.next();
      const Vtx<Tr> b = nx.cur();
      if (!((X)(b.c - k.c) < two_u0 * (X)(b.s - k.s))) { resolved = true; break; }
      cu = nx; k = b; moved = true;
      if (++steps == 6) break;
    }
    if (cu.v + 1 >= T) resolved = true;
    if (!resolved) lo = cu.v;                 // successor still better: the answer is at or right of here
    else if (moved) lo = hi = cu.v;
    else {                                    // left while the predecessor is at least as good
      resolved = false;
      steps = 0;
      while (cu.v > 0) {
        HullCursor<Tr> pv = cu;
        pv.prev();
        const Vtx<Tr> a = pv.cur();
        if ((X)(k.c - a.c) < two_u0 * (X)(k.s - a.s)) { resolved = true; break; }
        cu = pv; k = a;
        if (++steps == 6) break;
      }
      if (cu.v == 0) resolved = true;
      if (resolved) lo = hi = cu.v; else hi = cu.v;
    }
  }
  if (lo < hi) {
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      cu.seek(mid);
      const Vtx<Tr> a = cu.cur();
      cu.next();
      const Vtx<Tr> b = cu.cur();
      if ((X)(b.c - a.c) < two_u0 * (X)(b.s - a.s)) lo = mid + 1; else hi = mid;
    }
    cu.seek(lo);
    k = cu.cur();
  }
  Vtx<Tr> n = k;
  int left = T - 1 - cu.v;  // vertices after the current one
  // the successor n takes over at the first u with  n.c - k.c <= 2 u (n.s - k.s):  acc(u) = 2 u ds - dc >= 0,
  // acc(u + 1) = acc(u) + 2 ds; with no successor acc stays negative
  X acc = -1, step = 0;
  if (left > 0) {
    cu.next(); n = cu.cur();
    step = (X)(2 * (n.s - k.s));
    acc = step * (X)u0 - (X)(n.c - k.c);
  }
  int d = u0 - k.s;
  auto kp = prep(k.p);
  int32_t g = Tr::g_of(k.p);
  for (int u = u0; u < u1; ++u) {
    while (acc >= 0) {
      const int ks = n.s;
      const int32_t kc = n.c;
      kp = prep(n.p);
      --left;
      d = u - ks;
      g = Tr::g_of(n.p);
      if (left > 0) {
        cu.next(); n = cu.cur();
        step = (X)(2 * (n.s - ks));
        acc = step * (X)u - (X)(n.c - kc);
      } else {
        acc = -1; step = 0;
      }
    }
    emit(d, kp, d * d + g);
    ++d;
    acc += step;
  }
}

// ---------------------------------------------------------------------------------------------------
// z distance from the bit words of a column (z contiguous: bit z & 31 of word z >> 5).
//   column_outside(): nearest occupied z strictly below / above word wi, looking at the other words of the column and
//     at the z-slab halo (below_gap / above_gap: distance from the slab's first / last plane to the nearest occupied
//     voxel in the slabs below / above, <= 0 when none); packed as two int32 "far" sentinels when there is none.
//   site_dz(): distance of bit `bi` of the own word to the nearest occupied voxel of the column; >= kDzNone = none.
// ---------------------------------------------------------------------------------------------------
constexpr int kDzNone = 23171;  // dz^2 >= 2^29 = kDistInf: as far as the transform is concerned there is no obstacle
constexpr int kFar = 1 << 20;
EDT_HD int edt_clz(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __clz((int)v);
#else
  return __builtin_clz(v);
#endif
}
EDT_HD int edt_ffs(uint32_t v) {
#ifdef __CUDA_ARCH__
  return __ffs((int)v);
#else
  return __builtin_ffs((int)v);
#endif
}
EDT_HD void column_outside(const uint32_t* col, int wz, int Nz, int wi, int below_gap, int above_gap, int* below, int* above) {
  int lo = -kFar, hi = kFar;
  if (below_gap > 0) lo = -below_gap;
  if (above_gap > 0) hi = Nz - 1 + above_gap;
  for (int j = wi - 1; j >= 0; --j) {
    const uint32_t v = col[j];
    if (v) { lo = (j << 5) + 31 - edt_clz(v); break; }
  }
  for (int j = wi + 1; j < wz; ++j) {
    const uint32_t v = col[j];
    if (v) { hi = (j << 5) + edt_ffs(v) - 1; break; }
  }
  *below = lo; *above = hi;
}
EDT_HD int site_dz(uint32_t own, int bi, int z, int below, int above) {
  const uint32_t lowm = own & (0xffffffffu >> (31 - bi));  // bits <= bi
  const uint32_t highm = own & (0xffffffffu << bi);        // bits >= bi
  const int dn = lowm ? bi - (31 - edt_clz(lowm)) : z - below;
  const int up = highm ? edt_ffs(highm) - 1 - bi : above - z;
  return dn < up ? dn : up;
}

}  // namespace edtf
}  // namespace b200
