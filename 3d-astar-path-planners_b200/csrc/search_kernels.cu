// search_kernels.cu — batched independent path queries, one query per thread block (persistent blocks
// pull query ids from a global counter).
//
// Behaviour restated from the reference (not ported): src/Planners/AStar.cpp:80-179 (findPath),
// :56-78 (ComputeCost/UpdateVertex), src/Planners/ThetaStar.cpp:37-89, src/utils/LineOfSight.cpp:9-27,
// src/Planners/AlgorithmBase.cpp:123-212 (result object, retrievePath), src/utils/geometry_utils.cpp:10-20.
// Semantics are the CANONICAL ones of DESIGN.md §"Search semantics" (SURVEY Appendix B): node index =
// creation order, open list = exact priority queue on (f, creation index) with true decrease-key, guarded
// relaxations, hash keyed on the exact FP64 triple.  oracle/oracle.cpp (mode CANONICAL) is the checker.
//
// Mapping to the machine:
//   * warp 0 of the block is the control warp: lane L < 27 owns neighbour offset L of the current
//     expansion (same dx->dy->dz order as AStar.cpp:131-133), so the 26 isInMap tests, exact-key hash
//     probes, occupancy bit reads and relaxations of one expansion run in one warp-wide pass; node
//     creation order is recovered with a ballot prefix so pool indices match the sequential loop.
//   * open list: 32-ary implicit heap in the slot's HBM workspace (L1/L2 resident).  A pop reads each
//     level as one aligned 512 B row (one entry per lane) and finds the minimum with three REDUX ops.
//   * Theta*: the <= 26 fastLOS segments of an expansion are spread over all warps of the block,
//     warp-per-segment, lanes = consecutive 0.05 m samples, __ballot_sync early-out (common.cuh).
//   * FP64 everywhere positions are touched (voxel registration is rounding-dependent).
#include "kernels.h"
#include <cub/device/device_radix_sort.cuh>
#include "rbtree.cuh"

namespace b200 {

enum { NODE_NEW = 0, NODE_OPEN = 1, NODE_CLOSED = 2 };

__device__ __forceinline__ unsigned long long fkey(double f) {
  unsigned long long u = (unsigned long long)__double_as_longlong(f);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ULL);
}
__device__ __forceinline__ bool key_less(unsigned long long ka, uint32_t ia, unsigned long long kb, uint32_t ib) {
  return ka < kb || (ka == kb && ia < ib);
}
__device__ __forceinline__ uint32_t hash3(double x, double y, double z) {
  // -0.0 + 0.0 == +0.0: keys that compare equal under `==` hash equally (utils.hpp:63-92 semantics)
  unsigned long long a = (unsigned long long)__double_as_longlong(x + 0.0);
  unsigned long long b = (unsigned long long)__double_as_longlong(y + 0.0);
  unsigned long long c = (unsigned long long)__double_as_longlong(z + 0.0);
  unsigned long long h = a * 0x9E3779B97F4A7C15ULL;
  h ^= (b + 0x9E3779B97F4A7C15ULL + (h << 6) + (h >> 2)) * 0xC2B2AE3D27D4EB4FULL;
  h ^= (c + 0x165667B19E3779F9ULL + (h << 6) + (h >> 2)) * 0xFF51AFD7ED558CCDULL;
  h ^= h >> 33;
  h *= 0xC4CEB9FE1A85EC53ULL;
  h ^= h >> 29;
  return (uint32_t)h;
}

struct Slot {
  NodeRec* nodes;
  uint32_t* hash;
  HeapEnt* heap;  // physical index = logical + 31 (children of k: physical 32k+32 .. 32k+63)
  HeapEnt* top;   // shared memory: logical entries 0..32 (root + its 32 children) live here, not in `heap`
  uint32_t hmask;
  int heap_n;
  int use;
};
// The two top levels of the heap are the ancestors of every entry: keeping them in shared memory makes every
// sift-up step of a heap of <= 1057 entries, and the first level of every pop, an on-chip access instead of a
// dependent global-memory round trip (the ncu source view charged 15-20 % of all stall samples to those loads).
constexpr int kHeapTop = 33;
__device__ __forceinline__ HeapEnt heap_ld(const Slot& s, int k) { return k < kHeapTop ? s.top[k] : s.heap[k + 31]; }
__device__ __forceinline__ void heap_st(const Slot& s, int k, const HeapEnt& e) {
  if (k < kHeapTop) s.top[k] = e;
  else s.heap[k + 31] = e;
}

// ---- 32-ary heap, all 32 lanes of the control warp call these with warp-uniform arguments ----------
__device__ __forceinline__ void heap_place(Slot& s, int k, unsigned long long key, uint32_t idx, int lane) {
  if (lane == 0) {
    HeapEnt e;
    e.key = key; e.idx = idx; e.pad = 0;
    heap_st(s, k, e);
    s.nodes[idx].hpos = k;
  }
}
__device__ __forceinline__ void heap_sift_up(Slot& s, int k, unsigned long long key, uint32_t idx, int lane) {
  while (k > 0) {
    const int par = (k - 1) >> 5;
    const HeapEnt pe = heap_ld(s, par);
    if (!key_less(key, idx, pe.key, pe.idx)) break;
    heap_place(s, k, pe.key, pe.idx, lane);
    k = par;
  }
  heap_place(s, k, key, idx, lane);
  __syncwarp();
}
__device__ __forceinline__ void heap_push(Slot& s, unsigned long long key, uint32_t idx, int lane) {
  const int k = s.heap_n++;
  heap_sift_up(s, k, key, idx, lane);
}
// One expansion's open-list updates (<= 26, one per lane): `improved` lanes either decrease the key of the entry at
// `hpos` (>= 0, read before any of this expansion's updates) or append a new entry.
// The serial version costs two dependent global-memory round trips per update (the ncu source view put 20 % of the
// kernel's stall samples there).  Here every lane reads its parent entry at once; an update whose key is not below
// its parent's is written in place — that can never break the heap property, whatever the other lanes do: an
// entry's key only ever decreases, and appended leaves have old entries as parents once the heap is non-empty —
// and only the few that must move up are replayed one at a time with fresh reads.
__device__ __forceinline__ void heap_update_batch(Slot& s, bool improved, int hpos, unsigned long long key, uint32_t idx, int lane) {
  const unsigned im = __ballot_sync(0xffffffffu, improved);
  if (!im) return;
  const int n0 = s.heap_n;
  const bool is_new = improved && hpos < 0;
  const unsigned newmask = __ballot_sync(0xffffffffu, is_new);
  const int k = is_new ? n0 + __popc(newmask & ((1u << lane) - 1u)) : hpos;
  bool stay = false;
  if (improved && n0 > 0) {
    if (k == 0) {
      stay = true;
    } else {
      const HeapEnt pe = heap_ld(s, (k - 1) >> 5);
      stay = !key_less(key, idx, pe.key, pe.idx);
    }
  }
  s.heap_n = n0 + __popc(newmask);
  __syncwarp();
  if (stay) {
    HeapEnt e;
    e.key = key; e.idx = idx; e.pad = 0;
    heap_st(s, k, e);
    if (is_new) s.nodes[idx].hpos = k;
  }
  __syncwarp();
  unsigned mv = __ballot_sync(0xffffffffu, improved && !stay);
  while (mv) {
    const int b = __ffs(mv) - 1;
    mv &= mv - 1;
    const uint32_t bi = __shfl_sync(0xffffffffu, idx, b);
    const unsigned long long bk = __shfl_sync(0xffffffffu, key, b);
    const int bnew = __shfl_sync(0xffffffffu, (int)is_new, b);
    const int bk0 = __shfl_sync(0xffffffffu, k, b);
    // an appended entry starts from its reserved leaf; an existing one from wherever earlier moves left it
    const int from = bnew ? bk0 : s.nodes[bi].hpos;
    heap_sift_up(s, from, bk, bi, lane);
  }
}
// removes the root; returns its node index
__device__ __forceinline__ uint32_t heap_pop(Slot& s, int lane, const HeapEnt top) {
  const int n = --s.heap_n;
  if (lane == 0) s.nodes[top.idx].hpos = -1;
  if (n > 0) {
    const HeapEnt last = heap_ld(s, n);
    int k = 0;
    for (;;) {
      const int c0 = 32 * k + 1;
      if (c0 >= n) break;
      const int ci = c0 + lane;
      const bool valid = ci < n;
      HeapEnt ce;
      ce.key = ~0ULL; ce.idx = 0xffffffffu; ce.pad = 0;
      if (valid) ce = heap_ld(s, ci);
      const uint32_t hi = valid ? (uint32_t)(ce.key >> 32) : 0xffffffffu;
      const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
      const uint32_t lo = (valid && hi == mhi) ? (uint32_t)ce.key : 0xffffffffu;
      const uint32_t mlo = __reduce_min_sync(0xffffffffu, lo);
      const uint32_t id = (valid && hi == mhi && lo == mlo) ? ce.idx : 0xffffffffu;
      const uint32_t mid = __reduce_min_sync(0xffffffffu, id);
      const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
      if (!key_less(mkey, mid, last.key, last.idx)) break;
      const int wl = __ffs(__ballot_sync(0xffffffffu, valid && id == mid)) - 1;
      heap_place(s, k, mkey, mid, lane);
      k = c0 + wl;
    }
    heap_place(s, k, last.key, last.idx, lane);
  }
  __syncwarp();
  return top.idx;
}

// exact-key probe: returns node index or -1
__device__ __forceinline__ int hash_find(const Slot& s, double x, double y, double z) {
  uint32_t h = hash3(x, y, z) & s.hmask;
  for (;;) {
    const uint32_t e = s.hash[h];
    if (e == 0) return -1;
    const NodeRec* n = s.nodes + (e - 1);
    if (n->px == x && n->py == y && n->pz == z) return (int)(e - 1);
    h = (h + 1) & s.hmask;
  }
}
// same probe; on a hit also returns the fields UpdateVertex needs: state/hpos share the key's 32-byte sector and g
// is requested at once, so its round trip overlaps the rest of the neighbour phase
__device__ __forceinline__ int hash_find_ex(const Slot& s, double x, double y, double z, int& state, int& hpos, double& g) {
  uint32_t h = hash3(x, y, z) & s.hmask;
  for (;;) {
    const uint32_t e = s.hash[h];
    if (e == 0) return -1;
    const NodeRec* n = s.nodes + (e - 1);
    if (n->px == x && n->py == y && n->pz == z) {
      state = n->state; hpos = n->hpos; g = n->g;
      return (int)(e - 1);
    }
    h = (h + 1) & s.hmask;
  }
}
__device__ __forceinline__ uint32_t hash_insert(const Slot& s, double x, double y, double z, int idx) {
  uint32_t h = hash3(x, y, z) & s.hmask;
  for (;;) {
    if (atomicCAS(s.hash + h, 0u, (uint32_t)idx + 1u) == 0u) return h;
    h = (h + 1) & s.hmask;
  }
}

// In-kernel promotion (32-thread kernels): a query that outgrows its tier-0 workspace moves — nodes, open list, a rebuilt
// hash table — into one of the big promotion slots and simply goes on; registers and the shared-memory heap top carry the
// rest of its state, so nothing is recomputed.  (The alternative, still used by the 128-thread and faithful kernels and when
// the promotion slots run out, re-runs the query from scratch in a second launch: the BASELINE C3 batch of 10 000
// costlazythetastar queries spent 300 of its 600 ms re-running its one 36 000-node query.)  Returns the promotion slot, -1
// when none is left.  Called after the search loop has been left with ST_TIER_OVERFLOW (the interrupted expansion had no side
// effects yet and is redone), so the loop itself only carries one extra warp-uniform flag.
__device__ __forceinline__ int promote_slot(const NodeRec* nodes, uint32_t* hash, const HeapEnt* heap, int use, int heap_n, NodeRec* promo_nodes,
                                         uint32_t* promo_hash, HeapEnt* promo_heap, int promo_cap, uint32_t promo_hcap,
                                         uint32_t promo_open_entries, int promo_slots, int* promo_count, int lane) {
  int t = 0;
  if (lane == 0) t = atomicAdd(promo_count, 1);
  t = __shfl_sync(0xffffffffu, t, 0);
  if (t >= promo_slots) return -1;
  NodeRec* n1 = promo_nodes + (size_t)t * promo_cap;
  uint32_t* h1 = promo_hash + (size_t)t * promo_hcap;
  HeapEnt* hp1 = promo_heap + (size_t)t * promo_open_entries;
  for (int i = lane; i < use; i += 32) hash[nodes[i].slot] = 0u;  // leave the tier-0 table clean for the block's next query
  const uint4* src = reinterpret_cast<const uint4*>(nodes);
  uint4* dst = reinterpret_cast<uint4*>(n1);
  for (int i = lane; i < use * 4; i += 32) dst[i] = src[i];
  const int nh = heap_n + 32;  // physical entries: logical k >= kHeapTop lives at k + 31
  for (int i = lane; i < nh; i += 32) hp1[i] = heap[i];
  __syncwarp();
  Slot s1;
  s1.nodes = n1; s1.hash = h1; s1.heap = hp1; s1.top = nullptr; s1.hmask = promo_hcap - 1u; s1.heap_n = heap_n; s1.use = use;
  for (int i = lane; i < use; i += 32) n1[i].slot = hash_insert(s1, n1[i].px, n1[i].py, n1[i].pz, i);
  __syncwarp();
  return t;
}

struct Shared {
  long long q;
  int cap;         // physical node capacity of the workspace the current query lives in (tier 0, or a promotion slot)
  int state;       // 0 continue, 1 query finished (written by the control warp before the first barrier of an expansion)
  int stop2;       // 1: query finished in UpdateVertex.  A separate flag: the other warps may still be reading `state`
                   // when the control warp gets here (no barrier in between when the LoS phase is skipped)
  int used;        // nodes the finished query created (hash-clearing pass); not `state`: a warp that is late reading
                   // state == 1 at the loop exit must not see this value instead
  int los_phase;   // this expansion needs the block-wide LoS phase
  unsigned valid;  // neighbour lanes that reach UpdateVertex
  double ppx, ppy, ppz;
  double nbx[27], nby[27], nbz[27];
  int nsamp[27];
  unsigned char vis[27];
};

// cost-aware edge (ours, DESIGN.md "Cost-aware variants"): len + w * (len * ((qa + qb) / 131070))
__device__ __forceinline__ double costq_at(const MapDev& m, double x, double y, double z) {
  const int ix = min(max(pos_to_idx1(x, m.ox, m.inv_res), 0), m.Nx - 1);
  const int iy = min(max(pos_to_idx1(y, m.oy, m.inv_res), 0), m.Ny - 1);
  const int iz = min(max(pos_to_idx1(z, m.oz, m.inv_res), 0), m.Nz - 1);
  return (double)m.costq[vox_addr(m, ix, iy, iz)];
}
template <bool COST>
__device__ __forceinline__ double edge_cost(const SearchParams& P, double ax, double ay, double az, double bx, double by, double bz,
                                            double len) {
  if (!COST) return len;
  const double qa = costq_at(P.map, ax, ay, az), qb = costq_at(P.map, bx, by, bz);
  return dadd(len, dmul(P.cost_weight, dmul(len, __ddiv_rn(dadd(qa, qb), 131070.0))));
}

// ---- thetastaragr / thetastaragrpp (the fork's air-ground planners) -----------------------------------------
constexpr int kMotionBit = 1 << 16;
// start node: ThetaStarAGR.cpp:141-155 / ThetaStarAGRpp.cpp:146-155 (f was computed from the UNclamped position)
template <bool AGR>
__device__ __forceinline__ void agr_init_start(const SearchParams& P, NodeRec& n) {
  if (n.pz < 0) n.pz = 0;
  if (n.pz >= P.agr_ground_judge) {
    if (AGR) {
      const double pen = dadd(__ddiv_rn(dmul(P.agr_flying_cost, n.pz), P.agr_barrier), P.agr_flying_cost_default);
      n.g = dadd(n.g, pen);
      n.f = dadd(n.f, pen);
      n.penalty = pen;
    }
    n.state |= kMotionBit;
  } else if (AGR) {
    n.penalty = 0.0;
  }
}
__device__ __forceinline__ double std_max(double a, double b) { return (a < b) ? b : a; }  // std::max: NaN in `a` survives
struct AgrCand { double g, pen; int parent; bool motion; };
// ThetaStarAGR::ComputeCost (ThetaStarAGR.cpp:49-103): candidate the guard `< neighbor->g_score` is applied to
__device__ __forceinline__ AgrCand agr_candidate(const SearchParams& P, double cg, double cpen, int cur, int cpar, bool vis, double pg,
                                                 double ppen, double pdist, double dnorm, double nz) {
  AgrCand c;
  c.motion = nz >= P.agr_ground_judge;
  c.pen = c.motion ? dadd(__ddiv_rn(dmul(P.agr_flying_cost, nz), P.agr_barrier), P.agr_flying_cost_default) : 0.0;
  double t = dsub(dadd(cg, dnorm), cpen);
  if (c.motion) t = dadd(t, c.pen);
  c.g = t; c.parent = cur;
  if (cpar >= 0 && vis) {
    double u = dsub(dadd(pg, pdist), ppen);
    if (c.motion) u = dadd(u, c.pen);
    c.g = u; c.parent = cpar;
  }
  return c;
}
// ThetaStarAGRpp::ComputeCost (ThetaStarAGRpp.cpp:52-109)
__device__ __forceinline__ AgrCand agrpp_candidate(const SearchParams& P, double cg, double cz, bool cmotion, int cur, int cpar, bool vis,
                                                   double pg, double pz, bool pmotion, double pdist, double dnorm, double nz) {
  AgrCand c;
  c.motion = nz >= P.agr_ground_judge;
  c.pen = 0.0;
  const double alpha = std_max(dadd(1.0, __ddiv_rn(dmul(P.agr_flying_cost, dsub(nz, cz)), nz)), P.agr_epsilon);
  double t = dadd(cg, dmul(dnorm, (c.motion || cmotion) ? alpha : P.agr_epsilon));
  if (!cmotion && c.motion) t = dadd(t, P.agr_flying_cost_default);
  c.g = t; c.parent = cur;
  if (cpar >= 0 && vis) {
    const double al = std_max(dadd(1.0, __ddiv_rn(dmul(P.agr_flying_cost, dsub(nz, pz)), nz)), P.agr_epsilon);
    double u = dadd(pg, dmul(pdist, (c.motion || pmotion) ? al : P.agr_epsilon));
    if (!pmotion && c.motion) u = dadd(u, P.agr_flying_cost_default);
    if (u < c.g) { c.g = u; c.parent = cpar; }
  }
  return c;
}

// resident blocks per SM the register allocation is capped for.  Measured (1xB200, C2 world, queries/s): the
// 32-thread kernels at 28 blocks (<= 72 registers, a few spills) — Lazy Theta* 593k -> 723k, A* 992k -> 983k; at 32
// blocks (64 registers) Lazy Theta* gains another 6 % (700k -> 746k) while A* loses 4 %, so only the lazy variants use 32;
// the 128-thread LoS kernels at 6 blocks (<= 80 registers) — Theta* 79.7k -> 87.5k, 8 blocks no better.
#ifndef SEARCH_MINB1
#define SEARCH_MINB1 28
#endif
#ifndef SEARCH_MINB4
#define SEARCH_MINB4 6
#endif
#ifndef SEARCH_MINB_LAZY
#define SEARCH_MINB_LAZY 32
#endif
template <int ALGO, int NW>
__global__ void __launch_bounds__(NW * 32, NW != 1 ? SEARCH_MINB4
                                           : (ALGO == ALGO_LAZYTHETASTAR || ALGO == ALGO_COSTLAZYTHETASTAR) ? SEARCH_MINB_LAZY : SEARCH_MINB1)
search_kernel(const SearchParams P) {
  constexpr bool IS_AGR = ALGO == ALGO_THETASTARAGR, IS_AGRPP = ALGO == ALGO_THETASTARAGRPP;
  constexpr bool IS_THETA = ALGO == ALGO_THETASTAR || IS_AGR || IS_AGRPP;  // one fastLOS per generated neighbour
  constexpr bool IS_LAZY = ALGO == ALGO_LAZYTHETASTAR || ALGO == ALGO_COSTLAZYTHETASTAR;
  constexpr bool IS_COST = ALGO == ALGO_COSTASTAR || ALGO == ALGO_COSTLAZYTHETASTAR;
  __shared__ Shared sh;
  __shared__ HeapEnt heap_top[kHeapTop];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MapDev& m = P.map;
  Slot s;
  s.nodes = P.nodes + (size_t)blockIdx.x * P.cap;
  s.hash = P.hash + (size_t)blockIdx.x * P.hcap;
  s.heap = P.heap + (size_t)blockIdx.x * P.open_entries;
  s.top = heap_top;
  s.hmask = P.hcap - 1;
  if (threadIdx.x == 0) sh.cap = P.cap;

  // neighbour offset of this lane: L = 9a + 3b + c <-> (dx,dy,dz) = (dv[a],dv[b],dv[c])  (AStar.cpp:131-133)
  const int la = lane / 9, lb = (lane / 3) % 3, lc = lane % 3;
  const double ddx = P.dv[la % 3], ddy = P.dv[lb], ddz = P.dv[lc];
  const double dnorm = norm3(ddx, ddy, ddz);
  const bool lane_nb = lane < 27 && !(dnorm < 1e-3);  // AStar.cpp:136

  for (;;) {
    if (tid == 0) sh.q = (long long)atomicAdd(P.q_counter, 1ULL);
    __syncthreads();
    const long long qi = sh.q;
    if (qi >= P.nq) break;
    const long long q = P.qlist ? (long long)P.qlist[qi] : qi;
    const double sx = P.starts[3 * q], sy = P.starts[3 * q + 1], sz = P.starts[3 * q + 2];
    const double gx = P.goals[3 * q], gy = P.goals[3 * q + 1], gz = P.goals[3 * q + 2];
    unsigned long long t_begin = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));

    // ---------------- initialise (AStar.cpp:87-97) ----------------
    s.heap_n = 0;
    s.use = 1;
    long long iter_num = 0, los_checks = 0, los_samples = 0;
    int status = ST_OPEN_EMPTY, last = 0, term = -1, open_peak = 1;
    if (warp == 0) {
      if (lane == 0) {
        NodeRec n;
        n.px = sx; n.py = sy; n.pz = sz;
        n.g = 0.0;
        n.f = dmul(P.lambda, dmul(P.tie, norm3(dsub(gx, sx), dsub(gy, sy), dsub(gz, sz))));
        n.parent = -1; n.hpos = -1; n.via = -1; n.state = NODE_OPEN; n.pad = 0;
        if (IS_AGR || IS_AGRPP) agr_init_start<IS_AGR>(P, n);
        n.slot = hash_insert(s, n.px, n.py, n.pz, 0);
        s.nodes[0] = n;
      }
      __syncwarp();
      heap_push(s, fkey(s.nodes[0].f), 0u, lane);
      sh.state = 0;
      sh.stop2 = 0;
    }

    // ---------------- search loop (AStar.cpp:103-173) ----------------
    int cur = 0;
    int bypass = -1;  // node that skips the heap: see the push-then-pop bypass below
    double cx = 0, cy = 0, cz = 0, cg = 0, cpen = 0;
    bool cmotion = false;
    int cpar = -1;
    // per-lane neighbour state, live between the phases of one expansion
    int nb_idx = -1, nb_state = NODE_NEW, nb_hpos = -1;
    double nb_g = 0;
    double nx = 0, ny = 0, nz = 0;
    bool redo = false;  // re-entering after a promotion: the interrupted expansion of `cur` is repeated from its neighbour phase
    for (;;) {  // left once, or once more per promotion
    for (;;) {
      bool valid = false;
      if (warp == 0) {
        bool done = false;
        if (redo) {
        } else if (s.heap_n == 0 && bypass < 0) {  // AStar.cpp:175-178
          status = ST_OPEN_EMPTY; last = cur; done = true;
        } else {
          open_peak = max(open_peak, s.heap_n + (bypass >= 0 ? 1 : 0));
          NodeRec c;
          if (bypass >= 0) {  // the previous expansion's best new node is the open list's minimum: it never entered the heap
            cur = bypass;
            bypass = -1;
            c = s.nodes[cur];
          } else {
            const HeapEnt top = s.top[0];
            c = s.nodes[top.idx];  // in flight while the root is being replaced
            cur = (int)heap_pop(s, lane, top);
          }
          if (IS_LAZY && c.parent >= 0 && c.parent != c.via) {
            // SetVertex: verify the optimistic parent with one fastLOS; fall back to the best closed neighbour
            const NodeRec pr = s.nodes[c.parent];
            const double len = norm3(dsub(pr.px, c.px), dsub(pr.py, c.py), dsub(pr.pz, c.pz));
            ++los_checks;
            bool vis = false;
            int ns = 0;
            unsigned long long sumq = 0;
            if (!(P.max_los > 0 && len > P.max_los)) vis = los_warp<IS_COST>(m, pr.px, pr.py, pr.pz, c.px, c.py, c.pz, lane, &ns, &sumq);
            los_samples += ns;
            bool changed = false;
            if (vis) {
              if (IS_COST) {
                c.g = dadd(pr.g, dadd(len, dmul(P.cost_weight, dmul(kLosStep, __ddiv_rn((double)sumq, 65535.0)))));
                changed = true;
              }
            } else {
              // candidates: via first, then the 26 exact-key lookups in expansion order; strict < replaces
              const NodeRec vn = s.nodes[c.via];
              const int vcode = (c.state >> 8) & 0xff;
              const double vstep = norm3(P.dv[(vcode / 9) % 3], P.dv[(vcode / 3) % 3], P.dv[vcode % 3]);
              double best_g = dadd(vn.g, edge_cost<IS_COST>(P, vn.px, vn.py, vn.pz, c.px, c.py, c.pz, vstep));
              int best = c.via;
              double cand = __longlong_as_double(0x7ff0000000000000LL);
              int cand_idx = -1;
              if (lane_nb) {
                const double qx = dadd(c.px, ddx), qy = dadd(c.py, ddy), qz = dadd(c.pz, ddz);
                const int fi = hash_find(s, qx, qy, qz);
                if (fi >= 0 && (s.nodes[fi].state & 0xff) == NODE_CLOSED) {
                  const NodeRec fn = s.nodes[fi];
                  cand = dadd(fn.g, edge_cost<IS_COST>(P, fn.px, fn.py, fn.pz, c.px, c.py, c.pz, dnorm));
                  cand_idx = fi;
                }
              }
              // sequential-order argmin: smallest value, first lane on ties, and only if strictly below via's
              const unsigned long long ck = cand_idx >= 0 ? fkey(cand) : ~0ULL;
              const uint32_t hi = (uint32_t)(ck >> 32);
              const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
              const uint32_t lo = hi == mhi ? (uint32_t)ck : 0xffffffffu;
              const uint32_t mlo = __reduce_min_sync(0xffffffffu, lo);
              const unsigned wm = __ballot_sync(0xffffffffu, cand_idx >= 0 && hi == mhi && lo == mlo);
              if (wm) {
                const int wl = __ffs(wm) - 1;
                const double wg = __shfl_sync(0xffffffffu, cand, wl);
                const int wi = __shfl_sync(0xffffffffu, cand_idx, wl);
                if (wg < best_g) { best_g = wg; best = wi; }
              }
              c.parent = best;
              c.g = best_g;
              changed = true;
            }
            if (changed) {
              c.f = dadd(c.g, dmul(P.lambda, dmul(P.tie, norm3(dsub(gx, c.px), dsub(gy, c.py), dsub(gz, c.pz)))));
              if (lane == 0) {
                s.nodes[cur].g = c.g; s.nodes[cur].f = c.f; s.nodes[cur].parent = c.parent;
              }
            }
          }
          cx = c.px; cy = c.py; cz = c.pz; cg = c.g; cpar = c.parent;
          if (IS_AGR) cpen = c.penalty;
          if (IS_AGR || IS_AGRPP) cmotion = (c.state & kMotionBit) != 0;
          // termination (AStar.cpp:110-119)
          if (fabs(dsub(cx, gx)) <= P.r && fabs(dsub(cy, gy)) <= P.r && fabs(dsub(cz, gz)) <= P.r) {
            status = ST_SOLVED; last = cur; term = cur; done = true;
          }
        }
        if (!done) {
          if (!redo) {
            if (lane == 0) s.nodes[cur].state = (s.nodes[cur].state & ~0xff) | NODE_CLOSED;  // AStar.cpp:122
            ++iter_num;
          }
          redo = false;
          __syncwarp();
          // ---- neighbour resolution (AStar.cpp:131-169), one lane per offset ----
          nb_idx = -1; nb_state = NODE_NEW; nb_hpos = -1;
          nb_g = __longlong_as_double(0x7ff0000000000000LL);  // what a node created below holds
          valid = lane_nb;
          nx = dadd(cx, ddx); ny = dadd(cy, ddy); nz = dadd(cz, ddz);            // :138
          if (valid && !is_in_map(m, nx, ny, nz)) valid = false;                 // :141
          int occ = 0;
          if (valid) occ = inflate_occ(m, nx, ny, nz);  // issued before the hash probe (independent loads overlap), used at :151
          if (valid) {
            nb_idx = hash_find_ex(s, nx, ny, nz, nb_state, nb_hpos, nb_g);       // :146
            if (nb_idx >= 0 && (nb_state & 0xff) == NODE_CLOSED) valid = false;  // :147
            if (IS_AGRPP && nb_idx >= 0) valid = false;  // ThetaStarAGRpp.cpp:208-210: a node is relaxed once, at creation
          }
          if (valid && occ == 1) valid = false;                                  // :151
          const bool need = valid && nb_idx < 0;
          const unsigned cm = __ballot_sync(0xffffffffu, need);
          const int ncreate = __popc(cm);
          const int rank = __popc(cm & ((1u << lane) - 1u));
          bool oom = false, overflow = false;
          int fail_lane = 32;
          if (ncreate > 0) {
            if (sh.cap < P.allocate_num && s.use + ncreate > sh.cap) {
              overflow = true;  // physical tier too small: the query moves into a promotion slot (below, after the loop) or is re-run
            }
            if (!overflow && s.use + ncreate >= P.allocate_num) {  // :165-168: the creation that makes use == allocate_num bails out
              const int fail_rank = P.allocate_num - s.use - 1;
              const unsigned fm = __ballot_sync(0xffffffffu, need && rank == fail_rank);
              fail_lane = __ffs(fm) - 1;
              oom = true;
            }
          }
          if (overflow) {
            status = ST_TIER_OVERFLOW; last = cur; done = true;
          } else {
            if (need && lane <= fail_lane && s.use + rank < sh.cap) {
              const int ni = s.use + rank;
              NodeRec n;
              n.px = nx; n.py = ny; n.pz = nz;
              n.g = __longlong_as_double(0x7ff0000000000000LL);
              n.f = n.g;
              n.parent = -1; n.hpos = -1; n.via = -1; n.state = NODE_NEW; n.pad = 0;
              n.slot = hash_insert(s, nx, ny, nz, ni);
              s.nodes[ni] = n;
              nb_idx = ni;
            }
            if (oom) {
              if (lane >= fail_lane) valid = false;  // the failing neighbour and everything after it never reach UpdateVertex
              s.use = min(P.allocate_num, sh.cap);
              status = ST_POOL_EXHAUSTED; last = cur;
            } else {
              s.use += ncreate;
            }
            __syncwarp();
          }
          if (!done) {
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (IS_THETA) {
              const bool lp = cpar >= 0 && vm != 0;
              if (lp) {
                los_checks += __popc(vm);  // ThetaStar.cpp:50
                if (lane < 27) { sh.nbx[lane] = nx; sh.nby[lane] = ny; sh.nbz[lane] = nz; }
                if (lane == 0) {
                  const NodeRec pr = s.nodes[cpar];
                  sh.ppx = pr.px; sh.ppy = pr.py; sh.ppz = pr.pz;
                  sh.valid = vm;
                }
              }
              if (lane == 0) sh.los_phase = lp ? 1 : 0;
            }
          }
        }
        if (lane == 0) sh.state = done ? 1 : 0;
        if (done) valid = false;
      }
      if (NW > 1) __syncthreads(); else __syncwarp();
      if (sh.state == 1) break;

      // ---- Theta*: block-wide line-of-sight phase (ThetaStar.cpp:49-51) ----
      if (IS_THETA && sh.los_phase) {
        const unsigned vm = sh.valid;
        const double ppx = sh.ppx, ppy = sh.ppy, ppz = sh.ppz;
        for (int seg = warp; seg < 27; seg += NW) {
          if (!((vm >> seg) & 1u)) continue;
          const double ex = sh.nbx[seg], ey = sh.nby[seg], ez = sh.nbz[seg];
          bool vis = false;
          int ns = 0;
          if (P.max_los > 0) {
            const double len = norm3(dsub(ppx, ex), dsub(ppy, ey), dsub(ppz, ez));
            if (len > P.max_los) { if (lane == 0) { sh.vis[seg] = 0; sh.nsamp[seg] = 0; } continue; }
          }
          vis = los_warp<false>(m, ppx, ppy, ppz, ex, ey, ez, lane, &ns, nullptr);
          if (lane == 0) { sh.vis[seg] = vis ? 1 : 0; sh.nsamp[seg] = ns; }
        }
        if (NW > 1) __syncthreads(); else __syncwarp();
      }

      // ---- UpdateVertex (AStar.cpp:65-78 / ThetaStar.cpp:37-89), control warp ----
      if (warp == 0) {
        bool improved = false;
        double new_g = 0, new_f = 0, new_pen = 0;
        bool new_motion = false;
        if (valid) {
          int new_parent = cur;
          bool use_parent = false;
          if (IS_THETA && cpar >= 0) {
            los_samples += sh.nsamp[lane];
            use_parent = sh.vis[lane] != 0;
          }
          if (IS_AGR || IS_AGRPP) {
            double pg = 0, ppen = 0, ppz = 0, pdist = 0;
            bool pmotion = false;
            if (cpar >= 0) {
              const NodeRec pr = s.nodes[cpar];
              pg = pr.g; ppen = pr.penalty; ppz = pr.pz; pmotion = (pr.state & kMotionBit) != 0;
              pdist = norm3(dsub(pr.px, nx), dsub(pr.py, ny), dsub(pr.pz, nz));
            }
            const AgrCand cand = IS_AGR ? agr_candidate(P, cg, cpen, cur, cpar, use_parent, pg, ppen, pdist, dnorm, nz)
                                        : agrpp_candidate(P, cg, cz, cmotion, cur, cpar, use_parent, pg, ppz, pmotion, pdist, dnorm, nz);
            new_g = cand.g; new_parent = cand.parent; new_pen = cand.pen; new_motion = cand.motion;
          } else if (IS_THETA && use_parent) {  // ThetaStar.cpp:51-60
            const NodeRec pr = s.nodes[cpar];
            const double dist = norm3(dsub(pr.px, nx), dsub(pr.py, ny), dsub(pr.pz, nz));
            new_g = dadd(pr.g, dist);
            new_parent = cpar;
          } else if (IS_LAZY && cpar >= 0) {  // optimistic path 2 (Lazy Theta*)
            const NodeRec pr = s.nodes[cpar];
            const double dist = norm3(dsub(pr.px, nx), dsub(pr.py, ny), dsub(pr.pz, nz));
            new_g = dadd(pr.g, edge_cost<IS_COST>(P, pr.px, pr.py, pr.pz, nx, ny, nz, dist));
            new_parent = cpar;
          } else {  // AStar.cpp:57 / ThetaStar.cpp:39-47,61-69
            new_g = dadd(cg, edge_cost<IS_COST>(P, cx, cy, cz, nx, ny, nz, dnorm));
          }
          if (new_g < nb_g) {
            improved = true;
            new_f = dadd(new_g, dmul(P.lambda, dmul(P.tie, norm3(dsub(gx, nx), dsub(gy, ny), dsub(gz, nz)))));
            NodeRec* w = s.nodes + nb_idx;
            w->g = new_g; w->f = new_f; w->parent = new_parent;
            w->state = NODE_OPEN | (lane << 8) | (new_motion ? kMotionBit : 0);
            if (IS_LAZY) w->via = cur;
            if (IS_AGR) w->penalty = new_pen;
          }
        }
        __syncwarp();
        // open-list updates.  (f, index) is a total order, so the pop sequence does not depend on the heap's layout:
        // any valid heap holding the same entries will do.
        // Push-then-pop bypass.  With a weighted heuristic the best node generated by an expansion is usually the open
        // list's new minimum: pushing it (a climb to the root, the longest one) only to pop it next (a full sift-down)
        // is ~20 % of an expansion's instructions.  If the smallest (f, index) among this expansion's updates is a NEW
        // entry and beats the current root, it is handed straight to the next iteration instead — the same node the heap
        // would have returned, since every other update of this expansion and every old entry compares greater.
        bool to_heap = improved;
        if (status != ST_POOL_EXHAUSTED) {
          const unsigned long long k = improved ? fkey(new_f) : ~0ULL;
          const uint32_t hi = (uint32_t)(k >> 32);
          const uint32_t mhi = __reduce_min_sync(0xffffffffu, hi);
          const uint32_t lo = (improved && hi == mhi) ? (uint32_t)k : 0xffffffffu;
          const uint32_t mlo = __reduce_min_sync(0xffffffffu, lo);
          const uint32_t id = (improved && hi == mhi && lo == mlo) ? (uint32_t)nb_idx : 0xffffffffu;
          const uint32_t mid = __reduce_min_sync(0xffffffffu, id);
          if (mid != 0xffffffffu) {
            const bool mine = improved && id == mid;
            const unsigned wm = __ballot_sync(0xffffffffu, mine && nb_hpos < 0);
            const unsigned long long mkey = ((unsigned long long)mhi << 32) | mlo;
            if (wm && (s.heap_n == 0 || key_less(mkey, mid, s.top[0].key, s.top[0].idx))) {
              bypass = (int)mid;
              if (mine) to_heap = false;
            }
          }
        }
        heap_update_batch(s, to_heap, nb_hpos, fkey(new_f), (uint32_t)nb_idx, lane);
        if (status == ST_POOL_EXHAUSTED) {
          if (lane == 0) sh.stop2 = 1;
        }
      }
      if (NW > 1) __syncthreads(); else __syncwarp();
      if (sh.stop2 == 1) break;
    }
      // a 32-thread query that outgrew tier 0: move it into a promotion slot and re-enter the loop (else: reported for a re-run)
      if (!(NW == 1 && status == ST_TIER_OVERFLOW && P.promo_slots > 0 && sh.cap < P.promo_cap)) break;
      const int t = promote_slot(s.nodes, s.hash, s.heap, s.use, s.heap_n, P.promo_nodes, P.promo_hash, P.promo_heap, P.promo_cap, P.promo_hcap,
                                 P.promo_open_entries, P.promo_slots, P.promo_count, lane);
      if (t < 0) break;
      s.nodes = P.promo_nodes + (size_t)t * P.promo_cap;
      s.hash = P.promo_hash + (size_t)t * P.promo_hcap;
      s.heap = P.promo_heap + (size_t)t * P.promo_open_entries;
      s.hmask = P.promo_hcap - 1u;
      if (lane == 0) { sh.cap = P.promo_cap; sh.state = 0; }
      status = ST_OPEN_EMPTY;
      redo = true;
      __syncwarp();
    }

    // ---------------- result (AlgorithmBase.cpp:123-175, :202-212; geometry_utils.cpp:10-20) -------------
    if (warp == 0) {
      if (IS_THETA || IS_LAZY) {
        // los_samples was accumulated per lane; fold it
        long long v = los_samples;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        los_samples = IS_LAZY ? los_samples : v;
      }
      int n_path = 0;
      long long off = -1;
      float plen = 0.f;
      int* chain = reinterpret_cast<int*>(s.heap);  // the open list is dead now: reuse it as scratch
      float* seglen = reinterpret_cast<float*>(s.heap) + (size_t)sh.cap + 8;
      if (status == ST_SOLVED) {
        if (lane == 0) {
          int c = term;
          while (c >= 0) { chain[n_path++] = c; c = s.nodes[c].parent; }
          if (P.path_pool) {
            const unsigned long long o = atomicAdd(P.path_cursor, (unsigned long long)n_path);
            off = (o + (unsigned long long)n_path <= (unsigned long long)P.path_cap) ? (long long)o : -1;
          }
        }
        n_path = __shfl_sync(0xffffffffu, n_path, 0);
        off = __shfl_sync(0xffffffffu, off, 0);
        __syncwarp();
        for (int i = lane; i < n_path; i += 32) {
          const NodeRec a = s.nodes[chain[n_path - 1 - i]];
          if (off >= 0) {
            P.path_pool[3 * (off + i)] = a.px; P.path_pool[3 * (off + i) + 1] = a.py; P.path_pool[3 * (off + i) + 2] = a.pz;
          }
          if (i > 0) {
            const NodeRec b = s.nodes[chain[n_path - i]];
            const double ux = dmul(dsub(a.px, b.px), m.res), uy = dmul(dsub(a.py, b.py), m.res), uz = dmul(dsub(a.pz, b.pz), m.res);
            seglen[i] = __fsqrt_rn((float)dadd(dadd(dmul(ux, ux), dmul(uy, uy)), dmul(uz, uz)));
          }
        }
        __syncwarp();
        if (lane == 0)
          for (int i = 1; i < n_path; ++i) plen = __fadd_rn(plen, seglen[i]);
      }
      if (lane == 0) {
        b200_path_result_t r;
        const NodeRec ln = s.nodes[last];
        r.solved = status == ST_SOLVED ? 1 : 0;
        r.status = status;
        r.n_path = n_path;
        r.open_peak = open_peak;
        r.path_offset = off;
        r.g_final = ln.g;
        r.path_length = (double)plen;
        r.explored_nodes = s.use;
        r.iter_num = iter_num;
        r.los_checks = los_checks;
        r.los_samples = los_samples;
        r.goal_coords[0] = ln.px; r.goal_coords[1] = ln.py; r.goal_coords[2] = ln.pz;
        r.start_coords[0] = sx; r.start_coords[1] = sy; r.start_coords[2] = sz;
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        r.time_us = (double)(t_end - t_begin) * 1e-3;
        P.results[q] = r;
        if (status == ST_TIER_OVERFLOW && P.overflow_list) P.overflow_list[atomicAdd(P.overflow_count, 1)] = (int)q;
      }
      sh.used = s.use;  // broadcast the number of used nodes for the clearing pass
    }
    __syncthreads();
    // reset(): clear only the hash slots this query touched (AlgorithmBase.cpp:17-32 analogue)
    const int used = min(sh.used, sh.cap);
    for (int i = tid; i < used; i += NW * 32) s.hash[s.nodes[i].slot] = 0u;
    __syncthreads();
    if (NW == 1 && sh.cap != P.cap) {  // the query ended in a promotion slot (left clean above): back to the block's own workspace
      s.nodes = P.nodes + (size_t)blockIdx.x * P.cap;
      s.hash = P.hash + (size_t)blockIdx.x * P.hcap;
      s.heap = P.heap + (size_t)blockIdx.x * P.open_entries;
      s.hmask = P.hcap - 1;
      __syncthreads();  // every thread has read sh.cap
      if (tid == 0) sh.cap = P.cap;
    }
  }
}

// ===================================================================================================
// FAITHFUL semantics (astar, thetastar): bit-identical to the reference's own code — open list = the same
// red-black tree std::set builds (rbtree.cuh) with keys mutated in place, A* overwrites parent/g/f
// unconditionally (AStar.cpp:56-63), Theta* erases by the already-mutated key (ThetaStar.cpp:80-83), pops
// do not test the closed flag (AStar.cpp:105-122).  The per-neighbour arithmetic (isInMap, hash probe,
// occupancy, fastLOS, candidate g/f) is still done warp-/block-parallel; only the part whose ORDER the
// reference's result depends on — node writes interleaved with tree operations — is replayed sequentially
// by lane 0 in the reference's dx->dy->dz order.
// ===================================================================================================
struct FShared {
  int nb_idx[27];
  int new_parent[27];
  int flags[27];  // bit0 reaches UpdateVertex, bit1 write node, bit2 improved (g < old g), bit3 was IN_OPEN_SET, bit4 flying
  double new_g[27], new_f[27], new_pen[27];
};

// Shared-memory residency of the tree was measured and rejected (profiles/README.md "faithful"): 96 KB per block
// leaves 2 blocks/SM and HALVES throughput (thetastar 27k -> 14k q/s) — the kernel is bound by the number of
// serial instruction streams in flight, not by load latency.  The stores keep the mechanism (sizes 0 = off).
constexpr uint32_t kFaithfulSmemTree = 0;  // tree nodes (16 B) held in shared memory
constexpr uint32_t kFaithfulSmemKeys = 0;  // comparator keys (8 B) held in shared memory
constexpr size_t kFaithfulSmemBytes = kFaithfulSmemTree * sizeof(TreeNode) + kFaithfulSmemKeys * sizeof(double);

template <int ALGO, int NW>
__global__ void __launch_bounds__(NW * 32) search_faithful_kernel(const SearchParams P) {
  constexpr bool IS_AGR = ALGO == ALGO_THETASTARAGR;
  constexpr bool IS_THETA = ALGO == ALGO_THETASTAR || IS_AGR;
  __shared__ Shared sh;
  __shared__ FShared fs;
  extern __shared__ __align__(16) unsigned char fsmem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const MapDev& m = P.map;
  Slot s;
  s.nodes = P.nodes + (size_t)blockIdx.x * P.cap;
  s.hash = P.hash + (size_t)blockIdx.x * P.hcap;
  s.heap = P.heap + (size_t)blockIdx.x * P.open_entries;
  s.hmask = P.hcap - 1;
  const KeyStore fk{P.fkey + (size_t)blockIdx.x * P.cap, reinterpret_cast<double*>(fsmem + kFaithfulSmemTree * sizeof(TreeNode)),
                    kFaithfulSmemKeys};
  const TreeStore tstore{reinterpret_cast<TreeNode*>(s.heap), reinterpret_cast<TreeNode*>(fsmem), kFaithfulSmemTree};
  RbSet set;  // lane 0 of warp 0 owns the tree; the other lanes only see what it broadcasts

  const int la = lane / 9, lb = (lane / 3) % 3, lc = lane % 3;
  const double ddx = P.dv[la % 3], ddy = P.dv[lb], ddz = P.dv[lc];
  const double dnorm = norm3(ddx, ddy, ddz);
  const bool lane_nb = lane < 27 && !(dnorm < 1e-3);

  for (;;) {
    if (tid == 0) sh.q = (long long)atomicAdd(P.q_counter, 1ULL);
    __syncthreads();
    const long long qi = sh.q;
    if (qi >= P.nq) break;
    const long long q = P.qlist ? (long long)P.qlist[qi] : qi;
    const double sx = P.starts[3 * q], sy = P.starts[3 * q + 1], sz = P.starts[3 * q + 2];
    const double gx = P.goals[3 * q], gy = P.goals[3 * q + 1], gz = P.goals[3 * q + 2];
    unsigned long long t_begin = 0;
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));

    s.use = 1;
    long long iter_num = 0, los_checks = 0, los_samples = 0;
    int status = ST_OPEN_EMPTY, last = 0, term = -1, open_peak = 1;
    if (warp == 0) {
      if (lane == 0) {  // AStar.cpp:87-97
        NodeRec n;
        n.px = sx; n.py = sy; n.pz = sz;
        n.g = 0.0;
        n.f = dmul(P.lambda, dmul(P.tie, norm3(dsub(gx, sx), dsub(gy, sy), dsub(gz, sz))));
        n.parent = -1; n.hpos = -1; n.via = -1; n.state = NODE_OPEN; n.pad = 0;
        if (IS_AGR) agr_init_start<true>(P, n);
        n.slot = hash_insert(s, n.px, n.py, n.pz, 0);
        s.nodes[0] = n;
        fk.set(0u, n.f);
        set.init(tstore, fk, P.open_entries);
        set.insert(0u);
      }
      __syncwarp();
      sh.state = 0;
      sh.stop2 = 0;
    }

    int cur = 0;
    double cx = 0, cy = 0, cz = 0, cg = 0, cpen = 0;
    int cpar = -1;
    int nb_idx = -1;
    double nx = 0, ny = 0, nz = 0;
    for (;;) {
      bool valid = false;
      bool oom = false;
      if (warp == 0) {
        bool done = false;
        int popped = -1;
        if (lane == 0 && !set.empty()) {
          open_peak = max(open_peak, (int)set.count);
          popped = (int)set.pop_begin();  // AStar.cpp:105-106 (no closed test)
        }
        popped = __shfl_sync(0xffffffffu, popped, 0);
        if (popped < 0) {  // AStar.cpp:175-178
          status = ST_OPEN_EMPTY; last = cur; done = true;
        } else {
          cur = popped;
          const NodeRec c = s.nodes[cur];
          cx = c.px; cy = c.py; cz = c.pz; cg = c.g; cpar = c.parent;
          if (IS_AGR) cpen = c.penalty;
          if (fabs(dsub(cx, gx)) <= P.r && fabs(dsub(cy, gy)) <= P.r && fabs(dsub(cz, gz)) <= P.r) {  // :110-119
            status = ST_SOLVED; last = cur; term = cur; done = true;
          }
        }
        if (!done) {
          if (lane == 0) s.nodes[cur].state = (s.nodes[cur].state & ~0xff) | NODE_CLOSED;  // :122
          ++iter_num;
          __syncwarp();
          nb_idx = -1;
          valid = lane_nb;
          nx = dadd(cx, ddx); ny = dadd(cy, ddy); nz = dadd(cz, ddz);
          if (valid && !is_in_map(m, nx, ny, nz)) valid = false;
          if (valid) {
            nb_idx = hash_find(s, nx, ny, nz);
            if (nb_idx >= 0 && (s.nodes[nb_idx].state & 0xff) == NODE_CLOSED) valid = false;
          }
          if (valid && inflate_occ(m, nx, ny, nz) == 1) valid = false;
          const bool need = valid && nb_idx < 0;
          const unsigned cm = __ballot_sync(0xffffffffu, need);
          const int ncreate = __popc(cm);
          const int rank = __popc(cm & ((1u << lane) - 1u));
          bool overflow = false;
          int fail_lane = 32;
          if (ncreate > 0) {
            if (P.cap < P.allocate_num && s.use + ncreate > P.cap) overflow = true;
            else if (s.use + ncreate >= P.allocate_num) {
              const int fail_rank = P.allocate_num - s.use - 1;
              fail_lane = __ffs(__ballot_sync(0xffffffffu, need && rank == fail_rank)) - 1;
              oom = true;
            }
          }
          if (overflow) {
            status = ST_TIER_OVERFLOW; last = cur; done = true;
          } else {
            if (need && lane <= fail_lane && s.use + rank < P.cap) {
              const int ni = s.use + rank;
              NodeRec n;
              n.px = nx; n.py = ny; n.pz = nz;
              n.g = __longlong_as_double(0x7ff0000000000000LL);
              n.f = n.g;  // the reference leaves f stale here; it is always overwritten before its first use as a key
              n.parent = -1; n.hpos = -1; n.via = -1; n.state = NODE_NEW; n.pad = 0;
              n.slot = hash_insert(s, nx, ny, nz, ni);
              s.nodes[ni] = n;
              fk.set((uint32_t)ni, n.f);
              nb_idx = ni;
            }
            if (oom) {
              if (lane >= fail_lane) valid = false;
              s.use = min(P.allocate_num, P.cap);
              status = ST_POOL_EXHAUSTED; last = cur;
            } else s.use += ncreate;
            __syncwarp();
            const unsigned vm = __ballot_sync(0xffffffffu, valid);
            if (IS_THETA) {
              const bool lp = cpar >= 0 && vm != 0;
              if (lp) {
                los_checks += __popc(vm);
                if (lane < 27) { sh.nbx[lane] = nx; sh.nby[lane] = ny; sh.nbz[lane] = nz; }
                if (lane == 0) {
                  const NodeRec pr = s.nodes[cpar];
                  sh.ppx = pr.px; sh.ppy = pr.py; sh.ppz = pr.pz;
                  sh.valid = vm;
                }
              }
              if (lane == 0) sh.los_phase = lp ? 1 : 0;
            }
          }
        }
        if (lane == 0) sh.state = done ? 1 : 0;
        if (done) valid = false;
      }
      if (NW > 1) __syncthreads(); else __syncwarp();
      if (sh.state == 1) break;

      if (IS_THETA && sh.los_phase) {
        const unsigned vm = sh.valid;
        const double ppx = sh.ppx, ppy = sh.ppy, ppz = sh.ppz;
        for (int seg = warp; seg < 27; seg += NW) {
          if (!((vm >> seg) & 1u)) continue;
          int ns = 0;
          const bool vis = los_warp<false>(m, ppx, ppy, ppz, sh.nbx[seg], sh.nby[seg], sh.nbz[seg], lane, &ns, nullptr);
          if (lane == 0) { sh.vis[seg] = vis ? 1 : 0; sh.nsamp[seg] = ns; }
        }
        if (NW > 1) __syncthreads(); else __syncwarp();
      }

      if (warp == 0) {
        // ---- candidates, one lane per neighbour (pure functions of cur / parent(cur) / the neighbour's own record)
        int flags = 0, new_parent = cur;
        double new_g = 0, new_f = 0, new_pen = 0;
        if (valid) {
          const NodeRec nb = s.nodes[nb_idx];
          flags = 1 | (((nb.state & 0xff) == NODE_OPEN) ? 8 : 0);
          bool use_parent = false;
          if (IS_THETA && cpar >= 0) {
            los_samples += sh.nsamp[lane];
            use_parent = sh.vis[lane] != 0;
          }
          if (IS_AGR) {
            double pg = 0, ppen = 0, pdist = 0;
            if (cpar >= 0) {
              const NodeRec pr = s.nodes[cpar];
              pg = pr.g; ppen = pr.penalty;
              pdist = norm3(dsub(pr.px, nx), dsub(pr.py, ny), dsub(pr.pz, nz));
            }
            const AgrCand cand = agr_candidate(P, cg, cpen, cur, cpar, use_parent, pg, ppen, pdist, dnorm, nz);
            new_g = cand.g; new_parent = cand.parent; new_pen = cand.pen;
            if (cand.motion) flags |= 16;
          } else if (IS_THETA && use_parent) {
            const NodeRec pr = s.nodes[cpar];
            new_g = dadd(pr.g, norm3(dsub(pr.px, nx), dsub(pr.py, ny), dsub(pr.pz, nz)));
            new_parent = cpar;
          } else {
            new_g = dadd(cg, dnorm);
          }
          new_f = dadd(new_g, dmul(P.lambda, dmul(P.tie, norm3(dsub(gx, nx), dsub(gy, ny), dsub(gz, nz)))));
          const bool better = new_g < nb.g;
          if (IS_THETA) { if (better) flags |= 2 | 4; }   // ThetaStar.cpp: guarded write; improved iff written
          else { flags |= 2; if (better) flags |= 4; }    // AStar.cpp:60-62: unconditional write; :70 guards the insert
        }
        if (lane < 27) {
          fs.nb_idx[lane] = nb_idx; fs.new_parent[lane] = new_parent; fs.flags[lane] = flags;
          fs.new_g[lane] = new_g; fs.new_f[lane] = new_f;
          if (IS_AGR) fs.new_pen[lane] = new_pen;
        }
        __syncwarp();
        // ---- sequential replay in the reference's neighbour order (lane 0 owns the tree)
        if (lane == 0) {
          for (int k = 0; k < 27; ++k) {
            const int fl = fs.flags[k];
            if (!(fl & 1)) continue;
            NodeRec* w = s.nodes + fs.nb_idx[k];
            if (fl & 2) {
              w->parent = fs.new_parent[k]; w->g = fs.new_g[k]; w->f = fs.new_f[k]; fk.set((uint32_t)fs.nb_idx[k], fs.new_f[k]);
              if (IS_AGR) { w->penalty = fs.new_pen[k]; w->state = (w->state & ~kMotionBit) | ((fl & 16) ? kMotionBit : 0); }
            }
            if (fl & 4) {
              if (IS_THETA && (fl & 8)) set.erase_key((uint32_t)fs.nb_idx[k]);  // ThetaStar.cpp:80 (key already mutated)
              w->state = NODE_OPEN | (k << 8) | (w->state & kMotionBit);
              set.insert((uint32_t)fs.nb_idx[k]);                                 // AStar.cpp:71-76 / ThetaStar.cpp:81-86
            }
          }
        }
        const int ovf = __shfl_sync(0xffffffffu, set.overflow ? 1 : 0, 0);
        if (ovf) { status = P.cap < P.allocate_num ? ST_TIER_OVERFLOW : ST_OPEN_LIST_FULL; last = cur; }
        if (lane == 0 && (ovf || status == ST_POOL_EXHAUSTED)) sh.stop2 = 1;
      }
      if (NW > 1) __syncthreads(); else __syncwarp();
      if (sh.stop2 == 1) break;
    }

    // ---------------- result (same as the canonical kernel) ----------------
    if (warp == 0) {
      if (IS_THETA) {
        long long v = los_samples;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        los_samples = v;
      }
      int n_path = 0;
      long long off = -1;
      float plen = 0.f;
      int* chain = reinterpret_cast<int*>(s.heap);
      float* seglen = reinterpret_cast<float*>(s.heap) + (size_t)P.cap + 8;
      if (status == ST_SOLVED) {
        if (lane == 0) {
          int c = term;
          while (c >= 0) { chain[n_path++] = c; c = s.nodes[c].parent; }
          if (P.path_pool) {
            const unsigned long long o = atomicAdd(P.path_cursor, (unsigned long long)n_path);
            off = (o + (unsigned long long)n_path <= (unsigned long long)P.path_cap) ? (long long)o : -1;
          }
        }
        n_path = __shfl_sync(0xffffffffu, n_path, 0);
        off = __shfl_sync(0xffffffffu, off, 0);
        __syncwarp();
        for (int i = lane; i < n_path; i += 32) {
          const NodeRec a = s.nodes[chain[n_path - 1 - i]];
          if (off >= 0) {
            P.path_pool[3 * (off + i)] = a.px; P.path_pool[3 * (off + i) + 1] = a.py; P.path_pool[3 * (off + i) + 2] = a.pz;
          }
          if (i > 0) {
            const NodeRec b = s.nodes[chain[n_path - i]];
            const double ux = dmul(dsub(a.px, b.px), m.res), uy = dmul(dsub(a.py, b.py), m.res), uz = dmul(dsub(a.pz, b.pz), m.res);
            seglen[i] = __fsqrt_rn((float)dadd(dadd(dmul(ux, ux), dmul(uy, uy)), dmul(uz, uz)));
          }
        }
        __syncwarp();
        if (lane == 0)
          for (int i = 1; i < n_path; ++i) plen = __fadd_rn(plen, seglen[i]);
      }
      if (lane == 0) {
        b200_path_result_t r;
        const NodeRec ln = s.nodes[last];
        r.solved = status == ST_SOLVED ? 1 : 0;
        r.status = status;
        r.n_path = n_path;
        r.open_peak = open_peak;
        r.path_offset = off;
        r.g_final = ln.g;
        r.path_length = (double)plen;
        r.explored_nodes = s.use;
        r.iter_num = iter_num;
        r.los_checks = los_checks;
        r.los_samples = los_samples;
        r.goal_coords[0] = ln.px; r.goal_coords[1] = ln.py; r.goal_coords[2] = ln.pz;
        r.start_coords[0] = sx; r.start_coords[1] = sy; r.start_coords[2] = sz;
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        r.time_us = (double)(t_end - t_begin) * 1e-3;
        P.results[q] = r;
        if (status == ST_TIER_OVERFLOW && P.overflow_list) P.overflow_list[atomicAdd(P.overflow_count, 1)] = (int)q;
      }
      sh.used = s.use;
    }
    __syncthreads();
    const int used = min(sh.used, P.cap);
    for (int i = tid; i < used; i += NW * 32) s.hash[s.nodes[i].slot] = 0u;
    __syncthreads();
  }
}

size_t search_slot_bytes(int cap, int open_factor, uint32_t* hcap_out) {
  uint32_t h = 1024;
  while (h < 2u * (uint32_t)cap) h <<= 1;
  if (hcap_out) *hcap_out = h;
  return (size_t)cap * (sizeof(NodeRec) + sizeof(double)) + (size_t)h * 4 + open_region_entries(cap, open_factor) * sizeof(HeapEnt);
}

static inline bool algo_has_los_phase(int algo) { return algo == ALGO_THETASTAR || algo == ALGO_THETASTARAGR || algo == ALGO_THETASTARAGRPP; }
// ---------------------------------------------------------------------------------------------------
// launch order: longest-expected-first.  Blocks pull queries from a counter, so a long query that starts late
// leaves its slot busy while every other slot idles (measured on 131 072 Lazy Theta* queries: slot time sums to
// 141 ms per slot, the batch took 177 ms because a 100 ms query started late).  Predictor = the length of the
// straight start-goal segment that lies inside (slightly widened) obstacles: what a weighted search has to work its
// way around.  On 24 000 C2 queries, list scheduling by it ends 1.16x above the ideal makespan, against 1.41x for the
// squared start-goal distance used before (r = 0.5 with the expansion count) — profiles/README.md "issue order".
// One warp per query samples the segment every 0.1 m on five parallel lines (0, +-0.15, +-0.3 m sideways in xy).
// Results are indexed by the caller's query id whatever the order.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) query_order_keys_kernel(MapDev m, const double* __restrict__ starts, const double* __restrict__ goals,
                                                               int nq, float* __restrict__ keys, int* __restrict__ ids) {
  const int q = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (q >= nq) return;
  const double sx = starts[3 * (size_t)q], sy = starts[3 * (size_t)q + 1], sz = starts[3 * (size_t)q + 2];
  const double dx = goals[3 * (size_t)q] - sx, dy = goals[3 * (size_t)q + 1] - sy, dz = goals[3 * (size_t)q + 2] - sz;
  const float d = sqrtf((float)(dx * dx + dy * dy + dz * dz));
  float key = 0.0f;
  if (d == d && d < 1e6f) {  // NaN / absurd coordinates: any position in the order will do
    const int n = min(4096, max(1, (int)(d * 10.0f)));  // samples per line
    const double nxy = sqrt(dx * dx + dy * dy);
    const double px = nxy > 0 ? -dy / nxy : 1.0, py = nxy > 0 ? dx / nxy : 0.0;  // unit normal in the xy plane
    int hits = 0;
    for (int k = lane; k < n; k += 32) {
      const double t = (k + 0.5) / n;
      const double cx = sx + dx * t, cy = sy + dy * t, cz = sz + dz * t;
      bool hit = false;
#pragma unroll
      for (int j = -2; j <= 2; ++j) hit |= inflate_occ(m, cx + px * (0.15 * j), cy + py * (0.15 * j), cz) == 1;
      hits += hit ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
    key = (float)hits * (d / (float)n) + 1e-3f * d;  // blocked length (m); the distance only orders the unblocked queries
  }
  if (lane == 0) { keys[q] = key; ids[q] = q; }
}
static size_t align256(size_t b) { return (b + 255) & ~(size_t)255; }
static size_t query_order_sort_bytes(int nq) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const float*)nullptr, (float*)nullptr, (const int*)nullptr, (int*)nullptr, nq);
  return bytes;
}
size_t query_order_bytes(int nq) { return 4 * align256((size_t)nq * 4) + align256(query_order_sort_bytes(nq)); }
// returns the device array of nq query ids inside `buf` (query_order_bytes(nq) bytes)
const int* launch_query_order(const MapDev& m, const double* starts, const double* goals, int nq, void* buf, cudaStream_t st) {
  const size_t a = align256((size_t)nq * 4);
  float* k_in = (float*)buf;
  float* k_out = (float*)((char*)buf + a);
  int* v_in = (int*)((char*)buf + 2 * a);
  int* v_out = (int*)((char*)buf + 3 * a);
  void* tmp = (char*)buf + 4 * a;
  size_t tmp_bytes = query_order_sort_bytes(nq);
  query_order_keys_kernel<<<(unsigned)(((long long)nq * 32 + 255) / 256), 256, 0, st>>>(m, starts, goals, nq, k_in, v_in);
  cub::DeviceRadixSort::SortPairsDescending(tmp, tmp_bytes, k_in, k_out, v_in, v_out, nq, 0, 32, st);
  return v_out;
}

int search_threads_per_block(int algo) { return algo_has_los_phase(algo) ? 128 : 32; }
// workspace slots (= resident thread blocks) per SM.  A thread-per-query variant of faithful A* was measured too:
// 14x slower per expansion (495 vs 35 us) because SIMT divergence serialises the 32 different queries of a warp.
int search_slots_per_sm(int algo, int faithful) {
  (void)faithful;
  // measured on 1xB200, C2 world (queries/s): canonical A* 16 -> 685k, 24 -> 870k, 32 -> 911k slots/SM (32 resident
  // blocks is the SM's limit); faithful A* 16 -> 6.99k, 32 -> 7.08k (kept at 16: half the workspace); 128-thread LoS
  // kernels 4 -> 70.8k, 8 -> 78.2k, more is flat (registers allow 5 resident blocks).
  if (algo_has_los_phase(algo)) return 6;
  return faithful ? 16 : 32;
}

void launch_search(const SearchParams& p, int slots, cudaStream_t st) {
  if (p.faithful) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(search_faithful_kernel<ALGO_THETASTAR, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFaithfulSmemBytes);
      cudaFuncSetAttribute(search_faithful_kernel<ALGO_ASTAR, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFaithfulSmemBytes);
      attr_set = true;
    }
    if (p.algo == ALGO_THETASTAR) search_faithful_kernel<ALGO_THETASTAR, 4><<<slots, 128, kFaithfulSmemBytes, st>>>(p);
    else if (p.algo == ALGO_THETASTARAGR) search_faithful_kernel<ALGO_THETASTARAGR, 4><<<slots, 128, kFaithfulSmemBytes, st>>>(p);
    else search_faithful_kernel<ALGO_ASTAR, 1><<<slots, 32, kFaithfulSmemBytes, st>>>(p);
    return;
  }
  switch (p.algo) {
    case ALGO_ASTAR: search_kernel<ALGO_ASTAR, 1><<<slots, 32, 0, st>>>(p); break;
    case ALGO_THETASTAR: search_kernel<ALGO_THETASTAR, 4><<<slots, 128, 0, st>>>(p); break;
    case ALGO_LAZYTHETASTAR: search_kernel<ALGO_LAZYTHETASTAR, 1><<<slots, 32, 0, st>>>(p); break;
    case ALGO_COSTASTAR: search_kernel<ALGO_COSTASTAR, 1><<<slots, 32, 0, st>>>(p); break;
    case ALGO_THETASTARAGR: search_kernel<ALGO_THETASTARAGR, 4><<<slots, 128, 0, st>>>(p); break;
    case ALGO_THETASTARAGRPP: search_kernel<ALGO_THETASTARAGRPP, 4><<<slots, 128, 0, st>>>(p); break;
    default: search_kernel<ALGO_COSTLAZYTHETASTAR, 1><<<slots, 32, 0, st>>>(p); break;
  }
}

}  // namespace b200
