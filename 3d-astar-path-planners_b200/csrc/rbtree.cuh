// rbtree.cuh — device-side restatement of the ordered set the reference uses as its open list:
//   std::set<Node0*, NodeComparator> open_set_   (include/Planners/AlgorithmBase.hpp:77, utils.hpp:52-61)
//
// Why it exists: the reference MUTATES the key (f_score) of nodes that are inside the set, inserts without
// erasing (AStar.cpp:71-76) and erases by an already-mutated key (ThetaStar.cpp:80-83).  Its pop order is
// therefore a function of the red-black tree's SHAPE, not of a priority-queue abstraction (SURVEY Appendix
// B2).  To be bit-identical with the reference ("faithful" semantics) the open list has to be the same
// tree, rebalanced by the same rules: the algorithms below follow the published behaviour of libstdc++'s
// _Rb_tree (insert_unique position search, insert-and-rebalance, rebalance-for-erase, equal_range,
// increment/decrement) — GCC 13, the toolchain the oracle links against — written from scratch over
// 32-bit indices into the slot's workspace instead of pointers.
//
// One lane runs these (they are pointer chasing by nature); the other lanes of the control warp wait.
#pragma once
#include "kernels.h"

namespace b200 {

struct __align__(16) TreeNode {
  uint32_t parent, left, right;
  uint32_t vc;  // node index (low 31 bits) | red flag (bit 31)
};
static_assert(sizeof(TreeNode) == sizeof(HeapEnt), "tree nodes reuse the open-list region of the slot");

constexpr uint32_t RB_NIL = 0xffffffffu;
constexpr uint32_t RB_HEADER = 0u;
constexpr uint32_t RB_RED = 0x80000000u;

// Storage: pointer chasing at L2 latency (~350 cycles per dependent load) is what bounds this structure, so the
// first `kt` tree nodes and the first `kf` comparator keys live in SHARED memory (~30 cycles) and only the excess
// spills to the slot's global workspace.  Tree nodes are allocated lowest-index-first (free list, then bump), so
// a typical query (a few thousand live entries) never leaves shared memory.
struct TreeStore {
  TreeNode* g;   // global part (indexed by the absolute node id)
  TreeNode* s;   // shared part: ids [0, kt)
  uint32_t kt;
  __device__ __forceinline__ TreeNode& operator[](uint32_t x) const { return x < kt ? s[x] : g[x]; }
  // the whole node as ONE 16-byte load that the optimiser cannot split into per-field loads and sink below the uses
  // (it does that to `TreeNode n = t[x]`: vc, then the key, then — after the compare — left OR right: three dependent
  // round trips per level of a descent instead of two)
  __device__ __forceinline__ TreeNode load(uint32_t x) const {
    TreeNode n;
    if (x < kt) { n = s[x]; return n; }
    asm volatile("ld.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(n.parent), "=r"(n.left), "=r"(n.right), "=r"(n.vc) : "l"(g + x) : "memory");
    return n;
  }
};
struct KeyStore {  // compact copy of NodeRec::f — the comparator's only input
  double* g;
  double* s;
  uint32_t kf;
  __device__ __forceinline__ double get(uint32_t i) const { return i < kf ? s[i] : g[i]; }
  __device__ __forceinline__ void set(uint32_t i, double v) const { if (i < kf) s[i] = v; else g[i] = v; }
};

struct RbSet {
  TreeStore t;            // t[0] is the header: parent = root, left = leftmost, right = rightmost
  KeyStore fk;            // keys are compared through the nodes' CURRENT f (that is the point)
  uint32_t count, top, cap, free_head;
  bool overflow;

  __device__ void init(const TreeStore& mem, const KeyStore& keys, uint32_t capacity) {
    t = mem; fk = keys; cap = capacity;
    count = 0; top = 1; free_head = RB_NIL; overflow = false;
    t[RB_HEADER].parent = RB_NIL; t[RB_HEADER].left = RB_HEADER; t[RB_HEADER].right = RB_HEADER; t[RB_HEADER].vc = RB_RED;
  }
  __device__ bool empty() const { return count == 0; }
  __device__ uint32_t val(uint32_t x) const { return t[x].vc & ~RB_RED; }
  __device__ bool red(uint32_t x) const { return (t[x].vc & RB_RED) != 0; }
  __device__ bool black_or_nil(uint32_t x) const { return x == RB_NIL || !red(x); }
  __device__ void set_red(uint32_t x) { t[x].vc |= RB_RED; }
  __device__ void set_black(uint32_t x) { t[x].vc &= ~RB_RED; }
  __device__ void copy_color(uint32_t dst, uint32_t src) { t[dst].vc = (t[dst].vc & ~RB_RED) | (t[src].vc & RB_RED); }

  // NodeComparator (utils.hpp:52-61): f ascending, then node address — the pool index here (creation order,
  // what a fresh heap gives the reference's per-node `new`)
  __device__ bool key_less(uint32_t a, uint32_t b) const {
    const double fa = fk.get(a), fb = fk.get(b);
    if (fa != fb) return fa < fb;
    return a < b;
  }

  __device__ uint32_t alloc(uint32_t v) {
    uint32_t x;
    if (free_head != RB_NIL) { x = free_head; free_head = t[x].left; }
    else if (top < cap) x = top++;
    else { overflow = true; return RB_NIL; }
    t[x].vc = v;
    return x;
  }
  __device__ void release(uint32_t x) { t[x].left = free_head; free_head = x; }

  __device__ uint32_t minimum(uint32_t x) const { while (t[x].left != RB_NIL) x = t[x].left; return x; }
  __device__ uint32_t maximum(uint32_t x) const { while (t[x].right != RB_NIL) x = t[x].right; return x; }

  __device__ uint32_t increment(uint32_t x) const {
    if (t[x].right != RB_NIL) {
      x = t[x].right;
      while (t[x].left != RB_NIL) x = t[x].left;
    } else {
      uint32_t y = t[x].parent;
      while (x == t[y].right) { x = y; y = t[y].parent; }
      if (t[x].right != y) x = y;
    }
    return x;
  }
  __device__ uint32_t decrement(uint32_t x) const {
    if (red(x) && t[t[x].parent].parent == x) return t[x].right;  // x is the header
    if (t[x].left != RB_NIL) {
      uint32_t y = t[x].left;
      while (t[y].right != RB_NIL) y = t[y].right;
      return y;
    }
    uint32_t y = t[x].parent;
    while (x == t[y].left) { x = y; y = t[y].parent; }
    return y;
  }

  __device__ void rotate_left(uint32_t x) {
    const uint32_t y = t[x].right;
    t[x].right = t[y].left;
    if (t[y].left != RB_NIL) t[t[y].left].parent = x;
    t[y].parent = t[x].parent;
    if (x == t[RB_HEADER].parent) t[RB_HEADER].parent = y;
    else if (x == t[t[x].parent].left) t[t[x].parent].left = y;
    else t[t[x].parent].right = y;
    t[y].left = x;
    t[x].parent = y;
  }
  __device__ void rotate_right(uint32_t x) {
    const uint32_t y = t[x].left;
    t[x].left = t[y].right;
    if (t[y].right != RB_NIL) t[t[y].right].parent = x;
    t[y].parent = t[x].parent;
    if (x == t[RB_HEADER].parent) t[RB_HEADER].parent = y;
    else if (x == t[t[x].parent].right) t[t[x].parent].right = y;
    else t[t[x].parent].left = y;
    t[y].right = x;
    t[x].parent = y;
  }

  __device__ void insert_and_rebalance(bool insert_left, uint32_t x, uint32_t p) {
    t[x].parent = p; t[x].left = RB_NIL; t[x].right = RB_NIL; set_red(x);
    if (insert_left) {
      t[p].left = x;  // also makes leftmost = x when p is the header
      if (p == RB_HEADER) { t[RB_HEADER].parent = x; t[RB_HEADER].right = x; }
      else if (p == t[RB_HEADER].left) t[RB_HEADER].left = x;
    } else {
      t[p].right = x;
      if (p == t[RB_HEADER].right) t[RB_HEADER].right = x;
    }
    while (x != t[RB_HEADER].parent && red(t[x].parent)) {
      const uint32_t xpp = t[t[x].parent].parent;
      if (t[x].parent == t[xpp].left) {
        const uint32_t y = t[xpp].right;
        if (y != RB_NIL && red(y)) {
          set_black(t[x].parent); set_black(y); set_red(xpp);
          x = xpp;
        } else {
          if (x == t[t[x].parent].right) { x = t[x].parent; rotate_left(x); }
          set_black(t[x].parent); set_red(xpp);
          rotate_right(xpp);
        }
      } else {
        const uint32_t y = t[xpp].left;
        if (y != RB_NIL && red(y)) {
          set_black(t[x].parent); set_black(y); set_red(xpp);
          x = xpp;
        } else {
          if (x == t[t[x].parent].left) { x = t[x].parent; rotate_right(x); }
          set_black(t[x].parent); set_red(xpp);
          rotate_left(xpp);
        }
      }
    }
    set_black(t[RB_HEADER].parent);
  }

  // unlinks z, rebalances, returns the tree node to free (always z's storage after the relink)
  __device__ uint32_t rebalance_for_erase(uint32_t z) {
    uint32_t y = z, x = RB_NIL, x_parent = RB_NIL;
    if (t[y].left == RB_NIL) x = t[y].right;
    else if (t[y].right == RB_NIL) x = t[y].left;
    else {
      y = t[y].right;
      while (t[y].left != RB_NIL) y = t[y].left;
      x = t[y].right;
    }
    if (y != z) {  // relink y (z's successor) in place of z
      t[t[z].left].parent = y;
      t[y].left = t[z].left;
      if (y != t[z].right) {
        x_parent = t[y].parent;
        if (x != RB_NIL) t[x].parent = t[y].parent;
        t[t[y].parent].left = x;
        t[y].right = t[z].right;
        t[t[z].right].parent = y;
      } else x_parent = y;
      if (t[RB_HEADER].parent == z) t[RB_HEADER].parent = y;
      else if (t[t[z].parent].left == z) t[t[z].parent].left = y;
      else t[t[z].parent].right = y;
      t[y].parent = t[z].parent;
      const uint32_t cy = t[y].vc & RB_RED, cz = t[z].vc & RB_RED;  // swap colours
      t[y].vc = (t[y].vc & ~RB_RED) | cz;
      t[z].vc = (t[z].vc & ~RB_RED) | cy;
      y = z;
    } else {
      x_parent = t[y].parent;
      if (x != RB_NIL) t[x].parent = t[y].parent;
      if (t[RB_HEADER].parent == z) t[RB_HEADER].parent = x;
      else if (t[t[z].parent].left == z) t[t[z].parent].left = x;
      else t[t[z].parent].right = x;
      if (t[RB_HEADER].left == z) {
        if (t[z].right == RB_NIL) t[RB_HEADER].left = t[z].parent;  // header when z was the root
        else t[RB_HEADER].left = minimum(x);
      }
      if (t[RB_HEADER].right == z) {
        if (t[z].left == RB_NIL) t[RB_HEADER].right = t[z].parent;
        else t[RB_HEADER].right = maximum(x);
      }
    }
    if (!red(y)) {
      while (x != t[RB_HEADER].parent && black_or_nil(x)) {
        if (x == t[x_parent].left) {
          uint32_t w = t[x_parent].right;
          if (red(w)) { set_black(w); set_red(x_parent); rotate_left(x_parent); w = t[x_parent].right; }
          if (black_or_nil(t[w].left) && black_or_nil(t[w].right)) {
            set_red(w); x = x_parent; x_parent = t[x_parent].parent;
          } else {
            if (black_or_nil(t[w].right)) { set_black(t[w].left); set_red(w); rotate_right(w); w = t[x_parent].right; }
            copy_color(w, x_parent); set_black(x_parent);
            if (t[w].right != RB_NIL) set_black(t[w].right);
            rotate_left(x_parent);
            break;
          }
        } else {
          uint32_t w = t[x_parent].left;
          if (red(w)) { set_black(w); set_red(x_parent); rotate_right(x_parent); w = t[x_parent].left; }
          if (black_or_nil(t[w].right) && black_or_nil(t[w].left)) {
            set_red(w); x = x_parent; x_parent = t[x_parent].parent;
          } else {
            if (black_or_nil(t[w].left)) { set_black(t[w].right); set_red(w); rotate_left(w); w = t[x_parent].left; }
            copy_color(w, x_parent); set_black(x_parent);
            if (t[w].left != RB_NIL) set_black(t[w].left);
            rotate_right(x_parent);
            break;
          }
        }
      }
      if (x != RB_NIL) set_black(x);
    }
    return y;
  }

  __device__ void erase_node(uint32_t z) {
    const uint32_t y = rebalance_for_erase(z);
    release(y);
    --count;
  }

  // std::set::insert(v): unique-position search with the CURRENT keys, then insert-and-rebalance
  __device__ void insert(uint32_t v) {
    uint32_t x = t[RB_HEADER].parent, y = RB_HEADER;
    bool comp = true;
    // the descent is the hot loop of the faithful kernels (one lane, ~15 levels per insert): per level ONE 16-byte load of
    // the tree node and one load of its key; the inserted node's key cannot change during the descent and stays in a register
    const double fv = fk.get(v);
    if (x != RB_NIL) {
      // Per level the key of the node and BOTH child nodes are requested together (all three addresses come from the node
      // in hand), so a level costs one memory round trip instead of two dependent ones; the child not taken is a wasted
      // load on a lane that is waiting anyway.  The loop body is kept to what a level needs: the colour bit is masked in 32
      // bits before the key address is formed (one LOP + one IMAD.WIDE; the compiler's own 64-bit form took six), the
      // comparison is (fv < fx) || (fv == fx && v < xv) on predicates, and only left / right / vc of the chosen child move.
      uint32_t nleft, nright, nvc;
      {
        const TreeNode n = t.load(x);
        nleft = n.left; nright = n.right; nvc = n.vc;
      }
      for (;;) {
        y = x;
        uint32_t xv;
        asm("and.b32 %0, %1, 0x7fffffff;" : "=r"(xv) : "r"(nvc));  // opaque on purpose: keeps the mask out of the 64-bit address arithmetic
        const double fx = fk.get(xv);
        // (a missing child loads the header instead: unconditional loads, and what they return is never used — the loop
        // ends on x == RB_NIL before)
        const TreeNode cl = t.load(nleft != RB_NIL ? nleft : RB_HEADER), cr = t.load(nright != RB_NIL ? nright : RB_HEADER);
        const uint32_t ll = cl.left, lr = cl.right, lv = cl.vc, rl = cr.left, rr = cr.right, rv = cr.vc;
        comp = (fv < fx) || (fv == fx && v < xv);
        x = comp ? nleft : nright;
        if (x == RB_NIL) break;
        nleft = comp ? ll : rl; nright = comp ? lr : rr; nvc = comp ? lv : rv;
      }
    }
    uint32_t j = y;
    bool do_insert = false;
    if (comp) {
      if (j == t[RB_HEADER].left) do_insert = true;  // j == begin()
      else j = decrement(j);
    }
    if (!do_insert) do_insert = key_less(val(j), v);
    if (!do_insert) return;  // an equivalent element (the same node, found under its current key) is already there
    const bool insert_left = (y == RB_HEADER) || key_less(v, val(y));
    const uint32_t z = alloc(v);
    if (z == RB_NIL) return;
    insert_and_rebalance(insert_left, z, y);
    ++count;
  }

  // *begin() then erase(begin())
  __device__ uint32_t pop_begin() {
    const uint32_t b = t[RB_HEADER].left;
    const uint32_t v = val(b);
    erase_node(b);
    return v;
  }

  __device__ void clear() { init(t, fk, cap); }

  // std::set::erase(const key_type&): equal_range under the current keys, then erase the range
  __device__ void erase_key(uint32_t k) {
    uint32_t x = t[RB_HEADER].parent, y = RB_HEADER;
    uint32_t first = RB_HEADER, last = RB_HEADER;
    bool found = false;
    const double fkk = fk.get(k);  // the searched key: constant during the descents
    auto less_than_k = [&](uint32_t a, double fa) { return (fa != fkk) ? fa < fkk : a < k; };   // key(a) < key(k)
    auto k_less_than = [&](uint32_t a, double fa) { return (fkk != fa) ? fkk < fa : k < a; };   // key(k) < key(a)
    while (x != RB_NIL) {
      const TreeNode n = t.load(x);
      const uint32_t xv = n.vc & ~RB_RED;
      const double fx = fk.get(xv);
      if (less_than_k(xv, fx)) x = n.right;
      else if (k_less_than(xv, fx)) { y = x; x = n.left; }
      else {
        uint32_t xu = n.right, yu = y;
        y = x; x = n.left;
        while (x != RB_NIL) {  // lower bound
          const TreeNode m = t.load(x);
          const uint32_t mv = m.vc & ~RB_RED;
          if (!less_than_k(mv, fk.get(mv))) { y = x; x = m.left; } else x = m.right;
        }
        while (xu != RB_NIL) {  // upper bound
          const TreeNode m = t.load(xu);
          const uint32_t mv = m.vc & ~RB_RED;
          if (k_less_than(mv, fk.get(mv))) { yu = xu; xu = m.left; } else xu = m.right;
        }
        first = y; last = yu; found = true;
        break;
      }
    }
    if (!found) return;  // empty range (first == last == y)
    if (first == t[RB_HEADER].left && last == RB_HEADER) { clear(); return; }
    while (first != last) {
      const uint32_t nxt = increment(first);
      erase_node(first);
      first = nxt;
    }
  }
};

}  // namespace b200
