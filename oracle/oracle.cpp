// oracle/oracle.cpp — CPU restatement of the reference hot path.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may load this
// library, and only as the checker / the timed CPU baseline.  The product path
// (3d-astar-path-planners_b200/csrc) never links or calls it and has no CPU fallback.
//
// PARITY PIN: the reference (ChanJoon/3D-Astar-Path-Planners) ships no tests, golden vectors or
// fixtures.  This restatement is therefore pinned two ways:
//   (1) hand-derived known-answer facts (tests/test_oracle_facts.py), and
//   (2) oracle/_ref: the reference's OWN unmodified .cpp files compiled against a header shim
//       (oracle/shim/, see oracle/Makefile) and diffed against this file on seeded inputs
//       (tests/test_oracle_vs_ref.py; golden vectors from it are committed in tests/golden/).
// The only arithmetic the shim cannot pin is Eigen's own (un-vendored, Eigen 3.3.x assumed):
// squaredNorm() associates as (x*x + y*y) + z*z, normalized() divides each component by
// sqrt(squaredNorm()).  See DESIGN.md "Floating-point pins".
//
// All file:line citations are relative to /root/reference.
// Build: g++ -O2 -std=c++17 -ffp-contract=off -pthread -shared -fPIC  (oracle/Makefile)
//   -O2 / no -march / no fast-math mirrors CMakeLists.txt:4-7 (Release flags).

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>
#include <atomic>
#include <functional>
#include <thread>

namespace {

// Minimal thread pool-less parallel loop (std::thread; no OpenMP dependency).  fn(begin, end, tid).
static void parallel_chunks(int64_t n, int nthreads, int64_t chunk, const std::function<void(int64_t, int64_t, int)>& fn) {
  if (nthreads <= 1 || n <= chunk) { fn(0, n, 0); return; }
  std::atomic<int64_t> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&, t]() {
      for (;;) {
        int64_t b = next.fetch_add(chunk);
        if (b >= n) break;
        fn(b, std::min(n, b + chunk), t);
      }
    });
  for (auto& x : th) x.join();
}

struct V3 {
  double x, y, z;
};
static inline V3 operator+(const V3& a, const V3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(const V3& a, const V3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(const V3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
// Eigen 3.3 Vector3d::squaredNorm() with SSE2 packets: predux(x*x, y*y) then + z*z.
static inline double sqnorm(const V3& a) { return (a.x * a.x + a.y * a.y) + a.z * a.z; }
static inline double norm(const V3& a) { return std::sqrt(sqnorm(a)); }
// Eigen 3.3 normalized(): z = squaredNorm(); z > 0 ? n / sqrt(z) : n   (true division per component)
static inline V3 normalized(const V3& a) {
  double z = sqnorm(a);
  if (z > 0.0) {
    double s = std::sqrt(z);
    return {a.x / s, a.y / s, a.z / s};
  }
  return a;
}

// ------------------------------------------------------------------------------------------
// GridMap restatement
// ------------------------------------------------------------------------------------------
struct OMap {
  // include/plan_env/grid_map.h:50-58 (MappingParameters)
  double res, res_inv;
  V3 origin, size, minb, maxb, range;
  int N[3];
  double inflation, ground_height;
  // include/plan_env/grid_map.h:95 occupancy_buffer_inflate_ (char per voxel, z fastest)
  std::vector<char> inflate;
  int lb_min[3], lb_max[3];
  // new functionality (not in the reference): squared EDT + quantised cost field
  std::vector<int32_t> dist2;
  std::vector<uint16_t> costq;
  // depth-image fusion state (grid_map.h:89-135 MappingData), allocated by the first depth frame
  struct DepthParams {  // same layout as b200_depth_params_t
    double fx, fy, cx, cy, k_depth_scaling_factor, depth_filter_mindist, depth_filter_maxdist, max_ray_length;
    double p_hit, p_miss, p_min, p_max, p_occ, virtual_ceil_height;
    int32_t use_depth_filter, depth_filter_margin, skip_pixel, local_map_margin;
  } dp;
  bool dp_set = false;
  double prob_hit_log, prob_miss_log, clamp_min_log, clamp_max_log, min_occupancy_log, unknown_flag;
  std::vector<double> occ;              // occupancy_buffer_ (log-odds)
  std::vector<short> cnt_all, cnt_hit;  // count_hit_and_miss_, count_hit_
  std::vector<char> flag_rayend, flag_traverse;
  signed char raycast_num = 0;  // grid_map.h:124 `char raycast_num_`: wraps 127 -> -128, so flags alias every 256 fused frames
  bool has_first_depth = false, local_updated = false;
  int64_t depth_stats[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // [0] projected points, [1] rays cast, [2] ray steps, [3] voxels updated
  int64_t nvox() const { return (int64_t)N[0] * N[1] * N[2]; }
};

// grid_map.h:400-404  posToIndex: id = floor((pos - origin) * res_inv)
static inline void pos_to_index(const OMap& m, const V3& p, int id[3]) {
  id[0] = (int)std::floor((p.x - m.origin.x) * m.res_inv);
  id[1] = (int)std::floor((p.y - m.origin.y) * m.res_inv);
  id[2] = (int)std::floor((p.z - m.origin.z) * m.res_inv);
}
// grid_map.h:257-265  toAddress (z fastest). 64-bit here; the reference's int overflows only at C5 size.
static inline int64_t to_address(const OMap& m, int x, int y, int z) {
  return (int64_t)x * m.N[1] * m.N[2] + (int64_t)y * m.N[2] + z;
}
// grid_map.h:267-274 boundIndex
static inline void bound_index(const OMap& m, int id[3]) {
  for (int i = 0; i < 3; ++i) id[i] = std::max(std::min(id[i], m.N[i] - 1), 0);
}
// grid_map.h:370-385 isInMap(Vector3d)
static inline bool is_in_map(const OMap& m, const V3& p) {
  if (p.x < m.minb.x + 1e-4 || p.y < m.minb.y + 1e-4 || p.z < m.minb.z + 1e-4) return false;
  if (p.x > m.maxb.x - 1e-4 || p.y > m.maxb.y - 1e-4 || p.z > m.maxb.z - 1e-4) return false;
  return true;
}
// grid_map.h:387-398 isInMap(Vector3i)
static inline bool is_in_map_idx(const OMap& m, const int id[3]) {
  if (id[0] < 0 || id[1] < 0 || id[2] < 0) return false;
  if (id[0] > m.N[0] - 1 || id[1] > m.N[1] - 1 || id[2] > m.N[2] - 1) return false;
  return true;
}
// grid_map.h:350-359 getInflateOccupancy: -1 outside, else the byte
static inline int get_inflate_occupancy(const OMap& m, const V3& p) {
  if (!is_in_map(m, p)) return -1;
  int id[3];
  pos_to_index(m, p, id);
  return int(m.inflate[to_address(m, id[0], id[1], id[2])]);
}

// grid_map.cpp:155-171 resetBuffer(min,max): inclusive index box, clamped
static void reset_buffer(OMap& m, const V3& mn, const V3& mx) {
  int a[3], b[3];
  pos_to_index(m, mn, a);
  pos_to_index(m, mx, b);
  bound_index(m, a);
  bound_index(m, b);
  for (int x = a[0]; x <= b[0]; ++x)
    for (int y = a[1]; y <= b[1]; ++y)
      for (int z = a[2]; z <= b[2]; ++z) m.inflate[to_address(m, x, y, z)] = 0;
}

// grid_map.cpp:711-804 cloudCallback (after pcl::fromROSMsg: float32 x,y,z per point).
// Preconditions has_odom_ / non-empty / camera not NaN (:718-728) are the caller's.
static void cloud_update(OMap& m, const uint8_t* pts, int64_t n, int step, int ox, int oy, int oz, const V3& cam,
                         int nthreads) {
  if (n == 0) return;                                              // :724-725
  if (std::isnan(cam.x) || std::isnan(cam.y) || std::isnan(cam.z)) return;  // :727-728
  reset_buffer(m, cam - m.range, cam + m.range);                   // :730
  const int inf_step = (int)std::ceil(m.inflation / m.res);        // :735
  const int inf_step_z = 1;                                        // :736
  struct MM { double min_x, min_y, min_z, max_x, max_y, max_z; char pad[64]; };
  std::vector<MM> mm(std::max(1, nthreads), MM{m.maxb.x, m.maxb.y, m.maxb.z,      // :740-742
                                               m.minb.x, m.minb.y, m.minb.z, {0}});  // :744-746
  parallel_chunks(n, nthreads, 1 << 14, [&](int64_t i0, int64_t i1, int tid) {
  double min_x = mm[tid].min_x, min_y = mm[tid].min_y, min_z = mm[tid].min_z;
  double max_x = mm[tid].max_x, max_y = mm[tid].max_y, max_z = mm[tid].max_z;
  for (int64_t i = i0; i < i1; ++i) {                              // :748
    float fx, fy, fz;
    std::memcpy(&fx, pts + i * step + ox, 4);
    std::memcpy(&fy, pts + i * step + oy, 4);
    std::memcpy(&fz, pts + i * step + oz, 4);
    V3 p3d{(double)fx, (double)fy, (double)fz};                    // :751
    V3 devi = p3d - cam;                                           // :754
    if (std::fabs(devi.x) < m.range.x && std::fabs(devi.y) < m.range.y && std::fabs(devi.z) < m.range.z) {  // :757
      for (int x = -inf_step; x <= inf_step; ++x)
        for (int y = -inf_step; y <= inf_step; ++y)
          for (int z = -inf_step_z; z <= inf_step_z; ++z) {        // :761-763
            V3 q{fx + x * m.res, fy + y * m.res, fz + z * m.res};  // :765-767 (float promoted to double)
            max_x = std::max(max_x, q.x); max_y = std::max(max_y, q.y); max_z = std::max(max_z, q.z);  // :769-771
            min_x = std::min(min_x, q.x); min_y = std::min(min_y, q.y); min_z = std::min(min_z, q.z);  // :773-775
            int id[3];
            pos_to_index(m, q, id);                                // :777
            if (!is_in_map_idx(m, id)) continue;                   // :779
            m.inflate[to_address(m, id[0], id[1], id[2])] = 1;     // :782-784 (idempotent: same byte value from every thread)
          }
    }
  }
  mm[tid] = MM{min_x, min_y, min_z, max_x, max_y, max_z, {0}};
  });
  double min_x = m.maxb.x, min_y = m.maxb.y, min_z = m.maxb.z, max_x = m.minb.x, max_y = m.minb.y, max_z = m.minb.z;
  for (auto& t : mm) {
    min_x = std::min(min_x, t.min_x); min_y = std::min(min_y, t.min_y); min_z = std::min(min_z, t.min_z);
    max_x = std::max(max_x, t.max_x); max_y = std::max(max_y, t.max_y); max_z = std::max(max_z, t.max_z);
  }
  min_x = std::min(min_x, cam.x); min_y = std::min(min_y, cam.y); min_z = std::min(min_z, cam.z);  // :789-791
  max_x = std::max(max_x, cam.x); max_y = std::max(max_y, cam.y); max_z = std::max(max_z, cam.z);  // :793-795
  max_z = std::max(max_z, m.ground_height);                        // :797
  pos_to_index(m, V3{max_x, max_y, max_z}, m.lb_max);              // :799
  pos_to_index(m, V3{min_x, min_y, min_z}, m.lb_min);              // :800
  bound_index(m, m.lb_min);                                        // :802
  bound_index(m, m.lb_max);                                        // :803
}

// grid_map.cpp:851-900 publishMapInflate: occupied inflated voxels of the local-bounds box (+- margin when all_info),
// in x -> y -> z order, voxel centres (grid_map.h:406-410 indexToPos) not above visualization_truncate_height, as
// PointXYZ records (float x, y, z and PCL's padding float 1.0f).
static int64_t export_inflate_cloud(const OMap& m, int margin, double truncate_height, float* out, int64_t cap) {
  int a[3], b[3];
  for (int i = 0; i < 3; ++i) { a[i] = m.lb_min[i] - margin; b[i] = m.lb_max[i] + margin; }
  bound_index(m, a);
  bound_index(m, b);
  int64_t n = 0;
  for (int x = a[0]; x <= b[0]; ++x)
    for (int y = a[1]; y <= b[1]; ++y)
      for (int z = a[2]; z <= b[2]; ++z) {
        if (m.inflate[to_address(m, x, y, z)] == 0) continue;
        const double px = (x + 0.5) * m.res + m.origin.x, py = (y + 0.5) * m.res + m.origin.y, pz = (z + 0.5) * m.res + m.origin.z;
        if (pz > truncate_height) continue;
        if (n < cap) { out[4 * n] = (float)px; out[4 * n + 1] = (float)py; out[4 * n + 2] = (float)pz; out[4 * n + 3] = 1.0f; }
        ++n;
      }
  return n;
}

// src/utils/LineOfSight.cpp:9-27 fastLOS.  Returns visible; *nsamples = samples evaluated.
static const double kLosStep = 0.1 / 2.0;  // LineOfSight.cpp:13-14 (resolution hard-coded 0.1)
// `sumq` (extension, cost variants only): adds the quantised cost of every sampled voxel.
static bool fast_los(const OMap& m, const V3& start, const V3& end, int64_t* nsamples, uint64_t* sumq = nullptr) {
  double distance = norm(end - start);          // :16
  V3 direction = normalized(end - start);       // :17
  int64_t k = 0;
  for (double d = 0; d <= distance; d += kLosStep) {  // :19
    V3 sample = start + direction * d;          // :20
    ++k;
    if (get_inflate_occupancy(m, sample)) {     // :21 (-1 outside map counts as blocked)
      if (nsamples) *nsamples += k;
      return false;
    }
    if (sumq) {
      int id[3];
      pos_to_index(m, sample, id);
      *sumq += m.costq[to_address(m, id[0], id[1], id[2])];
    }
  }
  if (nsamples) *nsamples += k;
  return true;
}

// ------------------------------------------------------------------------------------------
// New functionality (not in the reference): exact squared EDT, cost field
// ------------------------------------------------------------------------------------------
// ===================================================================================================
// Depth-image fusion: projectDepthImage -> raycastProcess -> clearAndInflateLocalMap,
// restated sequentially in the reference's order, quirks included.
// ===================================================================================================
// grid_map.cpp:34-41,57-62 + grid_map.h:27 (logit is a macro over libm's log)
static void depth_set_params(OMap& m, const OMap::DepthParams& p) {
  m.dp = p;
  m.dp_set = true;
  m.prob_hit_log = std::log(p.p_hit / (1 - p.p_hit));
  m.prob_miss_log = std::log(p.p_miss / (1 - p.p_miss));
  m.clamp_min_log = std::log(p.p_min / (1 - p.p_min));
  m.clamp_max_log = std::log(p.p_max / (1 - p.p_max));
  m.min_occupancy_log = std::log(p.p_occ / (1 - p.p_occ));
  m.unknown_flag = 0.01;
}
// grid_map.cpp:77-86: buffers and their initial values
static void depth_alloc(OMap& m) {
  if (!m.occ.empty()) return;
  const size_t n = (size_t)m.nvox();
  m.occ.assign(n, m.clamp_min_log - m.unknown_flag);
  m.cnt_all.assign(n, 0);
  m.cnt_hit.assign(n, 0);
  m.flag_rayend.assign(n, (char)-1);
  m.flag_traverse.assign(n, (char)-1);
}
// camera_r * p + camera_pos with the product evaluated row by row as (r0*x + r1*y) + r2*z (the header stand-in's
// order, oracle/shim/shim_eigen.h operator*; Eigen itself is not in this image: unpinned)
static inline V3 cam_to_world(const double R[9], const V3& p, const V3& t) {
  return {((R[0] * p.x + R[1] * p.y) + R[2] * p.z) + t.x, ((R[3] * p.x + R[4] * p.y) + R[5] * p.z) + t.y,
          ((R[6] * p.x + R[7] * p.y) + R[8] * p.z) + t.z};
}
// grid_map.cpp:195-315 projectDepthImage.  Returns the projected points in the reference's order.
static void depth_project(OMap& m, const uint16_t* img, int rows, int cols, const V3& cam, const double R[9], std::vector<V3>& out) {
  const OMap::DepthParams& p = m.dp;
  out.clear();
  if (!p.use_depth_filter) {  // :212-231
    for (int v = 0; v < rows; ++v)
      for (int u = 0; u < cols; ++u) {
        const double depth = img[(size_t)v * cols + u] / p.k_depth_scaling_factor;
        V3 q{(u - p.cx) * depth / p.fx, (v - p.cy) * depth / p.fy, depth};
        out.push_back(cam_to_world(R, q, cam));
      }
    return;
  }
  if (!m.has_first_depth) { m.has_first_depth = true; return; }  // :236-239: the first filtered frame projects nothing
  const double inv_factor = 1.0 / p.k_depth_scaling_factor;
  for (int v = p.depth_filter_margin; v < rows - p.depth_filter_margin; v += p.skip_pixel) {
    size_t at = (size_t)v * cols + p.depth_filter_margin;
    for (int u = p.depth_filter_margin; u < cols - p.depth_filter_margin; u += p.skip_pixel) {
      double depth = img[at] * inv_factor;
      at += p.skip_pixel;
      // :258-271 — the zero test reads the pixel the pointer was just advanced TO, not the one whose depth was taken
      if (img[at] == 0) depth = p.max_ray_length + 0.1;
      else if (depth < p.depth_filter_mindist) continue;
      else if (depth > p.depth_filter_maxdist) depth = p.max_ray_length + 0.1;
      V3 q{(u - p.cx) * depth / p.fx, (v - p.cy) * depth / p.fy, depth};
      out.push_back(cam_to_world(R, q, cam));
    }
  }
}
// grid_map.cpp:484-507 closetPointInMap
static V3 closest_point_in_map(const OMap& m, const V3& pt, const V3& cam) {
  const double diff[3] = {pt.x - cam.x, pt.y - cam.y, pt.z - cam.z};
  const double max_tc[3] = {m.maxb.x - cam.x, m.maxb.y - cam.y, m.maxb.z - cam.z};
  const double min_tc[3] = {m.minb.x - cam.x, m.minb.y - cam.y, m.minb.z - cam.z};
  double min_t = 1000000;
  for (int i = 0; i < 3; ++i)
    if (std::fabs(diff[i]) > 0) {
      const double t1 = max_tc[i] / diff[i];
      if (t1 > 0 && t1 < min_t) min_t = t1;
      const double t2 = min_tc[i] / diff[i];
      if (t2 > 0 && t2 < min_t) min_t = t2;
    }
  const double k = min_t - 1e-3;
  return {cam.x + k * diff[0], cam.y + k * diff[1], cam.z + k * diff[2]};
}
// raycast.cpp:32-49 signum / mod / intbound
static inline int ray_signum(int x) { return x == 0 ? 0 : x < 0 ? -1 : 1; }
static inline double ray_mod(double v, double mod) { return std::fmod(std::fmod(v, mod) + mod, mod); }
static double ray_intbound(double s, double ds) {
  if (ds < 0) return ray_intbound(-s, -ds);
  s = ray_mod(s, 1);
  return (1 - s) / ds;
}
// raycast.cpp:253-346 RayCaster::setInput / step
struct RayWalk {
  int x, y, z, ex, ey, ez, sx, sy, sz;
  double tmx, tmy, tmz, tdx, tdy, tdz;
  void set(const V3& a, const V3& b) {
    x = (int)std::floor(a.x); y = (int)std::floor(a.y); z = (int)std::floor(a.z);
    ex = (int)std::floor(b.x); ey = (int)std::floor(b.y); ez = (int)std::floor(b.z);
    const double dx = ex - x, dy = ey - y, dz = ez - z;
    sx = ray_signum((int)dx); sy = ray_signum((int)dy); sz = ray_signum((int)dz);
    tmx = ray_intbound(a.x, dx); tmy = ray_intbound(a.y, dy); tmz = ray_intbound(a.z, dz);
    tdx = ((double)sx) / dx; tdy = ((double)sy) / dy; tdz = ((double)sz) / dz;
  }
  bool step(int& ox, int& oy, int& oz) {
    ox = x; oy = y; oz = z;
    if (x == ex && y == ey && z == ez) return false;
    if (tmx < tmy) {
      if (tmx < tmz) { x += sx; tmx += tdx; } else { z += sz; tmz += tdz; }
    } else {
      if (tmy < tmz) { y += sy; tmy += tdy; } else { z += sz; tmz += tdz; }
    }
    return true;
  }
};
// 3-D DDA line of sight (ours): visible iff both end points are inside the map and no voxel the
// RayCaster walk visits between them (start and end voxel included) is occupied in the inflated grid; the walk runs on
// map-relative voxel coordinates (pos - origin) * res_inv.
static bool dda_los(const OMap& m, const V3& s, const V3& e, int* ncells) {
  *ncells = 0;
  if (!is_in_map(m, s) || !is_in_map(m, e)) return false;
  RayWalk rc;
  rc.set(V3{(s.x - m.origin.x) * m.res_inv, (s.y - m.origin.y) * m.res_inv, (s.z - m.origin.z) * m.res_inv},
         V3{(e.x - m.origin.x) * m.res_inv, (e.y - m.origin.y) * m.res_inv, (e.z - m.origin.z) * m.res_inv});
  const int cap = m.N[0] + m.N[1] + m.N[2] + 4;
  for (;;) {
    int id[3];
    const bool more = rc.step(id[0], id[1], id[2]);
    ++*ncells;
    if (!is_in_map_idx(m, id) || m.inflate[to_address(m, id[0], id[1], id[2])]) return false;
    if (!more) return true;
    if (*ncells > cap) return false;
  }
}
// grid_map.cpp:173-193 setCacheOccupancy (no bounds test in the reference: callers keep `pos` inside the map; here an
// address outside the buffer is skipped and reported as -1 instead of corrupting memory)
static int64_t set_cache_occupancy(OMap& m, const V3& pos, int occ, std::vector<int64_t>& queue) {
  int id[3];
  pos_to_index(m, pos, id);
  const int64_t a = to_address(m, id[0], id[1], id[2]);
  if (a < 0 || a >= m.nvox()) return -1;
  m.cnt_all[a] = (short)(m.cnt_all[a] + 1);
  if (m.cnt_all[a] == 1) queue.push_back(a);
  if (occ == 1) m.cnt_hit[a] = (short)(m.cnt_hit[a] + 1);
  return a;
}
// grid_map.cpp:356,366: (pt - cam) / length * max_ray_length + cam, component-wise, left to right
static inline V3 shorten_ray(const V3& pt, const V3& cam, double length, double maxlen) {
  return {(pt.x - cam.x) / length * maxlen + cam.x, (pt.y - cam.y) / length * maxlen + cam.y, (pt.z - cam.z) / length * maxlen + cam.z};
}
// grid_map.cpp:317-482 raycastProcess
static void depth_raycast(OMap& m, const std::vector<V3>& pts, const V3& cam) {
  if (pts.empty()) return;
  m.raycast_num = (signed char)(m.raycast_num + 1);
  const double maxlen = m.dp.max_ray_length;
  double min_x = m.maxb.x, min_y = m.maxb.y, min_z = m.maxb.z, max_x = m.minb.x, max_y = m.minb.y, max_z = m.minb.z;
  std::vector<int64_t> queue;
  RayWalk rc;
  for (const V3& p0 : pts) {
    V3 pt = p0;
    int64_t vox;
    if (!is_in_map(m, pt)) {
      pt = closest_point_in_map(m, pt, cam);
      const double length = norm(pt - cam);
      if (length > maxlen) pt = shorten_ray(pt, cam, length, maxlen);
      vox = set_cache_occupancy(m, pt, 0, queue);
    } else {
      const double length = norm(pt - cam);
      if (length > maxlen) {
        pt = shorten_ray(pt, cam, length, maxlen);
        vox = set_cache_occupancy(m, pt, 0, queue);
      } else {
        vox = set_cache_occupancy(m, pt, 1, queue);
      }
    }
    max_x = std::max(max_x, pt.x); max_y = std::max(max_y, pt.y); max_z = std::max(max_z, pt.z);
    min_x = std::min(min_x, pt.x); min_y = std::min(min_y, pt.y); min_z = std::min(min_z, pt.z);
    if (vox >= 0) {  // :385-395 one ray per end voxel and round (flags persist across frames: a voxel whose flag still
                     // holds this round's number from 256 frames ago — or the initial -1 in round 255 — looks taken)
      if (m.flag_rayend[vox] == (char)m.raycast_num) continue;
      m.flag_rayend[vox] = (char)m.raycast_num;
    }
    m.depth_stats[1]++;
    rc.set(V3{pt.x / m.res, pt.y / m.res, pt.z / m.res}, V3{cam.x / m.res, cam.y / m.res, cam.z / m.res});  // :397
    int ix, iy, iz;
    // a ray between two points of the map takes at most N0+N1+N2 steps; the cap only guards a walk that misses its end
    // cell (the reference would run on, out of bounds)
    for (int steps = 0, cap = m.N[0] + m.N[1] + m.N[2] + 4; steps < cap && rc.step(ix, iy, iz); ++steps) {
      const V3 tmp{(ix + 0.5) * m.res, (iy + 0.5) * m.res, (iz + 0.5) * m.res};
      m.depth_stats[2]++;
      const int64_t v = set_cache_occupancy(m, tmp, 0, queue);
      if (v >= 0) {  // :408-418 stop at a voxel another ray of this round already crossed
        if (m.flag_traverse[v] == (char)m.raycast_num) break;
        m.flag_traverse[v] = (char)m.raycast_num;
      }
    }
  }
  min_x = std::min(min_x, cam.x); min_y = std::min(min_y, cam.y); min_z = std::min(min_z, cam.z);
  max_x = std::max(max_x, cam.x); max_y = std::max(max_y, cam.y); max_z = std::max(max_z, cam.z);
  max_z = std::max(max_z, m.ground_height);
  pos_to_index(m, V3{max_x, max_y, max_z}, m.lb_max);
  pos_to_index(m, V3{min_x, min_y, min_z}, m.lb_min);
  bound_index(m, m.lb_min);
  bound_index(m, m.lb_max);
  m.local_updated = true;
  // :448-481 log-odds update of every voxel touched this round
  int min_id[3], max_id[3];
  pos_to_index(m, cam - m.range, min_id);
  pos_to_index(m, cam + m.range, max_id);
  bound_index(m, min_id);
  bound_index(m, max_id);
  const int64_t nyz = (int64_t)m.N[1] * m.N[2];
  for (size_t qi = 0; qi < queue.size(); ++qi) {
    const int64_t a = queue[qi];
    const int idx[3] = {(int)(a / nyz), (int)((a % nyz) / m.N[2]), (int)(a % m.N[2])};
    const double upd = m.cnt_hit[a] >= m.cnt_all[a] - m.cnt_hit[a] ? m.prob_hit_log : m.prob_miss_log;
    m.cnt_hit[a] = m.cnt_all[a] = 0;
    m.depth_stats[3]++;
    if (upd >= 0 && m.occ[a] >= m.clamp_max_log) continue;
    if (upd <= 0 && m.occ[a] <= m.clamp_min_log) { m.occ[a] = m.clamp_min_log; continue; }
    const bool in_local = idx[0] >= min_id[0] && idx[0] <= max_id[0] && idx[1] >= min_id[1] && idx[1] <= max_id[1] &&
                          idx[2] >= min_id[2] && idx[2] <= max_id[2];
    if (!in_local) m.occ[a] = m.clamp_min_log;
    m.occ[a] = std::min(std::max(m.occ[a] + upd, m.clamp_min_log), m.clamp_max_log);
  }
}
// grid_map.cpp:509-627 clearAndInflateLocalMap
static void depth_clear_and_inflate(OMap& m) {
  const int vec_margin = 5, lm = m.dp.local_map_margin;
  int min_cut[3], max_cut[3], min_cut_m[3], max_cut_m[3];
  for (int i = 0; i < 3; ++i) { min_cut[i] = m.lb_min[i] - lm; max_cut[i] = m.lb_max[i] + lm; }
  bound_index(m, min_cut);
  bound_index(m, max_cut);
  for (int i = 0; i < 3; ++i) { min_cut_m[i] = min_cut[i] - vec_margin; max_cut_m[i] = max_cut[i] + vec_margin; }
  bound_index(m, min_cut_m);
  bound_index(m, max_cut_m);
  // :529-579 the three slab loops together write every voxel of box M that lies outside box C
  const double unknown = m.clamp_min_log - m.unknown_flag;
  for (int x = min_cut_m[0]; x <= max_cut_m[0]; ++x)
    for (int y = min_cut_m[1]; y <= max_cut_m[1]; ++y)
      for (int z = min_cut_m[2]; z <= max_cut_m[2]; ++z) {
        const bool inside = x >= min_cut[0] && x <= max_cut[0] && y >= min_cut[1] && y <= max_cut[1] && z >= min_cut[2] && z <= max_cut[2];
        if (!inside) m.occ[to_address(m, x, y, z)] = unknown;
      }
  // :581-593
  const int step = (int)std::ceil(m.inflation / m.res);
  for (int x = m.lb_min[0]; x <= m.lb_max[0]; ++x)
    for (int y = m.lb_min[1]; y <= m.lb_max[1]; ++y)
      for (int z = m.lb_min[2]; z <= m.lb_max[2]; ++z) m.inflate[to_address(m, x, y, z)] = 0;
  // :595-616 + grid_map.h:412-441 inflatePoint: the full (2*step+1)^3 cube, and only the LINEAR address is range-tested,
  // so a cube sticking out in y or z wraps into the neighbouring row/column
  const int64_t total = m.nvox();
  for (int x = m.lb_min[0]; x <= m.lb_max[0]; ++x)
    for (int y = m.lb_min[1]; y <= m.lb_max[1]; ++y)
      for (int z = m.lb_min[2]; z <= m.lb_max[2]; ++z) {
        if (!(m.occ[to_address(m, x, y, z)] > m.min_occupancy_log)) continue;
        for (int dx = -step; dx <= step; ++dx)
          for (int dy = -step; dy <= step; ++dy)
            for (int dz = -step; dz <= step; ++dz) {
              const int64_t a = to_address(m, x + dx, y + dy, z + dz);
              if (a < 0 || a >= total) continue;
              m.inflate[a] = 1;
            }
      }
  // :618-627 virtual ceiling (no range test in the reference; an address outside the buffer is skipped here)
  if (m.dp.virtual_ceil_height > -0.5) {
    const int ceil_id = (int)std::floor((m.dp.virtual_ceil_height - m.origin.z) * m.res_inv);
    for (int x = m.lb_min[0]; x <= m.lb_max[0]; ++x)
      for (int y = m.lb_min[1]; y <= m.lb_max[1]; ++y) {
        const int64_t a = to_address(m, x, y, ceil_id);
        if (a >= 0 && a < total) m.inflate[a] = 1;
      }
  }
}
// publishMap (grid_map.cpp:806-849; which 0) / publishUnknown (:902-939; which 1): PointXYZ records (x, y, z, 1.0f)
static int64_t export_occupancy_cloud(OMap& m, int which, int margin, double truncate_height, float* out, int64_t cap) {
  depth_alloc(m);
  int a[3], b[3];
  for (int i = 0; i < 3; ++i) { a[i] = m.lb_min[i] - margin; b[i] = m.lb_max[i] + margin; }
  bound_index(m, a);
  bound_index(m, b);
  int64_t n = 0;
  for (int x = a[0]; x <= b[0]; ++x)
    for (int y = a[1]; y <= b[1]; ++y)
      for (int z = a[2]; z <= b[2]; ++z) {
        const double o = m.occ[to_address(m, x, y, z)];
        if (which == 0 ? (o < m.min_occupancy_log) : !(o < m.clamp_min_log - 1e-3)) continue;
        const double px = (x + 0.5) * m.res + m.origin.x, py = (y + 0.5) * m.res + m.origin.y, pz = (z + 0.5) * m.res + m.origin.z;
        if (pz > truncate_height) continue;
        if (n < cap) { out[4 * n] = (float)px; out[4 * n + 1] = (float)py; out[4 * n + 2] = (float)pz; out[4 * n + 3] = 1.0f; }
        ++n;
      }
  return n;
}
// depthPoseCallback's gate (grid_map.cpp:688-697) + updateOccupancyCallback (grid_map.cpp:635-666)
static int depth_update(OMap& m, const uint16_t* img, int rows, int cols, const V3& cam, const double R[9]) {
  if (!is_in_map(m, cam)) return 0;
  depth_alloc(m);
  std::vector<V3> pts;
  depth_project(m, img, rows, cols, cam, R, pts);
  m.depth_stats[0] = (int64_t)pts.size();
  m.depth_stats[1] = m.depth_stats[2] = m.depth_stats[3] = 0;
  depth_raycast(m, pts, cam);
  if (m.local_updated) depth_clear_and_inflate(m);
  m.local_updated = false;
  return 1;
}

static const int32_t kDistInf = 1 << 29;  // "no obstacle anywhere on this line" sentinel

// Brute force: d2[v] = min over occupied u of |v-u|^2 in voxel units (for tiny grids).
static void edt_brute(const OMap& m, std::vector<int32_t>& out) {
  const int Nx = m.N[0], Ny = m.N[1], Nz = m.N[2];
  std::vector<int> ox, oy, oz;
  for (int x = 0; x < Nx; ++x)
    for (int y = 0; y < Ny; ++y)
      for (int z = 0; z < Nz; ++z)
        if (m.inflate[to_address(m, x, y, z)]) { ox.push_back(x); oy.push_back(y); oz.push_back(z); }
  out.assign(m.nvox(), kDistInf);
  for (int x = 0; x < Nx; ++x)
    for (int y = 0; y < Ny; ++y)
      for (int z = 0; z < Nz; ++z) {
        int32_t best = kDistInf;
        for (size_t k = 0; k < ox.size(); ++k) {
          int dx = x - ox[k], dy = y - oy[k], dz = z - oz[k];
          int32_t d = dx * dx + dy * dy + dz * dz;
          if (d < best) best = d;
        }
        out[to_address(m, x, y, z)] = best;
      }
}

// One exact 1-D min-plus pass  out[u] = min_i f[i] + (u-i)^2  via the integer lower envelope of
// parabolas (Felzenszwalb & Huttenlocher 2012, integer separator as in Meijster et al. 2000).
static void edt_line(const int32_t* f, int32_t* out, int n, int stride, std::vector<int>& s, std::vector<int>& t) {
  auto F = [&](int x, int i) -> int64_t { return (int64_t)(x - i) * (x - i) + f[(int64_t)i * stride]; };
  auto Sep = [&](int i, int u) -> int64_t {
    return ((int64_t)u * u - (int64_t)i * i + f[(int64_t)u * stride] - f[(int64_t)i * stride]) / (2 * (int64_t)(u - i));
  };
  int q = 0;
  s[0] = 0; t[0] = 0;
  for (int u = 1; u < n; ++u) {
    while (q >= 0 && F(t[q], s[q]) > F(t[q], u)) --q;
    if (q < 0) { q = 0; s[0] = u; }
    else {
      int64_t w = 1 + Sep(s[q], u);
      if (w < n) { ++q; s[q] = u; t[q] = (int)w; }
    }
  }
  for (int u = n - 1; u >= 0; --u) {
    int64_t v = F(u, s[q]);
    out[(int64_t)u * stride] = v >= kDistInf ? kDistInf : (int32_t)v;
    if (u == t[q]) --q;
  }
}

static void edt_exact(const OMap& m, std::vector<int32_t>& out, int nthreads) {
  const int Nx = m.N[0], Ny = m.N[1], Nz = m.N[2];
  std::vector<int32_t> a(m.nvox()), b(m.nvox());
  // z pass straight from occupancy (two sweeps)
  parallel_chunks((int64_t)Nx * Ny, nthreads, 1024, [&](int64_t c0, int64_t c1, int) {
  for (int64_t c = c0; c < c1; ++c) {
    const char* col = &m.inflate[c * Nz];
    int32_t* o = &a[c * Nz];
    int last = -1;
    for (int z = 0; z < Nz; ++z) {
      if (col[z]) last = z;
      o[z] = last < 0 ? kDistInf : (z - last) * (z - last);
    }
    last = -1;
    for (int z = Nz - 1; z >= 0; --z) {
      if (col[z]) last = z;
      if (last >= 0) o[z] = std::min(o[z], (last - z) * (last - z));
    }
  }
  });
  // y pass, then x pass
  parallel_chunks((int64_t)Nx * Nz, nthreads, 1024, [&](int64_t l0, int64_t l1, int) {
    std::vector<int> s(Ny + 1), t(Ny + 1);
    for (int64_t l = l0; l < l1; ++l) {
      int x = (int)(l / Nz), z = (int)(l % Nz);
      edt_line(&a[to_address(m, x, 0, z)], &b[to_address(m, x, 0, z)], Ny, Nz, s, t);
    }
  });
  parallel_chunks((int64_t)Ny * Nz, nthreads, 1024, [&](int64_t l0, int64_t l1, int) {
    std::vector<int> s(Nx + 1), t(Nx + 1);
    for (int64_t l = l0; l < l1; ++l) edt_line(&b[l], &a[l], Nx, Ny * Nz, s, t);
  });
  out.swap(a);
}

// Cost field definition (ours; DESIGN.md "Cost field"):  c = max(0, 1 - sqrt(d2)*res/d_safe) in
// float32, quantised q = (uint16)(c*65535 + 0.5).  Planners read q only (integer => order-free sums).
static inline float cost_from_d2(int32_t d2, float res, float d_safe) {
  if (d2 >= kDistInf) return 0.0f;
  float d = sqrtf((float)d2) * res;
  if (d >= d_safe) return 0.0f;
  return 1.0f - d / d_safe;
}
static inline uint16_t costq_from_cost(float c) { return (uint16_t)(c * 65535.0f + 0.5f); }

// ------------------------------------------------------------------------------------------
// Planner restatement
// ------------------------------------------------------------------------------------------
enum Algo { ASTAR = 0, THETASTAR = 1, LAZYTHETASTAR = 2, COSTASTAR = 3, COSTLAZYTHETASTAR = 4, THETASTARAGR = 5, THETASTARAGRPP = 6 };
enum Mode { FAITHFUL = 0, CANONICAL = 1 };
// include/utils/utils.hpp:19-21
static const char IN_CLOSE_SET = 'a', IN_OPEN_SET = 'b', NOT_EXPAND = 'c';

struct Node {  // include/utils/utils.hpp:23-49 (live fields only)
  V3 position;
  double g_score, f_score;
  Node* parent;
  char node_state;
  // lazy variants only: the expanded node that last relaxed this one, and that step's length
  Node* via;
  double via_step;
  // AGR variants (utils.hpp:28,33): flying penalty carried in g, and rolling(0)/flying(1)
  double penalty_g_score;
  int motion_state;
  Node() : parent(nullptr), node_state(NOT_EXPAND), via(nullptr), via_step(0), penalty_g_score(0), motion_state(0) {}
};

// include/utils/utils.hpp:52-61: (f_score, pointer). Pool is one contiguous array here, so pointer
// order == creation order (what a fresh glibc heap gives the reference's per-node `new`, AStar.cpp:24-26).
struct NodeComparator {
  bool operator()(const Node* a, const Node* b) const {
    if (a->f_score != b->f_score) return a->f_score < b->f_score;
    return a < b;
  }
};

// include/utils/utils.hpp:63-102: unordered_map<Vector3d, NodePtr>; key equality is exact `==` on
// three doubles (so -0.0 == +0.0, and libstdc++ hash<double> maps both to 0).
struct Key {
  double x, y, z;
  bool operator==(const Key& o) const { return x == o.x && y == o.y && z == o.z; }
};
struct KeyHash {
  size_t operator()(const Key& k) const {
    size_t seed = 0;
    const double e[3] = {k.x, k.y, k.z};
    for (int i = 0; i < 3; ++i) seed ^= std::hash<double>()(e[i]) + 0x9e3779b9 + (seed << 6) + (seed >> 2);
    return seed;
  }
};

struct PlannerParams {
  double resolution;        // <algo>/resolution         (AStar.cpp:32)
  double lambda_heuristic;  // <algo>/lambda_heuristic   (AStar.cpp:34)
  int allocate_num;         // <algo>/allocate_num       (AStar.cpp:35)
  double cost_weight;       // setCostFactor (AlgorithmBase.hpp:39); cost variants only
  double max_los;           // GetPath.srv:26 max_los; <=0 = unlimited (extension; lazy/theta variants)
  // thetastaragr/* (ThetaStarAGR.cpp:33-36, ThetaStarAGRpp.cpp:33-37)
  double barrier, flying_cost, flying_cost_default, ground_judge, epsilon;
};

struct Result {  // mirrors b200_path_result_t (include/b200_planner.h)
  int32_t solved;
  int32_t status;  // 0 ok, 1 open set empty, 2 pool exhausted
  int32_t n_path;
  int32_t open_peak;  // largest open-list size seen (diagnostic; sizes the GPU workspace)
  int64_t path_offset;
  double g_final;
  double path_length;
  int64_t explored_nodes;
  int64_t iter_num;
  int64_t los_checks;
  int64_t los_samples;
  double goal_coords[3];
  double start_coords[3];
  double time_us;
};

struct Planner {
  const OMap* map = nullptr;
  PlannerParams pp{};
  Algo algo = ASTAR;
  Mode mode = CANONICAL;
  double tie_breaker = 1.0 + 1.0 / 10000;  // AStar.cpp:40
  std::vector<Node> pool;                  // AStar.cpp:23-26
  int use_node_num = 0, iter_num = 0;
  std::unordered_map<Key, Node*, KeyHash> expanded;          // AlgorithmBase.hpp:74
  std::set<Node*, NodeComparator> open_faithful;             // AlgorithmBase.hpp:77
  std::set<std::pair<double, Node*>> open_canon;             // canonical: true PQ on (f, creation idx)
  std::vector<Node*> path_nodes;
  unsigned los_checks = 0;
  int64_t los_samples = 0;

  void init() {  // AStar.cpp:14-29
    pool.assign(pp.allocate_num, Node());
    use_node_num = 0;
    iter_num = 0;
  }
  void reset() {  // AlgorithmBase.cpp:17-32
    expanded.clear();
    path_nodes.clear();
    open_faithful.clear();
    open_canon.clear();
    for (int i = 0; i < use_node_num; ++i) {
      pool[i].parent = nullptr;
      pool[i].node_state = NOT_EXPAND;
      pool[i].via = nullptr;
    }
    use_node_num = 0;
    iter_num = 0;
  }
  double heu(const V3& a, const V3& b) const { return tie_breaker * norm(b - a); }  // AStar.cpp:43-45
  double costq_at(const V3& p) const {  // cost variants: quantised cost of the voxel containing p
    int id[3];
    pos_to_index(*map, p, id);
    bound_index(*map, id);
    return (double)map->costq[to_address(*map, id[0], id[1], id[2])];
  }
  bool is_cost() const { return algo == COSTASTAR || algo == COSTLAZYTHETASTAR; }
  bool is_lazy() const { return algo == LAZYTHETASTAR || algo == COSTLAZYTHETASTAR; }
  // edge cost for the cost-aware variants (ours): |d| + w * |d| * (c(a)+c(b))/2, c = q/65535
  double edge(const V3& a, const V3& b, double len) const {
    if (!is_cost()) return len;
    double qa = costq_at(a), qb = costq_at(b);
    return len + pp.cost_weight * (len * ((qa + qb) / 131070.0));
  }

  void open_insert(Node* n) {
    if (mode == FAITHFUL) open_faithful.insert(n);
    else open_canon.insert({n->f_score, n});
  }
  bool open_empty() const { return mode == FAITHFUL ? open_faithful.empty() : open_canon.empty(); }
  Node* open_pop() {
    if (mode == FAITHFUL) {
      Node* n = *open_faithful.begin();  // AStar.cpp:105-106
      open_faithful.erase(open_faithful.begin());
      return n;
    }
    Node* n = open_canon.begin()->second;
    open_canon.erase(open_canon.begin());
    return n;
  }

  // AStar.cpp:56-78 (faithful: unconditional overwrite, insert-without-erase) /
  // canonical: guarded write, true decrease-key.
  void update_vertex_astar(Node* cur, Node* nb, const V3& d_pos, const V3& target) {
    double old_g = nb->g_score, old_f = nb->f_score;
    double new_g = cur->g_score + edge(cur->position, nb->position, norm(d_pos));
    double new_f = new_g + pp.lambda_heuristic * heu(nb->position, target);
    if (mode == FAITHFUL) {
      nb->parent = cur; nb->g_score = new_g; nb->f_score = new_f;  // AStar.cpp:60-62
      if (nb->g_score < old_g) {                                   // AStar.cpp:70
        if (nb->node_state != IN_OPEN_SET) nb->node_state = IN_OPEN_SET;
        open_faithful.insert(nb);                                  // AStar.cpp:71-76
      }
      return;
    }
    if (new_g < old_g) {
      if (nb->node_state == IN_OPEN_SET) open_canon.erase({old_f, nb});
      nb->parent = cur; nb->g_score = new_g; nb->f_score = new_f;
      nb->node_state = IN_OPEN_SET;
      open_canon.insert({new_f, nb});
    }
  }

  // ThetaStar.cpp:37-89
  void update_vertex_theta(Node* cur, Node* nb, const V3& d_pos, const V3& target) {
    double old_g = nb->g_score, old_f = nb->f_score;
    // ComputeCost (:37-70)
    if (cur->parent == nullptr) {  // :39-47
      double direct = norm(d_pos);
      if ((cur->g_score + direct) < nb->g_score) {
        nb->parent = cur;
        nb->g_score = cur->g_score + direct;
        nb->f_score = nb->g_score + pp.lambda_heuristic * heu(nb->position, target);
      }
    } else {
      double distanceParent = norm(cur->parent->position - nb->position);  // :49
      los_checks++;                                                        // :50
      bool vis = (pp.max_los > 0 && distanceParent > pp.max_los) ? false
                                                                  : fast_los(*map, cur->parent->position, nb->position, &los_samples);
      if (vis) {  // :51-60
        double tmp = cur->parent->g_score + distanceParent;
        if (tmp < nb->g_score) {
          nb->parent = cur->parent;
          nb->g_score = tmp;
          nb->f_score = nb->g_score + pp.lambda_heuristic * heu(nb->position, target);
        }
      } else {  // :61-69
        if ((cur->g_score + norm(d_pos)) < nb->g_score) {
          nb->parent = cur;
          nb->g_score = cur->g_score + norm(d_pos);
          nb->f_score = nb->g_score + pp.lambda_heuristic * heu(nb->position, target);
        }
      }
    }
    // UpdateVertex (:72-89)
    if (nb->g_score < old_g) {
      if (mode == FAITHFUL) {
        if (nb->node_state == IN_OPEN_SET) {
          open_faithful.erase(nb);  // :80 — lookup by the ALREADY MUTATED key (may miss)
          open_faithful.insert(nb);
        } else {
          nb->node_state = IN_OPEN_SET;
          open_faithful.insert(nb);
        }
      } else {
        if (nb->node_state == IN_OPEN_SET) open_canon.erase({old_f, nb});
        nb->node_state = IN_OPEN_SET;
        open_canon.insert({nb->f_score, nb});
      }
    }
  }

  // shared tail of ThetaStar/AGR::UpdateVertex (ThetaStar.cpp:76-88, ThetaStarAGR.cpp:109-120, ThetaStarAGRpp.cpp:115-126)
  void reinsert_if_improved(Node* nb, double old_g, double old_f) {
    if (!(nb->g_score < old_g)) return;
    if (mode == FAITHFUL) {
      if (nb->node_state == IN_OPEN_SET) {
        open_faithful.erase(nb);  // lookup by the ALREADY MUTATED key (may miss)
        open_faithful.insert(nb);
      } else {
        nb->node_state = IN_OPEN_SET;
        open_faithful.insert(nb);
      }
    } else {
      if (nb->node_state == IN_OPEN_SET) open_canon.erase({old_f, nb});
      nb->node_state = IN_OPEN_SET;
      open_canon.insert({nb->f_score, nb});
    }
  }
  // ThetaStarAGR.cpp:49-103
  void update_vertex_agr(Node* cur, Node* nb, const V3& d_pos, const V3& target) {
    const double old_g = nb->g_score, old_f = nb->f_score;
    double direct_cost = norm(d_pos);                                            // :51
    double tmp_g_score = cur->g_score + direct_cost;                             // :52
    double penalty_g_score = 0.0;
    bool next_motion_state = (nb->position.z >= pp.ground_judge);                // :54
    if (next_motion_state) {                                                     // :56-62
      penalty_g_score = pp.flying_cost * nb->position.z / pp.barrier + pp.flying_cost_default;
      tmp_g_score -= cur->penalty_g_score;
      tmp_g_score += penalty_g_score;
    } else {
      tmp_g_score -= cur->penalty_g_score;
    }
    auto assign = [&](Node* parent, double g) {
      nb->parent = parent;
      nb->g_score = g;
      nb->f_score = nb->g_score + pp.lambda_heuristic * heu(nb->position, target);
      nb->motion_state = next_motion_state;
      nb->penalty_g_score = penalty_g_score;
    };
    if (cur->parent == nullptr) {                                                // :64-73
      if (tmp_g_score < nb->g_score) assign(cur, tmp_g_score);
    } else {
      los_checks++;                                                              // :77
      if (fast_los(*map, cur->parent->position, nb->position, &los_samples)) {   // :78
        double los_distance = norm(cur->parent->position - nb->position);        // :80
        tmp_g_score = cur->parent->g_score + los_distance;                       // :81
        if (next_motion_state) {                                                 // :82-87
          tmp_g_score -= cur->parent->penalty_g_score;
          tmp_g_score += penalty_g_score;
        } else {
          tmp_g_score -= cur->parent->penalty_g_score;
        }
        if (tmp_g_score < nb->g_score) assign(cur->parent, tmp_g_score);         // :88-95
      } else if (tmp_g_score < nb->g_score) {                                    // :96-102
        assign(cur, tmp_g_score);
      }
    }
    reinsert_if_improved(nb, old_g, old_f);
  }
  // ThetaStarAGRpp.cpp:52-109.  std::max(a, b) is (a < b) ? b : a — matters for the NaN of 0/0 at z == 0.
  static double stdmax(double a, double b) { return (a < b) ? b : a; }
  void update_vertex_agrpp(Node* cur, Node* nb, const V3& d_pos, const V3& target) {
    const double old_g = nb->g_score, old_f = nb->f_score;
    double direct_cost = norm(d_pos);                                            // :54
    bool next_motion_state = (nb->position.z >= pp.ground_judge);                // :56
    double delta_z = nb->position.z - cur->position.z;                           // :62
    double alpha = stdmax(1 + pp.flying_cost * delta_z / nb->position.z, pp.epsilon);   // :63
    direct_cost *= (next_motion_state || cur->motion_state == 1) ? alpha : pp.epsilon;  // :64
    double tmp_g_score = cur->g_score + direct_cost;                             // :66
    if (cur->motion_state == 0 && next_motion_state) tmp_g_score += pp.flying_cost_default;  // :69-71
    double best_cost = tmp_g_score;
    Node* best_parent = cur;
    if (cur->parent != nullptr) {                                                // :76
      los_checks++;
      if (fast_los(*map, cur->parent->position, nb->position, &los_samples)) {   // :78
        double los_distance = norm(cur->parent->position - nb->position);
        double delta_los_z = nb->position.z - cur->parent->position.z;
        double alpha_los = stdmax(1 + pp.flying_cost * delta_los_z / nb->position.z, pp.epsilon);
        los_distance *= (next_motion_state || cur->parent->motion_state == 1) ? alpha_los : pp.epsilon;
        double tmp_g_score_parent = cur->parent->g_score + los_distance;
        if (cur->parent->motion_state == 0 && next_motion_state) tmp_g_score_parent += pp.flying_cost_default;
        if (tmp_g_score_parent < best_cost) { best_cost = tmp_g_score_parent; best_parent = cur->parent; }
      }
    }
    if (best_cost < nb->g_score) {                                               // :101-108
      nb->parent = best_parent;
      nb->g_score = best_cost;
      nb->f_score = nb->g_score + pp.lambda_heuristic * heu(nb->position, target);
      nb->motion_state = next_motion_state ? 1 : 0;
    }
    reinsert_if_improved(nb, old_g, old_f);
  }

  // Lazy Theta* (Nash, Koenig, Tovey 2010) on this code base's conventions — OURS, parity unpinned.
  // Generation: optimistic path-2 from parent(cur) (no LoS); verification at expansion (set_vertex).
  void update_vertex_lazy(Node* cur, Node* nb, const V3& d_pos, const V3& target) {
    double old_g = nb->g_score, old_f = nb->f_score;
    Node* par = cur->parent ? cur->parent : cur;
    double len = cur->parent ? norm(par->position - nb->position) : norm(d_pos);
    double cand = par->g_score + edge(par->position, nb->position, len);
    if (cand < nb->g_score) {
      nb->parent = par;
      nb->g_score = cand;
      nb->f_score = cand + pp.lambda_heuristic * heu(nb->position, target);
      nb->via = cur;
      nb->via_step = norm(d_pos);
    }
    if (nb->g_score < old_g) {
      if (nb->node_state == IN_OPEN_SET) open_canon.erase({old_f, nb});
      nb->node_state = IN_OPEN_SET;
      open_canon.insert({nb->f_score, nb});
    }
  }
  // SetVertex at expansion: verify the optimistic parent with one fastLOS; on failure fall back to the
  // best CLOSED lattice neighbour (path 1).  `via` (always closed) is the first candidate, then the
  // 26 exact-key lookups in expansion order; strict < replaces.
  void set_vertex(Node* s, const V3& target, const double dvals[3]) {
    if (s->parent == nullptr || s->parent == s->via) return;  // start, or already a path-1 edge
    los_checks++;
    double len = norm(s->parent->position - s->position);
    uint64_t sumq = 0;
    bool vis = (pp.max_los > 0 && len > pp.max_los)
                   ? false
                   : fast_los(*map, s->parent->position, s->position, &los_samples, is_cost() ? &sumq : nullptr);
    if (vis) {
      if (is_cost()) {  // replace the optimistic trapezoid by the cost integrated over the LoS samples
        s->g_score = s->parent->g_score + (len + pp.cost_weight * (kLosStep * ((double)sumq / 65535.0)));
        s->f_score = s->g_score + pp.lambda_heuristic * heu(s->position, target);
      }
      return;
    }
    Node* best = s->via;
    double best_g = s->via->g_score + edge(s->via->position, s->position, s->via_step);
    for (int a = 0; a < 3; ++a)
      for (int b = 0; b < 3; ++b)
        for (int c = 0; c < 3; ++c) {
          V3 d{dvals[a], dvals[b], dvals[c]};
          if (norm(d) < 1e-3) continue;
          V3 np = s->position + d;
          auto it = expanded.find(Key{np.x, np.y, np.z});
          if (it == expanded.end() || it->second->node_state != IN_CLOSE_SET) continue;
          double g = it->second->g_score + edge(it->second->position, s->position, norm(d));
          if (g < best_g) { best_g = g; best = it->second; }
        }
    s->parent = best;
    s->g_score = best_g;
    s->f_score = best_g + pp.lambda_heuristic * heu(s->position, target);
  }

  // geometry_utils.cpp:10-20 calculatePathLength (float accumulation; delta is multiplied by res)
  static float path_length(const std::vector<V3>& p, double res) {
    float len = 0;
    for (size_t i = 1; i < p.size(); ++i) {
      double a = (p[i].x - p[i - 1].x) * res, b = (p[i].y - p[i - 1].y) * res, c = (p[i].z - p[i - 1].z) * res;
      len += sqrtf(a * a + b * b + c * c);  // pow(v,2) == v*v
    }
    return len;
  }

  // AStar.cpp:80-179 findPath (+ AlgorithmBase.cpp:123-175 createResultDataObject, :202-212 retrievePath)
  void find_path(const V3& source, const V3& target, Result* r, std::vector<V3>* path_out) {
    auto t0 = std::chrono::steady_clock::now();
    los_checks = 0;  // :84
    los_samples = 0;
    const double res = pp.resolution;
    // neighbour offsets exactly as the loop `for (double dx = -r; dx <= r + 1e-3; dx += r)` (:131-133)
    double dvals[8];
    int nd = 0;
    for (double dx = -res; dx <= res + 1e-3 && nd < 8; dx += res) dvals[nd++] = dx;

    Node* start = &pool[0];  // :87-92
    start->parent = nullptr;
    start->position = source;
    start->g_score = 0.0;
    start->f_score = pp.lambda_heuristic * heu(start->position, target);
    start->node_state = IN_OPEN_SET;
    start->via = nullptr;
    const bool agr = algo == THETASTARAGR || algo == THETASTARAGRPP;
    if (agr) {  // ThetaStarAGR.cpp:141-155, ThetaStarAGRpp.cpp:146-155 (the set holds one element, so mutating after
                // the insert, as the reference does, is the same as inserting the final key)
      if (start->position.z < 0) start->position.z = 0;
      if (start->position.z >= pp.ground_judge) {
        if (algo == THETASTARAGR) {
          const double pen = pp.flying_cost * start->position.z / pp.barrier + pp.flying_cost_default;
          start->g_score += pen;
          start->f_score += pen;
          start->penalty_g_score = pen;
        }
        start->motion_state = 1;
      } else {
        start->motion_state = 0;
        if (algo == THETASTARAGR) start->penalty_g_score = 0;
      }
    }
    open_insert(start);  // :94
    use_node_num += 1;   // :95
    expanded.insert({Key{start->position.x, start->position.y, start->position.z}, start});  // :97 (AGR: clamped z)

    Node* cur = nullptr;
    Node* last = nullptr;
    size_t open_peak = 1;
    bool solved = false;
    int status = 1;
    while (!open_empty()) {  // :103
      open_peak = std::max(open_peak, mode == FAITHFUL ? open_faithful.size() : open_canon.size());
      cur = open_pop();      // :105-106
      if (is_lazy()) set_vertex(cur, target, dvals);
      bool reach_end = std::abs(cur->position.x - target.x) <= res && std::abs(cur->position.y - target.y) <= res &&
                       std::abs(cur->position.z - target.z) <= res;  // :110-112
      if (reach_end) {  // :114-119
        Node* c = cur;  // retrievePath AlgorithmBase.cpp:202-212
        path_nodes.push_back(c);
        while (c->parent != nullptr) { c = c->parent; path_nodes.push_back(c); }
        std::reverse(path_nodes.begin(), path_nodes.end());
        solved = true; status = 0; last = cur;
        break;
      }
      cur->node_state = IN_CLOSE_SET;  // :122
      iter_num += 1;                   // :123
      V3 cur_pos = cur->position;      // :126
      bool oom = false;
      for (int a = 0; a < nd && !oom; ++a)
        for (int b = 0; b < nd && !oom; ++b)
          for (int c = 0; c < nd && !oom; ++c) {  // :131-133
            V3 d_pos{dvals[a], dvals[b], dvals[c]};
            if (norm(d_pos) < 1e-3) continue;      // :136
            V3 nb_pos = cur_pos + d_pos;           // :138
            if (!is_in_map(*map, nb_pos)) continue;  // :141
            Node* nb = nullptr;
            auto it = expanded.find(Key{nb_pos.x, nb_pos.y, nb_pos.z});  // :146
            if (it != expanded.end()) nb = it->second;
            if (nb != nullptr && nb->node_state == IN_CLOSE_SET) continue;  // :147
            if (nb != nullptr && algo == THETASTARAGRPP) continue;          // ThetaStarAGRpp.cpp:208-210: never revisit a node
            if (get_inflate_occupancy(*map, nb_pos) == 1) continue;         // :151 (== true)
            if (nb == nullptr) {  // :156-169
              nb = &pool[use_node_num];
              nb->position = nb_pos;
              nb->g_score = std::numeric_limits<double>::infinity();
              nb->f_score = std::numeric_limits<double>::infinity();
              if (algo == THETASTARAGRPP) nb->motion_state = (nb_pos.z >= pp.ground_judge) ? 1 : 0;  // ThetaStarAGRpp.cpp:216
              expanded.insert({Key{nb_pos.x, nb_pos.y, nb_pos.z}, nb});
              use_node_num += 1;
              if (use_node_num == pp.allocate_num) { oom = true; break; }  // :165-168
            }
            switch (algo) {  // :170 (virtual UpdateVertex)
              case ASTAR: case COSTASTAR: update_vertex_astar(cur, nb, d_pos, target); break;
              case THETASTAR: update_vertex_theta(cur, nb, d_pos, target); break;
              case THETASTARAGR: update_vertex_agr(cur, nb, d_pos, target); break;
              case THETASTARAGRPP: update_vertex_agrpp(cur, nb, d_pos, target); break;
              default: update_vertex_lazy(cur, nb, d_pos, target); break;
            }
          }
      if (oom) { status = 2; last = cur; break; }
    }
    if (!solved && status == 1) last = cur;  // :175-178 (open set empty)
    // createResultDataObject AlgorithmBase.cpp:123-175
    r->solved = solved;
    r->status = status;
    r->open_peak = (int32_t)std::min<size_t>(open_peak, 0x7fffffff);
    r->explored_nodes = use_node_num;
    r->iter_num = iter_num;
    r->los_checks = los_checks;
    r->los_samples = los_samples;
    r->g_final = last ? last->g_score : 0.0;
    const V3 lp = last ? last->position : source;
    r->goal_coords[0] = lp.x; r->goal_coords[1] = lp.y; r->goal_coords[2] = lp.z;  // :130 (last node, not the goal)
    r->start_coords[0] = source.x; r->start_coords[1] = source.y; r->start_coords[2] = source.z;
    std::vector<V3> path;
    if (solved) for (Node* n : path_nodes) path.push_back(n->position);  // AlgorithmBase.cpp:177-183
    r->n_path = (int)path.size();
    r->path_length = (double)path_length(path, map->res);  // AlgorithmBase.cpp:161 (grid resolution)
    r->time_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    if (path_out) *path_out = path;
  }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// C interface for ctypes (tests / bench cpu_baseline only)
// ------------------------------------------------------------------------------------------
extern "C" {

int orc_max_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

void* orc_map_create(const double origin[3], const double size[3], double res, double inflation, const double range[3],
                     double ground_height) {
  OMap* m = new OMap();
  m->res = res;
  m->res_inv = 1 / res;                                   // grid_map.cpp:52
  m->origin = {origin[0], origin[1], origin[2]};          // explicit origin; reference: grid_map.cpp:53
  m->size = {size[0], size[1], size[2]};                  // :54
  const double sz[3] = {size[0], size[1], size[2]};
  for (int i = 0; i < 3; ++i) m->N[i] = (int)std::ceil(sz[i] / res);  // :69-70
  m->minb = m->origin;                                    // :72
  m->maxb = m->origin + m->size;                          // :73
  m->range = {range[0], range[1], range[2]};
  m->inflation = inflation;
  m->ground_height = ground_height;
  m->inflate.assign(m->nvox(), 0);                        // :80
  for (int i = 0; i < 3; ++i) { m->lb_min[i] = 0; m->lb_max[i] = m->N[i] - 1; }
  return m;
}
void orc_map_destroy(void* h) { delete (OMap*)h; }
void orc_map_dims(void* h, int out[3]) { for (int i = 0; i < 3; ++i) out[i] = ((OMap*)h)->N[i]; }
void orc_map_update_cloud(void* h, const void* pts, int64_t n, int step, int ox, int oy, int oz, const double cam[3],
                          int nthreads) {
  cloud_update(*(OMap*)h, (const uint8_t*)pts, n, step, ox, oy, oz, V3{cam[0], cam[1], cam[2]}, nthreads < 1 ? 1 : nthreads);
}
void orc_map_clear_box(void* h, const double mn[3], const double mx[3]) {
  reset_buffer(*(OMap*)h, V3{mn[0], mn[1], mn[2]}, V3{mx[0], mx[1], mx[2]});
}
void orc_map_download_inflate(void* h, uint8_t* dst) {
  OMap* m = (OMap*)h;
  std::memcpy(dst, m->inflate.data(), m->nvox());
}
void orc_map_upload_inflate(void* h, const uint8_t* src) {
  OMap* m = (OMap*)h;
  std::memcpy(m->inflate.data(), src, m->nvox());
}
void orc_map_local_bounds(void* h, int mn[3], int mx[3]) {
  OMap* m = (OMap*)h;
  for (int i = 0; i < 3; ++i) { mn[i] = m->lb_min[i]; mx[i] = m->lb_max[i]; }
}
int64_t orc_map_export_inflate_cloud(void* h, int margin, double truncate_height, float* out, int64_t cap) {
  return export_inflate_cloud(*(OMap*)h, margin, truncate_height, out, cap);
}
void orc_map_set_depth_params(void* h, const void* p) { depth_set_params(*(OMap*)h, *(const OMap::DepthParams*)p); }
int orc_map_update_depth(void* h, const uint16_t* img, int rows, int cols, const double cam[3], const double R[9]) {
  OMap& m = *(OMap*)h;
  if (!m.dp_set) return -1;
  return depth_update(m, img, rows, cols, V3{cam[0], cam[1], cam[2]}, R);
}
void orc_map_download_occupancy(void* h, double* dst) {
  OMap& m = *(OMap*)h;
  depth_alloc(m);
  memcpy(dst, m.occ.data(), m.occ.size() * sizeof(double));
}
void orc_map_set_raycast_num(void* h, int v) { ((OMap*)h)->raycast_num = (signed char)v; }
int orc_map_raycast_num(void* h) { return ((OMap*)h)->raycast_num; }
void orc_map_depth_stats(void* h, int64_t out[8]) { memcpy(out, ((OMap*)h)->depth_stats, sizeof(int64_t) * 8); }
// grid_map.h:339-348 getOccupancy(Vector3d)
void orc_map_get_occupancy(void* h, const double* pos, int64_t n, int8_t* out) {
  OMap& m = *(OMap*)h;
  depth_alloc(m);
  for (int64_t i = 0; i < n; ++i) {
    const V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    if (!is_in_map(m, p)) { out[i] = -1; continue; }
    int id[3];
    pos_to_index(m, p, id);
    out[i] = m.occ[to_address(m, id[0], id[1], id[2])] > m.min_occupancy_log ? 1 : 0;
  }
}
int64_t orc_map_export_occupancy_cloud(void* h, int which, int margin, double truncate_height, float* out, int64_t cap) {
  return export_occupancy_cloud(*(OMap*)h, which, margin, truncate_height, out, cap);
}
void orc_los_batch_dda(void* h, const double* s, const double* e, int64_t n, uint8_t* vis, int32_t* ncells) {
  const OMap& m = *(OMap*)h;
  for (int64_t i = 0; i < n; ++i) {
    int nc = 0;
    vis[i] = dda_los(m, V3{s[3 * i], s[3 * i + 1], s[3 * i + 2]}, V3{e[3 * i], e[3 * i + 1], e[3 * i + 2]}, &nc) ? 1 : 0;
    if (ncells) ncells[i] = nc;
  }
}
// the voxel sequence of one walk (lattice coordinates in, integer cells out), for pinning against the reference's RayCaster
int64_t orc_ray_cells(const double a[3], const double b[3], int32_t* out, int64_t cap) {
  RayWalk rc;
  rc.set(V3{a[0], a[1], a[2]}, V3{b[0], b[1], b[2]});
  int64_t n = 0;
  int x, y, z;
  while (n < 1000000 && rc.step(x, y, z)) {
    if (n < cap) { out[3 * n] = x; out[3 * n + 1] = y; out[3 * n + 2] = z; }
    ++n;
  }
  return n;
}
void orc_map_get_inflate_occupancy(void* h, const double* pos, int64_t n, int8_t* out) {
  OMap* m = (OMap*)h;
  for (int64_t i = 0; i < n; ++i) out[i] = (int8_t)get_inflate_occupancy(*m, V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
}
// grid_map.h:310-321 setOccupied(Vector3d): inflated buffer <- 1 where isInMap
void orc_map_set_occupied(void* h, const double* pos, int64_t n) {
  OMap& m = *(OMap*)h;
  for (int64_t i = 0; i < n; ++i) {
    const V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    if (!is_in_map(m, p)) continue;
    int id[3];
    pos_to_index(m, p, id);
    m.inflate[(size_t)id[0] * m.N[1] * m.N[2] + (size_t)id[1] * m.N[2] + id[2]] = 1;
  }
}
// grid_map.h:323-337 setOccupancy(Vector3d, double): occupancy_buffer_ <- occ, values other than 0 / 1 rejected
void orc_map_set_occupancy(void* h, const double* pos, const double* occ, int64_t n) {
  OMap& m = *(OMap*)h;
  depth_alloc(m);
  for (int64_t i = 0; i < n; ++i) {
    if (occ[i] != 1 && occ[i] != 0) continue;
    const V3 p{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]};
    if (!is_in_map(m, p)) continue;
    int id[3];
    pos_to_index(m, p, id);
    m.occ[to_address(m, id[0], id[1], id[2])] = occ[i];
  }
}
void orc_map_is_in_map(void* h, const double* pos, int64_t n, uint8_t* out) {
  OMap* m = (OMap*)h;
  for (int64_t i = 0; i < n; ++i) out[i] = is_in_map(*m, V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]});
}
void orc_pos_to_index(void* h, const double* pos, int64_t n, int32_t* out) {
  OMap* m = (OMap*)h;
  for (int64_t i = 0; i < n; ++i) pos_to_index(*m, V3{pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]}, out + 3 * i);
}
void orc_los_batch(void* h, const double* s, const double* e, int64_t n, uint8_t* vis, int64_t* nsamples, int nthreads) {
  OMap* m = (OMap*)h;
  parallel_chunks(n, nthreads, 64, [&](int64_t i0, int64_t i1, int) {
    for (int64_t i = i0; i < i1; ++i) {
      int64_t k = 0;
      vis[i] = fast_los(*m, V3{s[3 * i], s[3 * i + 1], s[3 * i + 2]}, V3{e[3 * i], e[3 * i + 1], e[3 * i + 2]}, &k);
      if (nsamples) nsamples[i] = k;
    }
  });
}
// the d_k sequence of LineOfSight.cpp:19 (d += 0.05 by repeated addition)
void orc_los_dk(double* out, int64_t n) {
  double d = 0;
  for (int64_t i = 0; i < n; ++i) { out[i] = d; d += kLosStep; }
}

void orc_edt_brute(void* h, int32_t* out) {
  std::vector<int32_t> v;
  edt_brute(*(OMap*)h, v);
  std::memcpy(out, v.data(), v.size() * 4);
}
// builds dist2 (and costq when d_safe > 0) inside the map; copies dist2 to `out` when non-null
void orc_map_build_edt(void* h, double d_safe, int32_t* out, uint16_t* costq_out, float* cost_out, int nthreads) {
  OMap* m = (OMap*)h;
  edt_exact(*m, m->dist2, nthreads < 1 ? 1 : nthreads);
  if (out) std::memcpy(out, m->dist2.data(), m->dist2.size() * 4);
  if (d_safe > 0) {
    m->costq.resize(m->nvox());
    const float resf = (float)m->res, dsf = (float)d_safe;
    for (int64_t i = 0; i < m->nvox(); ++i) {
      float c = cost_from_d2(m->dist2[i], resf, dsf);
      m->costq[i] = costq_from_cost(c);
      if (cost_out) cost_out[i] = c;
    }
    if (costq_out) std::memcpy(costq_out, m->costq.data(), m->costq.size() * 2);
  }
}

void* orc_planner_create(int algo, int mode, double resolution, double lambda_heuristic, int allocate_num,
                         double cost_weight, double max_los, const double agr[5]) {
  Planner* p = new Planner();
  p->algo = (Algo)algo;
  p->mode = (Mode)mode;
  p->pp = {resolution, lambda_heuristic, allocate_num, cost_weight, max_los, agr[0], agr[1], agr[2], agr[3], agr[4]};
  p->init();
  return p;
}
void orc_planner_destroy(void* p) { delete (Planner*)p; }
void orc_planner_set_map(void* p, void* m) { ((Planner*)p)->map = (OMap*)m; }
// one query = reset() + findPath() as the node does (planner_ros_node.cpp:136,191-192)
int orc_planner_find_path(void* ph, const double s[3], const double g[3], Result* r, double* path, int64_t path_cap) {
  Planner* p = (Planner*)ph;
  p->reset();
  std::vector<V3> out;
  p->find_path(V3{s[0], s[1], s[2]}, V3{g[0], g[1], g[2]}, r, &out);
  r->path_offset = 0;
  int64_t n = std::min<int64_t>((int64_t)out.size(), path_cap);
  for (int64_t i = 0; i < n; ++i) { path[3 * i] = out[i].x; path[3 * i + 1] = out[i].y; path[3 * i + 2] = out[i].z; }
  return 0;
}
// AlgorithmBase.cpp:185-189 getVisitedNodes of the last find_path: pool[0 .. use_node_num - 1)
int64_t orc_planner_visited(void* ph, double* out, int64_t cap) {
  Planner* p = (Planner*)ph;
  const int64_t n = std::max(p->use_node_num - 1, 0);
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    out[3 * i] = p->pool[i].position.x; out[3 * i + 1] = p->pool[i].position.y; out[3 * i + 2] = p->pool[i].position.z;
  }
  return n;
}
// batched: queries are independent; one private planner per thread (reference is single-threaded;
// this is the "ref-NT" baseline of BASELINE.md §3).  Paths go to path[qi*path_stride ...].
int orc_planner_find_paths(int algo, int mode, double resolution, double lambda_heuristic, int allocate_num,
                           double cost_weight, double max_los, const double agr[5], void* map, const double* starts, const double* goals,
                           int64_t nq, Result* res, double* path, int64_t path_stride, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  std::vector<Planner> planners(nthreads);
  for (auto& p : planners) {
    p.algo = (Algo)algo;
    p.mode = (Mode)mode;
    p.pp = {resolution, lambda_heuristic, allocate_num, cost_weight, max_los, agr[0], agr[1], agr[2], agr[3], agr[4]};
    p.map = (OMap*)map;
    p.init();
  }
  parallel_chunks(nq, nthreads, 1, [&](int64_t q0, int64_t q1, int tid) {
    Planner& p = planners[tid];
    for (int64_t q = q0; q < q1; ++q) {
      p.reset();
      std::vector<V3> out;
      p.find_path(V3{starts[3 * q], starts[3 * q + 1], starts[3 * q + 2]}, V3{goals[3 * q], goals[3 * q + 1], goals[3 * q + 2]},
                  &res[q], &out);
      res[q].path_offset = q * path_stride;
      if (path) {
        int64_t n = std::min<int64_t>((int64_t)out.size(), path_stride);
        for (int64_t i = 0; i < n; ++i) {
          path[3 * (q * path_stride + i)] = out[i].x;
          path[3 * (q * path_stride + i) + 1] = out[i].y;
          path[3 * (q * path_stride + i) + 2] = out[i].z;
        }
      }
    }
  });
  return 0;
}

}  // extern "C"
