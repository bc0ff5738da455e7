// api.cu — the C-ABI of include/b200_planner.h: handles, device buffers, streams and launch sequencing.
// No torch types, no exceptions across the boundary, no CPU fallback.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

using namespace b200;

static thread_local std::string g_err;
static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  char buf[512];
  if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
  else snprintf(buf, sizeof buf, "%s", what);
  g_err = buf;
  return code;
}
#define CK(call)                                                        \
  do {                                                                  \
    cudaError_t e__ = (call);                                           \
    if (e__ != cudaSuccess) return fail(B200_ERR_CUDA, #call, e__);     \
  } while (0)
#define REQUIRE(cond, msg)                               \
  do {                                                   \
    if (!(cond)) return fail(B200_ERR_INVALID, msg);     \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 4 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); return -1; }
    cap = want;
    return 0;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct b200_map {
  b200_map_params_t prm;
  MapDev d;
  int N[3];
  double minb[3], maxb[3];
  int sm_count;
  cudaStream_t stream;
  long long* d_bounds;  // 6 order-preserving keys
  bool bounds_pending;
  double cam[3];
  int lb_min[3], lb_max[3];
  DevBuf cloud, stage_in, stage_out, stage_aux;
  int32_t* d_tmp;
  uint2* d_st;
  void* d_dz;
  int dz_bytes;
  int* d_counts;
  int4* d_cross;
  uint8_t* d_plane_empty;  // fused EDT: planes of pass 1 without any obstacle (edt_fused.cu)
  bool edt_fused;          // which pipeline the last build_edt ran
  int edt_mode;            // 0 automatic, 1 always the seven-kernel pipeline
  EdtOverlap ov;           // auxiliary streams / events of the fused EDT (created on first use)
  cudaStream_t copy_stream;  // host clouds: chunked upload overlapping the scatter
  cudaEvent_t copy_done[4], staging_free;
  bool copy_ready;
  bool ov_ready, ov_used;
  int slab_total_nz;       // z-slab sharding: height of the whole stacked grid (bounds the halo distances), 0 = unknown
  uint16_t* d_lut;
  int lut_cap;
  int lut_n;          // entries and d_safe the table currently holds
  double lut_d_safe;
  double* d_dk;
  double d_safe;
  bool has_edt;
  bool profiling;
  cudaEvent_t ev[12];
  double stage_ms[16];
  long long launches;
  // depth-image fusion
  b200_depth_params_t dprm;
  bool depth_set, depth_ready, has_first_depth;
  DepthDev dd;
  void* depth_state;  // one allocation: occ | cnt_all | cnt_hit | F | Fnew | E | touched | walked | ctl
  DevBuf depth_img, depth_rays, depth_paths;
  int raycast_num;
  int64_t depth_stats[8];
};

static inline int host_floor_idx(double pos, double origin, double inv, int n) {
  // grid_map.h:400-404 then :267-274 (boundIndex); the clamp is applied in double so huge ranges stay defined
  double v = std::floor((pos - origin) * inv);
  if (!(v >= 0.0)) return 0;
  if (v > (double)(n - 1)) return n - 1;
  return (int)v;
}

// shared shape of the three per-position readers
template <class T, class F>
static int per_position(b200_map_t* m, const double* pos, int64_t n, T* out, size_t out_elems_per_pos, int on_device, F launch) {
  REQUIRE(m && (n == 0 || (pos && out)) && n >= 0, "bad argument");
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  if (on_device) {
    launch(pos, out);
    m->launches++;
    CK(cudaGetLastError());
    return B200_OK;
  }
  if (m->stage_in.ensure((size_t)n * 24) || m->stage_out.ensure((size_t)n * out_elems_per_pos * sizeof(T)))
    return fail(B200_ERR_NOMEM, "position staging");
  CK(cudaMemcpyAsync(m->stage_in.p, pos, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  launch((const double*)m->stage_in.p, (T*)m->stage_out.p);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out, m->stage_out.p, (size_t)n * out_elems_per_pos * sizeof(T), cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}


extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }

// Host-only pose algebra of GridMap::depthOdomCallback (grid_map.cpp:965-982): body quaternion -> matrix, body2world *
// cam2body_ (grid_map.cpp:91), camera_q_ = Quaterniond(rotation block), and the rotation projectDepthImage then derives from
// camera_q_ (:208).  Quaternion <-> matrix in the published operation order of Eigen's QuaternionBase::toRotationMatrix and
// quaternionbase_assign_impl<Matrix3>; products as ((a0 b0 + a1 b1) + a2 b2) + a3 b3, every operation separately rounded.
static void quat_to_matrix(const double q[4], double m[9]) {  // q = (w, x, y, z)
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  const double tx = 2.0 * x, ty = 2.0 * y, tz = 2.0 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w;
  const double txx = tx * x, txy = ty * x, txz = tz * x;
  const double tyy = ty * y, tyz = tz * y, tzz = tz * z;
  m[0] = 1.0 - (tyy + tzz); m[1] = txy - twz; m[2] = txz + twy;
  m[3] = txy + twz; m[4] = 1.0 - (txx + tzz); m[5] = tyz - twx;
  m[6] = txz - twy; m[7] = tyz + twx; m[8] = 1.0 - (txx + tyy);
}
static void matrix_to_quat(const double m[9], double q[4]) {
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = std::sqrt(t + 1.0);
    q[0] = 0.5 * t;
    t = 0.5 / t;
    q[1] = (m[7] - m[5]) * t;
    q[2] = (m[2] - m[6]) * t;
    q[3] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    q[1 + i] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (m[3 * k + j] - m[3 * j + k]) * t;
    q[1 + j] = (m[3 * j + i] + m[3 * i + j]) * t;
    q[1 + k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
}
int b200_odom_to_camera_pose(const double body_pos[3], const double body_q_wxyz[4], double cam_pos[3], double cam_q_wxyz[4],
                             double cam_rot[9]) {
  REQUIRE(body_pos && body_q_wxyz && cam_pos && cam_q_wxyz && cam_rot, "NULL argument");
  static const double c2b[4][4] = {{0.0, 0.0, 1.0, 0.0}, {-1.0, 0.0, 0.0, 0.0}, {0.0, -1.0, 0.0, -0.02}, {0.0, 0.0, 0.0, 1.0}};  // :91
  double B[9], R[9];
  quat_to_matrix(body_q_wxyz, B);
  for (int i = 0; i < 3; ++i) {
    const double row[4] = {B[3 * i], B[3 * i + 1], B[3 * i + 2], body_pos[i]};  // body2world row i
    for (int j = 0; j < 4; ++j) {
      double acc = row[0] * c2b[0][j];
      for (int k = 1; k < 4; ++k) acc = acc + row[k] * c2b[k][j];
      if (j < 3) R[3 * i + j] = acc; else cam_pos[i] = acc;
    }
  }
  matrix_to_quat(R, cam_q_wxyz);
  quat_to_matrix(cam_q_wxyz, cam_rot);
  return B200_OK;
}
const char* b200_version(void) { return "b200-planner 0.1 (sm_100a)"; }

// streams for hosts that do not link the CUDA runtime themselves (the reference's ROS node does not)
int b200_stream_create(int32_t device, void** out) {
  REQUIRE(out, "NULL argument");
  CK(cudaSetDevice(device));
  cudaStream_t st;
  CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  *out = (void*)st;
  return B200_OK;
}
int b200_stream_destroy(void* stream) {
  if (stream) CK(cudaStreamDestroy((cudaStream_t)stream));
  return B200_OK;
}
int b200_host_alloc(void** ptr, int64_t bytes) {
  REQUIRE(ptr && bytes >= 0, "b200_host_alloc: bad args");
  CK(cudaMallocHost(ptr, (size_t)bytes));
  return B200_OK;
}
int b200_host_free(void* ptr) {
  CK(cudaFreeHost(ptr));
  return B200_OK;
}

int b200_map_create(const b200_map_params_t* p, b200_map_t** out) {
  REQUIRE(p && out, "b200_map_create: NULL argument");
  REQUIRE(p->resolution > 0 && p->size[0] > 0 && p->size[1] > 0 && p->size[2] > 0, "b200_map_create: bad geometry");
  CK(cudaSetDevice(p->device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, p->device));
  if (prop.major != 10) return fail(B200_ERR_UNSUPPORTED, "b200_map_create: device is not sm_100 (B200); no fallback path exists");
  b200_map* m = new b200_map();
  memset(&m->d, 0, sizeof m->d);
  m->prm = *p;
  m->sm_count = prop.multiProcessorCount;
  m->stream = cudaStreamLegacy;
  const double res = p->resolution;
  MapDev& d = m->d;
  d.res = res;
  d.inv_res = 1 / res;  // grid_map.cpp:52
  d.ox = p->origin[0]; d.oy = p->origin[1]; d.oz = p->origin[2];
  for (int i = 0; i < 3; ++i) {
    m->N[i] = (int)std::ceil(p->size[i] / res);  // grid_map.cpp:69-70
    m->minb[i] = p->origin[i];                   // :72
    m->maxb[i] = p->origin[i] + p->size[i];      // :73
    d.lo[i] = m->minb[i] + 1e-4;                 // grid_map.h:372
    d.hi[i] = m->maxb[i] - 1e-4;                 // grid_map.h:378
    m->lb_min[i] = 0;
    m->lb_max[i] = m->N[i] - 1;
  }
  if (m->N[0] > 16384 || m->N[1] > 16384 || m->N[2] > 16384) {
    delete m;
    return fail(B200_ERR_UNSUPPORTED, "b200_map_create: more than 16384 voxels along one axis");
  }
  d.Nx = m->N[0]; d.Ny = m->N[1]; d.Nz = m->N[2];
  d.wz = (d.Nz + 31) / 32;
  const size_t words = (size_t)d.Nx * d.Ny * d.wz;
  cudaError_t e = cudaMalloc(&d.occ, words * 4);
  if (e != cudaSuccess) { delete m; return fail(B200_ERR_NOMEM, "b200_map_create: occupancy grid", e); }
  cudaMemset(d.occ, 0, words * 4);  // grid_map.cpp:80
  // d_k table of fastLOS (LineOfSight.cpp:19): d += 0.05 by repeated addition, long enough for the map diagonal
  const double diag = std::sqrt(p->size[0] * p->size[0] + p->size[1] * p->size[1] + p->size[2] * p->size[2]);
  const int len = (int)std::ceil(diag / kLosStep) + 64;
  std::vector<double> dk(len);
  double acc = 0;
  for (int i = 0; i < len; ++i) { dk[i] = acc; acc += kLosStep; }
  e = cudaMalloc(&m->d_dk, (size_t)len * 8);
  if (e != cudaSuccess) { cudaFree(d.occ); delete m; return fail(B200_ERR_NOMEM, "b200_map_create: dk table", e); }
  cudaMemcpy(m->d_dk, dk.data(), (size_t)len * 8, cudaMemcpyHostToDevice);
  d.dk = m->d_dk;
  d.dk_len = len;
  e = cudaMalloc(&m->d_bounds, 6 * 8);
  if (e != cudaSuccess) { cudaFree(d.occ); cudaFree(m->d_dk); delete m; return fail(B200_ERR_NOMEM, "b200_map_create: bounds", e); }
  m->bounds_pending = false;
  m->d_tmp = nullptr;
  m->d_st = nullptr;
  m->d_dz = nullptr;
  m->dz_bytes = 0;
  m->d_counts = nullptr;
  m->d_cross = nullptr;
  m->d_plane_empty = nullptr;
  m->edt_fused = false;
  m->edt_mode = 0;
  m->slab_total_nz = 0;
  m->ov_ready = false;
  m->copy_ready = false;
  m->ov_used = false;
  m->d_lut = nullptr;
  m->lut_cap = 0;
  m->lut_n = 0;
  m->lut_d_safe = -1.0;
  m->has_edt = false;
  m->d_safe = 0;
  m->profiling = false;
  m->depth_set = m->depth_ready = m->has_first_depth = false;
  m->depth_state = nullptr;
  m->raycast_num = 0;
  memset(&m->dd, 0, sizeof m->dd);
  memset(m->depth_stats, 0, sizeof m->depth_stats);
  for (auto& ev : m->ev) cudaEventCreate(&ev);
  memset(m->stage_ms, 0, sizeof m->stage_ms);
  m->launches = 0;
  CK(cudaDeviceSynchronize());
  *out = m;
  return B200_OK;
}

int b200_map_destroy(b200_map_t* m) {
  if (!m) return B200_OK;
  cudaSetDevice(m->prm.device);
  cudaStreamSynchronize(m->stream);
  cudaFree(m->d.occ);
  if (m->d.dist2) cudaFree(m->d.dist2);
  if (m->d.costq) cudaFree(m->d.costq);
  if (m->d_tmp) cudaFree(m->d_tmp);
  if (m->d_st) cudaFree(m->d_st);
  if (m->d_dz) cudaFree(m->d_dz);
  if (m->d_counts) cudaFree(m->d_counts);
  if (m->d_cross) cudaFree(m->d_cross);
  if (m->d_plane_empty) cudaFree(m->d_plane_empty);
  if (m->copy_ready) {
    cudaStreamDestroy(m->copy_stream);
    for (int i = 0; i < 4; ++i) cudaEventDestroy(m->copy_done[i]);
    cudaEventDestroy(m->staging_free);
  }
  if (m->ov_ready) {
    for (int i = 0; i < 2; ++i) { cudaStreamDestroy(m->ov.s[i]); cudaEventDestroy(m->ov.y_done[i]); cudaEventDestroy(m->ov.x_done[i]); }
    cudaEventDestroy(m->ov.fork);
  }
  if (m->d_lut) cudaFree(m->d_lut);
  cudaFree(m->d_dk);
  cudaFree(m->d_bounds);
  m->cloud.release(); m->stage_in.release(); m->stage_out.release(); m->stage_aux.release();
  if (m->depth_state) cudaFree(m->depth_state);
  m->depth_img.release(); m->depth_rays.release(); m->depth_paths.release();
  for (auto& ev : m->ev) cudaEventDestroy(ev);
  delete m;
  return B200_OK;
}

int b200_map_set_stream(b200_map_t* m, void* st) {
  REQUIRE(m, "NULL map");
  m->stream = st ? (cudaStream_t)st : cudaStreamLegacy;
  return B200_OK;
}
int b200_map_sync(b200_map_t* m) {
  REQUIRE(m, "NULL map");
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_dims(const b200_map_t* m, int32_t n[3]) {
  REQUIRE(m && n, "NULL argument");
  for (int i = 0; i < 3; ++i) n[i] = m->N[i];
  return B200_OK;
}
int b200_map_get_region(const b200_map_t* m, double origin[3], double size[3]) {
  REQUIRE(m && origin && size, "NULL argument");
  for (int i = 0; i < 3; ++i) { origin[i] = m->prm.origin[i]; size[i] = m->prm.size[i]; }
  return B200_OK;
}
double b200_map_get_resolution(const b200_map_t* m) { return m ? m->prm.resolution : 0.0; }

int b200_map_clear_box(b200_map_t* m, const double mn[3], const double mx[3]) {
  REQUIRE(m && mn && mx, "NULL argument");
  CK(cudaSetDevice(m->prm.device));
  int a[3], b[3];
  const double o[3] = {m->d.ox, m->d.oy, m->d.oz};
  for (int i = 0; i < 3; ++i) {
    a[i] = host_floor_idx(mn[i], o[i], m->d.inv_res, m->N[i]);
    b[i] = host_floor_idx(mx[i], o[i], m->d.inv_res, m->N[i]);
  }
  launch_clear_box(m->d, a, b, m->sm_count, m->stream);
  m->has_edt = false;  // the distance / cost field describes the previous grid
  const bool whole = a[0] == 0 && a[1] == 0 && a[2] == 0 && b[0] == m->N[0] - 1 && b[1] == m->N[1] - 1 && b[2] == m->N[2] - 1;
  if (!whole) m->launches++;  // the whole map is one cudaMemsetAsync, not a kernel of ours
  CK(cudaGetLastError());
  return B200_OK;
}
int b200_map_reset(b200_map_t* m) {  // grid_map.cpp:144-153
  REQUIRE(m, "NULL map");
  int rc = b200_map_clear_box(m, m->minb, m->maxb);
  if (rc) return rc;
  m->bounds_pending = false;
  for (int i = 0; i < 3; ++i) { m->lb_min[i] = 0; m->lb_max[i] = m->N[i] - 1; }
  return B200_OK;
}

int b200_map_update_cloud(b200_map_t* m, const void* pts, int64_t n, int32_t step, int32_t ox, int32_t oy, int32_t oz,
                          const double cam[3], int32_t on_device) {
  REQUIRE(m && cam, "NULL argument");
  REQUIRE(n >= 0 && (n == 0 || pts), "b200_map_update_cloud: bad cloud");
  REQUIRE(step >= 12 && ox >= 0 && oy >= 0 && oz >= 0 && ox + 4 <= step && oy + 4 <= step && oz + 4 <= step && (ox | oy | oz | step) % 4 == 0,
          "b200_map_update_cloud: bad point layout");
  if (n == 0) return B200_OK;                                                  // grid_map.cpp:724-725
  if (std::isnan(cam[0]) || std::isnan(cam[1]) || std::isnan(cam[2])) return B200_OK;  // :727-728
  CK(cudaSetDevice(m->prm.device));
  const int s = (int)std::ceil(m->prm.obstacles_inflation / m->prm.resolution);  // :735
  if (s < 0 || s > 4) return fail(B200_ERR_UNSUPPORTED, "b200_map_update_cloud: inflation step > 4 voxels");
  const uint8_t* dpts = (const uint8_t*)pts;
  // A host cloud is uploaded in kCopyChunks pieces on a copy stream while the main stream clears the box and voxelises the
  // pieces that have arrived: the scatter (and the clear) hide behind the PCIe transfer instead of following it.
  constexpr int kCopyChunks = 4;
  // (pinned host memory only: a pageable source makes every cudaMemcpyAsync a staged, host-synchronous copy, and four of
  //  those measured 2.1 ms per frame against 1.3 ms for one)
  bool pipelined = false;
  if (!on_device && n >= 65536) {
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, pts) == cudaSuccess) pipelined = attr.type == cudaMemoryTypeHost;
    else cudaGetLastError();
  }
  if (!on_device) {
    if (m->cloud.ensure((size_t)n * step)) return fail(B200_ERR_NOMEM, "cloud staging");
    dpts = (const uint8_t*)m->cloud.p;
    if (pipelined) {
      if (!m->copy_ready) {
        CK(cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < kCopyChunks; ++i) CK(cudaEventCreateWithFlags(&m->copy_done[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&m->staging_free, cudaEventDisableTiming));
        m->copy_ready = true;
      }
      // the staging buffer may still be read by the previous frame's scatter (or by whatever the main stream runs)
      CK(cudaEventRecord(m->staging_free, m->stream));
      CK(cudaStreamWaitEvent(m->copy_stream, m->staging_free, 0));
      for (int i = 0; i < kCopyChunks; ++i) {
        const int64_t a = n * i / kCopyChunks, b = n * (i + 1) / kCopyChunks;
        CK(cudaMemcpyAsync((char*)m->cloud.p + (size_t)a * step, (const char*)pts + (size_t)a * step, (size_t)(b - a) * step,
                           cudaMemcpyHostToDevice, m->copy_stream));
        CK(cudaEventRecord(m->copy_done[i], m->copy_stream));
      }
    } else {
      CK(cudaMemcpyAsync(m->cloud.p, pts, (size_t)n * step, cudaMemcpyHostToDevice, m->stream));
    }
  }
  if (m->profiling) cudaEventRecord(m->ev[0], m->stream);
  double mn[3], mx[3];
  for (int i = 0; i < 3; ++i) { mn[i] = cam[i] - m->prm.local_update_range[i]; mx[i] = cam[i] + m->prm.local_update_range[i]; }
  int rc = b200_map_clear_box(m, mn, mx);  // :730
  if (rc) return rc;
  if (m->profiling) cudaEventRecord(m->ev[1], m->stream);
  // running min starts at map max, running max at map min (:740-746), as order-preserving keys
  long long keys[6];
  for (int i = 0; i < 3; ++i) {
    long long b;
    memcpy(&b, &m->maxb[i], 8); keys[i] = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
    memcpy(&b, &m->minb[i], 8); keys[3 + i] = b >= 0 ? b : (b ^ 0x7fffffffffffffffLL);
  }
  CK(cudaMemcpyAsync(m->d_bounds, keys, sizeof keys, cudaMemcpyHostToDevice, m->stream));
  if (pipelined) {
    for (int i = 0; i < kCopyChunks; ++i) {
      const int64_t a = n * i / kCopyChunks, b = n * (i + 1) / kCopyChunks;
      CK(cudaStreamWaitEvent(m->stream, m->copy_done[i], 0));
      if (launch_scatter(m->d, dpts + (size_t)a * step, b - a, step, ox, oy, oz, cam, m->prm.local_update_range, s, m->d_bounds, m->stream))
        return fail(B200_ERR_UNSUPPORTED, "scatter: inflation step");
      m->launches++;
    }
  } else {
    if (launch_scatter(m->d, dpts, n, step, ox, oy, oz, cam, m->prm.local_update_range, s, m->d_bounds, m->stream))
      return fail(B200_ERR_UNSUPPORTED, "scatter: inflation step");
    m->launches++;
  }
  m->has_edt = false;
  if (m->profiling) cudaEventRecord(m->ev[2], m->stream);
  CK(cudaGetLastError());
  m->bounds_pending = true;
  for (int i = 0; i < 3; ++i) m->cam[i] = cam[i];
  return B200_OK;
}

int b200_map_local_bounds(b200_map_t* m, int32_t mn[3], int32_t mx[3]) {
  REQUIRE(m && mn && mx, "NULL argument");
  if (m->bounds_pending) {
    CK(cudaSetDevice(m->prm.device));
    long long keys[6];
    CK(cudaMemcpyAsync(keys, m->d_bounds, sizeof keys, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
    double lo[3], hi[3];
    for (int i = 0; i < 6; ++i) {
      long long b = keys[i] >= 0 ? keys[i] : (keys[i] ^ 0x7fffffffffffffffLL);
      double v;
      memcpy(&v, &b, 8);
      if (i < 3) lo[i] = v; else hi[i - 3] = v;
    }
    for (int i = 0; i < 3; ++i) {  // grid_map.cpp:789-795
      lo[i] = std::min(lo[i], m->cam[i]);
      hi[i] = std::max(hi[i], m->cam[i]);
    }
    hi[2] = std::max(hi[2], m->prm.ground_height);  // :797
    const double o[3] = {m->d.ox, m->d.oy, m->d.oz};
    for (int i = 0; i < 3; ++i) {  // :799-803
      m->lb_max[i] = host_floor_idx(hi[i], o[i], m->d.inv_res, m->N[i]);
      m->lb_min[i] = host_floor_idx(lo[i], o[i], m->d.inv_res, m->N[i]);
    }
    m->bounds_pending = false;
  }
  for (int i = 0; i < 3; ++i) { mn[i] = m->lb_min[i]; mx[i] = m->lb_max[i]; }
  return B200_OK;
}

int b200_map_get_inflate_occupancy(b200_map_t* m, const double* pos, int64_t n, int8_t* out, int32_t on_device) {
  return per_position<int8_t>(m, pos, n, out, 1, on_device,
                              [&](const double* p, int8_t* o) { launch_occupancy(m->d, p, n, o, m->stream); });
}
int b200_map_is_in_map(b200_map_t* m, const double* pos, int64_t n, uint8_t* out, int32_t on_device) {
  return per_position<uint8_t>(m, pos, n, out, 1, on_device,
                               [&](const double* p, uint8_t* o) { launch_in_map(m->d, p, n, o, m->stream); });
}
int b200_map_pos_to_index(b200_map_t* m, const double* pos, int64_t n, int32_t* out, int32_t on_device) {
  return per_position<int32_t>(m, pos, n, out, 3, on_device,
                               [&](const double* p, int32_t* o) { launch_pos_to_index(m->d, p, n, o, m->stream); });
}

int b200_map_download_inflate(b200_map_t* m, uint8_t* dst) {
  REQUIRE(m && dst, "NULL argument");
  CK(cudaSetDevice(m->prm.device));
  const size_t nv = (size_t)m->N[0] * m->N[1] * m->N[2];
  if (m->stage_out.ensure(nv)) return fail(B200_ERR_NOMEM, "download staging");
  launch_unpack(m->d, (uint8_t*)m->stage_out.p, m->sm_count, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, m->stage_out.p, nv, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_upload_inflate(b200_map_t* m, const uint8_t* src) {
  REQUIRE(m && src, "NULL argument");
  CK(cudaSetDevice(m->prm.device));
  const size_t nv = (size_t)m->N[0] * m->N[1] * m->N[2];
  if (m->stage_in.ensure(nv)) return fail(B200_ERR_NOMEM, "upload staging");
  CK(cudaMemcpyAsync(m->stage_in.p, src, nv, cudaMemcpyHostToDevice, m->stream));
  launch_pack(m->d, (const uint8_t*)m->stage_in.p, m->sm_count, m->stream);
  m->has_edt = false;
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_device_bits(b200_map_t* m, void** p, int64_t* n_words) {
  REQUIRE(m && p && n_words, "NULL argument");
  *p = m->d.occ;
  *n_words = (int64_t)m->N[0] * m->N[1] * m->d.wz;
  return B200_OK;
}

static int build_edt_impl(b200_map_t* m, double d_safe, const int32_t* below_gap, const int32_t* above_gap, int on_device) {
  REQUIRE(m, "NULL map");
  CK(cudaSetDevice(m->prm.device));
  const size_t nv = (size_t)m->N[0] * m->N[1] * m->N[2];
  const size_t cols = (size_t)m->N[0] * m->N[1];
  auto need = [&](void** p, size_t bytes) -> bool {
    if (*p) return true;
    if (cudaMalloc(p, bytes) != cudaSuccess) { cudaGetLastError(); *p = nullptr; return false; }
    return true;
  };
  const bool halo = below_gap || above_gap;
  if (!need((void**)&m->d_tmp, nv * 4) || !need((void**)&m->d.dist2, nv * 4) ||
      !need((void**)&m->d_plane_empty, edt_fused_plane_flags_bytes(m->d)))
    return fail(B200_ERR_NOMEM, "distance-field buffers");
  const int32_t* d_below = below_gap;
  const int32_t* d_above = above_gap;
  if (halo && !on_device) {  // stage the two per-column halo arrays
    if (m->stage_in.ensure(cols * 8)) return fail(B200_ERR_NOMEM, "halo staging");
    int32_t* base = (int32_t*)m->stage_in.p;
    if (below_gap) { CK(cudaMemcpyAsync(base, below_gap, cols * 4, cudaMemcpyHostToDevice, m->stream)); d_below = base; }
    if (above_gap) { CK(cudaMemcpyAsync(base + cols, above_gap, cols * 4, cudaMemcpyHostToDevice, m->stream)); d_above = base + cols; }
  }
  int lut_n = 0;
  if (d_safe > 0) {
    if (!need((void**)&m->d.costq, nv * 2)) return fail(B200_ERR_NOMEM, "cost field");
    lut_n = edt_lut_size(m->prm.resolution, d_safe);
    if (lut_n > m->lut_cap) {
      if (m->d_lut) cudaFree(m->d_lut);
      m->d_lut = nullptr;
      if (!need((void**)&m->d_lut, (size_t)lut_n * 2)) return fail(B200_ERR_NOMEM, "cost table");
      m->lut_cap = lut_n;
      m->lut_d_safe = -1.0;
    }
    if (m->lut_d_safe != d_safe || m->lut_n != lut_n) {  // the table only depends on (resolution, d_safe)
      launch_cost_lut(m->d_lut, lut_n, m->prm.resolution, d_safe, m->stream);
      m->launches++;
      m->lut_d_safe = d_safe;
      m->lut_n = lut_n;
    }
  }
  // two shared-memory-staged kernels when a line's hull stack fits in shared memory (edt_fused.cu) ...
  static const bool force_legacy = getenv("B200_EDT_LEGACY") != nullptr;
  static const bool no_overlap = getenv("B200_EDT_NO_OVERLAP") != nullptr;
  if (!m->ov_ready && !no_overlap) {
    for (int i = 0; i < 2; ++i) {
      CK(cudaStreamCreateWithFlags(&m->ov.s[i], cudaStreamNonBlocking));
      CK(cudaEventCreate(&m->ov.y_done[i]));
      CK(cudaEventCreate(&m->ov.x_done[i]));
    }
    CK(cudaEventCreateWithFlags(&m->ov.fork, cudaEventDisableTiming));
    m->ov_ready = true;
  }
  m->edt_fused = !force_legacy && m->edt_mode == 0 && launch_edt_fused(m->d, m->d_tmp, m->d_plane_empty, m->d.dist2, d_safe > 0 ? m->d.costq : nullptr, m->d_lut,
                                                   lut_n, d_below, d_above, m->slab_total_nz, m->stream, m->profiling ? &m->ev[3] : nullptr,
                                                   m->ov_ready ? &m->ov : nullptr);
  if (m->edt_fused) m->launches += m->ov_ready ? 2 * std::max(1, m->ov.parts_used) : 2;  // z parts alternate between two streams
  else {
    // ... else the seven-kernel pipeline with its stacks in HBM (edt_kernels.cu)
    const int dz_bytes = edt_dz_is_wide(m->d, halo) ? 2 : 1;
    if (m->d_dz && m->dz_bytes < dz_bytes) { cudaFree(m->d_dz); m->d_dz = nullptr; }
    const size_t max_lines = (size_t)std::max(m->N[0], m->N[1]) * m->N[2];
    if (!need((void**)&m->d_st, nv * 8) || !need((void**)&m->d_dz, nv * dz_bytes) || !need((void**)&m->d_counts, max_lines * 4 * 4) ||
        !need((void**)&m->d_cross, max_lines * sizeof(int4)))
      return fail(B200_ERR_NOMEM, "distance-field buffers");
    m->dz_bytes = std::max(m->dz_bytes, dz_bytes);
    launch_edt(m->d, m->d_dz, m->d_tmp, m->d_st, m->d_counts, m->d_cross, m->d.dist2, m->d.costq, m->d_lut, lut_n, d_safe, d_below, d_above,
               m->stream, m->profiling ? &m->ev[3] : nullptr);
    m->launches += 7;
  }
  CK(cudaGetLastError());
  if (halo && !on_device) CK(cudaStreamSynchronize(m->stream));  // staging buffer is reused by other calls
  m->has_edt = true;
  m->d_safe = d_safe;
  return B200_OK;
}
// ---------------------------------------------------------------------------------------------------
// depth-image fusion
// ---------------------------------------------------------------------------------------------------
static double order_key_to_double(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffULL) : ~k;
  double d;
  memcpy(&d, &u, 8);
  return d;
}
static int depth_ensure_state(b200_map* m) {
  if (m->depth_ready) return B200_OK;
  const size_t nvox = (size_t)m->N[0] * m->N[1] * m->N[2];
  if (nvox >= 0xfffffffeULL) return fail(B200_ERR_UNSUPPORTED, "depth fusion: more than 2^32-2 voxels");
  const size_t bytes = nvox * 8 + 7 * nvox * 4 + 2 * nvox + 512;
  cudaError_t e = cudaMalloc(&m->depth_state, bytes);
  if (e != cudaSuccess) { m->depth_state = nullptr; return fail(B200_ERR_NOMEM, "depth fusion state (38 bytes per voxel)", e); }
  char* p = (char*)m->depth_state;
  DepthDev& D = m->dd;
  D.occ = (double*)p; p += nvox * 8;
  D.cnt_all = (uint32_t*)p; p += nvox * 4;
  D.cnt_hit = (uint32_t*)p; p += nvox * 4;
  D.F = (uint32_t*)p; p += nvox * 4;
  D.Fnew = (uint32_t*)p; p += nvox * 4;
  D.E = (uint32_t*)p; p += nvox * 4;
  D.touched = (uint32_t*)p; p += nvox * 4;
  D.walked = (uint32_t*)p; p += nvox * 4;
  D.flagT = (int8_t*)p; p += nvox;
  D.flagE = (int8_t*)p; p += nvox;
  p = (char*)(((uintptr_t)p + 255) & ~(uintptr_t)255);
  D.ctl = (DepthCtl*)p;
  static_assert(sizeof(DepthCtl) <= 256, "ctl block");
  CK(cudaMemsetAsync(D.cnt_all, 0, nvox * 8, m->stream));  // cnt_all + cnt_hit (grid_map.cpp:82-83)
  CK(cudaMemsetAsync(D.flagT, 0xff, nvox * 2, m->stream));  // flag_traverse_, flag_rayend_ = -1 (:84-85)
  launch_depth_init(D, D.unknown, (long long)nvox, m->stream);  // :79 occupancy_buffer_ = clamp_min - unknown_flag; flags = none
  m->launches++;
  CK(cudaGetLastError());
  m->depth_ready = true;
  return B200_OK;
}

int b200_map_set_depth_params(b200_map_t* m, const b200_depth_params_t* p) {
  REQUIRE(m && p, "NULL argument");
  REQUIRE(p->fx != 0 && p->fy != 0 && p->k_depth_scaling_factor != 0, "depth params: fx, fy, k_depth_scaling_factor must be non-zero");
  REQUIRE(p->skip_pixel >= 1 && p->depth_filter_margin >= 0, "depth params: skip_pixel >= 1, depth_filter_margin >= 0");
  for (double q : {p->p_hit, p->p_miss, p->p_min, p->p_max, p->p_occ}) REQUIRE(q > 0 && q < 1, "depth params: probabilities must be in (0, 1)");
  m->dprm = *p;
  DepthDev& D = m->dd;
  D.fx = p->fx; D.fy = p->fy; D.cx = p->cx; D.cy = p->cy;
  D.scale = p->k_depth_scaling_factor;
  D.inv_scale = 1.0 / p->k_depth_scaling_factor;  // grid_map.cpp:245
  D.mindist = p->depth_filter_mindist; D.maxdist = p->depth_filter_maxdist; D.max_ray_length = p->max_ray_length;
  // grid_map.h:27 `#define logit(x) (log((x) / (1 - (x))))` with the host's libm, exactly like the reference (:57-62)
  D.prob_hit_log = std::log(p->p_hit / (1 - p->p_hit));
  D.prob_miss_log = std::log(p->p_miss / (1 - p->p_miss));
  D.clamp_min_log = std::log(p->p_min / (1 - p->p_min));
  D.clamp_max_log = std::log(p->p_max / (1 - p->p_max));
  D.min_occ_log = std::log(p->p_occ / (1 - p->p_occ));
  const double unknown = D.clamp_min_log - 0.01;  // unknown_flag_, :62
  if (m->depth_ready && unknown != D.unknown) return fail(B200_ERR_UNSUPPORTED, "depth params: p_min cannot change once the log-odds grid exists");
  D.unknown = unknown;
  D.use_filter = p->use_depth_filter ? 1 : 0; D.margin = p->depth_filter_margin; D.skip = p->skip_pixel;
  for (int i = 0; i < 3; ++i) { D.minb[i] = m->minb[i]; D.maxb[i] = m->maxb[i]; }
  D.max_steps = m->N[0] + m->N[1] + m->N[2] + 4;
  m->depth_set = true;
  return B200_OK;
}

int b200_map_set_depth_frame_count(b200_map_t* m, int32_t raycast_num) {
  REQUIRE(m, "NULL map");
  m->raycast_num = (int)(signed char)raycast_num;
  return B200_OK;
}
int b200_map_depth_stats(b200_map_t* m, int64_t out[8]) {
  REQUIRE(m && out, "NULL argument");
  memcpy(out, m->depth_stats, sizeof m->depth_stats);
  out[5] = m->raycast_num;
  return B200_OK;
}

int b200_map_update_depth(b200_map_t* m, const uint16_t* depth, int32_t rows, int32_t cols, const double cam_pos[3],
                          const double cam_rot[9], int32_t on_device, int32_t* fused) {
  REQUIRE(m && depth && cam_pos && cam_rot && rows > 0 && cols > 0, "bad argument");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_update_depth: call b200_map_set_depth_params first");
  CK(cudaSetDevice(m->prm.device));
  cudaStream_t st = m->stream;
  if (fused) *fused = 0;
  memset(m->depth_stats, 0, sizeof m->depth_stats);
  // depthPoseCallback's gate (grid_map.cpp:688-697): a frame whose camera is outside the map is not fused
  for (int i = 0; i < 3; ++i)
    if (cam_pos[i] < m->d.lo[i] || cam_pos[i] > m->d.hi[i]) return B200_OK;
  if (fused) *fused = 1;
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  DepthDev& D = m->dd;
  const b200_depth_params_t& P = m->dprm;
  // which pixels become rays, in the reference's order (projectDepthImage, grid_map.cpp:212-231 / :233-275)
  int rows_s, cols_s;
  if (P.use_depth_filter) {
    if (!m->has_first_depth) { m->has_first_depth = true; return B200_OK; }  // :236-239: the first frame only arms the filter
    const int mg = P.depth_filter_margin, sk = P.skip_pixel;
    rows_s = rows - 2 * mg > 0 ? (rows - 2 * mg + sk - 1) / sk : 0;
    cols_s = cols - 2 * mg > 0 ? (cols - 2 * mg + sk - 1) / sk : 0;
    if (rows_s > 0 && cols_s > 0) {
       // :260 reads the pixel `skip_pixel` after the sampled one; it must exist
      const long long last = (long long)(mg + (rows_s - 1) * sk) * cols + mg + (long long)(cols_s - 1) * sk + sk;
      if (last >= (long long)rows * cols)
        return fail(B200_ERR_UNSUPPORTED, "depth filter: skip_pixel reaches past the image end (the reference reads out of bounds here)");
    }
  } else {
    rows_s = rows; cols_s = cols;
  }
  const long long n_rays = (long long)rows_s * cols_s;
  if (n_rays == 0) return B200_OK;  // raycastProcess :320: nothing projected, nothing happens
  REQUIRE(n_rays < (1LL << 24), "depth image too large (more than 2^24 sampled pixels)");  // 8 bits of round stamp remain
  const uint16_t* d_img = depth;
  if (!on_device) {
    if (m->depth_img.ensure((size_t)rows * cols * 2)) return fail(B200_ERR_NOMEM, "depth image staging");
    CK(cudaMemcpyAsync(m->depth_img.p, depth, (size_t)rows * cols * 2, cudaMemcpyHostToDevice, st));
    d_img = (const uint16_t*)m->depth_img.p;
  }
  if (m->depth_rays.ensure((size_t)n_rays * 44)) return fail(B200_ERR_NOMEM, "ray records");
  D.ray_pt = (double*)m->depth_rays.p;
  D.ray_addr = (uint32_t*)((char*)m->depth_rays.p + (size_t)n_rays * 24);
  D.cast_id = D.ray_addr + n_rays;
  D.path_off = D.cast_id + n_rays;
  D.path_len = D.path_off + n_rays;
  D.stop = D.path_len + n_rays;
  D.id_bits = 1;
  while ((1LL << D.id_bits) < n_rays) ++D.id_bits;  // < 31 (n_rays < 2^31): at least two stamp bits... checked below
  D.rows = rows; D.cols = cols; D.cols_s = cols_s; D.n_rays = (int)n_rays;
  for (int i = 0; i < 3; ++i) D.cam[i] = cam_pos[i];
  for (int i = 0; i < 9; ++i) D.R[i] = cam_rot[i];
  // `char raycast_num_` (grid_map.h:124): 127 is followed by -128
  const int round = (int)(signed char)(m->raycast_num + 1);
  D.round = (int8_t)round;
  // index box of camera +- local_update_range (:439-446)
  for (int i = 0; i < 3; ++i) {
    const double org = i == 0 ? m->d.ox : i == 1 ? m->d.oy : m->d.oz;
    D.loc_min[i] = host_floor_idx(cam_pos[i] - m->prm.local_update_range[i], org, m->d.inv_res, m->N[i]);
    D.loc_max[i] = host_floor_idx(cam_pos[i] + m->prm.local_update_range[i], org, m->d.inv_res, m->N[i]);
  }
  DepthCtl ctl;
  memset(&ctl, 0, sizeof ctl);
  for (int i = 0; i < 3; ++i) { ctl.bbox[i] = ~0ULL; ctl.bbox[i + 3] = 0ULL; }
  CK(cudaMemcpyAsync(D.ctl, &ctl, sizeof ctl, cudaMemcpyHostToDevice, st));
  launch_depth_project(D, m->d, d_img, st);
  launch_depth_path_count(D, m->d, st);
  m->launches += 2;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(&ctl, D.ctl, sizeof ctl, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  m->depth_stats[0] = ctl.n_valid;
  if (ctl.n_valid == 0) {  // every sampled pixel was filtered out: proj_points_cnt == 0, raycastProcess returns at once
    return B200_OK;
  }
  // a frame that cannot be completed must not leave its end-point counters and ray ids behind
  auto abandon = [&](int code, const char* what) {
    const size_t nvox = (size_t)m->N[0] * m->N[1] * m->N[2];
    cudaMemsetAsync(D.cnt_all, 0, nvox * 8, st);
    cudaMemsetAsync(D.F, 0xff, nvox * 12, st);  // F, Fnew, E are adjacent
    cudaStreamSynchronize(st);
    return fail(code, what);
  };
  if (ctl.n_path_total >= 0xffffffffULL) return abandon(B200_ERR_UNSUPPORTED, "depth fusion: more than 2^32 ray steps in one frame");
  if (m->depth_paths.ensure((size_t)ctl.n_path_total * 4 + 4)) return abandon(B200_ERR_NOMEM, "ray paths");
  m->raycast_num = round;  // :325
  D.paths = (uint32_t*)m->depth_paths.p;
  launch_depth_path_fill(D, m->d, ctl.n_cast_rays, st);
  m->launches++;
  // ray-order fixed point (see depth_kernels.cu): one cooperative launch
  {
    const int e = launch_depth_order(D, ctl.n_cast_rays, m->sm_count, st);
    if (e) return fail(B200_ERR_CUDA, "depth fusion: cooperative launch", (cudaError_t)e);
    m->launches++;
  }
  launch_depth_count(D, ctl.n_cast_rays, st);
  launch_depth_update(D, m->d, m->sm_count, st);
  m->launches += 3;
  CK(cudaGetLastError());
  // local bounds (:421-435): end-point box, extended by the camera position and the ground height
  double mn[3], mx[3];
  for (int i = 0; i < 3; ++i) {
    mn[i] = std::min(m->maxb[i], order_key_to_double(ctl.bbox[i]));      // :332-334 start values
    mx[i] = std::max(m->minb[i], order_key_to_double(ctl.bbox[i + 3]));  // :335-337
    mn[i] = std::min(mn[i], cam_pos[i]);
    mx[i] = std::max(mx[i], cam_pos[i]);
  }
  mx[2] = std::max(mx[2], m->prm.ground_height);
  for (int i = 0; i < 3; ++i) {
    const double org = i == 0 ? m->d.ox : i == 1 ? m->d.oy : m->d.oz;
    m->lb_max[i] = host_floor_idx(mx[i], org, m->d.inv_res, m->N[i]);
    m->lb_min[i] = host_floor_idx(mn[i], org, m->d.inv_res, m->N[i]);
  }
  m->bounds_pending = false;
  // clearAndInflateLocalMap (:509-627)
  int cut_min[3], cut_max[3], cutm_min[3], cutm_max[3];
  for (int i = 0; i < 3; ++i) {
    cut_min[i] = std::max(std::min(m->lb_min[i] - P.local_map_margin, m->N[i] - 1), 0);
    cut_max[i] = std::max(std::min(m->lb_max[i] + P.local_map_margin, m->N[i] - 1), 0);
    cutm_min[i] = std::max(std::min(cut_min[i] - 5, m->N[i] - 1), 0);
    cutm_max[i] = std::max(std::min(cut_max[i] + 5, m->N[i] - 1), 0);
  }
  const int inf_step = (int)std::ceil(m->prm.obstacles_inflation / m->prm.resolution);  // :581
  const int ceil_on = P.virtual_ceil_height > -0.5;
  const int ceil_id = (int)std::floor((P.virtual_ceil_height - m->d.oz) * m->d.inv_res);  // :620
  launch_depth_local_map(D, m->d, m->lb_min, m->lb_max, cut_min, cut_max, cutm_min, cutm_max, inf_step, ceil_on, ceil_id, m->sm_count, st);
  m->launches += 4;
  CK(cudaGetLastError());
  m->has_edt = false;
  CK(cudaMemcpyAsync(&ctl, D.ctl, sizeof ctl, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  m->depth_stats[1] = (int64_t)ctl.n_cast;
  m->depth_stats[2] = (int64_t)ctl.n_steps;
  m->depth_stats[3] = (int64_t)ctl.n_updated;
  m->depth_stats[4] = ctl.rounds;
  if (ctl.rounds > ctl.n_cast_rays + 3u) return fail(B200_ERR_CUDA, "depth fusion: ray-order iteration did not converge");
  return B200_OK;
}

int b200_map_get_occupancy(b200_map_t* m, const double* pos, int64_t n, int8_t* out, int32_t on_device) {
  REQUIRE(m, "NULL map");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_get_occupancy: call b200_map_set_depth_params first");
  CK(cudaSetDevice(m->prm.device));
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  return per_position<int8_t>(m, pos, n, out, 1, on_device, [&](const double* p, int8_t* o) {
    launch_depth_get_occupancy(m->d, m->dd.occ, m->dd.min_occ_log, p, n, o, m->stream);
  });
}
// GridMap::setOccupied (grid_map.h:310-321)
int b200_map_set_occupied(b200_map_t* m, const double* pos, int64_t n, int32_t on_device) {
  REQUIRE(m && (n == 0 || pos) && n >= 0, "b200_map_set_occupied: bad argument");
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  const double* dpos = pos;
  if (!on_device) {
    if (m->stage_in.ensure((size_t)n * 24)) return fail(B200_ERR_NOMEM, "position staging");
    CK(cudaMemcpyAsync(m->stage_in.p, pos, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
    dpos = (const double*)m->stage_in.p;
  }
  launch_set_occupied(m->d, dpos, n, m->stream);
  m->launches++;
  m->has_edt = false;
  CK(cudaGetLastError());
  if (!on_device) CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
// GridMap::setOccupancy (grid_map.h:323-337)
int b200_map_set_occupancy(b200_map_t* m, const double* pos, const double* occ, int64_t n, int32_t on_device) {
  REQUIRE(m && (n == 0 || (pos && occ)) && n >= 0 && n < 0xfffffffeLL, "b200_map_set_occupancy: bad argument");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_set_occupancy: call b200_map_set_depth_params first");
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  const double *dpos = pos, *docc = occ;
  if (!on_device) {
    if (m->stage_in.ensure((size_t)n * 32)) return fail(B200_ERR_NOMEM, "position staging");
    CK(cudaMemcpyAsync(m->stage_in.p, pos, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
    CK(cudaMemcpyAsync((char*)m->stage_in.p + (size_t)n * 24, occ, (size_t)n * 8, cudaMemcpyHostToDevice, m->stream));
    dpos = (const double*)m->stage_in.p;
    docc = (const double*)((char*)m->stage_in.p + (size_t)n * 24);
  }
  launch_set_occupancy(m->d, m->dd.occ, dpos, docc, n, m->dd.E, m->stream);  // E: per-frame scratch, 0xffffffff between frames
  m->launches += 2;
  CK(cudaGetLastError());
  if (!on_device) CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_get_log_odds(b200_map_t* m, const double* pos, int64_t n, double* out, int32_t on_device) {
  REQUIRE(m, "NULL map");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_get_log_odds: call b200_map_set_depth_params first");
  CK(cudaSetDevice(m->prm.device));
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  return per_position<double>(m, pos, n, out, 1, on_device, [&](const double* p, double* o) {
    launch_depth_get_log_odds(m->d, m->dd.occ, p, n, o, m->stream);
  });
}
int b200_map_download_occupancy(b200_map_t* m, double* dst) {
  REQUIRE(m && dst, "NULL argument");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_download_occupancy: call b200_map_set_depth_params first");
  CK(cudaSetDevice(m->prm.device));
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  const size_t nvox = (size_t)m->N[0] * m->N[1] * m->N[2];
  CK(cudaMemcpyAsync(dst, m->dd.occ, nvox * 8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}

int b200_map_build_edt(b200_map_t* m, double d_safe) { return build_edt_impl(m, d_safe, nullptr, nullptr, 0); }
int b200_map_has_distance_field(const b200_map_t* m, double d_safe) { return m && m->has_edt && m->d_safe == d_safe ? 1 : 0; }
int b200_map_set_edt_pipeline(b200_map_t* m, int32_t mode) {
  REQUIRE(m && (mode == 0 || mode == 1), "b200_map_set_edt_pipeline: mode must be 0 or 1");
  m->edt_mode = mode;
  return B200_OK;
}
int b200_map_edt_pipeline_used(const b200_map_t* m) { return m && m->edt_fused ? 1 : 0; }
int b200_map_build_edt_slab(b200_map_t* m, double d_safe, const int32_t* below_gap, const int32_t* above_gap, int32_t on_device) {
  return build_edt_impl(m, d_safe, below_gap, above_gap, on_device);
}
int b200_map_set_slab_extent(b200_map_t* m, int32_t total_nz) {
  REQUIRE(m && total_nz >= 0, "b200_map_set_slab_extent: bad argument");
  m->slab_total_nz = total_nz;
  return B200_OK;
}
int b200_map_column_extents(b200_map_t* m, int32_t* lowest, int32_t* highest, int32_t on_device) {
  REQUIRE(m && lowest && highest, "NULL argument");
  CK(cudaSetDevice(m->prm.device));
  const size_t cols = (size_t)m->N[0] * m->N[1];
  if (on_device) {
    launch_column_extents(m->d, lowest, highest, m->stream);
    m->launches++;
    CK(cudaGetLastError());
    return B200_OK;
  }
  if (m->stage_out.ensure(cols * 8)) return fail(B200_ERR_NOMEM, "extent staging");
  int32_t* base = (int32_t*)m->stage_out.p;
  launch_column_extents(m->d, base, base + cols, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(lowest, base, cols * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaMemcpyAsync(highest, base + cols, cols * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}

int b200_map_column_extents16(b200_map_t* m, int16_t* lo_hi, int32_t on_device) {
  REQUIRE(m && lo_hi, "NULL argument");
  REQUIRE(m->N[2] <= 32767, "b200_map_column_extents16: slab higher than 32767 voxels");
  CK(cudaSetDevice(m->prm.device));
  const size_t cols = (size_t)m->N[0] * m->N[1];
  if (on_device) {
    launch_column_extents16(m->d, (short2*)lo_hi, m->stream);
    m->launches++;
    CK(cudaGetLastError());
    return B200_OK;
  }
  if (m->stage_out.ensure(cols * 4)) return fail(B200_ERR_NOMEM, "extent staging");
  launch_column_extents16(m->d, (short2*)m->stage_out.p, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(lo_hi, m->stage_out.p, cols * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_slab_gaps(b200_map_t* m, const int16_t* gathered_lo_hi, const int32_t* slab_nz, int32_t n_slabs, int32_t rank,
                       int32_t* below_gap, int32_t* above_gap) {
  REQUIRE(m && gathered_lo_hi && slab_nz && below_gap && above_gap, "NULL argument");
  REQUIRE(n_slabs >= 1 && n_slabs <= kMaxSlabs && rank >= 0 && rank < n_slabs, "b200_map_slab_gaps: bad slab count / rank");
  CK(cudaSetDevice(m->prm.device));
  launch_slab_gaps((const short2*)gathered_lo_hi, slab_nz, n_slabs, rank, (long long)m->N[0] * m->N[1], below_gap, above_gap, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  return B200_OK;
}

static int download_field(b200_map_t* m, const void* src, void* dst, size_t bytes) {
  CK(cudaSetDevice(m->prm.device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}
int b200_map_device_fields(b200_map_t* m, void** dist2, void** costq) {
  REQUIRE(m && dist2 && costq, "NULL argument");
  *dist2 = m->has_edt ? (void*)m->d.dist2 : nullptr;
  *costq = m->has_edt && m->d_safe > 0 ? (void*)m->d.costq : nullptr;
  return B200_OK;
}
int b200_map_download_dist2(b200_map_t* m, int32_t* dst) {
  REQUIRE(m && dst && m->has_edt, "no distance field (call b200_map_build_edt)");
  return download_field(m, m->d.dist2, dst, (size_t)m->N[0] * m->N[1] * m->N[2] * 4);
}
// calculateDistancesMetrics / getAverage / getStdDev (src/utils/metrics.cpp:62-97): pure host arithmetic, same element
// order and operations (std::accumulate from 0.0, pow(x - mean, 2), sqrt(sum / n), minmax_element)
int b200_distance_metrics(const double* d, int64_t n, double out[4]) {
  REQUIRE(out && (n == 0 || d) && n >= 0, "b200_distance_metrics: bad argument");
  out[0] = out[1] = out[2] = out[3] = 0.0;  // metrics.cpp:77
  if (n == 0) return B200_OK;
  double sum = 0.0, mn = d[0], mx = d[0];
  for (int64_t i = 0; i < n; ++i) { sum += d[i]; if (d[i] < mn) mn = d[i]; if (!(d[i] < mx)) mx = d[i]; }
  const double mean = sum / (double)n;
  double sd = 0.0;
  for (int64_t i = 0; i < n; ++i) sd += std::pow(d[i] - mean, 2);
  sd = std::sqrt(sd / (double)n);
  out[0] = mean; out[1] = sd; out[2] = mn; out[3] = mx;
  return B200_OK;
}
// distance-to-obstacle statistics of a path: distances = sqrt(dist2[voxel(p_i)]) * resolution gathered on the device,
// then mean / population standard deviation / min / max accumulated on the host in the element order and arithmetic
// of the reference's calculateDistancesMetrics / getAverage / getStdDev (src/utils/metrics.cpp:62-97)
int b200_map_path_distance_metrics(b200_map_t* m, const double* path_xyz, int64_t n, double out[4], double* distances) {
  REQUIRE(m && out && (n == 0 || path_xyz) && n >= 0, "b200_map_path_distance_metrics: bad argument");
  REQUIRE(m->has_edt, "no distance field (call b200_map_build_edt)");
  out[0] = out[1] = out[2] = out[3] = 0.0;  // metrics.cpp:77
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  if (m->stage_in.ensure((size_t)n * 24) || m->stage_out.ensure((size_t)n * 8)) return fail(B200_ERR_NOMEM, "path staging");
  CK(cudaMemcpyAsync(m->stage_in.p, path_xyz, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  launch_distance_at(m->d, (const double*)m->stage_in.p, n, (double*)m->stage_out.p, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  std::vector<double> d((size_t)n);
  CK(cudaMemcpyAsync(d.data(), m->stage_out.p, (size_t)n * 8, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  b200_distance_metrics(d.data(), n, out);
  if (distances) memcpy(distances, d.data(), (size_t)n * 8);
  return B200_OK;
}
int b200_map_download_costq(b200_map_t* m, uint16_t* dst) {
  REQUIRE(m && dst && m->has_edt && m->d_safe > 0 && m->d.costq, "no cost field (build_edt with d_safe > 0)");
  return download_field(m, m->d.costq, dst, (size_t)m->N[0] * m->N[1] * m->N[2] * 2);
}

int b200_los_batch(b200_map_t* m, const double* s, const double* e, int64_t n, uint8_t* vis, int32_t* ns, int32_t on_device) {
  REQUIRE(m && n >= 0 && (n == 0 || (s && e && vis)), "bad argument");
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  if (on_device) {
    launch_los_batch(m->d, s, e, n, vis, ns, m->sm_count, m->stream);
    m->launches++;
    CK(cudaGetLastError());
    return B200_OK;
  }
  if (m->stage_in.ensure((size_t)n * 48) || m->stage_out.ensure((size_t)n) || m->stage_aux.ensure((size_t)n * 4))
    return fail(B200_ERR_NOMEM, "los staging");
  double* ds = (double*)m->stage_in.p;
  double* de = ds + 3 * n;
  CK(cudaMemcpyAsync(ds, s, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  CK(cudaMemcpyAsync(de, e, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  launch_los_batch(m->d, ds, de, n, (uint8_t*)m->stage_out.p, (int32_t*)m->stage_aux.p, m->sm_count, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vis, m->stage_out.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
  if (ns) CK(cudaMemcpyAsync(ns, m->stage_aux.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}

int b200_los_batch_dda(b200_map_t* m, const double* s, const double* e, int64_t n, uint8_t* vis, int32_t* nc, int32_t on_device) {
  REQUIRE(m && n >= 0 && (n == 0 || (s && e && vis)), "bad argument");
  if (n == 0) return B200_OK;
  CK(cudaSetDevice(m->prm.device));
  if (on_device) {
    launch_los_dda(m->d, s, e, n, vis, nc, m->stream);
    m->launches++;
    CK(cudaGetLastError());
    return B200_OK;
  }
  if (m->stage_in.ensure((size_t)n * 48) || m->stage_out.ensure((size_t)n) || m->stage_aux.ensure((size_t)n * 4))
    return fail(B200_ERR_NOMEM, "los staging");
  double* ds = (double*)m->stage_in.p;
  double* de = ds + 3 * n;
  CK(cudaMemcpyAsync(ds, s, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  CK(cudaMemcpyAsync(de, e, (size_t)n * 24, cudaMemcpyHostToDevice, m->stream));
  launch_los_dda(m->d, ds, de, n, (uint8_t*)m->stage_out.p, (int32_t*)m->stage_aux.p, m->stream);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(vis, m->stage_out.p, (size_t)n, cudaMemcpyDeviceToHost, m->stream));
  if (nc) CK(cudaMemcpyAsync(nc, m->stage_aux.p, (size_t)n * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}

int b200_map_export_inflate_cloud(b200_map_t* m, int32_t margin, double truncate_height, float* xyz1, int64_t cap, int64_t* n_out,
                                  int32_t on_device) {
  REQUIRE(m && n_out && cap >= 0 && (cap == 0 || xyz1) && margin >= 0, "bad argument");
  CK(cudaSetDevice(m->prm.device));
  int32_t lo[3], hi[3];
  int rc = b200_map_local_bounds(m, lo, hi);  // finishes the pending bounds reduction (grid_map.cpp:789-803)
  if (rc) return rc;
  int a[3], b[3];
  for (int i = 0; i < 3; ++i) {  // grid_map.cpp:858-869: +- local_map_margin when all_info, then boundIndex
    a[i] = std::max(std::min(lo[i] - margin, m->N[i] - 1), 0);
    b[i] = std::max(std::min(hi[i] + margin, m->N[i] - 1), 0);
  }
  *n_out = 0;
  if (a[0] > b[0] || a[1] > b[1] || a[2] > b[2]) return B200_OK;
  const long long cols = (long long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1);
  const size_t temp_bytes = export_scan_temp_bytes(cols + 1);
  const size_t ints = (size_t)(cols + 1) * 4;
  if (m->stage_aux.ensure(2 * ints + temp_bytes + 256)) return fail(B200_ERR_NOMEM, "export scratch");
  int* counts = (int*)m->stage_aux.p;
  int* offsets = counts + (cols + 1);
  void* temp = (char*)m->stage_aux.p + ((2 * ints + 255) & ~(size_t)255);
  float4* d_out = (float4*)xyz1;
  if (!on_device && cap > 0) {
    if (m->stage_out.ensure((size_t)cap * 16)) return fail(B200_ERR_NOMEM, "export staging");
    d_out = (float4*)m->stage_out.p;
  }
  launch_export(m->d, a, b, truncate_height, counts, offsets, temp, temp_bytes, d_out, cap, m->stream);
  m->launches += 3;
  CK(cudaGetLastError());
  int total = 0;
  CK(cudaMemcpyAsync(&total, offsets + cols, 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  *n_out = total;
  if (!on_device && cap > 0 && total > 0) {
    CK(cudaMemcpyAsync(xyz1, d_out, (size_t)std::min<int64_t>(total, cap) * 16, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  return B200_OK;
}

int b200_map_export_occupancy_cloud(b200_map_t* m, int32_t which, int32_t margin, double truncate_height, float* xyz1, int64_t cap,
                                    int64_t* n_out, int32_t on_device) {
  REQUIRE(m && n_out && cap >= 0 && (cap == 0 || xyz1) && margin >= 0 && (which == 0 || which == 1), "bad argument");
  if (!m->depth_set) return fail(B200_ERR_INVALID, "b200_map_export_occupancy_cloud: call b200_map_set_depth_params first");
  CK(cudaSetDevice(m->prm.device));
  int rc = depth_ensure_state(m);
  if (rc) return rc;
  int32_t lo[3], hi[3];
  rc = b200_map_local_bounds(m, lo, hi);
  if (rc) return rc;
  int a[3], b[3];
  for (int i = 0; i < 3; ++i) {  // grid_map.cpp:813-821 / :907-911: box +- margin, then boundIndex
    a[i] = std::max(std::min(lo[i] - margin, m->N[i] - 1), 0);
    b[i] = std::max(std::min(hi[i] + margin, m->N[i] - 1), 0);
  }
  *n_out = 0;
  if (a[0] > b[0] || a[1] > b[1] || a[2] > b[2]) return B200_OK;
  const long long cols = (long long)(b[0] - a[0] + 1) * (b[1] - a[1] + 1);
  const size_t temp_bytes = export_scan_temp_bytes(cols + 1);
  const size_t ints = (size_t)(cols + 1) * 4;
  if (m->stage_aux.ensure(2 * ints + temp_bytes + 256)) return fail(B200_ERR_NOMEM, "export scratch");
  int* counts = (int*)m->stage_aux.p;
  int* offsets = counts + (cols + 1);
  void* temp = (char*)m->stage_aux.p + ((2 * ints + 255) & ~(size_t)255);
  float4* d_out = (float4*)xyz1;
  if (!on_device && cap > 0) {
    if (m->stage_out.ensure((size_t)cap * 16)) return fail(B200_ERR_NOMEM, "export staging");
    d_out = (float4*)m->stage_out.p;
  }
  launch_depth_publish(m->dd, m->d, which, a, b, truncate_height, counts, offsets, temp, temp_bytes, cap > 0 ? d_out : nullptr, cap, m->stream);
  m->launches += 3;
  CK(cudaGetLastError());
  int total = 0;
  CK(cudaMemcpyAsync(&total, offsets + cols, 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  *n_out = total;
  if (!on_device && cap > 0 && total > 0) {
    CK(cudaMemcpyAsync(xyz1, d_out, (size_t)std::min<int64_t>(total, cap) * 16, cudaMemcpyDeviceToHost, m->stream));
    CK(cudaStreamSynchronize(m->stream));
  }
  return B200_OK;
}

int b200_map_launch_count(const b200_map_t* m, int64_t* n) {
  REQUIRE(m && n, "NULL argument");
  *n = m->launches;
  return B200_OK;
}
int b200_map_set_profiling(b200_map_t* m, int32_t enable) {
  REQUIRE(m, "NULL map");
  m->profiling = enable != 0;
  return B200_OK;
}
int b200_map_get_stage_ms(b200_map_t* m, double out[16]) {
  REQUIRE(m && out, "NULL argument");
  CK(cudaSetDevice(m->prm.device));
  CK(cudaStreamSynchronize(m->stream));
  memset(out, 0, 16 * sizeof(double));
  float t;
  if (cudaEventElapsedTime(&t, m->ev[0], m->ev[1]) == cudaSuccess) out[0] = t;
  if (cudaEventElapsedTime(&t, m->ev[1], m->ev[2]) == cudaSuccess) out[1] = t;
  if (m->edt_fused) {
    // with the two-stream overlap the split is: pass 1 = until BOTH halves' first kernels are done, pass 2 = the rest
    float y0 = -1.f, y1 = -1.f, all = -1.f;
    if (m->ov_ready && cudaEventElapsedTime(&y0, m->ev[3], m->ov.y_done[0]) == cudaSuccess &&
        cudaEventElapsedTime(&y1, m->ev[3], m->ov.y_done[1]) == cudaSuccess &&
        cudaEventElapsedTime(&all, m->ev[3], m->ev[5]) == cudaSuccess && y0 >= 0 && y1 >= 0 && std::max(y0, y1) <= all) {
      out[9] = std::max(y0, y1);
      out[10] = all - out[9];
    } else {
      cudaGetLastError();
      for (int i = 0; i < 2; ++i)
        if (cudaEventElapsedTime(&t, m->ev[3 + i], m->ev[4 + i]) == cudaSuccess) out[9 + i] = t;
    }
  } else {
    for (int i = 0; i < 7; ++i)
      if (cudaEventElapsedTime(&t, m->ev[3 + i], m->ev[4 + i]) == cudaSuccess) out[2 + i] = t;
  }
  cudaGetLastError();
  return B200_OK;
}

}  // extern "C"

// --------------------------------------------------------------------------------------------------
// cost (float32) download: derived from dist2 on the fly with the same formula the EDT pass quantises
// --------------------------------------------------------------------------------------------------
namespace b200 {
__global__ void __launch_bounds__(256) cost_f32_kernel(const int32_t* __restrict__ d2, float* __restrict__ out, long long n, float res, float d_safe) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int32_t v = d2[i];
    float c = 0.0f;
    if (v < kDistInf) {
      const float d = __fmul_rn(__fsqrt_rn((float)v), res);
      if (d < d_safe) c = __fsub_rn(1.0f, __fdiv_rn(d, d_safe));
    }
    out[i] = c;
  }
}
}  // namespace b200

extern "C" int b200_map_download_cost(b200_map_t* m, float* dst) {
  REQUIRE(m && dst && m->has_edt && m->d_safe > 0, "no cost field (build_edt with d_safe > 0)");
  CK(cudaSetDevice(m->prm.device));
  const size_t nv = (size_t)m->N[0] * m->N[1] * m->N[2];
  if (m->stage_out.ensure(nv * 4)) return fail(B200_ERR_NOMEM, "cost staging");
  cost_f32_kernel<<<m->sm_count * 8, 256, 0, m->stream>>>(m->d.dist2, (float*)m->stage_out.p, (long long)nv, (float)m->prm.resolution,
                                                           (float)m->d_safe);
  m->launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(dst, m->stage_out.p, nv * 4, cudaMemcpyDeviceToHost, m->stream));
  CK(cudaStreamSynchronize(m->stream));
  return B200_OK;
}

// --------------------------------------------------------------------------------------------------
// planners
// --------------------------------------------------------------------------------------------------
struct Tier {
  int cap = 0;
  int slots = 0;
  int factor = 0;
  uint32_t open_entries = 0;
  uint32_t hcap = 0;
  void* base = nullptr;
  NodeRec* nodes = nullptr;
  uint32_t* hash = nullptr;
  HeapEnt* heap = nullptr;
  double* fkey = nullptr;
};

struct b200_planner {
  int algo;
  int faithful;
  std::string name;
  b200_planner_params_t prm;
  b200_map* map;
  double dv[3];
  Tier tier[2];
  DevBuf q_in, q_res, q_path, q_order, ctl;  // ctl: [0] q_counter u64, [1] path_cursor u64, [2] overflow_count int, then overflow list
  int last_launches;
  int last_promoted;           // queries of the last call that moved into a promotion slot
  cudaStream_t stream;         // nullptr = the map's stream; own stream = batches of several planners on one map overlap
  const NodeRec* last_nodes;   // node records of the last single query (slot 0 of the tier that ran it), else nullptr
  int64_t last_used;           // use_node_num_ of that query
};

static void tier_free(Tier& t) {
  if (t.base) cudaFree(t.base);
  t = Tier();
}
static int tier_alloc(Tier& t, int cap, int slots, int factor) {
  if (t.base && t.cap == cap && t.slots == slots && t.factor == factor) return 0;
  tier_free(t);
  uint32_t hcap;
  search_slot_bytes(cap, factor, &hcap);
  const size_t nb = (size_t)slots * cap * sizeof(NodeRec);
  const size_t hb = (size_t)slots * hcap * 4;
  const size_t pb = (size_t)slots * open_region_entries(cap, factor) * sizeof(HeapEnt);
  const size_t fb = (size_t)slots * cap * sizeof(double);
  if (cudaMalloc(&t.base, nb + hb + pb + fb) != cudaSuccess) { cudaGetLastError(); t.base = nullptr; return -1; }
  t.nodes = (NodeRec*)t.base;
  t.heap = (HeapEnt*)((char*)t.base + nb);
  t.hash = (uint32_t*)((char*)t.base + nb + pb);
  t.fkey = (double*)((char*)t.base + nb + pb + hb);
  cudaMemset(t.hash, 0, hb);
  t.cap = cap; t.slots = slots; t.hcap = hcap; t.factor = factor;
  t.open_entries = (uint32_t)open_region_entries(cap, factor);
  return 0;
}

extern "C" {

int b200_planner_create(const char* algorithm, const b200_planner_params_t* p, b200_planner_t** out) {
  REQUIRE(algorithm && p && out, "NULL argument");
  REQUIRE(p->resolution > 1e-3 && p->allocate_num >= 2, "b200_planner_create: resolution must exceed 1e-3 and allocate_num >= 2");
  b200_planner* pl = new b200_planner();
  const std::string a(algorithm);
  // planner_ros_node.cpp:261-285: unknown names fall back to A*
  if (a == "thetastar") pl->algo = ALGO_THETASTAR;
  else if (a == "lazythetastar") pl->algo = ALGO_LAZYTHETASTAR;
  else if (a == "costastar") pl->algo = ALGO_COSTASTAR;
  else if (a == "costlazythetastar") pl->algo = ALGO_COSTLAZYTHETASTAR;
  else if (a == "thetastaragr") pl->algo = ALGO_THETASTARAGR;
  else if (a == "thetastaragrpp") pl->algo = ALGO_THETASTARAGRPP;
  else pl->algo = ALGO_ASTAR;
  pl->name = pl->algo == ALGO_ASTAR ? "astar" : a;
  pl->prm = *p;
  // semantics: the two algorithms the reference implements default to its exact (rb-tree) behaviour
  const bool has_reference = pl->algo == ALGO_ASTAR || pl->algo == ALGO_THETASTAR || pl->algo == ALGO_THETASTARAGR;
  const bool same_either_way = pl->algo == ALGO_THETASTARAGRPP;  // keys are never re-written inside the set
  if (p->semantics == B200_SEMANTICS_FAITHFUL && !has_reference && !same_either_way) {
    delete pl;
    return fail(B200_ERR_UNSUPPORTED, "b200_planner_create: faithful semantics exist only for astar / thetastar / thetastaragr(pp)");
  }
  if (p->semantics < 0 || p->semantics > 2) { delete pl; return fail(B200_ERR_INVALID, "b200_planner_create: bad semantics"); }
  if (p->semantics == B200_SEMANTICS_FAITHFUL && p->max_los > 0) {
    delete pl;
    return fail(B200_ERR_UNSUPPORTED, "b200_planner_create: max_los is an extension of the canonical semantics");
  }
  pl->faithful = has_reference && p->semantics != B200_SEMANTICS_CANONICAL && !(p->max_los > 0);
  pl->map = nullptr;
  pl->last_launches = 0;
  pl->stream = nullptr;
  // the neighbour offsets exactly as `for (double dx = -r; dx <= r + 1e-3; dx += r)` produces them (AStar.cpp:131)
  int nd = 0;
  double dv[8];
  for (double dx = -p->resolution; dx <= p->resolution + 1e-3 && nd < 8; dx += p->resolution) dv[nd++] = dx;
  if (nd != 3) { delete pl; return fail(B200_ERR_UNSUPPORTED, "b200_planner_create: resolution yields a neighbour stencil other than 3x3x3"); }
  for (int i = 0; i < 3; ++i) pl->dv[i] = dv[i];
  *out = pl;
  return B200_OK;
}

int b200_planner_destroy(b200_planner_t* pl) {
  if (!pl) return B200_OK;
  if (pl->map) cudaSetDevice(pl->map->prm.device);
  tier_free(pl->tier[0]);
  tier_free(pl->tier[1]);
  pl->q_in.release(); pl->q_res.release(); pl->q_path.release(); pl->q_order.release(); pl->ctl.release();
  delete pl;
  return B200_OK;
}
int b200_planner_set_map(b200_planner_t* pl, b200_map_t* m) {
  REQUIRE(pl && m, "NULL argument");
  pl->map = m;
  return B200_OK;
}
int b200_planner_set_stream(b200_planner_t* pl, void* st) {
  REQUIRE(pl, "NULL planner");
  pl->stream = (cudaStream_t)st;
  return B200_OK;
}
int b200_planner_set_cost_factor(b200_planner_t* pl, double w) {
  REQUIRE(pl, "NULL planner");
  pl->prm.cost_weight = w;
  return B200_OK;
}
int b200_planner_detect_collision(b200_planner_t* pl, const double pos[3], int32_t* collides) {
  REQUIRE(pl && pos && collides, "NULL argument");
  if (!pl->map) return fail(B200_ERR_NO_MAP, "planner has no map");
  int8_t v = 0;
  int rc = b200_map_get_inflate_occupancy(pl->map, pos, 1, &v, 0);
  if (rc) return rc;
  *collides = v != 0;  // AlgorithmBase.cpp:43: bool(int) — -1 (outside) counts as a collision
  return B200_OK;
}
int b200_planner_last_launches(b200_planner_t* pl, int32_t* n) {
  REQUIRE(pl && n, "NULL argument");
  *n = pl->last_launches;
  return B200_OK;
}

int b200_planner_find_paths(b200_planner_t* pl, const double* starts, const double* goals, int64_t nq, b200_path_result_t* results,
                            double* path_xyz, int64_t path_cap, int32_t on_device) {
  REQUIRE(pl && nq >= 0 && (nq == 0 || (starts && goals && results)), "bad argument");
  REQUIRE(nq < (1LL << 31), "too many queries in one call");
  if (!pl->map) return fail(B200_ERR_NO_MAP, "planner has no map");
  pl->last_launches = 0;
  pl->last_nodes = nullptr;
  if (nq == 0) return B200_OK;
  b200_map* m = pl->map;
  const bool cost_algo = pl->algo == ALGO_COSTASTAR || pl->algo == ALGO_COSTLAZYTHETASTAR;
  if (cost_algo && !(m->has_edt && m->d_safe > 0 && m->d.costq))
    return fail(B200_ERR_INVALID, "cost-aware planner needs b200_map_build_edt(map, d_safe > 0) first");
  CK(cudaSetDevice(m->prm.device));
  cudaStream_t st = pl->stream ? pl->stream : m->stream;
  if (!path_xyz) path_cap = 0;

  // workspace tiers: a small one for the common query, the full allocate_num one for the rare big query
  const int tpb = search_threads_per_block(pl->algo);
  const int per_sm = search_slots_per_sm(pl->algo, pl->faithful);
  int slots0 = m->sm_count * per_sm;
  if ((int64_t)slots0 > nq) slots0 = (int)nq;
  const int factor = open_region_factor(pl->algo, pl->faithful);
  // tier 0 holds 32 768 nodes per slot when that fits in a quarter of the free memory (a rerun in tier 1 costs a
  // second launch that waits for the first: 2 of 131 072 Lazy Theta* queries needed > 16 384 nodes and their
  // rerun took 75 of 290 ms), else 16 384, else fewer slots
  static const int cap0_env = getenv("B200_TIER0_CAP") ? atoi(getenv("B200_TIER0_CAP")) : 0;  // tuning knob
  int cap0 = std::min(pl->prm.allocate_num, cap0_env >= 1024 ? cap0_env : 32768);
  if (!(pl->tier[0].base && pl->tier[0].cap == cap0 && pl->tier[0].slots >= slots0)) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    if (pl->tier[0].base) free_b += (size_t)pl->tier[0].slots * search_slot_bytes(pl->tier[0].cap, pl->tier[0].factor, nullptr);
    const size_t budget = free_b / 4;
    if (cap0 > 16384 && (size_t)slots0 * search_slot_bytes(cap0, factor, nullptr) > budget) cap0 = 16384;
    while (slots0 > 1 && (size_t)slots0 * search_slot_bytes(cap0, factor, nullptr) > budget) slots0 = (slots0 + 1) / 2;
  }
  if (tier_alloc(pl->tier[0], cap0, std::max(slots0, pl->tier[0].cap == cap0 ? pl->tier[0].slots : 0), factor))
    return fail(B200_ERR_NOMEM, "search workspace (tier 0)");

  const double* d_starts = starts;
  const double* d_goals = goals;
  b200_path_result_t* d_res = results;
  double* d_path = path_xyz;
  if (!on_device) {
    if (pl->q_in.ensure((size_t)nq * 48) || pl->q_res.ensure((size_t)nq * sizeof(b200_path_result_t)) ||
        (path_cap > 0 && pl->q_path.ensure((size_t)path_cap * 24)))
      return fail(B200_ERR_NOMEM, "query staging");
    double* ds = (double*)pl->q_in.p;
    CK(cudaMemcpyAsync(ds, starts, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(ds + 3 * nq, goals, (size_t)nq * 24, cudaMemcpyHostToDevice, st));
    d_starts = ds;
    d_goals = ds + 3 * nq;
    d_res = (b200_path_result_t*)pl->q_res.p;
    d_path = path_cap > 0 ? (double*)pl->q_path.p : nullptr;
  }
  if (pl->ctl.ensure(64 + (size_t)nq * 4)) return fail(B200_ERR_NOMEM, "control block");
  CK(cudaMemsetAsync(pl->ctl.p, 0, 64, st));
  unsigned long long* q_counter = (unsigned long long*)pl->ctl.p;
  unsigned long long* path_cursor = q_counter + 1;
  int* overflow_count = (int*)(q_counter + 2);
  int* promo_count = overflow_count + 2;
  int* overflow_list = (int*)((char*)pl->ctl.p + 64);
  // Promotion slots: the canonical 32-thread kernels move a query that outgrows tier 0 into a full-size workspace and go
  // on (search_kernels.cu promote_slot) instead of reporting it for a re-run from scratch.  They are tier 1 allocated up
  // front (<= 32 slots, <= 4 GB and a quarter of the free memory, at least one); without them everything works as before,
  // through the re-run.
  const bool can_promote = !pl->faithful && tpb == 32 && pl->tier[0].cap < pl->prm.allocate_num;
  if (can_promote && !(pl->tier[1].base && pl->tier[1].cap == pl->prm.allocate_num)) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const size_t slot_b = search_slot_bytes(pl->prm.allocate_num, factor, nullptr);
    const char* cap_env = getenv("B200_PROMO_SLOTS");  // testing / tuning knob: upper bound on the promotion slots (0 = none)
    const size_t max_promo = cap_env ? (size_t)std::max(0, atoi(cap_env)) : 32;
    const size_t budget = std::min<size_t>(free_b / 4, (size_t)4 << 30);  // 32 slots of a 1 M-node workspace are 3.6 GB
    const int want = (int)std::min<size_t>({max_promo, (size_t)pl->tier[0].slots, std::max<size_t>(1, budget / std::max<size_t>(slot_b, 1))});
    if (want >= 1 && tier_alloc(pl->tier[1], pl->prm.allocate_num, want, factor)) cudaGetLastError();  // no memory: no promotion
  }

  SearchParams P;
  memset(&P, 0, sizeof P);
  P.map = m->d;
  P.r = pl->prm.resolution;
  P.lambda = pl->prm.lambda_heuristic;
  P.tie = 1.0 + 1.0 / 10000;  // AStar.cpp:40
  P.cost_weight = pl->prm.cost_weight;
  P.max_los = pl->prm.max_los;
  P.agr_barrier = pl->prm.agr_barrier; P.agr_flying_cost = pl->prm.agr_flying_cost;
  P.agr_flying_cost_default = pl->prm.agr_flying_cost_default; P.agr_ground_judge = pl->prm.agr_ground_judge;
  P.agr_epsilon = pl->prm.agr_epsilon;
  for (int i = 0; i < 3; ++i) P.dv[i] = pl->dv[i];
  P.allocate_num = pl->prm.allocate_num;
  P.algo = pl->algo;
  P.faithful = pl->faithful;
  P.starts = d_starts; P.goals = d_goals; P.qlist = nullptr; P.nq = nq;
  if (nq > 2 * (int64_t)pl->tier[0].slots) {  // more than two waves: issue the expected-longest queries first
    if (pl->q_order.ensure(query_order_bytes((int)nq))) return fail(B200_ERR_NOMEM, "query order");
    P.qlist = launch_query_order(P.map, d_starts, d_goals, (int)nq, pl->q_order.p, st);
  }
  P.results = d_res; P.path_pool = d_path; P.path_cap = path_cap; P.path_cursor = path_cursor;
  P.q_counter = q_counter; P.overflow_list = overflow_list; P.overflow_count = overflow_count;
  const Tier& t0 = pl->tier[0];
  P.nodes = t0.nodes; P.hash = t0.hash; P.heap = t0.heap; P.cap = t0.cap; P.hcap = t0.hcap; P.open_entries = t0.open_entries; P.fkey = t0.fkey;
  P.promo_count = promo_count;
  if (can_promote && pl->tier[1].base && pl->tier[1].cap == pl->prm.allocate_num) {
    const Tier& t1 = pl->tier[1];
    const char* cap_env = getenv("B200_PROMO_SLOTS");
    P.promo_nodes = t1.nodes; P.promo_hash = t1.hash; P.promo_heap = t1.heap; P.promo_cap = t1.cap;
    P.promo_slots = cap_env ? std::min(t1.slots, std::max(0, atoi(cap_env))) : t1.slots;
    P.promo_open_entries = t1.open_entries; P.promo_hcap = t1.hcap;
  }
  launch_search(P, std::min<int64_t>(t0.slots, nq), st);
  pl->last_launches++;
  CK(cudaGetLastError());
  const int promo_slots_used = P.promo_slots;
  P.promo_slots = 0;  // a re-run below uses tier 1 as its own workspace

  pl->last_promoted = 0;
  if (t0.cap < pl->prm.allocate_num) {
    int counts[3] = {0, 0, 0};  // overflow_count, (pad), promo_count
    CK(cudaMemcpyAsync(counts, overflow_count, sizeof counts, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const int n_over = counts[0];
    pl->last_promoted = std::min(counts[2], promo_slots_used);
    if (n_over > 0) {
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      uint32_t hc;
      const size_t slot_b = search_slot_bytes(pl->prm.allocate_num, factor, &hc);
      size_t budget = pl->tier[1].base ? (size_t)pl->tier[1].slots * slot_b : free_b / 2;
      int slots1 = (int)std::min<size_t>({(size_t)n_over, (size_t)m->sm_count * (tpb == 32 ? 4 : 2), std::max<size_t>(1, budget / slot_b)});
      if (pl->tier[1].base && pl->tier[1].cap == pl->prm.allocate_num && pl->tier[1].slots >= slots1) slots1 = pl->tier[1].slots;
      if (tier_alloc(pl->tier[1], pl->prm.allocate_num, slots1, factor)) return fail(B200_ERR_NOMEM, "search workspace (tier 1)");
      CK(cudaMemsetAsync(q_counter, 0, 8, st));
      const Tier& t1 = pl->tier[1];
      P.qlist = overflow_list; P.nq = n_over;
      P.overflow_list = nullptr; P.overflow_count = overflow_count;
      P.nodes = t1.nodes; P.hash = t1.hash; P.heap = t1.heap; P.cap = t1.cap; P.hcap = t1.hcap; P.open_entries = t1.open_entries; P.fkey = t1.fkey;
      launch_search(P, std::min(t1.slots, n_over), st);
      pl->last_launches++;
      CK(cudaGetLastError());
    }
  }
  if (!on_device) {
    CK(cudaMemcpyAsync(results, d_res, (size_t)nq * sizeof(b200_path_result_t), cudaMemcpyDeviceToHost, st));
    unsigned long long used = 0;
    CK(cudaMemcpyAsync(&used, path_cursor, 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (path_cap > 0) {
      const size_t npts = (size_t)std::min<unsigned long long>(used, (unsigned long long)path_cap);
      if (npts) CK(cudaMemcpyAsync(path_xyz, d_path, npts * 24, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
  }
  __atomic_fetch_add(&m->launches, (long long)pl->last_launches, __ATOMIC_RELAXED);  // planners on their own streams may run on several host threads
  return B200_OK;
}

int b200_planner_get_visited_nodes(b200_planner_t* pl, double* xyz, int64_t cap, int64_t* n_out) {
  REQUIRE(pl && n_out && cap >= 0 && (cap == 0 || xyz), "bad argument");
  if (!pl->last_nodes) return fail(B200_ERR_INVALID, "b200_planner_get_visited_nodes: the last call was not a single-query b200_planner_find_path");
  // AlgorithmBase.cpp:185-189: path_node_pool_[0 .. use_node_num_ - 1), i.e. every created node but the last, in creation order
  const int64_t n = std::max<int64_t>(pl->last_used - 1, 0);
  *n_out = n;
  const int64_t take = std::min(n, cap);
  if (take > 0) {
    CK(cudaSetDevice(pl->map->prm.device));
    cudaStream_t st = pl->stream ? pl->stream : pl->map->stream;
    CK(cudaMemcpy2DAsync(xyz, 24, pl->last_nodes, sizeof(NodeRec), 24, (size_t)take, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  return B200_OK;
}

int b200_planner_find_path(b200_planner_t* pl, const double start[3], const double goal[3], b200_path_result_t* result,
                           double* path_xyz, int64_t path_cap) {
  REQUIRE(pl && start && goal && result, "NULL argument");
  int rc = b200_planner_find_paths(pl, start, goal, 1, result, path_xyz, path_cap, 0);
  if (rc) return rc;
  if (result->path_offset < 0 && result->solved && path_xyz && path_cap > 0) {
    // path longer than the caller's buffer: rerun into a staging buffer of the right size and copy the prefix
    std::vector<double> tmp((size_t)result->n_path * 3);
    rc = b200_planner_find_paths(pl, start, goal, 1, result, tmp.data(), result->n_path, 0);
    if (rc) return rc;
    memcpy(path_xyz, tmp.data(), (size_t)std::min<int64_t>(path_cap, result->n_path) * 24);
  }
  if (result->path_offset > 0) result->path_offset = 0;
  // slot 0 of the tier that produced the result still holds the node records (positions first, creation order)
  // (a promoted single query sits in promotion slot 0 = the start of tier 1)
  const Tier& t = ((pl->last_launches >= 2 || pl->last_promoted > 0) && pl->tier[1].base) ? pl->tier[1] : pl->tier[0];
  pl->last_nodes = t.nodes;
  pl->last_used = result->explored_nodes;
  return B200_OK;
}

}  // extern "C"
