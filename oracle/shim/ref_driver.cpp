// ref_driver.cpp — drives the reference's OWN, UNMODIFIED sources (compiled from /root/reference where
// they lie, see oracle/Makefile target `ref`) through a small C interface, so tests can diff the oracle's
// restatement against what the reference code itself computes.  TEST INFRASTRUCTURE ONLY; output goes to
// oracle/_ref/libref.so (git-ignored).  No reference source text is copied here: this file only CALLS
// GridMap / Planners::AStar / Planners::ThetaStar through their own headers.
//
// The reference reads its configuration from ROS params and is fed by ROS callbacks; the header shim
// (oracle/shim/shim_ros.h) backs NodeHandle::param with a key->value table, and the callbacks
// (GridMap::odomCallback / cloudCallback, private members) are invoked directly — hence the
// `#define private public` below, which changes access only, not layout.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <numeric>
#include <queue>
#include <random>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <unordered_map>
#include <variant>
#include <vector>
#include <chrono>

#define private public
#define protected public
#include "plan_env/grid_map.h"
#include "utils/metrics.hpp"
#include "Planners/AStar.hpp"
#include "Planners/ThetaStar.hpp"
#include "Planners/ThetaStarAGR.hpp"
#include "Planners/ThetaStarAGRpp.hpp"
#undef private
#undef protected

namespace {
struct Result {  // same layout as oracle.cpp's Result / b200_path_result_t
  int32_t solved, status, n_path, open_peak;
  int64_t path_offset;
  double g_final, path_length;
  int64_t explored_nodes, iter_num, los_checks, los_samples;
  double goal_coords[3], start_coords[3];
  double time_us;
};
struct RefMap {
  GridMap::Ptr gm;
};
struct RefPlanner {
  std::unique_ptr<Planners::AlgorithmBase> algo;
  GridMap::Ptr gm;
};
struct Silence {  // the reference prints from initMap / failures; keep test logs readable
  std::streambuf* old;
  std::ostringstream sink;
  Silence() : old(std::cout.rdbuf(sink.rdbuf())) {}
  ~Silence() { std::cout.rdbuf(old); }
};
}  // namespace

extern "C" {

void ref_set_param(const char* key, double v) { shim::num_params()[key] = v; }
void ref_set_param_str(const char* key, const char* v) { shim::str_params()[key] = v; }

// GridMap::initMap reads grid_map/* params; its origin is always (-sx/2, -sy/2, ground_height)
void* ref_map_create(const double size[3], double res, double inflation, const double range[3], double ground_height) {
  Silence s;
  auto& P = shim::num_params();
  P["grid_map/resolution"] = res;
  P["grid_map/map_size_x"] = size[0]; P["grid_map/map_size_y"] = size[1]; P["grid_map/map_size_z"] = size[2];
  P["grid_map/local_update_range_x"] = range[0]; P["grid_map/local_update_range_y"] = range[1]; P["grid_map/local_update_range_z"] = range[2];
  P["grid_map/obstacles_inflation"] = inflation;
  P["grid_map/ground_height"] = ground_height;
  RefMap* m = new RefMap();
  m->gm.reset(new GridMap);
  ros::NodeHandle nh("~");
  m->gm->initMap(nh);
  return m;
}
void ref_map_destroy(void* h) { delete (RefMap*)h; }
void ref_map_dims(void* h, int out[3]) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int i = 0; i < 3; ++i) out[i] = g.mp_.map_voxel_num_(i);
}
void ref_map_region(void* h, double origin[3], double size[3]) {
  Eigen::Vector3d o, s;
  ((RefMap*)h)->gm->getRegion(o, s);
  for (int i = 0; i < 3; ++i) { origin[i] = o(i); size[i] = s(i); }
}
// one odometry message (sets camera_pos_ / has_odom_) followed by one PointCloud2 message
void ref_map_update_cloud(void* h, const void* pts, int64_t n, int step, int ox, int oy, int oz, const double cam[3]) {
  Silence s;
  GridMap& g = *((RefMap*)h)->gm;
  auto odom = std::make_shared<nav_msgs::Odometry>();
  odom->pose.pose.position.x = cam[0]; odom->pose.pose.position.y = cam[1]; odom->pose.pose.position.z = cam[2];
  g.odomCallback(odom);
  auto msg = std::make_shared<sensor_msgs::PointCloud2>();
  msg->point_step = step;
  msg->width = (uint32_t)n;
  msg->fields.resize(3);
  msg->fields[0].name = "x"; msg->fields[0].offset = ox;
  msg->fields[1].name = "y"; msg->fields[1].offset = oy;
  msg->fields[2].name = "z"; msg->fields[2].offset = oz;
  msg->data.assign((const uint8_t*)pts, (const uint8_t*)pts + (size_t)n * step);
  g.cloudCallback(msg);
}
void ref_map_clear_box(void* h, const double mn[3], const double mx[3]) {
  ((RefMap*)h)->gm->resetBuffer(Eigen::Vector3d(mn[0], mn[1], mn[2]), Eigen::Vector3d(mx[0], mx[1], mx[2]));
}
void ref_map_download_inflate(void* h, uint8_t* dst) {
  GridMap& g = *((RefMap*)h)->gm;
  std::memcpy(dst, g.md_.occupancy_buffer_inflate_.data(), g.md_.occupancy_buffer_inflate_.size());
}
void ref_map_upload_inflate(void* h, const uint8_t* src) {
  GridMap& g = *((RefMap*)h)->gm;
  std::memcpy(g.md_.occupancy_buffer_inflate_.data(), src, g.md_.occupancy_buffer_inflate_.size());
}
void ref_map_local_bounds(void* h, int mn[3], int mx[3]) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int i = 0; i < 3; ++i) { mn[i] = g.md_.local_bound_min_(i); mx[i] = g.md_.local_bound_max_(i); }
}
// GridMap::publishMapInflate (grid_map.cpp:851-900) with one fake subscriber; returns the PointXYZ records it published
int64_t ref_map_publish_inflate(void* h, int all_info, float* out, int64_t cap) {
  GridMap& g = *((RefMap*)h)->gm;
  shim::fake_subscribers() = 1;
  shim::last_cloud_blob().clear();
  g.publishMapInflate(all_info != 0);
  shim::fake_subscribers() = 0;
  const int64_t n = (int64_t)(shim::last_cloud_blob().size() / 16);
  std::memcpy(out, shim::last_cloud_blob().data(), (size_t)std::min<int64_t>(n, cap) * 16);
  return n;
}
// the reference's own RayCaster (raycast.cpp:253-346): cells returned by step() until it reports the end
int64_t ref_ray_cells(const double a[3], const double b[3], int32_t* out, int64_t cap) {
  RayCaster rc;
  rc.setInput(Eigen::Vector3d(a[0], a[1], a[2]), Eigen::Vector3d(b[0], b[1], b[2]));
  Eigen::Vector3d p;
  int64_t n = 0;
  while (n < 1000000 && rc.step(p)) {
    if (n < cap) { out[3 * n] = (int32_t)p(0); out[3 * n + 1] = (int32_t)p(1); out[3 * n + 2] = (int32_t)p(2); }
    ++n;
  }
  return n;
}
// GridMap::publishMap (grid_map.cpp:806-849, which = 0, one fake subscriber) / publishUnknown (:902-939, which = 1)
int64_t ref_map_publish_occupancy(void* h, int which, float* out, int64_t cap) {
  GridMap& g = *((RefMap*)h)->gm;
  shim::fake_subscribers() = 1;
  shim::last_cloud_blob().clear();
  if (which == 0) g.publishMap(); else g.publishUnknown();
  shim::fake_subscribers() = 0;
  const int64_t n = (int64_t)(shim::last_cloud_blob().size() / 16);
  std::memcpy(out, shim::last_cloud_blob().data(), (size_t)std::min<int64_t>(n, cap) * 16);
  return n;
}
// One depth frame through the reference's own callbacks: depthPoseCallback (image copy, pose, the in-map gate,
// grid_map.cpp:668-697) then the occupancy timer body updateOccupancyCallback (projectDepthImage -> raycastProcess
// -> clearAndInflateLocalMap, grid_map.cpp:635-666).  Returns 1 if the frame was fused (camera inside the map).
// rot_out (row-major 3x3) = the rotation the reference derived from the quaternion, for the caller to hand to the
// implementations under test (their boundary takes the matrix).
int ref_map_update_depth(void* h, const uint16_t* depth, int rows, int cols, const double cam_pos[3], const double quat_wxyz[4],
                         double rot_out[9]) {
  Silence s;
  GridMap& g = *((RefMap*)h)->gm;
  auto img = std::make_shared<sensor_msgs::Image>();
  img->height = (uint32_t)rows; img->width = (uint32_t)cols;
  img->encoding = sensor_msgs::image_encodings::TYPE_16UC1;
  img->data.resize((size_t)rows * cols * 2);
  std::memcpy(img->data.data(), depth, img->data.size());
  auto pose = std::make_shared<geometry_msgs::PoseStamped>();
  pose->pose.position.x = cam_pos[0]; pose->pose.position.y = cam_pos[1]; pose->pose.position.z = cam_pos[2];
  pose->pose.orientation.w = quat_wxyz[0]; pose->pose.orientation.x = quat_wxyz[1];
  pose->pose.orientation.y = quat_wxyz[2]; pose->pose.orientation.z = quat_wxyz[3];
  g.depthPoseCallback(img, pose);
  if (rot_out) {
    const Eigen::Matrix3d r = g.md_.camera_q_.toRotationMatrix();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) rot_out[3 * i + j] = r(i, j);
  }
  const int fused = g.md_.occ_need_update_ ? 1 : 0;
  g.updateOccupancyCallback(ros::TimerEvent());
  return fused;
}
// GridMap::depthOdomCallback (grid_map.cpp:965-994) with a 1x1 image: the camera pose the reference derives from a body
// odometry (quaternion -> matrix, body2world * cam2body_, matrix -> quaternion).  Marks the map for an occupancy update with
// that 1x1 image: use a scratch map.
void ref_map_depth_odom_pose(void* h, const double body_pos[3], const double body_q_wxyz[4], double cam_pos[3], double cam_q_wxyz[4]) {
  Silence s;
  GridMap& g = *((RefMap*)h)->gm;
  auto img = std::make_shared<sensor_msgs::Image>();
  img->height = 1; img->width = 1;
  img->encoding = sensor_msgs::image_encodings::TYPE_16UC1;
  img->data.assign(2, 0);
  auto odom = std::make_shared<nav_msgs::Odometry>();
  odom->pose.pose.position.x = body_pos[0]; odom->pose.pose.position.y = body_pos[1]; odom->pose.pose.position.z = body_pos[2];
  odom->pose.pose.orientation.w = body_q_wxyz[0]; odom->pose.pose.orientation.x = body_q_wxyz[1];
  odom->pose.pose.orientation.y = body_q_wxyz[2]; odom->pose.pose.orientation.z = body_q_wxyz[3];
  g.depthOdomCallback(img, odom);
  for (int i = 0; i < 3; ++i) cam_pos[i] = g.md_.camera_pos_(i);
  cam_q_wxyz[0] = g.md_.camera_q_.w_; cam_q_wxyz[1] = g.md_.camera_q_.x_; cam_q_wxyz[2] = g.md_.camera_q_.y_; cam_q_wxyz[3] = g.md_.camera_q_.z_;
}
void ref_map_download_occupancy(void* h, double* dst) {
  const auto& b = ((RefMap*)h)->gm->md_.occupancy_buffer_;
  std::memcpy(dst, b.data(), b.size() * sizeof(double));
}
int ref_map_raycast_num(void* h) { return ((RefMap*)h)->gm->md_.raycast_num_; }
int ref_map_proj_points(void* h) { return ((RefMap*)h)->gm->md_.proj_points_cnt; }
void ref_map_set_raycast_num(void* h, int v) { ((RefMap*)h)->gm->md_.raycast_num_ = v; }
void ref_map_get_occupancy(void* h, const double* pos, int64_t n, int8_t* out) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) out[i] = (int8_t)g.getOccupancy(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void ref_map_get_inflate_occupancy(void* h, const double* pos, int64_t n, int8_t* out) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) out[i] = (int8_t)g.getInflateOccupancy(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
// GridMap::setOccupied / setOccupancy (grid_map.h:310-337), one call per position, in order
void ref_map_set_occupied(void* h, const double* pos, int64_t n) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) g.setOccupied(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void ref_map_set_occupancy(void* h, const double* pos, const double* occ, int64_t n) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) g.setOccupancy(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), occ[i]);
}
// calculateDistancesMetrics (src/utils/metrics.cpp:62-78): {mean, std_dev, min, max} of the distances
void ref_distance_metrics(const double* d, int64_t n, double out[4]) {
  std::vector<std::pair<Eigen::Vector3i, double>> v;
  for (int64_t i = 0; i < n; ++i) v.push_back({Eigen::Vector3i(0, 0, 0), d[i]});
  const auto r = Planners::utils::metrics::calculateDistancesMetrics(v);
  out[0] = std::get<0>(r); out[1] = std::get<1>(r); out[2] = std::get<2>(r); out[3] = std::get<3>(r);
}
void ref_map_is_in_map(void* h, const double* pos, int64_t n, uint8_t* out) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) out[i] = g.isInMap(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]));
}
void ref_pos_to_index(void* h, const double* pos, int64_t n, int32_t* out) {
  GridMap& g = *((RefMap*)h)->gm;
  for (int64_t i = 0; i < n; ++i) {
    Eigen::Vector3i id;
    g.posToIndex(Eigen::Vector3d(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]), id);
    out[3 * i] = id(0); out[3 * i + 1] = id(1); out[3 * i + 2] = id(2);
  }
}
// LineOfSight::fastLOS takes two NodePtr
void ref_los_batch(void* h, const double* s, const double* e, int64_t n, uint8_t* vis) {
  GridMap::Ptr gm = ((RefMap*)h)->gm;
  Node0 a, b;
  for (int64_t i = 0; i < n; ++i) {
    a.position = Eigen::Vector3d(s[3 * i], s[3 * i + 1], s[3 * i + 2]);
    b.position = Eigen::Vector3d(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
    vis[i] = Planners::utils::LineOfSight::fastLOS(&a, &b, gm) ? 1 : 0;
  }
}

// algo: 0 astar, 1 thetastar, 2 thetastaragr, 3 thetastaragrpp (planner_ros_node.cpp:261-279: new X(); setParam();
// setGridMap; init()).  Both AGR variants read the thetastaragr/ namespace (ThetaStarAGRpp.cpp:28-37).
void* ref_planner_create(int algo, double resolution, double lambda_heuristic, int allocate_num, void* map) {
  Silence s;
  const std::string ns = algo == 0 ? "astar/" : (algo == 1 ? "thetastar/" : "thetastaragr/");
  auto& P = shim::num_params();
  P[ns + "resolution"] = resolution;
  P[ns + "time_resolution"] = 0.8;
  P[ns + "lambda_heuristic"] = lambda_heuristic;
  P[ns + "allocate_num"] = allocate_num;
  // The reference's open-list tie-break is the ADDRESS of its individually new-ed nodes (utils.hpp:55-60,
  // AStar.cpp:24-26), i.e. creation order on a fresh heap (what the ROS node has at start-up).  This process's
  // heap is full of holes (Python, numpy), so plug every hole a Node0 could land in before init() allocates
  // the pool, release the plugs afterwards, and retry (leaking the attempt) if the pool still came out unordered.
  std::vector<Node0*> plugs;
  RefPlanner* p = nullptr;
  for (int attempt = 0; attempt < 8; ++attempt) {
    p = new RefPlanner();
    if (algo == 1) p->algo.reset(new Planners::ThetaStar());
    else if (algo == 2) p->algo.reset(new Planners::ThetaStarAGR());
    else if (algo == 3) p->algo.reset(new Planners::ThetaStarAGRpp());
    else p->algo.reset(new Planners::AStar());
    // setParam is virtual in AlgorithmBase but ThetaStar and the AGR classes re-declare it non-virtually with the same
    // signature, so the call through the base pointer dispatches to the most derived one, as in the node
    p->algo->setParam();
    p->gm = ((RefMap*)map)->gm;
    p->algo->setGridMap(p->gm);
    int increasing = 0;
    const size_t start = plugs.size();
    while (increasing < 8192 && plugs.size() - start < 4000000) {
      Node0* n = new Node0();
      increasing = (!plugs.empty() && plugs.back() < n) ? increasing + 1 : 0;
      plugs.push_back(n);
    }
    p->algo->init();
    bool ordered = true;
    for (size_t i = 1; i < p->algo->path_node_pool_.size() && ordered; ++i)
      ordered = p->algo->path_node_pool_[i - 1] < p->algo->path_node_pool_[i];
    if (ordered) break;
    p->algo.release();  // leak: its nodes now plug the holes they fell into
    delete p;
    p = nullptr;
  }
  for (Node0* n : plugs) delete n;
  return p;
}
// Reference tie-break is the raw pointer value of individually new-ed nodes (utils.hpp:55-60, AStar.cpp:24-26):
// creation order only when the allocator hands out increasing addresses (fresh heap).  Returns the number of
// adjacent pool entries that are NOT in increasing address order.
int ref_planner_pool_disorder(void* h) {
  RefPlanner* p = (RefPlanner*)h;
  int bad = 0;
  for (size_t i = 1; i < p->algo->path_node_pool_.size(); ++i)
    if (!(p->algo->path_node_pool_[i - 1] < p->algo->path_node_pool_[i])) ++bad;
  return bad;
}
void ref_planner_destroy(void* h) {
  RefPlanner* p = (RefPlanner*)h;
  p->algo->reset();  // the destructor deletes use_node_num_ nodes only (AStar.cpp:8-12)
  delete p;
}
// requestPathService order: reset() -> findPath() -> read PathData -> getPath() (planner_ros_node.cpp:136-214)
int ref_planner_find_path(void* h, const double s[3], const double g[3], Result* r, double* path, int64_t path_cap) {
  Silence quiet;
  RefPlanner* p = (RefPlanner*)h;
  p->algo->reset();
  Planners::utils::PathData d = p->algo->findPath(Eigen::Vector3d(s[0], s[1], s[2]), Eigen::Vector3d(g[0], g[1], g[2]));
  std::memset(r, 0, sizeof *r);
  r->solved = std::get<bool>(d["solved"]) ? 1 : 0;
  r->g_final = std::get<double>(d["g_final_node"]);
  r->path_length = std::get<double>(d["path_length"]);
  r->explored_nodes = (int64_t)std::get<size_t>(d["explored_nodes"]);
  r->los_checks = std::get<unsigned int>(d["line_of_sight_checks"]);
  r->los_samples = -1;
  r->iter_num = p->algo->iter_num_;
  const Eigen::Vector3d gc = std::get<Eigen::Vector3d>(d["goal_coords"]), sc = std::get<Eigen::Vector3d>(d["start_coords"]);
  for (int i = 0; i < 3; ++i) { r->goal_coords[i] = gc(i); r->start_coords[i] = sc(i); }
  r->time_us = std::get<double>(d["time_spent"]);
  const std::vector<Eigen::Vector3d> pth = std::get<std::vector<Eigen::Vector3d>>(d["path"]);
  r->n_path = (int32_t)pth.size();
  r->status = r->solved ? 0 : (p->algo->use_node_num_ == p->algo->allocate_num_ ? 2 : 1);
  const int64_t n = std::min<int64_t>((int64_t)pth.size(), path_cap);
  for (int64_t i = 0; i < n; ++i) { path[3 * i] = pth[i](0); path[3 * i + 1] = pth[i](1); path[3 * i + 2] = pth[i](2); }
  return 0;
}

// AlgorithmBase::getVisitedNodes (AlgorithmBase.cpp:185-189) after the last findPath
int64_t ref_planner_visited(void* h, double* out, int64_t cap) {
  RefPlanner* p = (RefPlanner*)h;
  const auto v = p->algo->getVisitedNodes();
  const int64_t n = (int64_t)v.size();
  for (int64_t i = 0; i < std::min(n, cap); ++i) {
    out[3 * i] = v[i]->position(0); out[3 * i + 1] = v[i]->position(1); out[3 * i + 2] = v[i]->position(2);
  }
  return n;
}

}  // extern "C"
