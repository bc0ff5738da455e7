// b200_grid_map.hpp — host-side mirror of the reference's `GridMap` public interface
// (include/plan_env/grid_map.h:139-196 + the inline accessors :257-446) on top of the C-ABI
// (include/b200_planner.h).  Header-only, no ROS / Eigen / PCL needed: `Vec3d` / `Vec3i` are
// Eigen::Vector3d / Vector3i when B200_WITH_EIGEN is defined (drop-in for the ROS node), else the two
// small structs below (self-test, non-ROS hosts).
//
// What is mirrored: initMap (geometry + parameters of the cloud path), cloudCallback, odomCallback
// (camera position), resetBuffer (x2), posToIndex, indexToPos, toAddress, isInMap (x2), boundIndex,
// getInflateOccupancy, setOccupied, setOccupancy, getRegion, getResolution, getOrigin, odomValid, plus the new
// buildDistanceField / ensureDistanceField / pathDistanceMetrics, and the depth-image path (depthPoseCallback = the reference's callback plus its
// occupancy timer body; getOccupancy, isKnownOccupied, hasDepthObservation).
#pragma once
#include <cmath>
#include <cstdint>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "b200_planner.h"

#ifdef B200_WITH_EIGEN
#include <Eigen/Eigen>
namespace b200host {
using Vec3d = Eigen::Vector3d;
using Vec3i = Eigen::Vector3i;
}
#else
namespace b200host {
struct Vec3d {
  double v[3];
  Vec3d() : v{0, 0, 0} {}
  Vec3d(double x, double y, double z) : v{x, y, z} {}
  double& operator()(int i) { return v[i]; }
  const double& operator()(int i) const { return v[i]; }
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  const double* data() const { return v; }
  double* data() { return v; }
};
struct Vec3i {
  int v[3];
  Vec3i() : v{0, 0, 0} {}
  Vec3i(int x, int y, int z) : v{x, y, z} {}
  int& operator()(int i) { return v[i]; }
  const int& operator()(int i) const { return v[i]; }
};
}  // namespace b200host
#endif

namespace b200host {

struct B200Error : std::runtime_error {
  int code;
  B200Error(int c, const std::string& what) : std::runtime_error(what + ": " + b200_last_error()), code(c) {}
};
inline void b200_check(int rc, const char* what) {
  if (rc != B200_OK) throw B200Error(rc, what);
}

// ROS-private-parameter stand-in: the same keys the reference reads with node_.param(...)
// (grid_map.cpp:12-50, AStar.cpp:32-35, ThetaStar.cpp:26-29).  With ROS present, fill it from the
// NodeHandle (see planner_ros_node_b200.cpp).
struct ParamMap {
  std::map<std::string, double> num;
  std::map<std::string, std::string> str;
  template <class T> void param(const std::string& key, T& out, const T& def) const {
    auto it = num.find(key);
    out = it == num.end() ? def : (T)it->second;
  }
};

class GridMap {
 public:
  typedef std::shared_ptr<GridMap> Ptr;
  enum { POSE_STAMPED = 1, ODOMETRY = 2, INVALID_IDX = -10000 };

  GridMap() {}
  ~GridMap() { if (h_) b200_map_destroy(h_); }
  GridMap(const GridMap&) = delete;
  GridMap& operator=(const GridMap&) = delete;

  // GridMap::initMap (grid_map.cpp:6-142): same parameter names and defaults
  void initMap(const ParamMap& nh, int device = 0) {
    double x_size, y_size, z_size;
    nh.param("grid_map/resolution", prm_.resolution, 0.1);
    nh.param("grid_map/map_size_x", x_size, 40.0);
    nh.param("grid_map/map_size_y", y_size, 40.0);
    nh.param("grid_map/map_size_z", z_size, 5.0);
    nh.param("grid_map/local_update_range_x", prm_.local_update_range[0], 5.0);
    nh.param("grid_map/local_update_range_y", prm_.local_update_range[1], 5.0);
    nh.param("grid_map/local_update_range_z", prm_.local_update_range[2], 5.0);
    nh.param("grid_map/obstacles_inflation", prm_.obstacles_inflation, 0.099);
    nh.param("grid_map/ground_height", prm_.ground_height, -0.01);
    prm_.origin[0] = -x_size / 2.0; prm_.origin[1] = -y_size / 2.0; prm_.origin[2] = prm_.ground_height;  // grid_map.cpp:53
    prm_.size[0] = x_size; prm_.size[1] = y_size; prm_.size[2] = z_size;
    prm_.device = device;
    prm_.reserved = 0;
    // depth-fusion parameters (grid_map.cpp:21-43), same keys and defaults
    bool use_filter;
    nh.param("grid_map/fx", dprm_.fx, 387.229248046875);
    nh.param("grid_map/fy", dprm_.fy, 387.229248046875);
    nh.param("grid_map/cx", dprm_.cx, 321.04638671875);
    nh.param("grid_map/cy", dprm_.cy, 243.44969177246094);
    nh.param("grid_map/use_depth_filter", use_filter, true);
    dprm_.use_depth_filter = use_filter ? 1 : 0;
    nh.param("grid_map/depth_filter_maxdist", dprm_.depth_filter_maxdist, 5.0);
    nh.param("grid_map/depth_filter_mindist", dprm_.depth_filter_mindist, 0.2);
    nh.param("grid_map/depth_filter_margin", dprm_.depth_filter_margin, (int32_t)1);
    nh.param("grid_map/k_depth_scaling_factor", dprm_.k_depth_scaling_factor, 1000.0);
    nh.param("grid_map/skip_pixel", dprm_.skip_pixel, (int32_t)2);
    nh.param("grid_map/p_hit", dprm_.p_hit, 0.70);
    nh.param("grid_map/p_miss", dprm_.p_miss, 0.35);
    nh.param("grid_map/p_min", dprm_.p_min, 0.12);
    nh.param("grid_map/p_max", dprm_.p_max, 0.97);
    nh.param("grid_map/p_occ", dprm_.p_occ, 0.80);
    nh.param("grid_map/max_ray_length", dprm_.max_ray_length, 4.5);
    nh.param("grid_map/virtual_ceil_height", dprm_.virtual_ceil_height, 2.5);
    nh.param("grid_map/local_map_margin", dprm_.local_map_margin, (int32_t)1);
    create();
    b200_check(b200_map_set_depth_params(h_, &dprm_), "initMap (depth parameters)");
  }
  // explicit-origin variant (the C-ABI does not force the centred origin)
  void initMap(const b200_map_params_t& p) { prm_ = p; create(); }
  void setDepthParams(const b200_depth_params_t& d) { dprm_ = d; b200_check(b200_map_set_depth_params(h_, &dprm_), "setDepthParams"); }

  // odomCallback (grid_map.cpp:699-709)
  void odomCallback(const Vec3d& position) { camera_pos_ = position; has_odom_ = true; }
  // cloudCallback (grid_map.cpp:711-804) on a raw PointCloud2 payload
  void cloudCallback(const void* data, int64_t n_points, int point_step, int off_x = 0, int off_y = 4, int off_z = 8) {
    has_cloud_ = true;
    if (!has_odom_) return;  // "no odom!" (:718-722)
    const double cam[3] = {camera_pos_(0), camera_pos_(1), camera_pos_(2)};
    b200_check(b200_map_update_cloud(h_, data, n_points, point_step, off_x, off_y, off_z, cam, 0), "cloudCallback");
  }
#ifdef B200_WITH_ROS
  void cloudCallback(const sensor_msgs::PointCloud2ConstPtr& img) {
    int ox = 0, oy = 4, oz = 8;
    for (const auto& f : img->fields) { if (f.name == "x") ox = f.offset; else if (f.name == "y") oy = f.offset; else if (f.name == "z") oz = f.offset; }
    cloudCallback(img->data.data(), (int64_t)img->width * img->height, img->point_step, ox, oy, oz);
  }
  void odomCallback(const nav_msgs::OdometryConstPtr& odom) {
    odomCallback(Vec3d(odom->pose.pose.position.x, odom->pose.pose.position.y, odom->pose.pose.position.z));
  }
#endif

  // depthPoseCallback (grid_map.cpp:668-697) + the 20 Hz occupancy timer body updateOccupancyCallback (:635-666) in one
  // call: a CV_16UC1 image (a 32FC1 image is converted by the caller with k_depth_scaling_factor, :674-677), the camera
  // position and orientation quaternion (w, x, y, z).  The rotation matrix is formed here with the textbook
  // unit-quaternion formula; with Eigen present pass camera_q.toRotationMatrix() to the matrix overload instead.
  bool depthPoseCallback(const uint16_t* depth, int rows, int cols, const Vec3d& pos, double qw, double qx, double qy, double qz) {
    const double R[9] = {1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw),     2 * (qx * qz + qy * qw),
                         2 * (qx * qy + qz * qw),     1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw),
                         2 * (qx * qz - qy * qw),     2 * (qy * qz + qx * qw),     1 - 2 * (qx * qx + qy * qy)};
    return depthPoseCallback(depth, rows, cols, pos, R);
  }
  bool depthPoseCallback(const uint16_t* depth, int rows, int cols, const Vec3d& pos, const double cam_rot_row_major[9]) {
    camera_pos_ = pos;
    const double cam[3] = {pos(0), pos(1), pos(2)};
    int32_t fused = 0;
    b200_check(b200_map_update_depth(h_, depth, rows, cols, cam, cam_rot_row_major, 0, &fused), "depthPoseCallback");
    if (fused) { has_odom_ = true; has_depth_ = true; }  // :688-693
    return fused != 0;
  }
  // depthOdomCallback (grid_map.cpp:965-994): the camera pose is the body odometry composed with cam2body_
  // (grid_map.cpp:91) and then goes matrix -> quaternion -> matrix like the reference's md_.camera_q_ (:982, :208);
  // b200_odom_to_camera_pose restates that algebra in Eigen's operation order (pinned to the reference's own
  // depthOdomCallback in tests/test_oracle_vs_ref.py).  A frame whose camera is outside the map is dropped like in
  // depthPoseCallback (the reference fuses it: undefined behaviour there).
  bool depthOdomCallback(const uint16_t* depth, int rows, int cols, const Vec3d& body_pos, double qw, double qx, double qy, double qz) {
    const double p[3] = {body_pos(0), body_pos(1), body_pos(2)}, q[4] = {qw, qx, qy, qz};
    double cp[3], cq[4], R[9];
    b200_check(b200_odom_to_camera_pose(p, q, cp, cq, R), "depthOdomCallback");
    return depthPoseCallback(depth, rows, cols, Vec3d(cp[0], cp[1], cp[2]), R);
  }
#ifdef B200_WITH_ROS
  void depthPoseCallback(const sensor_msgs::ImageConstPtr& img, const geometry_msgs::PoseStampedConstPtr& pose) {
    cv_bridge::CvImagePtr cv_ptr = cv_bridge::toCvCopy(img, img->encoding);
    if (img->encoding == sensor_msgs::image_encodings::TYPE_32FC1) (cv_ptr->image).convertTo(cv_ptr->image, CV_16UC1, dprm_.k_depth_scaling_factor);
    const auto& o = pose->pose.orientation;
    depthPoseCallback(cv_ptr->image.ptr<uint16_t>(0), cv_ptr->image.rows, cv_ptr->image.cols,
                      Vec3d(pose->pose.position.x, pose->pose.position.y, pose->pose.position.z), o.w, o.x, o.y, o.z);
  }
#endif
#ifdef B200_WITH_ROS
  void depthOdomCallback(const sensor_msgs::ImageConstPtr& img, const nav_msgs::OdometryConstPtr& odom) {  // grid_map.cpp:965-994
    cv_bridge::CvImagePtr cv_ptr = cv_bridge::toCvCopy(img, img->encoding);
    if (img->encoding == sensor_msgs::image_encodings::TYPE_32FC1) (cv_ptr->image).convertTo(cv_ptr->image, CV_16UC1, dprm_.k_depth_scaling_factor);
    const auto& o = odom->pose.pose.orientation;
    const auto& p = odom->pose.pose.position;
    depthOdomCallback(cv_ptr->image.ptr<uint16_t>(0), cv_ptr->image.rows, cv_ptr->image.cols, Vec3d(p.x, p.y, p.z), o.w, o.x, o.y, o.z);
  }
#endif
  // grid_map.h:310-337 — setOccupied writes the inflated grid, setOccupancy the log-odds grid (values 0 / 1 only)
  inline void setOccupied(Vec3d pos) {
    const double p[3] = {pos(0), pos(1), pos(2)};
    b200_check(b200_map_set_occupied(h_, p, 1, 0), "setOccupied");
  }
  inline void setOccupancy(Vec3d pos, double occ = 1) {
    if (occ != 1 && occ != 0) { std::cout << "occ value error!" << std::endl; return; }  // grid_map.h:325-329
    const double p[3] = {pos(0), pos(1), pos(2)};
    b200_check(b200_map_set_occupancy(h_, p, &occ, 1, 0), "setOccupancy");
  }
  // grid_map.h:339-348 / :276-311 — single reads of occupancy_buffer_ (log-odds) and the inflated grid
  inline int getOccupancy(Vec3d pos) const {
    int8_t v = 0;
    const double p[3] = {pos(0), pos(1), pos(2)};
    b200_check(b200_map_get_occupancy(h_, p, 1, &v, 0), "getOccupancy");
    return (int)v;
  }
  inline int getOccupancy(Vec3i id) const {
    if (id(0) < 0 || id(0) >= n_[0] || id(1) < 0 || id(1) >= n_[1] || id(2) < 0 || id(2) >= n_[2]) return -1;
    Vec3d pos;
    indexToPos(id, pos);
    return getOccupancy(pos);
  }
  // grid_map.h:276-300 — isUnknown / isKnownFree read the raw log-odds of the (clamped) voxel
  inline double logOdds(const Vec3d& pos) const {
    double v = 0;
    const double p[3] = {pos(0), pos(1), pos(2)};
    b200_check(b200_map_get_log_odds(h_, p, 1, &v, 0), "log odds");
    return v;
  }
  inline bool isUnknown(const Vec3d& pos) const { return logOdds(pos) < clampMinLog() - 1e-3; }
  inline bool isUnknown(const Vec3i& id) const {
    Vec3i id1 = id;
    boundIndex(id1);
    Vec3d pos;
    indexToPos(id1, pos);
    return isUnknown(pos);
  }
  inline bool isKnownFree(const Vec3i& id) const {
    Vec3i id1 = id;
    boundIndex(id1);
    Vec3d pos;
    indexToPos(id1, pos);
    return logOdds(pos) >= clampMinLog() && getInflateOccupancy(pos) == 0;
  }
  inline bool isKnownOccupied(const Vec3i& id) const {
    Vec3i id1 = id;
    boundIndex(id1);
    Vec3d pos;
    indexToPos(id1, pos);
    return getInflateOccupancy(pos) == 1;
  }
  bool hasDepthObservation() const { return has_depth_; }
  const b200_depth_params_t& depthParams() const { return dprm_; }
  double clampMinLog() const { return std::log(dprm_.p_min / (1 - dprm_.p_min)); }  // grid_map.h:27, grid_map.cpp:59

  void resetBuffer() { b200_check(b200_map_reset(h_), "resetBuffer"); }
  void resetBuffer(Vec3d min_pos, Vec3d max_pos) {
    const double a[3] = {min_pos(0), min_pos(1), min_pos(2)}, b[3] = {max_pos(0), max_pos(1), max_pos(2)};
    b200_check(b200_map_clear_box(h_, a, b), "resetBuffer(min,max)");
  }

  // grid_map.h:400-404 / :257-265 / :267-274 / :370-398 — pure host arithmetic, same expressions
  inline void posToIndex(const Vec3d& pos, Vec3i& id) const {
    for (int i = 0; i < 3; ++i) id(i) = (int)std::floor((pos(i) - prm_.origin[i]) * res_inv_);
  }
  inline void indexToPos(const Vec3i& id, Vec3d& pos) const {
    for (int i = 0; i < 3; ++i) pos(i) = (id(i) + 0.5) * prm_.resolution + prm_.origin[i];
  }
  inline int toAddress(const Vec3i& id) const { return id(0) * n_[1] * n_[2] + id(1) * n_[2] + id(2); }
  inline int toAddress(int& x, int& y, int& z) const { return x * n_[1] * n_[2] + y * n_[2] + z; }
  inline void boundIndex(Vec3i& id) const {
    for (int i = 0; i < 3; ++i) id(i) = std::max(std::min(id(i), n_[i] - 1), 0);
  }
  inline bool isInMap(const Vec3d& pos) const {
    for (int i = 0; i < 3; ++i)
      if (pos(i) < prm_.origin[i] + 1e-4 || pos(i) > prm_.origin[i] + prm_.size[i] - 1e-4) return false;
    return true;
  }
  inline bool isInMap(const Vec3i& idx) const {
    for (int i = 0; i < 3; ++i)
      if (idx(i) < 0 || idx(i) > n_[i] - 1) return false;
    return true;
  }
  // grid_map.h:350-359 — one device read (the node calls this twice per request, via detectCollision)
  inline int getInflateOccupancy(Vec3d pos) const {
    const double p[3] = {pos(0), pos(1), pos(2)};
    int8_t v = 0;
    b200_check(b200_map_get_inflate_occupancy(h_, p, 1, &v, 0), "getInflateOccupancy");
    return (int)v;
  }
  void getInflateOccupancy(const std::vector<Vec3d>& pos, std::vector<int8_t>& out) const {
    std::vector<double> flat(pos.size() * 3);
    for (size_t i = 0; i < pos.size(); ++i) for (int k = 0; k < 3; ++k) flat[3 * i + k] = pos[i](k);
    out.resize(pos.size());
    b200_check(b200_map_get_inflate_occupancy(h_, flat.data(), (int64_t)pos.size(), out.data(), 0), "getInflateOccupancy");
  }
  void getRegion(Vec3d& ori, Vec3d& size) const {
    ori = Vec3d(prm_.origin[0], prm_.origin[1], prm_.origin[2]);
    size = Vec3d(prm_.size[0], prm_.size[1], prm_.size[2]);
  }
  inline double getResolution() const { return prm_.resolution; }
  Vec3d getOrigin() const { return Vec3d(prm_.origin[0], prm_.origin[1], prm_.origin[2]); }
  int getVoxelNum() const { return n_[0] * n_[1] * n_[2]; }
  bool odomValid() const { return has_odom_; }
  void getLocalBounds(Vec3i& mn, Vec3i& mx) const {
    int32_t a[3], b[3];
    b200_check(b200_map_local_bounds(h_, a, b), "local bounds");
    mn = Vec3i(a[0], a[1], a[2]); mx = Vec3i(b[0], b[1], b[2]);
  }
  // occupancy_buffer_inflate_ (reference byte layout) for publishers / parity
  std::vector<uint8_t> downloadInflate() const {
    std::vector<uint8_t> g((size_t)getVoxelNum());
    b200_check(b200_map_download_inflate(h_, g.data()), "download");
    return g;
  }
  // publishMapInflate (grid_map.cpp:851-900): the PointXYZ records (x, y, z, 1.0f) of the cloud the reference publishes
  std::vector<float> publishMapInflate(bool all_info = false, int local_map_margin = 1, double visualization_truncate_height = 2.4) const {
    int64_t n = 0;
    b200_check(b200_map_export_inflate_cloud(h_, all_info ? local_map_margin : 0, visualization_truncate_height, nullptr, 0, &n, 0), "publishMapInflate");
    std::vector<float> pts((size_t)n * 4);
    if (n) b200_check(b200_map_export_inflate_cloud(h_, all_info ? local_map_margin : 0, visualization_truncate_height, pts.data(), n, &n, 0), "publishMapInflate");
    return pts;
  }
  // publishMap (grid_map.cpp:806-849) / publishUnknown (:902-939): the PointXYZ records of the clouds the reference publishes
  std::vector<float> publishMap(double visualization_truncate_height = 2.4) const {
    return exportOccupancy(0, dprm_.local_map_margin / 2, visualization_truncate_height);
  }
  std::vector<float> publishUnknown(double visualization_truncate_height = 2.4) const {
    return exportOccupancy(1, 0, visualization_truncate_height);
  }
  // new: exact squared EDT + safety cost field (replaces the commented-out evaluateCoarseEDT hook)
  void buildDistanceField(double d_safe) { b200_check(b200_map_build_edt(h_, d_safe), "build_edt"); }
  // (re)build only when a map update has invalidated the field or d_safe changed
  void ensureDistanceField(double d_safe) {
    if (!b200_map_has_distance_field(h_, d_safe)) buildDistanceField(d_safe);
  }
  // GetPath.srv:20-23 — {mean, std_dev, min, max} of the distance to the nearest obstacle (m) over the path points,
  // DistancesData order of calculateDistancesMetrics (metrics.cpp:62-78)
  struct DistancesData { double mean, std_dev, min, max; };
  DistancesData pathDistanceMetrics(const std::vector<Vec3d>& path) const {
    std::vector<double> flat(path.size() * 3);
    for (size_t i = 0; i < path.size(); ++i) for (int k = 0; k < 3; ++k) flat[3 * i + k] = path[i](k);
    double o[4];
    b200_check(b200_map_path_distance_metrics(h_, flat.data(), (int64_t)path.size(), o, nullptr), "pathDistanceMetrics");
    return DistancesData{o[0], o[1], o[2], o[3]};
  }

  b200_map_t* handle() const { return h_; }
  const int* voxelNum() const { return n_; }

 private:
  std::vector<float> exportOccupancy(int which, int margin, double truncate) const {
    int64_t n = 0;
    b200_check(b200_map_export_occupancy_cloud(h_, which, margin, truncate, nullptr, 0, &n, 0), "publishMap");
    std::vector<float> pts((size_t)n * 4);
    if (n) b200_check(b200_map_export_occupancy_cloud(h_, which, margin, truncate, pts.data(), n, &n, 0), "publishMap");
    return pts;
  }
  void create() {
    if (h_) { b200_map_destroy(h_); h_ = nullptr; }
    res_inv_ = 1 / prm_.resolution;
    b200_check(b200_map_create(&prm_, &h_), "initMap");
    int32_t n[3];
    b200_map_dims(h_, n);
    for (int i = 0; i < 3; ++i) n_[i] = n[i];
  }
  b200_map_params_t prm_{};
  b200_depth_params_t dprm_{};
  b200_map_t* h_ = nullptr;
  int n_[3] = {0, 0, 0};
  double res_inv_ = 0;
  Vec3d camera_pos_;
  bool has_odom_ = false, has_cloud_ = false, has_depth_ = false;
};

}  // namespace b200host
