// b200_planners.hpp — host-side mirror of the reference's planner class family on top of the C-ABI:
//   Planners::AlgorithmBase (include/Planners/AlgorithmBase.hpp:28-53), AStar (AStar.hpp:14-37),
//   ThetaStar (ThetaStar.hpp), and the north-star extensions LazyThetaStar / CostAStar / CostLazyThetaStar.
// Same method names, same call order as HeuristicPlannerROS uses them (planner_ros_node.cpp:136-214):
//   reset() -> detectCollision() x2 -> findPath() -> std::get<T>(path_data[key]) -> getPath() -> reset()
// and the same PathData keys / variant alternatives as createResultDataObject (AlgorithmBase.cpp:123-175).
#pragma once
#include <functional>
#include <iostream>
#include <map>
#include <string>
#include <variant>
#include <vector>

#include "b200_grid_map.hpp"

namespace b200host {
namespace Planners {

// include/utils/utils.hpp:110-111
using DataVariant = std::variant<std::string, Vec3d, std::vector<Vec3d>, double, size_t, int, bool, unsigned int>;
using PathData = std::map<std::string, DataVariant>;
using HeuristicFunction = std::function<unsigned int(Vec3d, Vec3d)>;

class AlgorithmBase {
 public:
  explicit AlgorithmBase(const std::string& name) : algorithm_name_(name) {}
  virtual ~AlgorithmBase() {
    if (h_) b200_planner_destroy(h_);
    if (own_stream_) b200_stream_destroy(own_stream_);
  }
  enum { REACH_HORIZON = 1, REACH_END = 2, NO_PATH = 3, NEAR_END = 4 };

  void setGridMap(GridMap::Ptr& grid_map) {
    grid_map_ = grid_map;
    if (h_) b200_check(b200_planner_set_map(h_, grid_map_->handle()), "setGridMap");
  }
  // <algo>/resolution, time_resolution, lambda_heuristic, allocate_num (AStar.cpp:31-41, ThetaStar.cpp:25-35)
  virtual void setParam(const ParamMap& nh) {
    const std::string ns = (algorithm_name_ == "thetastaragrpp" ? std::string("thetastaragr") : algorithm_name_) + "/";
    nh.param(ns + "resolution", prm_.resolution, -1.0);
    nh.param(ns + "lambda_heuristic", prm_.lambda_heuristic, -1.0);
    nh.param(ns + "allocate_num", prm_.allocate_num, -1);
    nh.param(ns + "max_line_of_sight_distance", prm_.max_los, 0.0);
    // both air-ground variants read the thetastaragr/ namespace (ThetaStarAGR.cpp:33-36, ThetaStarAGRpp.cpp:28-37)
    nh.param(std::string("thetastaragr/barrier"), prm_.agr_barrier, 0.0);
    nh.param(std::string("thetastaragr/flying_cost"), prm_.agr_flying_cost, 0.0);
    nh.param(std::string("thetastaragr/flying_cost_default"), prm_.agr_flying_cost_default, 0.0);
    nh.param(std::string("thetastaragr/ground_judge"), prm_.agr_ground_judge, 0.0);
    nh.param(std::string("thetastaragr/epsilon"), prm_.agr_epsilon, 0.0);
    double sem = 0;
    nh.param(ns + "semantics", sem, 0.0);
    prm_.semantics = (int32_t)sem;
    prm_.cost_weight = cost_weight_;
  }
  // allocates the device-side planner (the reference allocates its node pool here, AStar.cpp:14-29)
  virtual void init() {
    if (h_) { b200_planner_destroy(h_); h_ = nullptr; }
    b200_check(b200_planner_create(algorithm_name_.c_str(), &prm_, &h_), "init");
    if (grid_map_) b200_check(b200_planner_set_map(h_, grid_map_->handle()), "init/set_map");
    if (own_stream_) b200_check(b200_planner_set_stream(h_, own_stream_), "init/set_stream");
  }
  // new: give this planner a CUDA stream of its own (after init()), so that planners bound to the same GridMap and driven
  // from different threads overlap their batches on the device (b200_planner_set_stream); the map must be up to date first
  void useOwnStream(int device = 0) {
    if (!own_stream_) b200_check(b200_stream_create(device, &own_stream_), "useOwnStream");
    b200_check(b200_planner_set_stream(h_, own_stream_), "useOwnStream");
  }
  // AlgorithmBase.cpp:17-32 — per-query state lives on the device and is reset by every findPath
  void reset() { path_.clear(); }
  // stored and never used, exactly like the reference's Node0 planners (SURVEY §2 "utils / Heuristic")
  void setHeuristic(HeuristicFunction h) { heuristic_ = h; }
  virtual void setCostFactor(const float& f) {
    cost_weight_ = f;
    prm_.cost_weight = f;
    if (h_) b200_check(b200_planner_set_cost_factor(h_, f), "setCostFactor");
  }
  // AlgorithmBase.cpp:38-44: bool(getInflateOccupancy) — outside the map (-1) counts as a collision
  bool detectCollision(Vec3d& c) {
    const double p[3] = {c(0), c(1), c(2)};
    int32_t hit = 0;
    b200_check(b200_planner_detect_collision(h_, p, &hit), "detectCollision");
    return hit != 0;
  }

  virtual PathData findPath(Vec3d source, Vec3d target) {
    const double s[3] = {source(0), source(1), source(2)}, g[3] = {target(0), target(1), target(2)};
    b200_path_result_t r;
    std::vector<double> buf(3 * 4096);
    b200_check(b200_planner_find_path(h_, s, g, &r, buf.data(), 4096), "findPath");
    if (r.n_path > 4096) {
      buf.resize(3 * (size_t)r.n_path);
      b200_check(b200_planner_find_path(h_, s, g, &r, buf.data(), r.n_path), "findPath");
    }
    path_.clear();
    for (int i = 0; i < r.n_path; ++i) path_.push_back(Vec3d(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]));
    last_ = r;
    if (!r.solved) {  // the reference's stdout diagnostics (AStar.cpp:165-178, AlgorithmBase.cpp:136)
      std::cout << (r.status == 2 ? "run out of memory." : r.status == 4 ? "open-list workspace exhausted (faithful semantics)." : "open set empty, no path!") << std::endl;
      std::cout << "Error impossible to calculate a solution" << std::endl;
    }
    return toPathData(r, path_);
  }
  virtual PathData findPath(Vec3d source, Vec3d target, Vec3d, Vec3d, Vec3d) { return findPath(source, target); }
  std::vector<Vec3d> getPath() { return path_; }
  // AlgorithmBase.cpp:185-189: the reference hands out Node0 pointers; positions are what its callers could read from them
  std::vector<Vec3d> getVisitedNodes() {
    int64_t n = 0;
    b200_check(b200_planner_get_visited_nodes(h_, nullptr, 0, &n), "getVisitedNodes");
    std::vector<double> flat((size_t)n * 3);
    if (n) b200_check(b200_planner_get_visited_nodes(h_, flat.data(), n, &n), "getVisitedNodes");
    std::vector<Vec3d> out((size_t)n);
    for (int64_t i = 0; i < n; ++i) out[(size_t)i] = Vec3d(flat[3 * i], flat[3 * i + 1], flat[3 * i + 2]);
    return out;
  }

  // batched form (new): nq independent reset()+findPath() pairs, one thread block each
  void findPaths(const std::vector<Vec3d>& sources, const std::vector<Vec3d>& targets, std::vector<b200_path_result_t>& results,
                 std::vector<double>& path_xyz, int64_t path_cap) {
    const size_t n = sources.size();
    std::vector<double> s(3 * n), g(3 * n);
    for (size_t i = 0; i < n; ++i) for (int k = 0; k < 3; ++k) { s[3 * i + k] = sources[i](k); g[3 * i + k] = targets[i](k); }
    results.resize(n);
    path_xyz.resize(3 * (size_t)path_cap);
    b200_check(b200_planner_find_paths(h_, s.data(), g.data(), (int64_t)n, results.data(), path_xyz.data(), path_cap, 0), "findPaths");
  }
  const b200_path_result_t& lastResult() const { return last_; }
  b200_planner_t* handle() const { return h_; }

 protected:
  // AlgorithmBase.cpp:123-175 — same keys, same variant alternative per key
  PathData toPathData(const b200_path_result_t& r, const std::vector<Vec3d>& path) const {
    PathData d;
    d["solved"] = (bool)(r.solved != 0);
    d["goal_coords"] = Vec3d(r.goal_coords[0], r.goal_coords[1], r.goal_coords[2]);
    d["g_final_node"] = (double)r.g_final;
    d["algorithm"] = algorithm_name_;
    d["path"] = path;
    d["time_spent"] = (double)r.time_us;  // real device time; the reference always reports 0
    d["explored_nodes"] = (size_t)r.explored_nodes;
    d["start_coords"] = Vec3d(r.start_coords[0], r.start_coords[1], r.start_coords[2]);
    d["path_length"] = (double)r.path_length;
    for (const char* k : {"total_cost1", "total_cost2", "h_cost", "g_cost1", "g_cost2", "c_cost", "grid_cost1", "grid_cost2"})
      d[k] = (unsigned int)0;
    d["line_of_sight_checks"] = (unsigned int)r.los_checks;
    return d;
  }
  HeuristicFunction heuristic_;
  GridMap::Ptr grid_map_;
  double cost_weight_{0};
  const std::string algorithm_name_;
  b200_planner_params_t prm_{};
  b200_planner_t* h_ = nullptr;
  void* own_stream_ = nullptr;
  std::vector<Vec3d> path_;
  b200_path_result_t last_{};
};

class AStar : public AlgorithmBase {
 public:
  AStar() : AlgorithmBase("astar") {}
  explicit AStar(std::string name) : AlgorithmBase(name) {}
};
class ThetaStar : public AStar {
 public:
  ThetaStar() : AStar("thetastar") {}
  explicit ThetaStar(std::string name) : AStar(name) {}
};
class ThetaStarAGR : public ThetaStar { public: ThetaStarAGR() : ThetaStar("thetastaragr") {} };
class ThetaStarAGRpp : public ThetaStar { public: ThetaStarAGRpp() : ThetaStar("thetastaragrpp") {} };
class LazyThetaStar : public ThetaStar { public: LazyThetaStar() : ThetaStar("lazythetastar") {} };
class CostAStar : public AStar { public: CostAStar() : AStar("costastar") {} };
class CostLazyThetaStar : public ThetaStar { public: CostLazyThetaStar() : ThetaStar("costlazythetastar") {} };

}  // namespace Planners
}  // namespace b200host
