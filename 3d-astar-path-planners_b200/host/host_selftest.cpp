// host_selftest.cpp — exercises the C++ adapter classes (GridMap / Planners::*) the way the reference's ROS
// node drives its own classes (planner_ros_node.cpp:135-214, :261-296) and prints one JSON line per query so
// tests/test_gpu_host_adapter.py can compare with the ctypes path.  No ROS, no Eigen.
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <thread>

#include "b200_planners.hpp"

using namespace b200host;

int main(int argc, char** argv) {
  const char* algo = argc > 1 ? argv[1] : "thetastar";
  try {
    ParamMap nh;
    nh.num["grid_map/resolution"] = 0.1;
    nh.num["grid_map/map_size_x"] = 12.8; nh.num["grid_map/map_size_y"] = 12.8; nh.num["grid_map/map_size_z"] = 3.2;
    nh.num["grid_map/local_update_range_x"] = 100; nh.num["grid_map/local_update_range_y"] = 100; nh.num["grid_map/local_update_range_z"] = 100;
    for (const char* a : {"astar", "thetastar", "thetastaragr", "lazythetastar", "costastar", "costlazythetastar"}) {
      nh.num[std::string(a) + "/resolution"] = 0.1;
      nh.num[std::string(a) + "/lambda_heuristic"] = 5.0;
      nh.num[std::string(a) + "/allocate_num"] = 100000;
    }
    nh.num["thetastaragr/barrier"] = 2.0; nh.num["thetastaragr/flying_cost"] = 10.0; nh.num["thetastaragr/flying_cost_default"] = 0.0;
    nh.num["thetastaragr/ground_judge"] = 0.1; nh.num["thetastaragr/epsilon"] = 0.1;   // launch/planner.launch:71-77
    // configureAlgorithm (planner_ros_node.cpp:261-296)
    std::unique_ptr<Planners::AlgorithmBase> algorithm;
    const std::string name(algo);
    if (name == "astar") algorithm.reset(new Planners::AStar());
    else if (name == "thetastar") algorithm.reset(new Planners::ThetaStar());
    else if (name == "thetastaragr") algorithm.reset(new Planners::ThetaStarAGR());
    else if (name == "thetastaragrpp") algorithm.reset(new Planners::ThetaStarAGRpp());
    else if (name == "lazythetastar") algorithm.reset(new Planners::LazyThetaStar());
    else if (name == "costastar") algorithm.reset(new Planners::CostAStar());
    else if (name == "costlazythetastar") algorithm.reset(new Planners::CostLazyThetaStar());
    else if (name == "depth") {
      // depth-image path: a wall 2 m in front of a camera at (0,0,1) looking along +x, seen in 6 frames
      GridMap::Ptr gm(new GridMap);
      gm->initMap(nh);
      const int rows = 120, cols = 160;
      b200_depth_params_t dp = gm->depthParams();
      dp.fx = dp.fy = 96.8; dp.cx = 80.3; dp.cy = 60.9;
      gm->setDepthParams(dp);
      std::vector<uint16_t> img((size_t)rows * cols, 2000);
      // optical frame (x right, y down, z forward) looking along world +x: q = (0.5, -0.5, 0.5, -0.5)
      int fused = 0;
      for (int f = 0; f < 6; ++f) fused += gm->depthPoseCallback(img.data(), rows, cols, Vec3d(0.0, 0.0, 1.0), 0.5, -0.5, 0.5, -0.5) ? 1 : 0;
      int64_t st[8];
      b200_map_depth_stats(gm->handle(), st);
      std::printf("{\"mode\": \"depth\", \"fused\": %d, \"wall_occupancy\": %d, \"wall_inflated\": %d, \"free_occupancy\": %d, "
                  "\"outside\": %d, \"projected\": %lld, \"rays_cast\": %lld, \"ray_steps\": %lld, \"updates\": %lld, \"has_depth\": %d}\n",
                  fused, gm->getOccupancy(Vec3d(2.02, 0.0, 1.0)), gm->getInflateOccupancy(Vec3d(2.02, 0.0, 1.0)),
                  gm->getOccupancy(Vec3d(1.0, 0.0, 1.0)), gm->getOccupancy(Vec3d(100.0, 0.0, 1.0)), (long long)st[0], (long long)st[1],
                  (long long)st[2], (long long)st[3], gm->hasDepthObservation() ? 1 : 0);
      return 0;
    }
    else if (name == "pool") {
      // three LazyThetaStar planners on ONE map, each on a stream of its own and driven by its own thread, two batches
      // each: every batch must equal the batch a single planner computes on the map's stream
      GridMap::Ptr gm(new GridMap);
      gm->initMap(nh);
      std::vector<float> cloud;
      for (int j = 40; j < 88; ++j)
        for (int k = 0; k < 32; ++k) { cloud.push_back(0.05f); cloud.push_back(-6.4f + (j + 0.5f) * 0.1f); cloud.push_back(-0.01f + (k + 0.5f) * 0.1f); cloud.push_back(0.f); }
      gm->odomCallback(Vec3d(0, 0, 1));
      gm->cloudCallback(cloud.data(), (int64_t)cloud.size() / 4, 16);
      std::vector<Vec3d> src, dst;
      for (int i = 0; i < 400; ++i) {
        src.push_back(Vec3d(-5.0 + 0.01 * i, -4.0 + 0.02 * i, 0.5 + 0.004 * i));
        dst.push_back(Vec3d(5.0 - 0.015 * i, 4.0 - 0.018 * i, 2.5 - 0.004 * i));
      }
      auto make = [&]() {
        std::unique_ptr<Planners::LazyThetaStar> p(new Planners::LazyThetaStar());
        p->setParam(nh);
        p->setGridMap(gm);
        p->init();
        return p;
      };
      auto base = make();
      std::vector<b200_path_result_t> want;
      std::vector<double> want_path;
      base->findPaths(src, dst, want, want_path, 400 * 256);
      b200_map_sync(gm->handle());
      const int lanes = 3;
      std::vector<std::unique_ptr<Planners::LazyThetaStar>> pl;
      for (int l = 0; l < lanes; ++l) { pl.push_back(make()); pl.back()->useOwnStream(0); }
      std::vector<int> bad(lanes, 0);
      std::vector<std::thread> th;
      for (int l = 0; l < lanes; ++l)
        th.emplace_back([&, l]() {
          for (int rep = 0; rep < 2; ++rep) {
            std::vector<b200_path_result_t> got;
            std::vector<double> got_path;
            pl[l]->findPaths(src, dst, got, got_path, 400 * 256);
            for (size_t i = 0; i < got.size(); ++i)
              if (got[i].solved != want[i].solved || got[i].n_path != want[i].n_path || got[i].g_final != want[i].g_final ||
                  got[i].path_length != want[i].path_length || got[i].explored_nodes != want[i].explored_nodes)
                ++bad[l];
          }
        });
      for (auto& t : th) t.join();
      int solved = 0, mism = 0;
      for (const auto& r : want) solved += r.solved;
      for (int b : bad) mism += b;
      std::printf("{\"mode\": \"pool\", \"lanes\": %d, \"queries\": %zu, \"solved\": %d, \"mismatches\": %d}\n", lanes, want.size(), solved, mism);
      return mism ? 4 : 0;
    }
    else algorithm.reset(new Planners::AStar());
    algorithm->setCostFactor(2.0f);
    algorithm->setParam(nh);
    GridMap::Ptr grid_map(new GridMap);
    grid_map->initMap(nh);
    algorithm->setGridMap(grid_map);
    algorithm->init();

    // a deterministic free-standing wall (PointXYZ stride 16): y in [-2.4, 2.4), full height; paths go round its end.
    // (A wall that seals the map except for one gap makes the reference's own A* take 5.8e7 iterations and
    // 1.9e7 open-set entries — DESIGN.md "Reference pathologies" — so it is not a useful smoke scenario.)
    std::vector<float> cloud;
    for (int j = 40; j < 88; ++j)
      for (int k = 0; k < 32; ++k) {
        cloud.push_back(0.05f); cloud.push_back(-6.4f + (j + 0.5f) * 0.1f); cloud.push_back(-0.01f + (k + 0.5f) * 0.1f); cloud.push_back(0.f);
      }
    grid_map->odomCallback(Vec3d(0, 0, 1));
    grid_map->cloudCallback(cloud.data(), (int64_t)cloud.size() / 4, 16);
    if (name.rfind("cost", 0) == 0) grid_map->buildDistanceField(1.0);

    // requestPathService (planner_ros_node.cpp:135-214)
    algorithm->reset();
    Vec3d start(-3.0, -1.0, 1.0), goal(3.0, 1.0, 1.5), blocked(0.05, 0.0, 1.0), outside(100, 0, 0);
    if (algorithm->detectCollision(start) || algorithm->detectCollision(goal)) { std::printf("{\"error\": \"endpoints\"}\n"); return 1; }
    const bool c1 = algorithm->detectCollision(blocked), c2 = algorithm->detectCollision(outside);
    auto path_data = algorithm->findPath(start, goal, Vec3d(), Vec3d(), Vec3d());
    const bool solved = std::get<bool>(path_data["solved"]);
    const auto path = algorithm->getPath();
    const Vec3d gc = std::get<Vec3d>(path_data["goal_coords"]);
    // GetPath.srv:20-23 from the distance field; then setOccupied must invalidate that field
    grid_map->ensureDistanceField(1.0);
    const GridMap::DistancesData dd = grid_map->pathDistanceMetrics(path);
    const int before = grid_map->getInflateOccupancy(Vec3d(3.0, 3.0, 1.0));
    grid_map->setOccupied(Vec3d(3.0, 3.0, 1.0));
    grid_map->setOccupied(outside);
    const int after = grid_map->getInflateOccupancy(Vec3d(3.0, 3.0, 1.0));
    const int stale = b200_map_has_distance_field(grid_map->handle(), 1.0) ? 0 : 1;
    std::printf("{\"algorithm\": \"%s\", \"solved\": %d, \"n_path\": %zu, \"g_final_node\": %.17g, \"path_length\": %.17g, "
                "\"dist_mean\": %.17g, \"dist_std\": %.17g, \"dist_min\": %.17g, \"dist_max\": %.17g, \"set_occupied\": [%d, %d], \"edt_stale\": %d, "
                "\"explored_nodes\": %zu, \"line_of_sight_checks\": %u, \"goal_coords\": [%.17g, %.17g, %.17g], "
                "\"collide_wall\": %d, \"collide_outside\": %d, \"total_cost1\": %u, \"time_spent_us\": %.3f}\n",
                std::get<std::string>(path_data["algorithm"]).c_str(), solved ? 1 : 0, path.size(), std::get<double>(path_data["g_final_node"]),
                std::get<double>(path_data["path_length"]), dd.mean, dd.std_dev, dd.min, dd.max, before, after, stale, std::get<size_t>(path_data["explored_nodes"]),
                std::get<unsigned int>(path_data["line_of_sight_checks"]), gc(0), gc(1), gc(2), c1 ? 1 : 0, c2 ? 1 : 0,
                std::get<unsigned int>(path_data["total_cost1"]), std::get<double>(path_data["time_spent"]));
    algorithm->reset();
    return solved ? 0 : 2;
  } catch (const std::exception& e) {
    std::printf("{\"error\": \"%s\"}\n", e.what());
    return 3;
  }
}
