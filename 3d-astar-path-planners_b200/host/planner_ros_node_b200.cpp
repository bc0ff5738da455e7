// planner_ros_node_b200.cpp — the ROS-side glue a maintainer adds to run the reference's node
// (src/planner_ros_node.cpp, class HeuristicPlannerROS) on the B200 path.  It only compiles where ROS,
// Eigen and the package's generated srv headers exist (build with -DB200_WITH_ROS -DB200_WITH_EIGEN inside the
// catkin package) — none of which are present in this repository's image, so this file is
// NOT part of the tested build.  Everything it calls (b200host::GridMap, b200host::Planners::*) is.
//
// Differences from the reference node are confined to three places, marked  // B200:
//   1. parameters are copied from the NodeHandle into a ParamMap once (configureAlgorithm)
//   2. the cloud / odom / depth+pose subscribers live here instead of inside GridMap::initMap, and a depth frame is
//      fused when it arrives instead of from the reference's 20 Hz timer (grid_map.cpp:119, :635-666)
//   3. GetPath.time_spent is the real device time (the reference's is always 0)
//   4. the distance field is (re)built before a cost-aware request and whenever the reply's distance-to-obstacle
//      fields are filled (srv/GetPath.srv:20-23; the reference leaves them 0 because its EDT hook is commented out)
// Wire format (srv/GetPath.srv, srv/SetAlgorithm.srv), topic names and the request flow are unchanged.
#if defined(B200_WITH_ROS) && defined(B200_WITH_EIGEN)
#include <ros/ros.h>
#include <sensor_msgs/PointCloud2.h>
#include <nav_msgs/Odometry.h>
#include <nav_msgs/Path.h>
#include <sensor_msgs/Image.h>
#include <sensor_msgs/image_encodings.h>
#include <geometry_msgs/PoseStamped.h>
#include <cv_bridge/cv_bridge.h>
#include <message_filters/subscriber.h>
#include <message_filters/synchronizer.h>
#include <message_filters/sync_policies/approximate_time.h>
#include <heuristic_planners/GetPath.h>
#include <heuristic_planners/SetAlgorithm.h>

#include "b200_planners.hpp"

using namespace b200host;

class HeuristicPlannerROS {
 public:
  HeuristicPlannerROS() {
    std::string algorithm_name, heuristic;
    lnh_.param("algorithm", algorithm_name, std::string("astar"));
    lnh_.param("heuristic", heuristic, std::string("euclidean"));
    configureAlgorithm(algorithm_name, heuristic);
    request_path_server_ = lnh_.advertiseService("request_path", &HeuristicPlannerROS::requestPathService, this);
    change_planner_server_ = lnh_.advertiseService("set_algorithm", &HeuristicPlannerROS::setAlgorithm, this);
    global_path_pub_ = lnh_.advertise<nav_msgs::Path>("global_path", 1);
    // B200 (2): GridMap::initMap's own subscriptions (grid_map.cpp:116-117)
    cloud_sub_ = lnh_.subscribe<sensor_msgs::PointCloud2>("/grid_map/cloud", 10, &HeuristicPlannerROS::cloudCallback, this);
    odom_sub_ = lnh_.subscribe<nav_msgs::Odometry>("/grid_map/odom", 10, &HeuristicPlannerROS::odomCallback, this);
    // depth image + camera pose / body odometry, synchronised like grid_map.cpp:96-113 according to grid_map/pose_type
    // (the shipped launch file sets 2 = ODOMETRY)
    int pose_type = GridMap::POSE_STAMPED;
    lnh_.param("grid_map/pose_type", pose_type, 1);  // grid_map.cpp:46
    depth_sub_.reset(new message_filters::Subscriber<sensor_msgs::Image>(lnh_, "/grid_map/depth", 50));
    if (pose_type == GridMap::POSE_STAMPED) {
      pose_sub_.reset(new message_filters::Subscriber<geometry_msgs::PoseStamped>(lnh_, "/grid_map/pose", 25));
      sync_image_pose_.reset(new message_filters::Synchronizer<SyncPolicyImagePose>(SyncPolicyImagePose(100), *depth_sub_, *pose_sub_));
      sync_image_pose_->registerCallback(boost::bind(&HeuristicPlannerROS::depthPoseCallback, this, _1, _2));
    } else if (pose_type == GridMap::ODOMETRY) {
      sync_odom_sub_.reset(new message_filters::Subscriber<nav_msgs::Odometry>(lnh_, "/grid_map/odom", 100));
      sync_image_odom_.reset(new message_filters::Synchronizer<SyncPolicyImageOdom>(SyncPolicyImageOdom(100), *depth_sub_, *sync_odom_sub_));
      sync_image_odom_->registerCallback(boost::bind(&HeuristicPlannerROS::depthOdomCallback, this, _1, _2));
    }
  }

 private:
  void cloudCallback(const sensor_msgs::PointCloud2ConstPtr& msg) { grid_map_->cloudCallback(msg); }
  void odomCallback(const nav_msgs::OdometryConstPtr& msg) { grid_map_->odomCallback(msg); }
  // depthPoseCallback + updateOccupancyCallback of the reference in one step (b200host::GridMap::depthPoseCallback)
  void depthPoseCallback(const sensor_msgs::ImageConstPtr& img, const geometry_msgs::PoseStampedConstPtr& pose) {
    grid_map_->depthPoseCallback(img, pose);
  }

  void depthOdomCallback(const sensor_msgs::ImageConstPtr& img, const nav_msgs::OdometryConstPtr& odom) {
    grid_map_->depthOdomCallback(img, odom);
  }

  bool setAlgorithm(heuristic_planners::SetAlgorithmRequest& req, heuristic_planners::SetAlgorithmResponse& rep) {
    configureAlgorithm(req.algorithm.data, req.heuristic.data);
    rep.result.data = true;
    return true;
  }

  bool requestPathService(heuristic_planners::GetPathRequest& req, heuristic_planners::GetPathResponse& rep) {
    algorithm_->reset();
    if (!req.algorithm.data.empty()) configureAlgorithm(req.algorithm.data, req.heuristic.data.empty() ? heuristic_ : req.heuristic.data);
    Vec3d start(req.start.x, req.start.y, req.start.z), goal(req.goal.x, req.goal.y, req.goal.z);
    if (algorithm_->detectCollision(start)) { std::cout << start << ": Start not valid" << std::endl; return false; }
    if (algorithm_->detectCollision(goal)) { std::cout << goal << ": Goal not valid" << std::endl; return false; }
    Vec3d origin, size;
    grid_map_->getRegion(origin, size);
    auto within = [&](Vec3d& p) {  // literal, including the reference's size-vs-(origin+size) quirk
      return p.x() >= origin.x() && p.y() >= origin.y() && p.z() >= origin.z() && p.x() <= size.x() && p.y() <= size.y() && p.z() <= size.z();
    };
    if (!within(start) || !within(goal)) return false;
    int tries = req.tries.data == 0 ? 1 : req.tries.data;
    // B200 (4): the cost-aware planners read the cost field, the reply reads the distance field; both describe the
    // CURRENT grid, so refresh them if a cloud / depth frame arrived since the last build (0.3 ms at 512x512x128)
    grid_map_->ensureDistanceField(d_safe_);
    for (int i = 0; i < tries; ++i) {
      auto path_data = algorithm_->findPath(start, goal, Vec3d::Zero(), Vec3d::Zero(), Vec3d::Zero());
      if (!std::get<bool>(path_data["solved"])) { ROS_INFO("Could not calculate path between request points"); continue; }
      rep.time_spent.data = std::get<double>(path_data["time_spent"]) / 1000;  // B200 (3): real time, in ms
      rep.path_length.data = std::get<double>(path_data["path_length"]);
      rep.explored_nodes.data = std::get<size_t>(path_data["explored_nodes"]);
      rep.line_of_sight_checks.data = std::get<unsigned int>(path_data["line_of_sight_checks"]);
      rep.total_cost1.data = std::get<unsigned int>(path_data["total_cost1"]);
      rep.total_cost2.data = std::get<unsigned int>(path_data["total_cost2"]);
      rep.h_cost.data = std::get<unsigned int>(path_data["h_cost"]);
      rep.g_cost1.data = std::get<unsigned int>(path_data["g_cost1"]);
      rep.g_cost2.data = std::get<unsigned int>(path_data["g_cost2"]);
      rep.c_cost.data = std::get<unsigned int>(path_data["c_cost"]);
      rep.n_points.data = (int)algorithm_->getPath().size();
      const GridMap::DistancesData dd = grid_map_->pathDistanceMetrics(algorithm_->getPath());  // B200 (4)
      rep.min_distance_to_obstacle.data = dd.min;
      rep.max_distance_to_obstacle.data = dd.max;
      rep.mean_distance_to_obstacle.data = dd.mean;
      rep.mean_std_dev_to_obstacle.data = dd.std_dev;
      rep.cost_weight.data = cost_weight_;
      rep.max_los.data = max_los_;
      nav_msgs::Path global_path;
      global_path.header.frame_id = "world";
      global_path.header.stamp = ros::Time::now();
      for (const auto& it : algorithm_->getPath()) {
        geometry_msgs::PoseStamped pose;
        pose.header = global_path.header;
        pose.pose.position.x = it.x(); pose.pose.position.y = it.y(); pose.pose.position.z = it.z();
        pose.pose.orientation.w = 1.0;
        global_path.poses.push_back(pose);
      }
      algorithm_->reset();
      global_path_pub_.publish(global_path);
    }
    return true;
  }

  void configureAlgorithm(const std::string& algorithm_name, const std::string& heuristic) {
    heuristic_ = heuristic;
    ParamMap nh;  // B200 (1): snapshot of the private parameters the reference reads lazily
    auto snap = [&](const std::string& key) {  // a ROS parameter may be stored as double, int or bool
      double v; int iv; bool bv;
      if (lnh_.getParam(key, v)) nh.num[key] = v;
      else if (lnh_.getParam(key, iv)) nh.num[key] = iv;
      else if (lnh_.getParam(key, bv)) nh.num[key] = bv ? 1.0 : 0.0;
    };
    // every key GridMap::initMap reads (grid_map.cpp:12-50)
    for (const char* k : {"resolution", "map_size_x", "map_size_y", "map_size_z", "local_update_range_x", "local_update_range_y",
                          "local_update_range_z", "obstacles_inflation", "fx", "fy", "cx", "cy", "use_depth_filter",
                          "depth_filter_tolerance", "depth_filter_maxdist", "depth_filter_mindist", "depth_filter_margin",
                          "k_depth_scaling_factor", "skip_pixel", "p_hit", "p_miss", "p_min", "p_max", "p_occ", "min_ray_length",
                          "max_ray_length", "visualization_truncate_height", "virtual_ceil_height", "show_occ_time", "pose_type",
                          "local_map_margin", "ground_height"})
      snap(std::string("grid_map/") + k);
    for (const char* k : {"thetastaragr/barrier", "thetastaragr/flying_cost", "thetastaragr/flying_cost_default",
                          "thetastaragr/ground_judge", "thetastaragr/epsilon"})
      snap(k);
    for (const char* a : {"astar", "thetastar", "thetastaragr", "thetastaragrpp", "lazythetastar", "costastar", "costlazythetastar"})
      for (const char* k : {"resolution", "lambda_heuristic", "allocate_num", "max_line_of_sight_distance", "semantics"})
        snap(std::string(a) + "/" + k);
    if (algorithm_name == "thetastar") algorithm_.reset(new Planners::ThetaStar());
    else if (algorithm_name == "thetastaragr") algorithm_.reset(new Planners::ThetaStarAGR());
    else if (algorithm_name == "thetastaragrpp") algorithm_.reset(new Planners::ThetaStarAGRpp());
    else if (algorithm_name == "lazythetastar") algorithm_.reset(new Planners::LazyThetaStar());
    else if (algorithm_name == "costastar") algorithm_.reset(new Planners::CostAStar());
    else if (algorithm_name == "costlazythetastar") algorithm_.reset(new Planners::CostLazyThetaStar());
    else { if (algorithm_name != "astar") ROS_WARN("Wrong algorithm name parameter. Using A* by default"); algorithm_.reset(new Planners::AStar()); }
    double cw = 0, ml = 0;
    lnh_.param("cost_weight", cw, 0.0);
    lnh_.param("max_line_of_sight_distance", ml, 0.0);
    lnh_.param("d_safe", d_safe_, 1.0);  // B200 (4): safety distance of the cost field c = max(0, 1 - d / d_safe)
    cost_weight_ = (float)cw; max_los_ = (float)ml;
    algorithm_->setCostFactor((float)cw);
    algorithm_->setParam(nh);
    grid_map_.reset(new GridMap);  // like the reference, switching algorithms discards the map
    grid_map_->initMap(nh);
    algorithm_->setGridMap(grid_map_);
    algorithm_->init();
  }

  ros::NodeHandle lnh_{"~"};
  ros::ServiceServer request_path_server_, change_planner_server_;
  ros::Subscriber cloud_sub_, odom_sub_;
  typedef message_filters::sync_policies::ApproximateTime<sensor_msgs::Image, geometry_msgs::PoseStamped> SyncPolicyImagePose;  // grid_map.h:228
  std::shared_ptr<message_filters::Subscriber<sensor_msgs::Image>> depth_sub_;
  std::shared_ptr<message_filters::Subscriber<geometry_msgs::PoseStamped>> pose_sub_;
  std::shared_ptr<message_filters::Synchronizer<SyncPolicyImagePose>> sync_image_pose_;
  typedef message_filters::sync_policies::ApproximateTime<sensor_msgs::Image, nav_msgs::Odometry> SyncPolicyImageOdom;  // grid_map.h:226
  std::shared_ptr<message_filters::Subscriber<nav_msgs::Odometry>> sync_odom_sub_;
  std::shared_ptr<message_filters::Synchronizer<SyncPolicyImageOdom>> sync_image_odom_;
  double d_safe_ = 1.0;
  float cost_weight_ = 0, max_los_ = 0;
  ros::Publisher global_path_pub_;
  std::unique_ptr<Planners::AlgorithmBase> algorithm_;
  GridMap::Ptr grid_map_;
  std::string heuristic_;
};

int main(int argc, char** argv) {
  ros::init(argc, argv, "heuristic_planner_ros_node");
  HeuristicPlannerROS node;
  ros::spin();  // single callback thread, as in the reference (planner_ros_node.cpp:417)
  return 0;
}
#endif
