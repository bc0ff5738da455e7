// shim_ros.h — hand-written stand-ins for the ROS / PCL / OpenCV / message_filters / Boost declarations the
// reference's grid_map.{h,cpp} and planner headers mention.  TEST INFRASTRUCTURE (oracle/_ref build only):
// just enough surface for the reference's UNMODIFIED sources to compile; only the pieces the cloud path
// executes (NodeHandle::param, PointCloud2 -> pcl::PointCloud<PointXYZ>) do real work.  Nothing is copied
// from those libraries.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace shim {
// what the last ros::Publisher::publish(sensor_msgs::PointCloud2) carried (the map publishers' output)
inline std::vector<uint8_t>& last_cloud_blob() { static std::vector<uint8_t> b; return b; }
inline int& fake_subscribers() { static int n = 0; return n; }
inline std::map<std::string, double>& num_params() { static std::map<std::string, double> m; return m; }
inline std::map<std::string, std::string>& str_params() { static std::map<std::string, std::string> m; return m; }
template <class T> struct ParamGet {
  static bool get(const std::string& k, T& out) {
    auto it = num_params().find(k);
    if (it == num_params().end()) return false;
    out = (T)it->second;
    return true;
  }
};
template <> struct ParamGet<std::string> {
  static bool get(const std::string& k, std::string& out) {
    auto it = str_params().find(k);
    if (it == str_params().end()) return false;
    out = it->second;
    return true;
  }
};
}  // namespace shim

#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_INFO_STREAM(x) ((void)0)
#define ROS_WARN_STREAM(x) ((void)0)
#define ROS_ERROR_STREAM(x) ((void)0)
#define ROS_DEBUG(...) ((void)0)

namespace ros {
struct Time { double t = 0; static Time now() { return Time(); } double toSec() const { return t; } };
struct Duration { double d; explicit Duration(double v = 0) : d(v) {} double toSec() const { return d; } };
inline Duration operator-(const Time& a, const Time& b) { return Duration(a.t - b.t); }
struct TimerEvent {};
struct Timer {};
struct Subscriber {};
struct Publisher {
  int getNumSubscribers() const { return shim::fake_subscribers(); }
  template <class M> void publish(const M& m) const { capture(m, 0); }
 private:
  template <class M> static auto capture(const M& m, int) -> decltype((void)m.data, (void)m.point_step) { shim::last_cloud_blob() = m.data; }
  template <class M> static void capture(const M&, long) {}
};
class NodeHandle {
 public:
  NodeHandle() {}
  NodeHandle(const char*) {}
  NodeHandle(const std::string&) {}
  template <class T> bool param(const std::string& key, T& out, const T& def) const {
    if (shim::ParamGet<T>::get(key, out)) return true;
    out = def;
    return false;
  }
  template <class M, class C> Subscriber subscribe(const std::string&, int, void (C::*)(const std::shared_ptr<const M>&), C*) { return Subscriber(); }
  template <class C> Timer createTimer(Duration, void (C::*)(const TimerEvent&), C*) { return Timer(); }
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace ros

namespace std_msgs { struct Header { std::string frame_id; ros::Time stamp; unsigned seq = 0; }; }
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseWithCovariance { Pose pose; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
typedef std::shared_ptr<const PoseStamped> PoseStampedConstPtr;
}
namespace nav_msgs {
struct Odometry { std_msgs::Header header; geometry_msgs::PoseWithCovariance pose; };
typedef std::shared_ptr<const Odometry> OdometryConstPtr;
}
namespace visualization_msgs { struct Marker { std_msgs::Header header; std::vector<geometry_msgs::Point> points; }; }
namespace sensor_msgs {
struct PointField { std::string name; uint32_t offset = 0; uint8_t datatype = 7; uint32_t count = 1; };
struct PointCloud2 {
  std_msgs::Header header;
  uint32_t height = 1, width = 0;
  std::vector<PointField> fields;
  bool is_bigendian = false;
  uint32_t point_step = 0, row_step = 0;
  std::vector<uint8_t> data;
  bool is_dense = true;
};
typedef std::shared_ptr<const PointCloud2> PointCloud2ConstPtr;
struct Image { std_msgs::Header header; uint32_t height = 0, width = 0; std::string encoding; std::vector<uint8_t> data; };
typedef std::shared_ptr<const Image> ImageConstPtr;
namespace image_encodings { static const std::string TYPE_32FC1 = "32FC1"; static const std::string TYPE_16UC1 = "16UC1"; }
}

#define CV_16UC1 2
namespace cv {
class Mat {
 public:
  int rows = 0, cols = 0;
  std::shared_ptr<std::vector<uint8_t>> buf;
  template <class T> T* ptr(int r) { return reinterpret_cast<T*>(buf->data()) + (size_t)r * cols; }
  template <class T> T& at(int r, int c) { return ptr<T>(r)[c]; }
  void convertTo(Mat& o, int, double = 1.0) const { o = *this; }
  void copyTo(Mat& o) const { o = *this; }
};
}
namespace cv_bridge {
struct CvImage { cv::Mat image; };
typedef std::shared_ptr<CvImage> CvImagePtr;
// what cv_bridge::toCvCopy does for a 16UC1 image: a deep copy of the payload, rows = height, cols = width
inline CvImagePtr toCvCopy(const sensor_msgs::ImageConstPtr& img, const std::string&) {
  auto p = std::make_shared<CvImage>();
  if (img) {
    p->image.rows = (int)img->height;
    p->image.cols = (int)img->width;
    p->image.buf = std::make_shared<std::vector<uint8_t>>(img->data);
  }
  return p;
}
}

namespace pcl {
struct PCLHeader { std::string frame_id; };
struct PointXYZ { float x = 0, y = 0, z = 0, pad = 1.0f; };  // PCL initialises data[3] to 1
template <class P> struct PointCloud {
  std::vector<P> points;
  uint32_t width = 0, height = 0;
  bool is_dense = true;
  PCLHeader header;
  void push_back(const P& p) { points.push_back(p); }
};
// what pcl::fromROSMsg does for PointXYZ: float32 x, y, z read at the offsets the message's fields give
inline void fromROSMsg(const sensor_msgs::PointCloud2& msg, PointCloud<PointXYZ>& cloud) {
  uint32_t ox = 0, oy = 4, oz = 8;
  for (const auto& f : msg.fields) { if (f.name == "x") ox = f.offset; else if (f.name == "y") oy = f.offset; else if (f.name == "z") oz = f.offset; }
  const size_t n = msg.point_step ? msg.data.size() / msg.point_step : 0;
  cloud.points.resize(n);
  for (size_t i = 0; i < n; ++i) {
    const uint8_t* p = msg.data.data() + i * msg.point_step;
    std::memcpy(&cloud.points[i].x, p + ox, 4);
    std::memcpy(&cloud.points[i].y, p + oy, 4);
    std::memcpy(&cloud.points[i].z, p + oz, 4);
  }
  cloud.width = (uint32_t)n; cloud.height = 1;
}
// what pcl::toROSMsg does for PointXYZ: the point array copied verbatim (16-byte records, padding float included)
template <class P> inline void toROSMsg(const PointCloud<P>& c, sensor_msgs::PointCloud2& msg) {
  msg.point_step = sizeof(P);
  msg.width = (uint32_t)c.points.size();
  msg.height = 1;
  msg.data.resize(c.points.size() * sizeof(P));
  if (!c.points.empty()) std::memcpy(msg.data.data(), c.points.data(), msg.data.size());
}
}

namespace message_filters {
template <class M> struct Subscriber { Subscriber(ros::NodeHandle&, const std::string&, int) {} };
namespace sync_policies {
template <class A, class B> struct ApproximateTime { explicit ApproximateTime(int) {} };
template <class A, class B> struct ExactTime { explicit ExactTime(int) {} };
}
template <class Policy> struct Synchronizer {
  template <class A, class B> Synchronizer(const Policy&, A&, B&) {}
  template <class F> void registerCallback(F) {}
};
}

namespace boost {
template <class... A> int bind(A&&...) { return 0; }
namespace multi_index {
template <class... A> struct indexed_by {};
template <class... A> struct hashed_unique {};
template <class... A> struct ordered_non_unique {};
template <class T> struct tag {};
template <class T> struct identity {};
template <class C, class T, T C::*P> struct member {};
}
template <class V, class I> struct multi_index_container { template <class Tag> struct index { typedef int type; }; };
}
struct shim_placeholder {};
static shim_placeholder _1, _2;
