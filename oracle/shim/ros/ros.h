// oracle/shim/ros/ros.h — TEST INFRASTRUCTURE ONLY: the subset of roscpp that the reference's three lidar nodes use
// (ros::init / NodeHandle::param, subscribe, advertise / Publisher::publish / spin, spinOnce, ok / Rate / Time), backed by
// lvo_shim_bus.h.  Parameters come from the harness (refnode_start), as the launch files would set them
// (launch/aloam_velodyne_HDL_64.launch:3-13).
#pragma once
#include "../lvo_shim_bus.h"
#include <sstream>
#include <thread>
#include <vector>

#define ROS_WARN(...) do { if (lvo_shim::bus().verbose) { fprintf(stderr, "[ WARN] " __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#define ROS_INFO(...) do { if (lvo_shim::bus().verbose) { fprintf(stderr, "[ INFO] " __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)
#define ROS_ERROR(...) do { fprintf(stderr, "[ERROR] " __VA_ARGS__); fprintf(stderr, "\n"); } while (0)
#define ROS_BREAK() do { fprintf(stderr, "ROS_BREAK at %s:%d\n", __FILE__, __LINE__); abort(); } while (0)

namespace ros {

class Time {
 public:
  Time() : t_(0) {}
  Time& fromSec(double t) { t_ = t; return *this; }
  double toSec() const { return t_; }
  bool operator==(const Time& o) const { return t_ == o.t_; }
  bool operator<(const Time& o) const { return t_ < o.t_; }
 private:
  double t_;
};

struct TransportHints { TransportHints& tcpNoDelay(bool = true) { return *this; } };
class Subscriber {};

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& topic) : topic_(topic) {}
  template <class M> void publish(const M& msg) const { lvo_shim::publish(topic_, std::make_shared<const M>(msg)); }
  const std::string& getTopic() const { return topic_; }
 private:
  std::string topic_;
};

inline void init(int&, char**, const std::string&) {}

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <class T, class D> bool param(const std::string& name, T& var, const D& def) const {
    lvo_shim::Bus& b = lvo_shim::bus();
    std::lock_guard<std::mutex> lk(b.mu);
    auto it = b.params.find(name);
    if (it == b.params.end()) { var = static_cast<T>(def); return false; }   // NodeHandle::param<T>: the default is converted to T
    std::istringstream is(it->second);
    T v; is >> v; var = v;
    return true;
  }
  template <class T> bool getParam(const std::string& name, T& var) const { T d = var; return param<T>(name, var, d); }
  template <class M> Subscriber subscribe(const std::string& topic, unsigned, void (*cb)(const std::shared_ptr<M const>&), const TransportHints& = TransportHints()) {
    lvo_shim::Bus& b = lvo_shim::bus();
    std::lock_guard<std::mutex> lk(b.mu);
    b.subscribers[topic] = [cb](const std::shared_ptr<const void>& m) { cb(std::static_pointer_cast<const M>(m)); };
    return Subscriber();
  }
  template <class M> Publisher advertise(const std::string& topic, unsigned, bool = false) { return Publisher(topic); }
};

inline bool ok() { lvo_shim::Bus& b = lvo_shim::bus(); std::lock_guard<std::mutex> lk(b.mu); return !b.shutdown; }
inline void spinOnce() { lvo_shim::dispatch_pending(); }
inline void spin() {
  while (ok()) { lvo_shim::dispatch_pending(); lvo_shim::wait_inbound(50); }
  // a node whose main() returns would destroy joinable std::threads (laserMapping.cpp:934): park here for good
  for (;;) std::this_thread::sleep_for(std::chrono::hours(1));
}
class Rate {
 public:
  explicit Rate(double hz) : ms_(hz > 0 ? (int)(1000.0 / hz) : 10) {}
  bool sleep() { lvo_shim::wait_inbound(ms_); return true; }   // wakes as soon as a message arrives
 private:
  int ms_;
};
namespace package { inline std::string getPath(const std::string&) { return "."; } }

}  // namespace ros
