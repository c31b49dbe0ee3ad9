// oracle/shim/lvo_shim_bus.h — TEST INFRASTRUCTURE ONLY.
// A single-process stand-in for the ROS transport so that the reference's node sources (src/scanRegistration.cpp,
// src/laserOdometry.cpp, src/laserMapping.cpp) compile and run UNMODIFIED inside oracle/_ref (ROS itself is absent from this
// image).  One node per shared object: inbound messages are queued by the harness (refnode_push_*), delivered to the node's
// subscriber callbacks by ros::spin / ros::spinOnce on the node's own main thread, and everything the node publishes is kept
// per topic for the harness to collect (refnode_take_*).
#pragma once
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <chrono>

namespace lvo_shim {

struct Bus {
  std::mutex mu;
  std::condition_variable cv_in, cv_out;
  std::map<std::string, std::string> params;
  std::map<std::string, std::function<void(const std::shared_ptr<const void>&)>> subscribers;  // topic -> typed delivery
  std::deque<std::pair<std::string, std::shared_ptr<const void>>> inbound;
  std::map<std::string, std::deque<std::shared_ptr<const void>>> outbound;
  std::map<std::string, long> published;   // messages published per topic so far
  bool shutdown = false;
  const std::set<std::string>* keep = nullptr;   // topics whose messages are stored (others are only counted); null = all
  bool verbose = false;
};
inline Bus& bus() { static Bus* b = new Bus(); return *b; }   // leaked on purpose: node threads outlive static destruction

// deliver everything queued so far to the subscriber callbacks (caller = the node's spinning thread)
inline int dispatch_pending() {
  int n = 0;
  Bus& b = bus();
  for (;;) {
    std::pair<std::string, std::shared_ptr<const void>> m;
    std::function<void(const std::shared_ptr<const void>&)> fn;
    {
      std::lock_guard<std::mutex> lk(b.mu);
      if (b.inbound.empty()) break;
      m = b.inbound.front();
      b.inbound.pop_front();
      auto it = b.subscribers.find(m.first);
      if (it != b.subscribers.end()) fn = it->second;
    }
    if (fn) { fn(m.second); ++n; }
  }
  return n;
}
inline void wait_inbound(int timeout_ms) {
  Bus& b = bus();
  std::unique_lock<std::mutex> lk(b.mu);
  b.cv_in.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return !b.inbound.empty() || b.shutdown; });
}
inline void post(const std::string& topic, std::shared_ptr<const void> msg) {
  Bus& b = bus();
  { std::lock_guard<std::mutex> lk(b.mu); b.inbound.emplace_back(topic, std::move(msg)); }
  b.cv_in.notify_all();
}
inline void publish(const std::string& topic, std::shared_ptr<const void> msg) {
  Bus& b = bus();
  { std::lock_guard<std::mutex> lk(b.mu); if (!b.keep || b.keep->count(topic)) b.outbound[topic].push_back(std::move(msg)); b.published[topic]++; }
  b.cv_out.notify_all();
}
// the reference prints ~20 lines per frame; silenced unless LVO_REFNODE_VERBOSE is set
inline int quiet_printf(const char* fmt, ...) {
  if (!bus().verbose) return 0;
  va_list ap; va_start(ap, fmt); int r = vprintf(fmt, ap); va_end(ap); return r;
}

}  // namespace lvo_shim
