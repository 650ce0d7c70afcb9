// Stand-in for <ros/ros.h> (test infrastructure, our code): an in-process message bus.  See oracle/ref_shim/README.md.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace boost {   // the reference spells a few smart pointers through boost
template <typename T> using shared_ptr = std::shared_ptr<T>;
}

namespace ros {

struct Time {
  double sec_ = 0.0;
  Time() {}
  explicit Time(double s) : sec_(s) {}
  double toSec() const { return sec_; }
  Time& fromSec(double s) { sec_ = s; return *this; }
  static Time now() { return Time(0.0); }
  bool operator==(const Time& o) const { return sec_ == o.sec_; }
  bool operator!=(const Time& o) const { return sec_ != o.sec_; }
};
struct Duration { double sec_ = 0; explicit Duration(double s = 0) : sec_(s) {} void sleep() const {} };
struct Rate { explicit Rate(double) {} void sleep() {} };

namespace shim {
struct Stored { std::shared_ptr<const void> msg; };
struct Bus {
  std::map<std::string, std::vector<std::function<void(const std::shared_ptr<const void>&)>>> subscribers;
  std::deque<std::pair<std::string, std::shared_ptr<const void>>> pending;
  std::map<std::string, std::shared_ptr<const void>> last;       // last message published on a topic
  std::map<std::string, long> count;                             // messages published per topic
  std::map<std::string, double> params_d;
  std::map<std::string, int> params_i;
  std::map<std::string, std::string> params_s;
  int ok_budget = 0;                                             // ros::ok() returns true this many more times
  static Bus& get() { static Bus b; return b; }
};
}  // namespace shim

class Publisher {
 public:
  Publisher() {}
  explicit Publisher(const std::string& t) : topic_(t) {}
  template <typename M> void publish(const M& m) const {
    auto p = std::make_shared<const M>(m);
    shim::Bus& b = shim::Bus::get();
    b.last[topic_] = p;
    ++b.count[topic_];
    if (b.subscribers.count(topic_)) b.pending.emplace_back(topic_, p);
  }
  std::string getTopic() const { return topic_; }
 private:
  std::string topic_;
};
class Subscriber {};

class NodeHandle {
 public:
  NodeHandle() {}
  explicit NodeHandle(const std::string&) {}
  template <typename M> Publisher advertise(const std::string& topic, int) { return Publisher(topic); }
  template <typename M> Subscriber subscribe(const std::string& topic, int, void (*cb)(const std::shared_ptr<M const>&)) {
    auto& subs = shim::Bus::get().subscribers[topic];
    subs.clear();   // one subscriber per topic and process: a driver may enter the node's main() more than once
    subs.push_back([cb](const std::shared_ptr<const void>& p) { cb(std::static_pointer_cast<const M>(p)); });
    return Subscriber();
  }
  template <typename T> bool param(const std::string& name, T& var, const T& def) const { return lookup(name, var, def); }
 private:
  static std::string base(const std::string& n) { size_t k = n.find_last_of('/'); return k == std::string::npos ? n : n.substr(k + 1); }
  static bool lookup(const std::string& n, int& v, const int& d) { auto& m = shim::Bus::get().params_i; auto it = m.find(base(n)); v = it == m.end() ? d : it->second; return it != m.end(); }
  static bool lookup(const std::string& n, double& v, const double& d) { auto& m = shim::Bus::get().params_d; auto it = m.find(base(n)); v = it == m.end() ? d : it->second; return it != m.end(); }
  static bool lookup(const std::string& n, float& v, const float& d) { auto& m = shim::Bus::get().params_d; auto it = m.find(base(n)); v = it == m.end() ? d : (float)it->second; return it != m.end(); }
  static bool lookup(const std::string& n, std::string& v, const std::string& d) { auto& m = shim::Bus::get().params_s; auto it = m.find(base(n)); v = it == m.end() ? d : it->second; return it != m.end(); }
  static bool lookup(const std::string& n, bool& v, const bool& d) { auto& m = shim::Bus::get().params_i; auto it = m.find(base(n)); v = it == m.end() ? d : it->second != 0; return it != m.end(); }
};

inline void init(int&, char**, const std::string&) {}
inline bool ok() { shim::Bus& b = shim::Bus::get(); if (b.ok_budget > 0) { --b.ok_budget; return true; } return false; }
inline void spinOnce() {
  shim::Bus& b = shim::Bus::get();
  while (!b.pending.empty()) {
    auto item = b.pending.front();
    b.pending.pop_front();
    auto it = b.subscribers.find(item.first);
    if (it != b.subscribers.end()) for (auto& cb : it->second) cb(item.second);
  }
}
inline void spin() {}   // the harness delivers messages itself (spinOnce)

}  // namespace ros

namespace std_msgs {
struct Header { uint32_t seq = 0; ros::Time stamp; std::string frame_id; };
}

#define ROS_INFO(...) do { } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_ERROR(...) do { } while (0)
#define ROS_INFO_STREAM(x) do { } while (0)
#define ROS_BREAK() std::abort()
#define ROS_ASSERT(c) do { if (!(c)) std::abort(); } while (0)
