// type-check stub of <ros/ros.h> (see ros/stubs/README.md)
#pragma once
#include <cstdint>
#include <string>
#include <memory>
namespace ros {
struct Time { uint32_t sec = 0, nsec = 0; static Time now() { return Time(); } double toSec() const { return sec + 1e-9 * nsec; } };
struct Duration { double d = 0; Duration() {} explicit Duration(double s) : d(s) {} };
class Publisher { public: template <class M> void publish(const M&) const {} };
class NodeHandle {
public:
    template <class M> Publisher advertise(const std::string&, uint32_t, bool = false) { return Publisher(); }
    template <class T> bool getParam(const std::string&, T&) const { return false; }
    template <class T> void param(const std::string&, T& v, const T& d) const { v = d; }
};
inline void init(int&, char**, const std::string&) {}
inline void spin() {}
inline bool ok() { return true; }
}  // namespace ros
#define ROS_INFO(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_WARN(...) ((void)0)
