// ref_shim "ros/ros.h" -- TEST INFRASTRUCTURE.  NOT ROS.  Inert stand-ins for the few ROS names that the reference's control
// sources mention (A1CtrlStates.h:9,135-...; A1RobotControl.{h,cpp} debug publishers), so that they compile unmodified into
// oracle/_ref/.  Parameters always take their defaults, publishers drop what they are given.
#pragma once
#include <math.h>   // the real header chain brings the C names (isnan, ...) into the global namespace; A1RobotControl.cpp:315,559 rely on it
#include <string>
namespace ros {
struct Duration { Duration() {} explicit Duration(double) {} };
struct Time { static Time now() { return Time(); } };
class Publisher {
 public:
  template <class M> void publish(const M&) const {}
};
class NodeHandle {
 public:
  template <class T, class U> bool param(const std::string&, T& var, const U& def) const { var = (T)def; return false; }
  template <class T> bool param(const std::string&, T&) const { return false; }
  template <class T> bool getParam(const std::string&, T&) const { return false; }
  template <class M> Publisher advertise(const std::string&, int) { return Publisher(); }
};
}  // namespace ros
