// ref_shim -- TEST INFRASTRUCTURE, see ros/ros.h.  Field names of visualization_msgs/Marker that A1RobotControl.cpp:60-140 sets.
#pragma once
#include <ros/ros.h>
#include <string>
#include <vector>
namespace visualization_msgs {
struct Marker {
  enum { ARROW = 0, CUBE = 1, SPHERE = 2, CYLINDER = 3, LINE_STRIP = 4 };
  enum { ADD = 0, MODIFY = 0, DELETE = 2 };
  struct { std::string frame_id; } header;
  std::string ns;
  int id = 0, type = 0, action = 0;
  struct XYZ { double x = 0, y = 0, z = 0; };
  struct { XYZ position; struct { double x = 0, y = 0, z = 0, w = 1; } orientation; } pose;
  XYZ scale;
  struct RGBA { float r = 0, g = 0, b = 0, a = 0; };
  RGBA color;
  ros::Duration lifetime;
  std::vector<XYZ> points;
  std::vector<RGBA> colors;
};
}  // namespace visualization_msgs
