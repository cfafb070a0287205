// ref_shim -- TEST INFRASTRUCTURE, see ros/ros.h.
#pragma once
namespace geometry_msgs { struct PointStamped { struct { double x = 0, y = 0, z = 0; } point; }; }
