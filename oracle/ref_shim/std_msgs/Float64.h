// ref_shim -- TEST INFRASTRUCTURE, see ros/ros.h.
#pragma once
namespace std_msgs { struct Float64 { double data = 0; }; }
