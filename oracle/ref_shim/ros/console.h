// ref_shim -- TEST INFRASTRUCTURE, see ros/ros.h.
#pragma once
#define ROS_INFO(...) ((void)0)
#define ROS_WARN(...) ((void)0)
#define ROS_ERROR(...) ((void)0)
#define ROS_DEBUG(...) ((void)0)
