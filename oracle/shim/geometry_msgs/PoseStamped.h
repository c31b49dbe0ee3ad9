// shim (test infrastructure): geometry_msgs/Pose, PoseStamped
#pragma once
#include <std_msgs/Header.h>
namespace geometry_msgs {
struct Point { double x = 0, y = 0, z = 0; };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; };
struct Pose { Point position; Quaternion orientation; };
struct PoseStamped { std_msgs::Header header; Pose pose; };
struct PoseWithCovariance { Pose pose; double covariance[36] = {0}; };
}
