// shim (test infrastructure): std_msgs/Header
#pragma once
#include <ros/ros.h>
namespace std_msgs { struct Header { unsigned seq = 0; ros::Time stamp; std::string frame_id; }; }
