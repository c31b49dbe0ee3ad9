// shim (test infrastructure): sensor_msgs/Imu (only queued by laserOdometry.cpp:139-144, never read)
#pragma once
#include <std_msgs/Header.h>
#include <memory>
namespace sensor_msgs {
struct Imu { std_msgs::Header header; typedef std::shared_ptr<const Imu> ConstPtr; };
typedef std::shared_ptr<const Imu> ImuConstPtr;
}
