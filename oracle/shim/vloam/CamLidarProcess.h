// shim (test infrastructure): laserOdometry.cpp:62,243,262-265,310 constructs the vloam camera-lidar process and starts it.  That
// subsystem (visual odometry; OpenCV, GTSAM, Sophus) is OUT OF SCOPE and has no effect on the lidar odometry state,
// so it is an inert stand-in here; found before the reference's own include/vloam/CamLidarProcess.h through -Ishim.
#pragma once
#include <ros/ros.h>
#include <array>
#include <memory>
struct Transformd { double pos[3] = {0, 0, 0}; double rot[4] = {0, 0, 0, 1}; };   // include/vloam/Twist.h:22 (only .pos is touched, :257-260)
class CamLidarProcess {
 public:
  CamLidarProcess(std::shared_ptr<ros::NodeHandle>, std::shared_ptr<ros::NodeHandle>) {}
  void run() {}
};
