// shim (test infrastructure): the tf types laserMapping.cpp:877-888 fills before broadcasting (the broadcast itself is a no-op here)
#pragma once
#include <ros/ros.h>
namespace tf {
struct Vector3 { double x, y, z; Vector3(double a = 0, double b = 0, double c = 0) : x(a), y(b), z(c) {} };
struct Quaternion { double x = 0, y = 0, z = 0, w = 1; void setX(double v) { x = v; } void setY(double v) { y = v; } void setZ(double v) { z = v; } void setW(double v) { w = v; } };
struct Transform { Vector3 origin; Quaternion rotation; void setOrigin(const Vector3& v) { origin = v; } void setRotation(const Quaternion& q) { rotation = q; } };
struct StampedTransform : Transform {
  ros::Time stamp; std::string frame_id, child_frame_id;
  StampedTransform() {}
  StampedTransform(const Transform& t, const ros::Time& s, const std::string& f, const std::string& c) : Transform(t), stamp(s), frame_id(f), child_frame_id(c) {}
};
}
