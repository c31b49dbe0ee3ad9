// shim (test infrastructure): laserOdometry.cpp:233-235 initialises gflags/glog; nothing else of them is used
#pragma once
namespace google { inline void ParseCommandLineFlags(int*, char***, bool) {} inline void InitGoogleLogging(const char*) {} }
static bool FLAGS_alsologtostderr __attribute__((unused)) = false;
