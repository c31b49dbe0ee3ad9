// shim: nothing from <ceres/rotation.h> is used by lidarFactor.hpp
