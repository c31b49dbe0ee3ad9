// shim: lidarFactor.hpp includes this PCL header but uses nothing from it
