// shim (test infrastructure): tf::TransformBroadcaster — the transform duplicates /aft_mapped_to_init, nothing is sent
#pragma once
#include <tf/transform_datatypes.h>
namespace tf { struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} }; }
