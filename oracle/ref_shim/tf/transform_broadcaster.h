#pragma once
#include <tf/transform_datatypes.h>
namespace tf { struct TransformBroadcaster { void sendTransform(const StampedTransform&) {} }; }
