// CollisionDetectorAllocator for the B200 detector -- the template supplies every allocateEnv overload given NAME
// (moveit_core/collision_detection/include/moveit/collision_detection/collision_detector_allocator.hpp:71-99), the
// same way CollisionDetectorAllocatorDistanceField does
// (collision_distance_field/include/moveit/collision_distance_field/collision_detector_allocator_distance_field.hpp:44-49).
#pragma once
#include <moveit/collision_detection/collision_detector_allocator.hpp>

#include "collision_env_b200.hpp"

namespace collision_detection
{
class CollisionDetectorAllocatorB200
  : public CollisionDetectorAllocatorTemplate<CollisionEnvB200, CollisionDetectorAllocatorB200>
{
public:
  static const std::string NAME;  // "B200_DISTANCE_FIELD"
};
}  // namespace collision_detection
