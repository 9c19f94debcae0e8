// pluginlib entry: collision_detection::CollisionPlugin::initialize must hand the scene an allocator
// (moveit_core/collision_detection/include/moveit/collision_detection/collision_plugin.hpp:80-94; FCL's loader at
// collision_detection_fcl/src/collision_detector_fcl_plugin_loader.cpp:42-50 is the pattern).  The reference ships
// NO loader for its distance-field detector, so this one is new.
#include <moveit/collision_detection/collision_plugin.hpp>
#include <pluginlib/class_list_macros.hpp>

#include "collision_detector_allocator_b200.hpp"

namespace collision_detection
{
const std::string CollisionDetectorAllocatorB200::NAME("B200_DISTANCE_FIELD");

class CollisionDetectorB200PluginLoader : public CollisionPlugin
{
public:
  bool initialize(const planning_scene::PlanningScenePtr& scene) const override
  {
    scene->allocateCollisionDetector(CollisionDetectorAllocatorB200::create());
    return true;
  }
};
}  // namespace collision_detection

PLUGINLIB_EXPORT_CLASS(collision_detection::CollisionDetectorB200PluginLoader, collision_detection::CollisionPlugin)
