// MoveIt 2 side of the drop-in: a collision_detection::CollisionEnv whose distance-field checks run on a B200
// through libb200sv.so (include/b200sv.h).  Compiled only where MoveIt 2 headers exist (see CMakeLists.txt in
// this directory).  The authoring image has no ROS 2 / Eigen / geometric_shapes, so there it is compiled against minimal
// stand-in declarations (tests/mock_moveit/, tests/test_adapter_compiles.py: g++ -std=c++17 -Wall -Werror -c): every
// member it calls, every override and every const-qualification is checked by a compiler; it is not linked or run.
// INTEGRATION.md walks through it.
//
// Mirrors collision_detection::CollisionEnvDistanceField
// (moveit_core/collision_distance_field/include/moveit/collision_distance_field/collision_env_distance_field.hpp:57-310):
// same constructors, same overrides of the CollisionEnv pure virtuals (collision_env.hpp:84-195), same World
// observer wiring (collision_env_distance_field.cpp:74-75, 92-95, 1704-1779), plus the batched entry point.
#pragma once

#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include <moveit/collision_detection/collision_env.hpp>
#include <moveit/collision_distance_field/collision_distance_field_types.hpp>

#include "b200sv.h"

namespace collision_detection
{
class CollisionEnvB200 : public CollisionEnv
{
public:
  // defaults of CollisionEnvDistanceField (collision_env_distance_field.hpp:49-55)
  static constexpr double DEFAULT_SIZE_X = 3.0, DEFAULT_SIZE_Y = 3.0, DEFAULT_SIZE_Z = 4.0;
  static constexpr double DEFAULT_RESOLUTION = 0.02, DEFAULT_COLLISION_TOLERANCE = 0.0;
  static constexpr double DEFAULT_MAX_PROPOGATION_DISTANCE = 0.25;

  CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, double padding = 0.0, double scale = 1.0);
  CollisionEnvB200(const moveit::core::RobotModelConstPtr& robot_model, const WorldPtr& world, double padding = 0.0,
                   double scale = 1.0);
  CollisionEnvB200(const CollisionEnvB200& other, const WorldPtr& world);
  ~CollisionEnvB200() override;

  // ---- CollisionEnv interface (collision_env.hpp:84-195) ---------------------------------------------------
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                          const moveit::core::RobotState& state) const override;
  void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                          const AllowedCollisionMatrix& acm) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                           const moveit::core::RobotState& state) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                           const AllowedCollisionMatrix& acm) const override;
  // continuous checks: "not implemented" in the reference too (collision_env_distance_field.cpp:1502-1515)
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2) const override;
  void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                           const moveit::core::RobotState& state2, const AllowedCollisionMatrix& acm) const override;
  void checkCollision(const CollisionRequest& req, CollisionResult& res,
                      const moveit::core::RobotState& state) const override;
  void checkCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                      const AllowedCollisionMatrix& acm) const override;
  // stubs in the reference (collision_env_distance_field.hpp:109-124, 179-199)
  void distanceSelf(const DistanceRequest&, DistanceResult&, const moveit::core::RobotState&) const override {}
  void distanceRobot(const DistanceRequest&, DistanceResult&, const moveit::core::RobotState&) const override {}
  void setWorld(const WorldPtr& world) override;

  /// The world and self fields reproduce PropagationDistanceField's propagation voxel for voxel
  /// (B200SV_PROPAGATION_REFERENCE) instead of the exact transform.  Takes effect for contexts created afterwards: the
  /// cached ones are dropped, the world is replayed into their successors object by object, as the reference adds it.
  void setReferencePropagation(bool on);

  // ---- the new batched entry point --------------------------------------------------------------------------
  /// N group-state vectors (layout of RobotState::setJointGroupPositions(group, const double*), row-major N x V)
  /// -> validity bitmap, bit (i % 8) of byte (i / 8) = 1 iff state i passes PlanningScene::checkCollision's checks
  /// (world + self field + intra-group).  Every variable outside the group comes from `base_state`.
  /// Returns false (and logs) on error; `valid_bitmap` must hold ceil(N / 8) bytes.
  bool checkCollisionBatch(const CollisionRequest& req, const moveit::core::RobotState& base_state,
                           const AllowedCollisionMatrix* acm, const double* q, std::size_t n_states,
                           uint8_t* valid_bitmap, uint32_t flags = B200SV_CHECK_ALL) const;

  /// Batched CollisionEnvDistanceField::getCollisionGradients (collision_env_distance_field.cpp:1517-1538) -- the call
  /// CHOMP makes per trajectory waypoint (chomp_optimizer.cpp:881-915) -- for N group states at once.  gradients[s][e] is
  /// what gsr->gradients_[e] holds in the reference after the call for state s: one entry per updated link of the group
  /// (empty for links without geometry, :1198-1231), then one per attached body (all its shapes' spheres, :1236-1251);
  /// distances DBL_MAX / types NONE where nothing within max_propogation_distance was seen; `collision` stays false (the
  /// reference's gradient path never raises it).  Returns false (and logs, gradients cleared) on error.
  bool getCollisionGradientsBatch(const CollisionRequest& req, const moveit::core::RobotState& base_state,
                                  const AllowedCollisionMatrix* acm, const double* q, std::size_t n_states,
                                  std::vector<std::vector<GradientInfo>>& gradients) const;

private:
  /// One DistanceFieldCacheEntry's worth of state on the device.  Owns its b200sv context: the context is destroyed when
  /// the LAST holder of the shared_ptr lets go, so a thread that is still inside a check keeps it alive even if the
  /// cache has moved on.  `mutex` is held from "configure for this (state, ACM)" to "check done": two threads whose
  /// states differ outside the group cannot reconfigure the context under each other.
  struct GroupContext
  {
    ~GroupContext();
    b200sv_ctx* ctx = nullptr;
    std::mutex mutex;
    // what the context is currently configured for (compareCacheEntryToState / ...ToAllowedCollisionMatrix,
    // collision_env_distance_field.cpp:1254-1370)
    std::vector<int> non_group_indices;      // dfce->state_check_indices_
    std::vector<double> non_group_values;    // dfce->state_values_
    std::vector<uint8_t> self_enabled, intra_enabled;  // the masks the ACM produced
    std::string attached_signature;          // names, touch links, shapes and link-frame poses of the attached bodies
    std::vector<std::string> entry_names;    // dfce entries: the links, then one per shape of every attached body
    std::vector<bool> entry_is_attached;
    // layout of the flat per-sphere outputs (b200sv_collision_gradients): entry e owns spheres [sphere_start[e], [e + 1])
    std::vector<int32_t> sphere_start, body_id;
    std::vector<double> sphere_radii;
    std::vector<std::string> entry_joint_names;  // GradientInfo::joint_name (:1203, :1250)
    std::size_t world_version = 0;           // which state of world_points_ its world field holds
  };
  /// What generateDistanceFieldCacheEntry derives from (group, state, ACM): the flat collision model
  struct FlatModel
  {
    std::vector<int32_t> glink_index, sphere_start, body_id;
    std::vector<uint8_t> self_enabled, intra_enabled;
    std::vector<double> sphere_xyzr, bounding_sphere;
    std::vector<std::string> entry_names;
    std::vector<bool> entry_is_attached;
    std::string attached_signature;
  };
  void flattenModel(const moveit::core::JointModelGroup* jmg, const moveit::core::RobotState& state,
                    const AllowedCollisionMatrix* acm, FlatModel& out) const;
  std::shared_ptr<GroupContext> createContext(const moveit::core::JointModelGroup* jmg,
                                              const moveit::core::RobotState& state, const FlatModel& fm) const;
  void initialize();
  /// getDistanceFieldCacheEntry (collision_env_distance_field.cpp:177-233): the group's context, configured for
  /// (state outside the group, ACM, attached bodies).  Returns it LOCKED (`lock` owns GroupContext::mutex) or nullptr.
  std::shared_ptr<GroupContext> getContext(const std::string& group, const moveit::core::RobotState& state,
                                           const AllowedCollisionMatrix* acm, std::unique_lock<std::mutex>& lock) const;
  void checkSingle(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                   const AllowedCollisionMatrix* acm, uint32_t flags) const;
  void notifyObjectChange(const ObjectConstPtr& obj, World::Action action);
  bool replayWorld(b200sv_ctx* ctx) const;

  Eigen::Vector3d size_, origin_;
  double resolution_, collision_tolerance_, max_propogation_distance_;
  bool use_signed_distance_field_ = false;  // DEFAULT_USE_SIGNED_DISTANCE_FIELD (collision_env_distance_field.hpp:52)
  /// B200SV_PROPAGATION_REFERENCE for the world and self fields: every voxel equals PropagationDistanceField's (tens of
  /// milliseconds per world change instead of a tenth of one).  Default: the exact transform; the environment variable
  /// B200SV_PROPAGATION=reference turns it on for a whole process, setReferencePropagation for one environment.
  bool reference_propagation_ = false;
  std::vector<BodyDecompositionConstPtr> link_body_decomposition_vector_;  // as the reference (:1049-1078)
  std::map<std::string, unsigned int> link_body_decomposition_index_map_;
  // world points currently in the field, per object id (DistanceFieldCacheEntryWorld::posed_body_point_decompositions_)
  // per world object: what it contributed to the world field -- points decomposed on the host (meshes, cones, planes,
  // octree node centres) and primitives decomposed on the device (b200sv_field_add_shape, B200SV_LATTICE_BODY)
  struct WorldObjectRecord
  {
    EigenSTL::vector_Vector3d points;
    std::vector<b200sv_shape> shapes;
  };
  std::map<std::string, WorldObjectRecord> world_points_;
  std::size_t world_version_ = 0;  // bumped by every world change
  mutable std::mutex mutex_;       // guards contexts_ and world_points_ (never held across a device call of a check)
  mutable std::map<std::string, std::shared_ptr<GroupContext>> contexts_;
  World::ObserverHandle observer_handle_;
};
}  // namespace collision_detection
