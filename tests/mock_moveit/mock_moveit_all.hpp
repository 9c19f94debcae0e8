// TEST INFRASTRUCTURE, not product code.  Minimal stand-ins for the MoveIt 2 / Eigen / geometric_shapes / rclcpp / pluginlib /
// OMPL declarations that moveit2_b200/adapter/ touches, so the adapter can be COMPILED in an image that has none of them
// (tests/test_adapter_compiles.py runs g++ -std=c++17 -Wall -c over every adapter source against this directory).
// Only declarations: nothing here is linked or run, so a member the adapter misuses (wrong name, wrong argument types, a
// const violation, a missing override) is a compile error exactly as it would be against the real headers.  Each block
// names the reference header it stands for (paths relative to /root/reference); signatures are the reference's.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <set>
#include <type_traits>
#include <string>
#include <utility>
#include <vector>

// ---- Eigen (external): the few operations the adapter uses ---------------------------------------------------------------
namespace Eigen
{
struct Vector3d
{
  double v[3];
  Vector3d() : v{ 0, 0, 0 } {}
  Vector3d(double x, double y, double z) : v{ x, y, z } {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  const double* data() const { return v; }
  double* data() { return v; }
};
struct Matrix4d
{
  double m[16];  // column-major
  double operator()(int r, int c) const { return m[4 * c + r]; }
};
struct Isometry3d
{
  Matrix4d mat;
  const Matrix4d& matrix() const { return mat; }
  double operator()(int r, int c) const { return mat(r, c); }
  Vector3d operator*(const Vector3d& p) const;
  Isometry3d operator*(const Isometry3d& o) const;
};
template <typename M>
struct Map
{
  explicit Map(double* p) : p_(p) {}
  Map& operator=(const M& m)
  {
    for (int i = 0; i < 16; ++i)
      p_[i] = m.m[i];
    return *this;
  }
  double* p_;
};
}  // namespace Eigen
namespace EigenSTL  // eigen_stl_containers (external)
{
using vector_Vector3d = std::vector<Eigen::Vector3d>;
using vector_Isometry3d = std::vector<Eigen::Isometry3d>;
}  // namespace EigenSTL

// ---- rclcpp + moveit/utils/logger.hpp (moveit_core/utils/include/moveit/utils/logger.hpp:51) -------------------------------
namespace rclcpp
{
class Logger
{
};
}  // namespace rclcpp
namespace moveit
{
rclcpp::Logger getLogger(const std::string& name);
}
void mock_rclcpp_log(const rclcpp::Logger&, const char* fmt, ...) __attribute__((format(printf, 2, 3)));
#define RCLCPP_WARN(logger, ...) mock_rclcpp_log(logger, __VA_ARGS__)
#define RCLCPP_ERROR(logger, ...) mock_rclcpp_log(logger, __VA_ARGS__)
#define RCLCPP_INFO(logger, ...) mock_rclcpp_log(logger, __VA_ARGS__)

// ---- geometric_shapes (external): shapes.h, bodies.h ------------------------------------------------------------------------
namespace octomap
{
class OcTree;
}
namespace shapes
{
enum ShapeType { UNKNOWN_SHAPE, SPHERE, CYLINDER, CONE, BOX, PLANE, MESH, OCTREE };
class Shape
{
public:
  virtual ~Shape() {}
  ShapeType type;
};
class Sphere : public Shape
{
public:
  double radius;
};
class Cylinder : public Shape
{
public:
  double length, radius;
};
class Box : public Shape
{
public:
  double size[3];
};
class OcTree : public Shape
{
public:
  std::shared_ptr<const octomap::OcTree> octree;
};
using ShapeConstPtr = std::shared_ptr<const Shape>;
}  // namespace shapes
namespace bodies
{
struct BoundingSphere
{
  Eigen::Vector3d center;
  double radius;
};
}  // namespace bodies

// ---- moveit_core/robot_model ---------------------------------------------------------------------------------------------------
namespace moveit
{
namespace core
{
class LinkModel;
class JointModelGroup;
class JointModel  // robot_model/include/moveit/robot_model/joint_model.hpp:117-
{
public:
  enum JointType { UNKNOWN, REVOLUTE, PRISMATIC, PLANAR, FLOATING, FIXED };
  virtual ~JointModel() {}
  JointType getType() const;
  std::size_t getVariableCount() const;
  std::size_t getFirstVariableIndex() const;
  const JointModel* getMimic() const;
  double getMimicOffset() const;
  double getMimicFactor() const;
  double getDistanceFactor() const;
  const std::string& getName() const;
};
class RevoluteJointModel : public JointModel  // revolute_joint_model.hpp:46-
{
public:
  const Eigen::Vector3d& getAxis() const;
  bool isContinuous() const;
};
class PrismaticJointModel : public JointModel  // prismatic_joint_model.hpp:46-
{
public:
  const Eigen::Vector3d& getAxis() const;
};
class LinkModel  // link_model.hpp:71-
{
public:
  const std::string& getName() const;
  std::size_t getLinkIndex() const;
  const JointModel* getParentJointModel() const;
  const LinkModel* getParentLinkModel() const;
  const Eigen::Isometry3d& getJointOriginTransform() const;
  const std::vector<shapes::ShapeConstPtr>& getShapes() const;
  const EigenSTL::vector_Isometry3d& getCollisionOriginTransforms() const;
};
class JointModelGroup  // joint_model_group.hpp:63-
{
public:
  const std::string& getName() const;
  const std::vector<int>& getVariableIndexList() const;
  const std::vector<const JointModel*>& getMimicJointModels() const;
  const std::vector<const JointModel*>& getActiveJointModels() const;
  const std::vector<const JointModel*>& getJointModels() const;
  const std::vector<const LinkModel*>& getUpdatedLinkModels() const;
  bool isLinkUpdated(const std::string& name) const;
  unsigned int getVariableCount() const;
};
class RobotModel  // robot_model.hpp:76-
{
public:
  const JointModelGroup* getJointModelGroup(const std::string& name) const;
  const std::vector<const LinkModel*>& getLinkModels() const;
  const std::vector<const LinkModel*>& getLinkModelsWithCollisionGeometry() const;
  std::size_t getVariableCount() const;
};
using RobotModelConstPtr = std::shared_ptr<const RobotModel>;
using RobotModelPtr = std::shared_ptr<RobotModel>;

class AttachedBody  // robot_state/include/moveit/robot_state/attached_body.hpp:57-
{
public:
  const std::string& getName() const;
  const std::vector<shapes::ShapeConstPtr>& getShapes() const;
  const EigenSTL::vector_Isometry3d& getShapePosesInLinkFrame() const;
  const std::set<std::string>& getTouchLinks() const;
};
class RobotState  // robot_state/include/moveit/robot_state/robot_state.hpp:89-
{
public:
  const double* getVariablePositions() const;
  double getVariablePosition(int index) const;
  void getAttachedBodies(std::vector<const AttachedBody*>& attached_bodies, const LinkModel* link_model) const;
  const AttachedBody* getAttachedBody(const std::string& name) const;
  const Eigen::Isometry3d& getFrameTransform(const std::string& frame_id, bool* frame_found = nullptr) const;
  void copyJointGroupPositions(const JointModelGroup* group, double* gstate) const;
};
}  // namespace core
}  // namespace moveit

// ---- moveit_core/collision_detection ------------------------------------------------------------------------------------------
namespace collision_detection
{
class World  // world.hpp:56-
{
public:
  struct Object  // :78-117
  {
    std::string id_;
    Eigen::Isometry3d pose_;
    std::vector<shapes::ShapeConstPtr> shapes_;
    EigenSTL::vector_Isometry3d shape_poses_, global_shape_poses_;
  };
  using ObjectPtr = std::shared_ptr<Object>;
  using ObjectConstPtr = std::shared_ptr<const Object>;
  enum ActionBits { UNINITIALIZED = 0, CREATE = 1, DESTROY = 2, MOVE_SHAPE = 4, ADD_SHAPE = 8, REMOVE_SHAPE = 16 };  // :254-262
  class Action  // :266-283
  {
  public:
    Action() : action_(UNINITIALIZED) {}
    Action(int v) : action_(v) {}
    operator ActionBits() const { return ActionBits(action_); }

  private:
    int action_;
  };
  class ObserverHandle  // :289-300
  {
  };
  using ObserverCallbackFn = std::function<void(const ObjectConstPtr&, Action)>;
  ObserverHandle addObserver(const ObserverCallbackFn& callback);
  void removeObserver(const ObserverHandle observer_handle);
  void notifyObserverAllObjects(const ObserverHandle observer_handle, Action action) const;
  bool hasObject(const std::string& object_id) const;
  ObjectConstPtr getObject(const std::string& object_id) const;
  const Eigen::Isometry3d& getGlobalShapeTransform(const std::string& object_id, int shape_index) const;
};
using WorldPtr = std::shared_ptr<World>;
using WorldConstPtr = std::shared_ptr<const World>;

namespace AllowedCollision  // collision_matrix.hpp:56-73
{
enum Type { NEVER, ALWAYS, CONDITIONAL };
}
class AllowedCollisionMatrix  // collision_matrix.hpp:84-
{
public:
  bool getEntry(const std::string& name1, const std::string& name2, AllowedCollision::Type& allowed_collision_type) const;
};
namespace BodyTypes  // collision_common.hpp:56-70
{
enum Type { ROBOT_LINK, ROBOT_ATTACHED, WORLD_OBJECT };
}
using BodyType = BodyTypes::Type;
struct Contact  // collision_common.hpp:80-145
{
  Eigen::Vector3d pos, normal;
  double depth;
  std::string body_name_1, body_name_2;
  BodyType body_type_1, body_type_2;
};
struct CollisionRequest  // :147-186
{
  std::string group_name;
  bool distance = false, cost = false, contacts = false;
  std::size_t max_contacts = 1, max_contacts_per_pair = 1, max_cost_sources = 1;
  bool verbose = false;
};
struct CollisionResult  // :334-370
{
  using ContactMap = std::map<std::pair<std::string, std::string>, std::vector<Contact>>;
  bool collision = false;
  double distance = 0.0;
  std::size_t contact_count = 0;
  ContactMap contacts;
};
struct DistanceRequest
{
};
struct DistanceResult
{
};

class CollisionEnv  // collision_env.hpp:52-330
{
public:
  CollisionEnv(const moveit::core::RobotModelConstPtr& model, double padding = 0.0, double scale = 1.0);
  CollisionEnv(const moveit::core::RobotModelConstPtr& model, const WorldPtr& world, double padding = 0.0, double scale = 1.0);
  CollisionEnv(const CollisionEnv& other, const WorldPtr& world);
  virtual ~CollisionEnv() {}
  virtual void checkSelfCollision(const CollisionRequest& req, CollisionResult& res,
                                  const moveit::core::RobotState& state) const = 0;  // :84
  virtual void checkSelfCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                                  const AllowedCollisionMatrix& acm) const = 0;  // :93
  virtual void checkCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state) const;
  virtual void checkCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                              const AllowedCollisionMatrix& acm) const;
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res,
                                   const moveit::core::RobotState& state) const = 0;  // :120
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state,
                                   const AllowedCollisionMatrix& acm) const = 0;  // :130
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                                   const moveit::core::RobotState& state2, const AllowedCollisionMatrix& acm) const = 0;  // :143
  virtual void checkRobotCollision(const CollisionRequest& req, CollisionResult& res, const moveit::core::RobotState& state1,
                                   const moveit::core::RobotState& state2) const = 0;  // :155
  virtual void distanceSelf(const DistanceRequest& req, DistanceResult& res, const moveit::core::RobotState& state) const = 0;
  virtual void distanceRobot(const DistanceRequest& req, DistanceResult& res, const moveit::core::RobotState& state) const = 0;
  virtual void setWorld(const WorldPtr& world);  // :237
  const WorldPtr& getWorld() { return world_; }
  const WorldConstPtr& getWorld() const { return world_const_; }
  using ObjectPtr = World::ObjectPtr;
  using ObjectConstPtr = World::ObjectConstPtr;
  const moveit::core::RobotModelConstPtr& getRobotModel() const { return robot_model_; }
  double getLinkPadding(const std::string& link_name) const;

protected:
  virtual void updatedPaddingOrScaling(const std::vector<std::string>& links);  // :312
  moveit::core::RobotModelConstPtr robot_model_;
  std::map<std::string, double> link_padding_, link_scale_;

private:
  WorldPtr world_;
  WorldConstPtr world_const_;
};
using CollisionEnvPtr = std::shared_ptr<CollisionEnv>;
using CollisionEnvConstPtr = std::shared_ptr<const CollisionEnv>;

class CollisionDetectorAllocator  // collision_detector_allocator.hpp:47-68
{
public:
  virtual ~CollisionDetectorAllocator() {}
  virtual const std::string& getName() const = 0;
  virtual CollisionEnvPtr allocateEnv(const WorldPtr& world, const moveit::core::RobotModelConstPtr& robot_model) const = 0;
  virtual CollisionEnvPtr allocateEnv(const CollisionEnvConstPtr& orig, const WorldPtr& world) const = 0;
  virtual CollisionEnvPtr allocateEnv(const moveit::core::RobotModelConstPtr& robot_model) const = 0;
};
using CollisionDetectorAllocatorPtr = std::shared_ptr<CollisionDetectorAllocator>;
template <class Env, class Alloc>
class CollisionDetectorAllocatorTemplate : public CollisionDetectorAllocator  // :71-99: what the template instantiates
{
public:
  const std::string& getName() const override { return Alloc::NAME; }
  CollisionEnvPtr allocateEnv(const WorldPtr& w, const moveit::core::RobotModelConstPtr& m) const override
  {
    return std::make_shared<Env>(m, w);
  }
  CollisionEnvPtr allocateEnv(const CollisionEnvConstPtr& orig, const WorldPtr& w) const override
  {
    return std::make_shared<Env>(dynamic_cast<const Env&>(*orig), w);
  }
  CollisionEnvPtr allocateEnv(const moveit::core::RobotModelConstPtr& m) const override { return std::make_shared<Env>(m); }
  static CollisionDetectorAllocatorPtr create() { return std::make_shared<Alloc>(); }
};

// ---- moveit_core/collision_distance_field/.../collision_distance_field_types.hpp ----------------------------------------------
struct CollisionSphere  // :62-73
{
  Eigen::Vector3d relative_vec_;
  double radius_;
};
enum CollisionType  // :75-81
{
  NONE = 0,
  SELF = 1,
  INTRA = 2,
  ENVIRONMENT = 3,
};
struct GradientInfo  // :84-113
{
  GradientInfo() : closest_distance(DBL_MAX), collision(false) {}
  double closest_distance;
  bool collision;
  EigenSTL::vector_Vector3d sphere_locations;
  std::vector<double> distances;
  EigenSTL::vector_Vector3d gradients;
  std::vector<CollisionType> types;
  std::vector<double> sphere_radii;
  std::string joint_name;
  void clear();
};
class BodyDecomposition  // :178-240
{
public:
  BodyDecomposition(const shapes::ShapeConstPtr& shape, double resolution, double padding = 0.01);
  BodyDecomposition(const std::vector<shapes::ShapeConstPtr>& shapes, const EigenSTL::vector_Isometry3d& poses,
                    double resolution, double padding);
  const std::vector<CollisionSphere>& getCollisionSpheres() const;
  const EigenSTL::vector_Vector3d& getCollisionPoints() const;
  const bodies::BoundingSphere& getRelativeBoundingSphere() const;
};
using BodyDecompositionPtr = std::shared_ptr<BodyDecomposition>;
using BodyDecompositionConstPtr = std::shared_ptr<const BodyDecomposition>;
class PosedBodyPointDecomposition  // :300-327
{
public:
  PosedBodyPointDecomposition(const BodyDecompositionConstPtr& body_decomposition);
  PosedBodyPointDecomposition(const BodyDecompositionConstPtr& body_decomposition, const Eigen::Isometry3d& pose);
  PosedBodyPointDecomposition(const std::shared_ptr<const octomap::OcTree>& octree);
  const EigenSTL::vector_Vector3d& getCollisionPoints() const;
};
}  // namespace collision_detection

// ---- moveit_core/planning_scene + collision_plugin.hpp:80-94 + pluginlib ----------------------------------------------------------
namespace planning_scene
{
class PlanningScene  // planning_scene.hpp:  allocateCollisionDetector(const CollisionDetectorAllocatorPtr&)
{
public:
  void allocateCollisionDetector(const collision_detection::CollisionDetectorAllocatorPtr& allocator);
};
using PlanningScenePtr = std::shared_ptr<PlanningScene>;
}  // namespace planning_scene
namespace collision_detection
{
class CollisionPlugin
{
public:
  CollisionPlugin() {}
  virtual ~CollisionPlugin() {}
  virtual bool initialize(const planning_scene::PlanningScenePtr& scene) const = 0;
};
}  // namespace collision_detection
// pluginlib/class_list_macros.hpp: registration needs a concrete, default-constructible class derived from the base
#define PLUGINLIB_EXPORT_CLASS(class_type, base_class_type)                                                        \
  static_assert(std::is_base_of<base_class_type, class_type>::value, "plugin class must derive from its base");   \
  base_class_type* mock_pluginlib_factory() { return new class_type(); }

// ---- OMPL (external) + moveit_planners/ompl/ompl_interface ------------------------------------------------------------------------
namespace ompl
{
namespace base
{
class State
{
public:
  template <class T>
  const T* as() const
  {
    return static_cast<const T*>(this);
  }
  template <class T>
  T* as()
  {
    return static_cast<T*>(this);
  }
};
class StateSpace
{
public:
  double getLongestValidSegmentLength() const;
  virtual double distance(const State* state1, const State* state2) const;
  virtual void interpolate(const State* from, const State* to, double t, State* state) const;
  virtual ~StateSpace() {}
};
using StateSpacePtr = std::shared_ptr<StateSpace>;
class SpaceInformation
{
public:
  const StateSpacePtr& getStateSpace() const;
  bool satisfiesBounds(const State* state) const;
};
using SpaceInformationPtr = std::shared_ptr<SpaceInformation>;
class MotionValidator  // ompl/base/MotionValidator.h
{
public:
  MotionValidator(const SpaceInformationPtr& si) : si_(si.get()), valid_(0), invalid_(0) {}
  virtual ~MotionValidator() {}
  virtual bool checkMotion(const State* s1, const State* s2) const = 0;
  virtual bool checkMotion(const State* s1, const State* s2, std::pair<State*, double>& lastValid) const = 0;

protected:
  SpaceInformation* si_;
  mutable unsigned int valid_, invalid_;
};
class StateValidityChecker  // ompl/base/StateValidityChecker.h
{
public:
  StateValidityChecker(const SpaceInformationPtr& si) : si_(si.get()) {}
  virtual ~StateValidityChecker() {}
  virtual bool isValid(const State* state) const = 0;

protected:
  SpaceInformation* si_;
};
}  // namespace base
}  // namespace ompl
namespace ompl_interface
{
class ModelBasedStateSpace : public ompl::base::StateSpace  // parameterization/model_based_state_space.hpp:  StateType::values
{
public:
  class StateType : public ompl::base::State
  {
  public:
    double* values;
  };
};
class ModelBasedPlanningContext  // model_based_planning_context.hpp: getJointModelGroup()
{
public:
  const moveit::core::JointModelGroup* getJointModelGroup() const;
};
}  // namespace ompl_interface
