// Host-side (C++) model of the path's inputs: the flat robot / collision model the C-ABI takes, a URDF+SRDF
// loader that produces them the way MoveIt's RobotModel / JointModelGroup / CollisionEnvDistanceField would, and
// the double-precision FK used once per context for everything that does not move with the group.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "device_model.hpp"

namespace b200sv
{
inline int jointVarCount(int t) { return t == REVOLUTE || t == PRISMATIC ? 1 : t == PLANAR ? 3 : t == FLOATING ? 7 : 0; }

using Mat4 = std::array<double, 16>;  // column-major, as Eigen::Isometry3d

struct HostRobot
{
  int32_t n_links = 0, n_vars = 0;
  std::vector<int32_t> parent, joint_type, first_var;
  std::vector<double> joint_axis;    // n_links*3
  std::vector<double> joint_origin;  // n_links*16
  std::vector<double> base_positions;
  std::vector<int32_t> group_var_index;
  std::vector<int32_t> mimic_dst, mimic_src;
  std::vector<double> mimic_factor, mimic_offset;
  std::vector<std::string> link_names;          // optional (URDF path)
  std::vector<double> group_lower, group_upper;  // optional (URDF path)
};

struct HostCollisionModel
{
  int32_t n_glinks = 0;
  std::vector<int32_t> glink_index;
  std::vector<uint8_t> self_enabled;
  std::vector<uint8_t> intra_enabled;  // n_glinks*n_glinks
  std::vector<int32_t> sphere_start;   // n_glinks+1
  std::vector<double> sphere_xyzr;     // n_spheres*4
  std::vector<double> bounding_sphere; // n_glinks*4
  std::vector<int32_t> body_id;        // n_glinks or empty: -1 for links, shapes of one attached body share an id
};

// RobotState::updateLinkTransformsInternal in double (robot_state.cpp:793-848): all links, link-index order.
void hostForwardKinematics(const HostRobot& r, const std::vector<double>& positions, std::vector<Mat4>& out);
Mat4 mulAffine(const Mat4& a, const Mat4& b);
bool isIdentityTransform(const double* m16);

// URDF + SRDF text -> HostRobot / HostCollisionModel for `group`, plus the self-field points
// (collision_env_distance_field.cpp:907-953).  SPHERE collision geometry only.  Returns false and sets err.
// interior lattice points (link frame) of every link with geometry that the group does not move: (link index, xyz...)
using SelfLocalPoints = std::vector<std::pair<int, std::vector<double>>>;
bool loadUrdfSrdf(const std::string& urdf, const std::string& srdf, const std::string& group, double resolution,
                  HostRobot& robot, HostCollisionModel& cm, std::vector<double>& self_points, SelfLocalPoints& self_local,
                  SelfLocalPoints& link_points, std::string& err, bool& parse_error);
// The self field's generating points at `positions` (all model variables): every link's points moved by its global transform
// (generateDistanceFieldCacheEntry, collision_env_distance_field.cpp:907-953).
void poseSelfPoints(const HostRobot& robot, const std::vector<double>& positions, const SelfLocalPoints& self_local,
                    std::vector<double>& self_points);
}  // namespace b200sv
