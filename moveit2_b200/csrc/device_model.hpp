// Records shared by the host table builder and the kernels.  One blob per precision (float for the fast pass,
// double for the exact pass) lives in global memory; every CTA copies it into shared memory once.
#pragma once
#include <cstdint>

namespace b200sv
{
// moveit::core::JointModel::JointType subset; values match B200SV_JOINT_* in include/b200sv.h
enum JointType : int32_t { FIXED = 0, REVOLUTE = 1, PRISMATIC = 2, PLANAR = 3, FLOATING = 4 };

// 3x4 row-major affine transforms everywhere on the device (the last row of an isometry is implicit).
template <typename R>
struct alignas(16) LinkRec
{
  R o[12];     // joint origin; for a link whose parent does not move with the group: T_parent(base state) * origin
  R ax[3];     // joint axis
  int parent;  // device-link slot of the parent, -1 if `o` is already global
  int type;    // JointType
  int var;     // first VarRec of the joint
  int has_origin;  // 0: `o` is the identity (LinkModel::jointOriginTransformIsIdentity), skip the product
  int s0, s1;      // collision spheres carried by this link: SphereRec range [s0, s1) (spheres are sorted by link)
  int pad;
};

// Value of one full-state variable as an affine function of the group state: v = a * q[k] + b (k < 0: v = b).
// Encodes setJointGroupPositions + updateMimicJoints(group) (robot_state.cpp:571-584, 210-220).
struct VarRec
{
  double a, b;
  int k, pad;
};

template <typename R>
struct alignas(16) SphereRec
{
  R x, y, z, r;  // CollisionSphere::relative_vec_ / radius_
  int slot;      // device-link slot
  int thr;       // collide with the WORLD field iff d2[voxel] < thr  (integer form of types.cpp:229)
  int thr_self;  // same for the SELF field; 0 when dfce->self_collision_enabled_ is false for the link
  int pad;
};

// A CLUSTER: a contiguous run of one link's collision spheres.  Carries the LINK's reference bounding sphere (for
// the reference's own cull, identical for every cluster of the link) and the cluster's tight sphere (ours).
template <typename R>
struct alignas(16) BsRec
{
  R x, y, z, r;      // BodyDecomposition::getRelativeBoundingSphere() of the link: doBoundingSpheresIntersect (types.cpp:408-418)
  R tx, ty, tz, tr;  // tight bounding sphere of the cluster's spheres (centroid, max |c_k - c| + r_k): a conservative
                     // extra cull -- it can only skip pairs that cannot collide
  int slot, s0, s1, pad;  // device-link slot, sphere range [s0, s1)
};

// Enabled pair of clusters of two different links whose link pair is enabled in
// dfce->intra_group_collision_enabled_ (i < j in dfce order).
struct alignas(16) PairRec
{
  int a, b;         // BsRec indices
  int quirk_free;   // 1: tight spheres this close imply the reference's link-level cull passes (proved on the host)
  int pad;
};

// Pairs [p0, p1) share cluster `a`: the level-1 loop poses it once.
struct alignas(16) GroupRec
{
  int a, p0, p1, pad;
};

template <typename R>
struct ModelHeader
{
  int n_links, n_vars, n_spheres, n_bs, n_pairs, n_group_vars;
  int off_links, off_vars, off_spheres, off_bs, off_pairs, total_bytes;
  int state_stride, link_stride;  // shared-memory strides (in R) of a state's / a link's transform
  int nx, ny, nz, n_groups;
  int off_groups, n_spheres_padded, pad1, pad2;  // spheres are padded to a multiple of 4 with thr = 0 dummies
  R om[3];      // VoxelGrid::origin_minus_
  R oo_res;     // VoxelGrid::oo_resolution_
  R margin;     // FP32 pass: voxel-coordinate uncertainty (cells); 0 in the FP64 pass
  R pair_eps;   // FP32 pass: uncertainty of a centre distance (m)
  R bs_eps;     // FP32 pass: uncertainty of a SQUARED bounding-centre distance (m^2)
  R tight_eps;  // slack (m) of the conservative tight-sphere cull (FP32: covers rounding; FP64: 1e-9)
};

// Shared-memory layout of one state's transforms: link stride 12 words (3x4 row-major), every row 16-byte aligned
// so a row is ONE 128-bit access.  The state stride is padded to 4 (mod 32) words in the FP32 blob: the 8 lanes of a
// quarter-warp then touch 8 distinct 16-byte bank groups (conflict-free 128-bit accesses).
constexpr int kLinkStrideWords = 12;
}  // namespace b200sv
