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
  int slot, s0, s1, pad;  // device-link slot, sphere range [s0, s1), pad = index of the link in dfce order (glink)
};

// Enabled pair of clusters of two different links whose link pair is enabled in
// dfce->intra_group_collision_enabled_ (i < j in dfce order).
struct alignas(16) PairRec
{
  int a, b;         // BsRec indices
  int quirk_free;   // 1: tight spheres this close imply the reference's link-level cull passes (proved on the host)
  int alt0;         // -1, or the head of an AltRec chain: further bounding-sphere pairs any ONE of which passes the cull
};

// An attached body of several shapes is culled as a whole by the reference: ANY shape's bounding sphere passing
// doBoundingSpheresIntersect lets ALL the body's spheres be tested (collision_env_distance_field.cpp:455-507).  The ABI
// carries one entry per shape, so a pair that involves such a body chains the other (shape, shape) combinations here.
struct alignas(16) AltRec
{
  int a, b;  // BsRec indices (any cluster of the entry: the link-level bounding sphere is the same for all of them)
  int next;  // -1 ends the chain
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
  int off_groups, n_spheres_padded, off_alts, n_alts;  // spheres are padded to a multiple of 4 with thr = 0 dummies
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

// ======================================================================================================================
// Tables of the FAST (FP32) pass, fast_check.cu.  One blob, copied to shared memory by every CTA.  The fast pass works
// in VOXEL coordinates: the host folds f = (p - origin_minus) * oo_resolution (VoxelGrid::getCellFromLocation,
// voxel_grid.hpp:522) into every root transform, so a posed sphere centre IS its (fractional) cell coordinate and all
// lengths below carry the suffix _v (cells).  Joint types: FIXED, REVOLUTE, PRISMATIC (a group with a planar or
// floating joint runs the generic checkKernel<float> instead).
namespace b200sv
{
namespace fast
{
constexpr int kParentRoot = -1;  // the link's transform is [M | m] itself (constants already in the voxel frame)
constexpr int kParentPrev = -2;  // parent = the previous link of the table: its transform is still in registers
// Link::type of a link that does not move relative to an ancestor (fixed joint, or a joint outside the group at its base-state
// value): the host re-expressed its spheres and clusters in that ancestor's frame, the kernel only fetches the ancestor's
// transform -- no joint, no 3x4 product, no store.
constexpr int kAlias = 100;

// T_link = T_parent * [M | m]
//   REVOLUTE   M = a0 + sin(q) * a1 + (1 - cos(q)) * a2,  m = ot        (a0 = O, a1 = O*K, a2 = O*K^2: Rodrigues)
//   PRISMATIC  M = a0,                                     m = ot + q * a1[0..2]   (a1[0..2] = O * axis)
//   FIXED      M = a0,                                     m = ot
// O = rotation of the joint origin, K = cross-product matrix of the axis, q = a * group_q[k] + b evaluated in double
// (setJointGroupPositions + updateMimicJoints) and narrowed.
struct alignas(16) Link
{
  int type, k;     // JointType; index into the group state (k < 0: the joint value is the constant b)
  int parent_off;  // kParentRoot, kParentPrev or the word offset of the parent's transform in the state's block
  int store_off;   // word offset this link's transform is stored at, -1: nobody reads it back
  int h0, h1;      // Hot range of the link's spheres, padded to a multiple of the field loop's unroll
  int c0, c1;      // Cluster range of the link
  double a, b;
  float ot[3];
  int range_check;  // 1: a prismatic joint upstream: verify the parked cluster centres stay within Header::rep_limit
  float a0[9], a1[9], a2[9];
  float pad3;  // the link's cell-face margin in cells (its spheres' worst-case FP32 error; <= Header::margin)
};
static_assert(sizeof(Link) == 176, "fast::Link layout");

// One collision sphere of the field loop, ONE 16-byte broadcast read.
struct alignas(16) Hot
{
  float x, y, z;  // CollisionSphere::relative_vec_ (link frame, metres)
  unsigned t;     // 8-bit fields: 0x01000100 - (thr | thr_self << 16); 16-bit fields: thr | thr_self << 16.
                  // thr = 0 never collides (padding, or the self check disabled for the link)
};

constexpr int kClusterSpheres = 4;  // a cluster holds at most this many spheres; PSphere rows are padded to it
constexpr int kPSphereStride = 5;   // PSphere entries per cluster row: 80 bytes, so 8 lanes reading 8 different rows
                                    // with one 128-bit access touch 8 distinct 16-byte bank groups

// A cluster's tight-sphere centre in its link frame: posed in the link loop and parked in shared memory
// ([cluster][x, y, z][lane]) for the level-1 cull of the pair phase.
struct alignas(16) Cluster
{
  float tx, ty, tz;
  int slot;  // hierarchical cull only: >= 0 = row of the shared-memory arrays (a link-level centre), -1 = global scratch, row = index
};

// Pair phase: sphere k of cluster c is PSphere[kPSphereStride * c + k].  Padding entries have r_v = -1e30: their
// distance test can never pass.
struct alignas(16) PSphere
{
  float x, y, z;  // link frame, metres
  float r_v;      // radius in cells
};

struct alignas(16) Group  // pairs [p0, p1) share cluster a
{
  int a, p0, p1, pad;
};

struct alignas(16) Pair  // 16 bytes: ONE 128-bit read per level-1 lane and per work item
{
  unsigned rt2_h;    // (tr_a + tr_b + slack)^2 in cells^2 as half bits, rounded up (the level-1 cull runs in half)
  unsigned ab;       // Cluster indices a | b << 16
  unsigned offs;     // transform offsets of the two links (words), a_off | b_off << 16
  int quirk;         // -1: the reference's link-level cull provably passes whenever the tight spheres are this close;
                     // else index of the Quirk record to evaluate
};
static_assert(sizeof(Pair) == 16, "fast::Pair layout");

// Hierarchical cull, level 0: a pair of links (bodies) whose cluster pairs [cp0, cp0 + ncp) are contiguous in the Pair
// table.  Same first two fields as Pair (the transposed cull reads only those); la / lb index the parked LINK-level tight
// spheres.  A link pair with more than 32 cluster pairs is split into several records.
struct alignas(16) LinkPair
{
  unsigned rt2_h;
  unsigned ab;  // la | lb << 16
  unsigned cp0, ncp;
};

// doBoundingSpheresIntersect on the LINKS' bounding spheres (types.cpp:408-418), quirk kept: squared centre distance
// (m^2) against the plain radius sum (m).  In cells: |Ca - Cb|_v^2 < (Ra + Rb) * oo^2.
struct alignas(16) Quirk
{
  float ax, ay, az, lim_v2;
  float bx, by, bz, eps_v2;
  int a_off, b_off;
  int next;  // -1, or the next record of an any-shape chain (AltRec above): the cull passes if ANY record of the chain does
  int pad1;
};

struct alignas(16) Header
{
  int n_links, n_groups, n_pairs, n_group_vars;
  int off_links, off_hot, off_psphere, off_clusters;
  int off_groups, off_pairs, off_quirks, total_bytes;
  int state_stride;  // words of shared memory per state (transforms)
  int n_clusters;
  int nz, nynz;
  unsigned idx_bias;  // -(0x4B400000 * (nynz + nz + 1)) mod 2^32: undoes the float-magic bias of the three cell indices
  int t_rows;  // float4 rows of transforms per state (3 per stored link): size of a warp's global-memory region / 32
  int n_linkpairs, off_linkpairs;  // hierarchical cull (many cluster pairs): LinkPair records; 0 = level 1 directly
  float lo[3], margin;      // clamp of (centre - margin) to [0.5 - margin, n - 0.5 - margin]: a centre outside the
  float hi[3], two_margin;  // field lands in the 1-cell border shell, where nothing collides
  float pair_eps_v, tight_eps_v, pad0;
  int n_parked;  // centres parked in SHARED memory: every cluster, or with the hierarchical cull only the link-level ones (the
                 // cluster-level centres then live in a per-warp global scratch: Cluster::slot says where an entry goes)
  float rep_base[3];  // cluster centres are parked relative to this point (cells): small magnitudes for half
  float rep_limit;    // the level-1 slack assumes |centre - rep_base| < rep_limit per axis
  float ident[12];    // identity 3x4: the "parent transform" of a root link
};
}  // namespace fast
}  // namespace b200sv
