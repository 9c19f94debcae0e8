/*
 * b200sv -- B200-native batched state-validity engine: the C-ABI drop-in boundary.
 *
 * libb200sv.so replaces ONE path of MoveIt 2 (reference paths relative to moveit_core/):
 *   RobotState::setJointGroupPositions + updateLinkTransforms        robot_state/src/robot_state.cpp:571-584, 793-848
 *   CollisionEnvDistanceField::checkSelfCollision / checkRobotCollision / checkCollision
 *                                                                    collision_distance_field/src/collision_env_distance_field.cpp:235-272, 1382-1500
 *   PropagationDistanceField::addPointsToField / removePointsFromField / reset
 *                                                                    distance_field/src/propagation_distance_field.cpp:177-219, 506-524
 * A MoveIt-side CollisionEnv subclass (moveit2_b200/adapter/) binds exactly these entry points; INTEGRATION.md
 * shows the binding.  Plain C types only: no STL, Eigen or torch types cross this boundary.
 *
 * Conventions
 *   - All matrices are 4x4 column-major doubles, the memory layout of Eigen::Isometry3d.
 *   - Link order is the RobotModel link index (DFS pre-order, RobotModel::getLinkModels()).
 *   - A state is the JointModelGroup's variable vector, exactly what RobotState::setJointGroupPositions(group,
 *     const double*) takes: the group's joints sorted by joint index, each joint's variables in declaration
 *     order, mimic-joint slots present but overwritten (joint_model_group.cpp:127-171, robot_state.cpp:210-220).
 *   - Every function returns B200SV_OK (0) or a negative error; b200sv_last_error() has the message.  There is no
 *     CPU fallback: without a usable CUDA device creation fails with B200SV_ERR_NO_DEVICE.
 *   - Calls on one context are serialised by an internal mutex; separate contexts are independent.
 *   - The *_device entry points run on the CALLER's stream but share the context's scratch buffers: the library orders each
 *     such call behind the context's own stream and behind the previous call on a different caller stream (events, no host
 *     synchronisation), and orders the context's next own-stream work behind it.
 */
#ifndef B200SV_H
#define B200SV_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200SV_OK 0
#define B200SV_ERR_INVALID (-1)     /* bad argument / inconsistent model */
#define B200SV_ERR_CUDA (-2)        /* a CUDA call failed */
#define B200SV_ERR_UNSUPPORTED (-3) /* valid in the reference, not implemented here (signed fields, ...) */
#define B200SV_ERR_NO_DEVICE (-4)   /* no CUDA device: the product path never falls back to the CPU */
#define B200SV_ERR_PARSE (-5)       /* URDF / SRDF text could not be parsed */

/* device ordinal for a HOST-ONLY context: the model is parsed and the lookup tables are built so they can be
 * inspected (b200sv_get_*), but every compute entry point returns B200SV_ERR_NO_DEVICE. */
#define B200SV_DEVICE_NONE (-1)

typedef struct b200sv_ctx b200sv_ctx;

/* moveit::core::JointModel::JointType, the subset with a computeTransform (robot_model/src/{revolute,prismatic,planar,floating,fixed}_joint_model.cpp) */
enum
{
  B200SV_JOINT_FIXED = 0,
  B200SV_JOINT_REVOLUTE = 1, /* also URDF "continuous" */
  B200SV_JOINT_PRISMATIC = 2,
  B200SV_JOINT_PLANAR = 3,
  B200SV_JOINT_FLOATING = 4
};

/* Flat view of moveit::core::RobotModel + one JointModelGroup + the RobotState that supplies every variable the
 * group does not own.  Arrays are indexed by link index; entry i describes link i and its parent joint. */
typedef struct b200sv_robot
{
  int32_t n_links;              /* RobotModel::getLinkModelCount() */
  int32_t n_vars;               /* RobotModel::getVariableCount() */
  const int32_t* parent;        /* LinkModel::getParentLinkModel()->getLinkIndex(), -1 for the root link */
  const int32_t* joint_type;    /* type of LinkModel::getParentJointModel() */
  const double* joint_axis;     /* [n_links*3] RevoluteJointModel/PrismaticJointModel::getAxis() (revolute: normalised) */
  const double* joint_origin;   /* [n_links*16] LinkModel::getJointOriginTransform(), column-major */
  const int32_t* first_var;     /* JointModel::getFirstVariableIndex() of the parent joint */
  const double* base_positions; /* [n_vars] RobotState::getVariablePositions() for everything outside the group */
  int32_t n_group_vars;         /* JointModelGroup::getVariableCount() */
  const int32_t* group_var_index; /* [n_group_vars] JointModelGroup::getVariableIndexList() */
  int32_t n_mimic;              /* JointModelGroup::getMimicJointModels().size() */
  const int32_t* mimic_dst;     /* first variable index of each group mimic joint */
  const int32_t* mimic_src;     /* first variable index of the joint it mimics */
  const double* mimic_factor;   /* JointModel::getMimicFactor() */
  const double* mimic_offset;   /* JointModel::getMimicOffset() */
} b200sv_robot;

/* Flat view of DistanceFieldCacheEntry + the BodyDecompositions of the group's updated links
 * (collision_distance_field/include/moveit/collision_distance_field/collision_common_distance_field.hpp:89-124,
 *  collision_env_distance_field.cpp:735-838). */
typedef struct b200sv_collision_model
{
  int32_t n_glinks;             /* dfce->link_names_.size() + attached-body entries */
  const int32_t* glink_index;   /* model link index of each.  The links come first, in getUpdatedLinkModels() order; an
                                   index that already appeared denotes an ATTACHED BODY carried by that link
                                   (dfce->attached_body_names_, collision_env_distance_field.cpp:798-822): its spheres and
                                   bounding sphere are given in the LINK frame (attach pose folded in), one entry per
                                   shape of the body */
  const uint8_t* self_enabled;  /* dfce->self_collision_enabled_ [n_glinks] */
  const uint8_t* intra_enabled; /* dfce->intra_group_collision_enabled_ [n_glinks*n_glinks]; only j > i is read */
  const int32_t* sphere_start;  /* [n_glinks+1] range of each link's collision spheres; empty range = no geometry */
  const double* sphere_xyzr;    /* [n_spheres*4] CollisionSphere::relative_vec_ (link frame) and radius_ */
  const double* bounding_sphere; /* [n_glinks*4] BodyDecomposition::getRelativeBoundingSphere() centre, radius */
  const int32_t* body_id;       /* [n_glinks] or NULL.  -1 for links; the entries (shapes) of ONE attached body share an id
                                   and must ride on one link.  The reference culls a (link, body) or (body, body) pair as a
                                   whole: if ANY shape's bounding sphere passes doBoundingSpheresIntersect, ALL of the
                                   body's spheres are tested (collision_env_distance_field.cpp:455-507).  The entries here
                                   are per shape, so with body_id given every pair that involves a several-shape body is
                                   either PROVED to pass that test whenever its spheres are in reach (the ordinary case:
                                   the squared-distance-vs-radius-sum test of types.cpp:416-417 only rejects near spheres
                                   when bounding radii sum to about a metre) or carries the other shapes' bounding spheres
                                   and evaluates the any-shape rule as the reference does.  NULL: every attached entry is
                                   a body of its own. */
} b200sv_collision_model;

/* Constructor arguments of CollisionEnvDistanceField (collision_env_distance_field.hpp:49-84): the world and the
 * self field share this geometry; the grid corner is origin - size/2 (collision_env_distance_field.cpp:1786-1787). */
typedef struct b200sv_field_cfg
{
  double size[3];
  double origin[3];
  double resolution;
  double max_propagation_distance;
  double collision_tolerance;
  int32_t use_signed_distance_field; /* 0 = the reference default (collision_env_distance_field.hpp:52); 1 = propagate_negative_:
                                        every rebuild also transforms the complement (distance of an occupied voxel to the
                                        nearest free one), getDistance = sqrt(d2) * res - sqrt(neg d2) * res
                                        (propagation_distance_field.hpp:604-607) feeds the gradients */
} b200sv_field_cfg;

typedef struct b200sv_info
{
  int32_t n_links, n_vars, n_group_vars, n_glinks, n_spheres, n_link_pairs;
  int32_t num_cells[3]; /* VoxelGrid::getNumCells */
  int32_t max_distance_sq;
  int32_t device;
  int32_t field_bytes_per_voxel; /* compact field the check kernel gathers from: 1 or 2 */
  int32_t sm_count;
} b200sv_info;

typedef struct b200sv_stats
{
  uint64_t n_states;    /* states in the last check call */
  uint64_t n_valid;     /* collision-free states */
  uint64_t n_rechecked; /* states the FP32 pass could not decide and the FP64 pass recomputed */
  float check_ms;       /* device time of the last check call's kernels (CUDA events) */
  float edt_ms;         /* device time of the last field rebuild */
} b200sv_stats;

enum { B200SV_FIELD_WORLD = 0, B200SV_FIELD_SELF = 1 };

/* flags of b200sv_check_batch */
#define B200SV_CHECK_WORLD 1u /* getEnvironmentCollisions  (checkRobotCollision) */
#define B200SV_CHECK_SELF 2u  /* getSelfCollisions         (checkSelfCollision, field part) */
#define B200SV_CHECK_INTRA 4u /* getIntraGroupCollisions   (checkSelfCollision, sphere pairs) */
#define B200SV_CHECK_ALL 7u   /* PlanningScene::checkCollision */
#define B200SV_CHECK_FP64 256u /* run every state through the FP64 kernel (default: FP32 pass + FP64 recheck of
                                  the states whose result could depend on FP32 rounding) */

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* Replaces the CollisionEnvDistanceField constructor (collision_env_distance_field.cpp:59-96). */
int b200sv_create(int device, const b200sv_robot* robot, const b200sv_collision_model* model,
                  const b200sv_field_cfg* field, b200sv_ctx** out);
/* Same, from URDF + SRDF text, for hosts without a RobotModel (tests, bench): restates RobotModel construction
 * (robot_model/src/robot_model.cpp:848-890, 1173-1254; joint_model_group.cpp:114-300), the SRDF-derived ACM
 * (collision_detection/src/collision_matrix.cpp:67-78) and, for SPHERE collision geometry, BodyDecomposition
 * (collision_distance_field_types.cpp:59-85, 292-335).  The self field is built from the non-group links as
 * generateDistanceFieldCacheEntry does (collision_env_distance_field.cpp:907-953). */
int b200sv_create_from_urdf(int device, const char* urdf_xml, const char* srdf_xml, const char* group_name,
                            const b200sv_field_cfg* field, b200sv_ctx** out);
void b200sv_destroy(b200sv_ctx* ctx);
/* ctx == NULL returns the message of the last failed create on this thread. */
const char* b200sv_last_error(const b200sv_ctx* ctx);
int b200sv_get_info(const b200sv_ctx* ctx, b200sv_info* info);
int b200sv_get_stats(b200sv_ctx* ctx, b200sv_stats* stats);
/* Flat model accessors for the URDF path (so a caller can cross-check against its own RobotModel). Arrays may be
 * NULL to skip. link_names: '\n'-joined, written into buf (cap bytes). */
int b200sv_get_robot_arrays(const b200sv_ctx* ctx, int32_t* parent, int32_t* joint_type, double* joint_axis,
                            double* joint_origin, int32_t* first_var, double* base_positions,
                            int32_t* group_var_index);
int b200sv_get_model_arrays(const b200sv_ctx* ctx, int32_t* glink_index, uint8_t* self_enabled,
                            uint8_t* intra_enabled, int32_t* sphere_start, double* sphere_xyzr,
                            double* bounding_sphere);
int b200sv_get_link_names(const b200sv_ctx* ctx, char* buf, size_t cap);
/* Variable bounds of the group, as RobotState::setToRandomPositions(group) samples them. */
int b200sv_get_group_bounds(const b200sv_ctx* ctx, double* lower, double* upper);
/* Integer collision thresholds the check kernel compares voxel d^2 against, one per sphere (dfce order):
 * collide iff d2 < threshold -- the exact integer form of collision_distance_field_types.cpp:229. */
int b200sv_get_sphere_thresholds(const b200sv_ctx* ctx, int32_t* thresholds);

/* ---- distance fields: PropagationDistanceField (distance_field/src/propagation_distance_field.cpp) -------- */
/* reset(): :506-524 */
int b200sv_field_reset(b200sv_ctx* ctx, int which);
/* addPointsToField(): :177-196 + propagatePositive :390-444, as ONE parallel exact truncated EDT.  xyz: host. */
int b200sv_field_add_points(b200sv_ctx* ctx, int which, const double* xyz, size_t n_points);
/* removePointsFromField(): :198-219 (result equals a fresh build without those voxels, test_distance_field.cpp:508) */
int b200sv_field_remove_points(b200sv_ctx* ctx, int which, const double* xyz, size_t n_points);
/* updatePointsInField(): :130-175 -- voxels of old_xyz that are not voxels of new_xyz are removed, voxels of new_xyz that
 * are not voxels of old_xyz are added (a voxel in both is left as it is), then the field is rebuilt. */
int b200sv_field_update_points(b200sv_ctx* ctx, int which, const double* old_xyz, size_t n_old, const double* new_xyz,
                               size_t n_new);
/* Several occupancy changes, ONE transform: between begin and end every add / remove / update / add_shape call only edits the
 * occupancy bits; end runs the EDT once if anything changed.  (CollisionEnvDistanceField::updateDistanceObject,
 * collision_env_distance_field.cpp:1713-1725, removes an object's old points and adds its new ones -- two propagations in
 * the reference, one rebuild here.)  Reads of the field inside the bracket see the state before it. */
int b200sv_field_begin_update(b200sv_ctx* ctx, int which);
int b200sv_field_end_update(b200sv_ctx* ctx, int which);
/* ---- which propagation a field runs --------------------------------------------------------------------------------------
 * B200SV_PROPAGATION_EXACT (default): every change rebuilds the field with the exact truncated Euclidean transform.  Its
 *   values are never above the reference's and differ from them in a fraction of a percent of the voxels, because
 *   PropagationDistanceField::propagatePositive (propagation_distance_field.cpp:390-444) hands closest points along a
 *   direction-limited neighbourhood in queue order and can miss the true nearest obstacle.
 * B200SV_PROPAGATION_REFERENCE: the field keeps the reference's per-voxel record (distance_square_, closest_point_,
 *   update_direction_, propagation_distance_field.hpp:82-116) and its bucket queue on the device, and b200sv_field_add_points
 *   / remove_points / update_points / add_points_device run addNewObstacleVoxels (:221-301), removeObstacleVoxels (:303-388)
 *   and propagatePositive in data-parallel phases whose outcome is the serial queue's: every voxel's squared distance equals
 *   the reference's after the same sequence of calls with the same point arrays (the order of the points and of the calls
 *   matters, as it does in the reference).  Milliseconds per call instead of a tenth of one.  Signed fields run the
 *   negative half the same way (the flood of :255-298, propagateNegative :446-504).
 *   b200sv_field_read_df is readFromStream (:675-760: a reset, then the stream's voxels in x, y, z order).  The
 *   shape / octree / set_d2 / import entry points refuse with B200SV_ERR_UNSUPPORTED while a field is in this mode
 *   (feed their points through add_points), and begin/end_update brackets are refused too (the reference propagates per
 *   call).  b200sv_set_base_state(.., rebuild_self_field = 1) builds a self field in this mode as the reference builds a
 *   new cache entry's: a fresh field, the links' points added in one call.
 * Switching the mode resets the field; the self field is then refilled from the link points the context holds. */
#define B200SV_PROPAGATION_EXACT 0
#define B200SV_PROPAGATION_REFERENCE 1
int b200sv_field_set_propagation(b200sv_ctx* ctx, int which, int mode);
typedef struct b200sv_propagation_stats
{
  uint64_t rounds;               /* data-parallel rounds so far (at least one per non-empty bucket and call) */
  uint64_t hazard_rounds;        /* rounds cut short because an offer changed a voxel queued later in the same bucket */
  uint64_t queue_entries;        /* bucket-queue entries drained so far */
  uint64_t backward_offers;      /* accepted offers into a bucket at or below the one being drained (:426-437 pushes them
                                    behind the loop: never run in that call) */
  uint64_t queued_for_next_call; /* such entries still waiting in the queue, as in the reference */
} b200sv_propagation_stats;
int b200sv_field_propagation_stats(b200sv_ctx* ctx, int which, b200sv_propagation_stats* out);

/* DistanceField::addShapeToField for a shapes::Sphere (distance_field.cpp:208-238): the interior lattice points of
 * findInternalPointsConvex (find_internal_points.cpp:39-64; bodies::Sphere::containsPoint = |c - p|^2 <= r^2) are marked on
 * the device and the field is rebuilt.  This is the construction of the reference's fixture moveit_core/test_big.df
 * (test_distance_field.cpp:957-963).  Boxes, cylinders and meshes go through the caller's bodies::Body (their
 * containsPoint lives in the un-vendored geometric_shapes) and b200sv_field_add_points. */
int b200sv_field_add_sphere(b200sv_ctx* ctx, int which, const double* center_xyz, double radius);
/* Spheres, cylinders and boxes decomposed into their interior lattice points ON THE DEVICE, in either of the two ways the
 * reference does it:
 *   B200SV_LATTICE_WORLD  DistanceField::addShapeToField / removeShapeFromField (distance_field.cpp:206-238, 316-326): the body
 *                         is posed, the lattice of findInternalPointsConvex is laid out in the world frame;
 *   B200SV_LATTICE_BODY   what CollisionEnvDistanceField does with world objects (updateDistanceObject :1730-1779 ->
 *                         BodyDecomposition::init, types.cpp:292-335, + PosedBodyPointDecomposition::updatePose :388-406): the
 *                         lattice is laid out around the body at the identity pose, the kept points are moved by `pose`.
 * type = shapes::ShapeType; dims: sphere { radius }, cylinder { radius, length }, box { x, y, z }; pose = the top three rows of
 * the Isometry3d, row-major [R | t].  containsPoint follows geometric_shapes' bodies.cpp (scale 1, padding 0), which the
 * reference does not vendor: for cylinders and boxes parity is unpinned (spheres: test_big.df).  remove != 0 clears the
 * voxels instead.  The field is rebuilt. */
#define B200SV_SHAPE_SPHERE 1
#define B200SV_SHAPE_CYLINDER 2
#define B200SV_SHAPE_BOX 4
#define B200SV_LATTICE_WORLD 0
#define B200SV_LATTICE_BODY 1
typedef struct b200sv_shape
{
  int32_t type;
  int32_t lattice_frame;
  double dims[3];
  double pose[12];
} b200sv_shape;
int b200sv_field_add_shape(b200sv_ctx* ctx, int which, const b200sv_shape* shape, int remove);
/* A convex mesh (bodies::ConvexMesh of geometric_shapes) decomposed on the device: the caller -- who owns the mesh and the
 * convex-hull code -- passes the hull's planes, the library lays out findInternalPointsConvex's lattice around the bounding
 * sphere and keeps the points on the inner side of every plane (ConvexMesh::containsPoint -> isPointInsidePlanes:
 * normal . (i_pose * p) + w - 1e-9 <= 0; geometric_shapes is un-vendored in the reference, parity unpinned).
 *   planes           [n_planes * 4] unit normal x, y, z and w in the body's base frame (w as the library recomputes it from a
 *                    scaled + padded vertex of the plane)
 *   bounding_sphere  centre (body frame) and radius: ConvexMesh::computeBoundingSphere
 *   lattice_frame, pose, remove: as b200sv_field_add_shape */
typedef struct b200sv_convex
{
  int32_t n_planes;
  int32_t lattice_frame;
  const double* planes;
  double bounding_sphere[4];
  double pose[12];
} b200sv_convex;
int b200sv_field_add_convex(b200sv_ctx* ctx, int which, const b200sv_convex* body, int remove);
/* DistanceField::addOcTreeToField (distance_field.cpp:240-291) behind octomap's leaf iteration, which stays with the caller
 * (octomap is its dependency): leaves [n_leaves * 4] = centre x, y, z and size of every OCCUPIED leaf inside the field's
 * bounding box.  A leaf no larger than a cell marks its centre's voxel, a larger one the lattice the reference fills it with
 * (x = cx - c .. cx + c in accumulated steps of the resolution, c = ceil(size / resolution) * resolution / 2).  Rebuilds.
 * (World octrees of CollisionEnvDistanceField go through PosedBodyPointDecomposition(octree) instead, types.cpp:355-365:
 * every tree node's centre as a point -- that is b200sv_field_add_points.) */
int b200sv_field_add_octree_leaves(b200sv_ctx* ctx, int which, const double* leaves, size_t n_leaves);
/* Same two with xyz already in device memory, asynchronous on `cuda_stream` (a cudaStream_t). */
int b200sv_field_add_points_device(b200sv_ctx* ctx, int which, const double* d_xyz, size_t n_points,
                                   void* cuda_stream);
/* int32 squared cell distances, [nx*ny*nz], index x*ny*nz + y*nz + z (voxel_grid.hpp:425); host buffers. */
int b200sv_field_get_d2(b200sv_ctx* ctx, int which, int32_t* d2_out);
/* Inject an externally built field (e.g. the reference's own propagation result) so check parity can be tested
 * independently of EDT parity. */
int b200sv_field_set_d2(b200sv_ctx* ctx, int which, const int32_t* d2);
/* Signed fields (propagateNegative, propagation_distance_field.cpp:446-504): negative_distance_square_ per voxel, same
 * layout; 0 in free voxels and in every voxel of an unsigned field.  set: inject, like b200sv_field_set_d2. */
int b200sv_field_get_neg_d2(b200sv_ctx* ctx, int which, int32_t* neg_d2_out);
int b200sv_field_set_neg_d2(b200sv_ctx* ctx, int which, const int32_t* neg_d2);
/* PropagationDistanceField::writeToStream / readFromStream (propagation_distance_field.cpp:635-760): the ".df" format of
 * the reference's fixtures (moveit_core/test_small.df, test_big.df) -- 7 ASCII header lines, then zlib data with the
 * occupancy, x-major, then y, then z packed 8 voxels per byte LSB-first.  write: call with out = NULL to learn the size.
 * read: resets the field and rebuilds it from the stream's occupied voxels; the stream's geometry must equal the
 * context's (the reference resizes its field instead). */
int b200sv_field_write_df(b200sv_ctx* ctx, int which, uint8_t* out, size_t cap, size_t* n_out);
int b200sv_field_read_df(b200sv_ctx* ctx, int which, const uint8_t* bytes, size_t n_bytes);
/* The self field's generating points (URDF path only): n_out receives the count; xyz may be NULL. */
int b200sv_get_self_points(const b200sv_ctx* ctx, double* xyz, size_t cap_points, size_t* n_out);

/* ---- the hot path ------------------------------------------------------------------------------------------ */
/* Batched RobotState::setJointGroupPositions + updateLinkTransforms: q [n*n_group_vars] (host) ->
 * link_T [n*n_links*16] (host), global link transforms, column-major.  precision_bits: 32 or 64. */
int b200sv_fk_batch(b200sv_ctx* ctx, const double* q, size_t n_states, int precision_bits, double* link_T);
/* Batched PlanningScene::checkCollision over n group-state vectors.  valid_bitmap: ceil(n/8) bytes (host),
 * bit (i % 8) of byte (i / 8) is 1 iff state i is collision-free under the selected checks.
 * Timed region of the end-to-end benchmark: H2D of q, kernels, D2H of the bitmap.
 * Small batches take short cuts that change the latency, never the answer: up to 32 states are ONE kernel launch with the
 * states in the launch parameters and the results polled in mapped host memory; ONE state (a planner's per-state isValid)
 * is answered by a kernel that stays resident on the context's stream for 200 us after its last answer (B200SV_LINGER_US;
 * B200SV_LINGER=0 turns it off) and is fed through a mailbox in mapped host memory, so back-to-back calls pay no launch.
 * Every other entry point of the context asks that kernel to leave first; it never outlives 20 ms. */
int b200sv_check_batch(b200sv_ctx* ctx, const double* q, size_t n_states, uint32_t flags, uint8_t* valid_bitmap);
/* Same with q and the bitmap resident in device memory; asynchronous on `cuda_stream`.
 * d_valid_bitmap must hold 4*ceil(n/32) bytes (whole 32-bit words are written). */
int b200sv_check_batch_device(b200sv_ctx* ctx, const double* d_q, size_t n_states, uint32_t flags,
                              uint8_t* d_valid_bitmap, void* cuda_stream);
/* Same with float32 states -- NOT the reference's layout (RobotState holds doubles): for callers that keep their samples in
 * float and want half the host-to-device traffic.  Every value is widened on the device, q = (double)q_f32, so the result is
 * bit-for-bit what b200sv_check_batch returns for those widened states. */
int b200sv_check_batch_f32(b200sv_ctx* ctx, const float* q, size_t n_states, uint32_t flags, uint8_t* valid_bitmap);
/* b200sv_fk_batch with q and link_T resident in device memory; asynchronous on `cuda_stream`. */
int b200sv_fk_batch_device(b200sv_ctx* ctx, const double* d_q, size_t n_states, int precision_bits, double* d_link_T,
                           void* cuda_stream);
/* Batched edge ("motion") validation -- new; the consumer is ompl::base::MotionValidator::checkMotion as MoveIt
 * configures it (DiscreteMotionValidator over ModelBasedStateSpace; longest_valid_segment_fraction:
 * moveit_planners/ompl/ompl_interface/src/model_based_planning_context.cpp:294-317).  For edge e the states checked are
 * JointModelGroup::interpolate(from, to, j / nd) (robot_model/src/joint_model_group.cpp:474-486; continuous revolute
 * joints along the shorter arc, revolute_joint_model.cpp:138-171) for j = 1 .. nd - 1, then `to` itself, with
 * nd = ceil(JointModelGroup::distance(from, to) / max_segment_length) (:462-472; max_segment_length =
 * longest_valid_segment_fraction * JointModelGroup::getMaximumExtent, :451-460).  `from` is assumed valid, as OMPL does.
 *   var_continuous        [n_group_vars] 1 for the variable of a continuous revolute joint; NULL = none
 *   var_distance_factor   [n_group_vars] JointModel::getDistanceFactor of the variable's joint, 0 for variables that
 *                         are not active (mimic); NULL = all 1
 *   max_states_per_edge   0 = no cap; otherwise an edge is checked at no more than this many states
 *   valid_bitmap          ceil(n_edges / 8) bytes, bit e = 1 iff every checked state of edge e is collision-free
 *   first_invalid         [n_edges] or NULL: j of the first colliding state (1 .. nd, nd = `to`), -1 if the edge is valid
 *   n_states_checked      or NULL: total states the call expanded to and checked
 * OMPL is not vendored in the reference tree (moveit_planners/ompl/package.xml:25, no version pin): the discretisation
 * follows its published algorithm; parity for it is pinned by this repository's own tests only. */
int b200sv_check_motions(b200sv_ctx* ctx, const double* q_from, const double* q_to, size_t n_edges,
                         double max_segment_length, const uint8_t* var_continuous, const double* var_distance_factor,
                         uint32_t max_states_per_edge, uint32_t flags, uint8_t* valid_bitmap, int32_t* first_invalid,
                         uint64_t* n_states_checked);
/* Batched CollisionEnvDistanceField::getCollisionGradients (collision_env_distance_field.cpp:1517-1538), the call
 * CHOMP's optimisation loop makes per waypoint (moveit_planners/chomp/chomp_motion_planner/src/chomp_optimizer.cpp:881-915,
 * acm = nullptr): per collision sphere the smallest distance seen, its gradient and what produced it, in the reference's
 * order -- self field (SELF = 1, :355-428), every sphere pair of every enabled link pair (INTRA = 2, gradient = the
 * un-normalised centre difference, :644-709), world field (ENVIRONMENT = 3, :1645-1681); field distances are
 * sqrt(d2) * resolution with central-difference gradients times the reference's INTEGER inv_twice_resolution_
 * (distance_field.cpp:73-97, distance_field.hpp:614); distances >= max_propagation_distance are not recorded.
 * FP64, the reference's operation order.  The per-link PosedDistanceFields of getSelfProximityGradients (:386-411), which the
 * reference consults when the cache entry's ACM is not empty, are reproduced once b200sv_set_gradient_link_mask switched
 * them on (see "per-link PosedDistanceFields" below).
 *   distances         [n * n_spheres]      DBL_MAX where nothing was recorded (GradientInfo::distances)
 *   gradients         [n * n_spheres * 3]  0 where nothing was recorded (the reference leaves those uninitialised)
 *   types             [n * n_spheres]      CollisionType: NONE 0, SELF 1, INTRA 2, ENVIRONMENT 3
 *   closest_distance  [n * n_glinks] or NULL   GradientInfo::closest_distance per link (field reads only)
 *   sphere_locations  [n * n_spheres * 3] or NULL   posed sphere centres (GradientInfo::sphere_locations) */
int b200sv_collision_gradients(b200sv_ctx* ctx, const double* q, size_t n_states, double* distances, double* gradients,
                               uint8_t* types, double* closest_distance, double* sphere_locations);
/* One reported contact (collision_detection::Contact as the distance-field detector fills it,
 * collision_env_distance_field.cpp:313-331, 596-621, 1604-1623): pos = centre of the colliding sphere of body 1. */
typedef struct b200sv_contact
{
  int32_t body_1;   /* dfce entry (link or attached body) */
  int32_t body_2;   /* the other entry (intra-group), -1 = "self" (self field), -2 = "environment" (world field) */
  int32_t sphere_1; /* index into body 1's sphere list */
  int32_t sphere_2; /* index into body 2's sphere list, -1 for field contacts */
  double pos[3];
} b200sv_contact;
/* Batched CollisionEnvDistanceField::checkCollision with req.contacts = true (:1389-1411): getSelfCollisions ->
 * getIntraGroupCollisions -> getEnvironmentCollisions in that order, each bounded by max_contacts_per_pair per body pair
 * and ending the sequence at max_contacts -- so WHICH contacts are reported is the reference's choice.  `flags` selects
 * the passes (B200SV_CHECK_SELF | INTRA = checkSelfCollision, WORLD = checkRobotCollision).  FP64.
 *   collision      [n]  res.collision
 *   contact_count  [n]  res.contact_count
 *   contacts       [n * max_contacts], the first contact_count[i] of state i are valid */
int b200sv_check_contacts(b200sv_ctx* ctx, const double* q, size_t n_states, uint32_t flags, uint32_t max_contacts,
                          uint32_t max_contacts_per_pair, uint8_t* collision, int32_t* contact_count,
                          b200sv_contact* contacts);
/* Posed collision-sphere centres of ONE state, FP64 kernel: [n_spheres*3] (host).  Parity hook for
 * PosedBodySphereDecomposition::updatePose (collision_distance_field_types.cpp:388-395). */
int b200sv_sphere_centers(b200sv_ctx* ctx, const double* q_one, double* centers);

/* ---- DistanceFieldCacheEntry updates on a live context (collision_env_distance_field.cpp:202-233, 1254-1370) --------- */
/* The reference regenerates its cache entry when the ACM changes (compareCacheEntryToAllowedCollisionMatrix :1314-1370) or a
 * joint outside the group moves (compareCacheEntryToState :1254-1312).  These two re-derive the same things in place -- the
 * enable masks and everything built from them, the constant link transforms, and the self field -- without re-allocating the
 * context or touching its world field.
 *   self_enabled   [n_glinks]            dfce->self_collision_enabled_
 *   intra_enabled  [n_glinks * n_glinks] dfce->intra_group_collision_enabled_ */
int b200sv_set_acm_masks(b200sv_ctx* ctx, const uint8_t* self_enabled, const uint8_t* intra_enabled);
/* base_positions [n_vars]: RobotState::getVariablePositions() for everything outside the group.  rebuild_self_field != 0:
 * the self field is rebuilt from the non-group links' interior points posed at the new state (generateDistanceFieldCacheEntry
 * :907-953) -- the points the URDF path derived, or those given with b200sv_set_self_link_points; 0: the caller rebuilds it
 * (b200sv_field_reset + b200sv_field_add_points on B200SV_FIELD_SELF). */
int b200sv_set_base_state(b200sv_ctx* ctx, const double* base_positions, int rebuild_self_field);
/* Interior points of a link the group does not move, in the LINK frame (BodyDecomposition::getCollisionPoints,
 * collision_distance_field_types.hpp:212-215), for contexts made with b200sv_create. */
int b200sv_set_self_link_points(b200sv_ctx* ctx, int32_t link_index, const double* xyz_link_frame, size_t n_points);

/* ---- field replication (the field is built on ONE GPU and broadcast) --------------------------------- */
/* A built field as one device buffer (occupancy bits, 16-bit d^2, this field's compact elements) that another context with the
 * same geometry imports: between the two calls the bytes can travel by ncclBroadcast (one process per GPU, see
 * moveit2_b200/sharding.py) or a peer copy.  Asynchronous on `cuda_stream`.  Unsigned fields only. */
int b200sv_field_export_size(const b200sv_ctx* ctx, size_t* n_bytes);
int b200sv_field_export_device(b200sv_ctx* ctx, int which, void* d_buf, void* cuda_stream);
int b200sv_field_import_device(b200sv_ctx* ctx, int which, const void* d_buf, void* cuda_stream);

/* ---- several GPUs of one box behind one handle (one process, e.g. a C++ MoveIt node loading the plugin) ----------------- */
/* One context per listed device (a device may be listed twice).  States are sharded in contiguous blocks of whole bitmap
 * words, fields are built on the first device and copied to the others device-to-device, every block's bitmap lands in the
 * caller's buffer: there is no other exchange on this path. */
typedef struct b200sv_multi b200sv_multi;
int b200sv_create_multi(const int* devices, int n_dev, const b200sv_robot* robot, const b200sv_collision_model* model,
                        const b200sv_field_cfg* field, b200sv_multi** out);
int b200sv_create_multi_from_urdf(const int* devices, int n_dev, const char* urdf_xml, const char* srdf_xml,
                                  const char* group_name, const b200sv_field_cfg* field, b200sv_multi** out);
void b200sv_destroy_multi(b200sv_multi* m);
const char* b200sv_multi_last_error(const b200sv_multi* m);
int b200sv_multi_device_count(const b200sv_multi* m);
/* Context i, for every single-context call; after changing a field of context 0, b200sv_multi_field_sync replicates it. */
b200sv_ctx* b200sv_multi_ctx(b200sv_multi* m, int i);
int b200sv_multi_field_sync(b200sv_multi* m, int which);
int b200sv_multi_field_reset(b200sv_multi* m, int which);
/* b200sv_field_set_propagation for the handle: the first device builds every field (in the chosen mode), the others receive
 * its values.  b200sv_multi_field_add_points then yields the reference's voxels on every device. */
int b200sv_multi_field_set_propagation(b200sv_multi* m, int which, int mode);
int b200sv_multi_field_add_points(b200sv_multi* m, int which, const double* xyz, size_t n_points);
/* b200sv_check_batch over all devices: block i of the states goes to context i; pinned host buffers make every copy a DMA. */
int b200sv_check_batch_multi(b200sv_multi* m, const double* q, size_t n_states, uint32_t flags, uint8_t* valid_bitmap);
int b200sv_multi_get_stats(b200sv_multi* m, b200sv_stats* stats);

/* ---- per-link PosedDistanceFields of the gradient path ---------------------------------------------------------------- */
/* getSelfProximityGradients (collision_env_distance_field.cpp:386-411) also consults, for every entry i, the
 * PosedDistanceField of every OTHER link j whose name differs and for which the cache entry's ACM has no entry other than
 * NEVER -- but only when that ACM is not empty (CHOMP creates its GroupStateRepresentation with the scene's ACM,
 * chomp_optimizer.cpp:102, so for CHOMP it is).  A link's field is a PropagationDistanceField of the size of its bounding
 * sphere's diameter, cornered at the sphere's centre - size / 2 IN THE LINK FRAME, filled with the link's interior points
 * (getGroupStateRepresentation :1177-1197).  The query is PosedDistanceField::getDistanceGradient exactly as written
 * (collision_distance_field_types.hpp:163-176): the point is rotated by the pose's rotation transposed WITHOUT subtracting
 * the translation, the gradient comes back as pose * g WITH the translation added; consequently an out-of-bounds query ends
 * the entry's sphere loop (types.cpp:104).  b200sv_collision_gradients reproduces that when a mask is set.
 *   b200sv_set_link_field        the field of entry `glink` as the caller built it (a MoveIt workspace: the reference's own
 *                                PosedDistanceField, voxel by voxel): size and corner as above, d2 [nx * ny * nz] squared cell
 *                                distances with n = int(size / resolution); d2 = NULL: the link has none
 *   b200sv_build_link_fields     URDF path: every entry's field built on the device (exact EDT of the link's interior points)
 *   b200sv_get_link_field        dims (0 0 0: none), corner, and -- if d2 is not NULL -- the voxels
 *   b200sv_set_gradient_link_mask  enabled [n_glinks * n_glinks]: entry i consults link j's field; NULL = the ACM is empty,
 *                                no field is consulted (the default) */
int b200sv_set_link_field(b200sv_ctx* ctx, int32_t glink, const double* size, const double* corner, const int32_t* d2);
int b200sv_build_link_fields(b200sv_ctx* ctx);
int b200sv_get_link_field(b200sv_ctx* ctx, int32_t glink, int32_t* dims, double* corner, int32_t* d2);
int b200sv_set_gradient_link_mask(b200sv_ctx* ctx, const uint8_t* enabled);

/* ---- the FP32 error budget -------------------------------------------------------------------------------------------- */
/* The fast pass decides in FP32; every decision carries a budget eps (metres) and whatever could flip within it is decided
 * by the exact FP64 pass.  eps_derived is the worst-case FP32 error of a posed sphere centre, derived from the model (chain
 * depth, joint origins, joint ranges, field extent: see deriveFp32Eps in csrc/api.cu); eps_used is what the tables were built
 * with (= eps_derived unless b200sv_set_tuning or B200SV_FP32_EPS fixed it). */
int b200sv_get_fp32_budget(const b200sv_ctx* ctx, double* eps_used_m, double* eps_derived_m);
/* The probe that holds the budget to account: the fast kernel's own FK and sphere posing (same device function), no checks;
 * centres [n * n_spheres * 3] in metres.  tests/ compares them with the FP64 centres (b200sv_sphere_centers). */
int b200sv_debug_fast_centres(b200sv_ctx* ctx, const double* q, size_t n_states, double* centres);

/* ---- measurement helpers ------------------------------------------------------------------------------------ */
/* Random 32-byte-sector gather over the world field's footprint: the "gather-bandwidth roofline" the check
 * kernel is compared against.  Reports sectors/s * 32 B as GB/s. */
int b200sv_gather_roofline(b200sv_ctx* ctx, size_t n_reads, int iters, double* gbytes_per_s);
/* Tuning knobs (0 keeps the default): FP32 sphere-centre error bound in metres, warps per CTA cap, CTAs per SM. */
int b200sv_set_tuning(b200sv_ctx* ctx, double fp32_eps_m, int warps_per_block, int blocks_per_sm);

#ifdef __cplusplus
}
#endif
#endif /* B200SV_H */
