// Launchers of the CUDA kernels (definitions in check_kernels.cu / edt_kernels.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#ifndef B200SV_SPHERE_CHUNK
#define B200SV_SPHERE_CHUNK 4  // spheres per unrolled pass of the field loop; the sphere table is padded to a multiple
#endif

#ifndef B200SV_FAST_UNROLL
#define B200SV_FAST_UNROLL 2  // spheres per unrolled chunk of the fast pass's field loop; Hot ranges are padded per link
#endif

namespace b200sv
{
template <typename FT>
struct CheckArgs
{
  const uint8_t* model;  // ModelHeader<R> blob (float blob for the fast pass, double blob for the exact pass)
  int model_bytes;
  const double* q;       // [n_states * n_group_vars]
  unsigned long long n_states;
  unsigned long long state_base;  // index of q[0] within the caller's batch (chunked launches); list entries are global
  const void* field;     // interleaved compact d^2: element [voxel] = world | self << (8 * sizeof(FT)),
                         // voxel = x*ny*nz + y*nz + z; one gather serves both checks
  uint32_t flags;
  uint32_t* bitmap;      // validity words
  uint32_t* flag_list;   // fast pass: states to recheck (out)
  unsigned* flag_count;
  unsigned list_cap;
  const uint32_t* list;  // exact pass from list (in)
  const unsigned* list_count;
  float* t_scratch;      // fast pass only: global-memory home of the link transforms (null: shared memory)
  uint8_t* rep_scratch;  // fast pass, hierarchical cull only: per-warp home of the CLUSTER-level parked centres (halves)
  int hier;              // fast pass only: the blob carries LinkPair records (hierarchical cull)
  int n_group_vars;      // fast pass only: selects the kernel instance (register prefetch of the states)
  int wide;              // fast pass only: the wide instance (more than 16 variables, or a planar joint in the group)
};

struct FkArgs
{
  const uint8_t* model;
  int model_bytes;
  const double* q;
  unsigned long long n_states;
  int n_model_links;
  const int* link_slot;   // [n_model_links] device-link slot or -1
  const double* const_T;  // [n_model_links*16] base-state global transforms (column-major)
  double* out;            // [n_states * n_model_links * 16]
};

template <typename FT>
cudaError_t launchCheck(const CheckArgs<FT>& a, bool fp64, bool from_list, int grid, int block, size_t smem_bytes,
                        cudaStream_t stream);
// exact pass over the recheck list, one warp per state (latency of a short list)
template <typename FT>
cudaError_t launchRecheckWarp(const CheckArgs<FT>& a, int grid, int state_stride, cudaStream_t stream);
// the latency path: up to 32 states, exact FP64 check, one warp per state, results as bytes in mapped host memory
template <typename FT>
struct LatencyArgs
{
  const uint8_t* model;  // double blob
  int model_bytes;
  const void* field;
  uint32_t flags;
  int n_states;
  int state_stride;      // doubles of link transforms per state (the double blob's)
  int n_links;           // device links (the double blob's): sin / cos scratch per warp
  int use_inline;        // the states are q_inline (n_states * n_group_vars <= 32 doubles); else q
  const double* q;       // device-visible (mapped pinned host) states
  uint8_t* result;       // device-visible (mapped pinned host) [n_states]: 1 = valid, 0 = in collision
  double q_inline[32];
};
template <typename FT>
cudaError_t launchLatencyCheck(const LatencyArgs<FT>& a, cudaStream_t stream);

// The same check for ONE state at a time, by a kernel that stays resident for a short while after it has answered: a planner
// validates states back to back (OMPL's per-state isValid), and every call after the first finds the kernel already
// running -- no launch, the request and the answer travel through a mailbox in mapped pinned host memory.
struct LingerMailbox  // mapped pinned host memory, 16-byte aligned
{
  unsigned long long req;   // host -> kernel: (sequence number << 3) | check flags of the newest request
  unsigned long long quit;  // host -> kernel: the epoch of the kernel that should exit now
  unsigned long long done;  // kernel -> host: (sequence number << 8) | result (1 = valid, 0 = in collision)
  unsigned long long pad;
  double q[32];             // the request's state
};
// A few edges (an OMPL MotionValidator validates ONE per checkMotion): the host sizes the batch itself, one CTA per visited
// state interpolates its own state from the edge's ends and runs the exact check; result bytes polled in mapped host memory.
constexpr int kEdgeLatencyEdges = 64;
template <typename FT>
struct EdgeLatencyArgs
{
  const uint8_t* model;  // double blob
  int model_bytes;
  const void* field;
  uint32_t flags;
  int state_stride, n_links, n_edges, n_vars;
  uint32_t cont_mask;        // bit k: group variable k is a continuous revolute joint (interpolated along the shorter arc)
  int use_inline;            // one edge: its ends are ends_inline (from, then to); else `ends`
  const double* ends;        // device-visible (mapped pinned host) [2 * n_edges * n_vars]: all from rows, then all to rows
  uint8_t* result;           // device-visible (mapped pinned host) [offsets[n_edges]]: 1 = valid, 0 = in collision
  uint32_t offsets[kEdgeLatencyEdges + 1];  // exclusive scan of the edges' state counts
  double ends_inline[64];
};
template <typename FT>
cudaError_t launchEdgeLatencyCheck(const EdgeLatencyArgs<FT>& a, cudaStream_t stream);

template <typename FT>
struct LingerArgs
{
  const uint8_t* model;  // double blob
  int model_bytes;
  const void* field;
  int state_stride, n_links;
  LingerMailbox* mail;               // device-visible
  unsigned long long first_req;      // the request this launch answers first (its state is q_inline)
  unsigned long long epoch;          // exit when mail->quit == epoch
  unsigned long long linger_ns;      // exit this long after the last answer
  unsigned long long life_ns;        // exit this long after the launch, whatever happens
  double q_inline[32];
};
template <typename FT>
cudaError_t launchLingerCheck(const LingerArgs<FT>& a, cudaStream_t stream);
int checkKernelRegs(bool fp64, bool from_list, int field_bytes);
int checkPerWarpBytes(bool fp64, int state_stride);  // shared memory one warp of the check kernel needs
// One link's PosedDistanceField on the device: 16-bit d^2 on its own small grid in the LINK frame
struct LinkFieldDev
{
  const uint16_t* d2;  // nullptr: the link has no field
  int n[3];
  double om[3];        // VoxelGrid origin_minus_ = corner - resolution / 2
};
// K5: per-sphere distance / gradient / type (getCollisionGradients), FP64, exact blob
struct GradArgs
{
  const uint8_t* model;  // double blob
  int model_bytes;
  const double* q;
  unsigned long long n_states;
  const uint16_t* d2_world;  // min(d^2, max_d2) per voxel
  const uint16_t* d2_self;
  const uint16_t* nd2_world;  // signed fields: negative_distance_square_ per voxel; nullptr for unsigned fields
  const uint16_t* nd2_self;
  const uint8_t* self_enabled;  // [n_glinks]
  int n_glinks, n_spheres, state_stride;  // state_stride: doubles of link transforms per state (the double blob's)
  // per-link PosedDistanceFields (GroupStateRepresentation::link_distance_fields_): consulted for entry i and link j where
  // lf_enabled[i * n_glinks + j]; lf == nullptr: the cache entry's ACM is empty, none is consulted
  const struct LinkFieldDev* lf;   // [n_glinks]
  const uint8_t* lf_enabled;       // [n_glinks * n_glinks]
  const int32_t* sphere_start;     // [n_glinks + 1]
  const int32_t* glink_slot;       // [n_glinks] device-link slot of each entry's link
  const double* sqrt_table;  // [max_d2 + 1]: sqrt(double(i)) * resolution (PropagationDistanceField::sqrt_table_)
  double resolution, inv_twice_resolution, max_distance, max_propagation;
  double* distances;  // [n * n_spheres]
  double* gradients;  // [n * n_spheres * 3]
  uint8_t* types;     // [n * n_spheres]
  double* closest;    // [n * n_glinks]
  double* centers;    // [n * n_spheres * 3] scratch / output: posed sphere centres (GradientInfo::sphere_locations)
};
cudaError_t launchGradients(const GradArgs& a, int grid, int block, size_t smem_bytes, cudaStream_t stream);

// contacts (checkCollision with req.contacts), FP64, exact blob
struct b200sv_contact_rec  // = b200sv_contact of include/b200sv.h
{
  int32_t body_1, body_2, sphere_1, sphere_2;
  double pos[3];
};
struct ContactArgs
{
  const uint8_t* model;
  int model_bytes;
  const double* q;
  unsigned long long n_states;
  const uint16_t* d2_world;
  const uint16_t* d2_self;
  uint32_t flags;
  unsigned max_contacts, max_per_pair;
  int n_glinks, n_spheres, state_stride;  // state_stride: doubles of link transforms per state (the double blob's)
  const int32_t* sphere_start;   // [n_glinks + 1]
  const int32_t* glink_slot;     // [n_glinks] device-link slot
  const double* glink_bs;        // [n_glinks * 4] link-frame bounding sphere
  const int32_t* body_id;        // [n_glinks] -1 for links, shapes of one attached body share an id
  const uint8_t* self_enabled;   // [n_glinks]
  const uint8_t* intra;          // [n_glinks * n_glinks]
  double* centers;               // [n * n_spheres * 3] scratch
  uint8_t* collision;            // [n]
  int32_t* contact_count;        // [n]
  b200sv_contact_rec* contacts;  // [n * max_contacts]
};
cudaError_t launchContacts(const ContactArgs& a, int grid, int block, size_t smem_bytes, cudaStream_t stream);

// fast_check.cu: the FP32 pass for groups of fixed / revolute / prismatic joints (a.model = the fast blob)
template <typename FT>
cudaError_t launchFastCheck(const CheckArgs<FT>& a, int grid, int block, size_t smem_bytes, cudaStream_t stream);
// FP32 error-budget probe: the fast kernel's FK + sphere posing alone, every Hot record's centre in cells
cudaError_t launchFastCentres(const uint8_t* model, int model_bytes, const double* q, unsigned long long n, int n_group_vars,
                              bool wide, float* scratch, float* out, int n_hot, cudaStream_t stream);
int fastCheckKernelRegs(int field_bytes, int n_group_vars);
int fastCheckPerWarpBytes(int state_stride, int n_clusters, int n_group_vars, int n_linkpairs, bool wide);
int fastCheckRepBytes(int n_rows);  // bytes of one warp's global scratch of parked cluster centres (hierarchical cull)
int fastCheckMaxVars();
int fastCheckRegisterPrefetchVars();
int fastCheckMaxWarps(int n_group_vars);  // __launch_bounds__ of the fast kernel instance that takes the group
int fastCheckUnroll();
cudaError_t launchFk(const FkArgs& a, bool fp64, int grid, int block, size_t smem_bytes, cudaStream_t stream);
cudaError_t launchSphereCenters(const FkArgs& a, double* d_centers, size_t smem_bytes, cudaStream_t stream);
cudaError_t launchGather(const uint8_t* field, unsigned long long n_sectors, int per_thread, int grid, int block,
                         unsigned* sink, cudaStream_t stream);

// ---- batched edge validation (motion_kernels.cu) ------------------------------------------------------------
struct MotionArgs
{
  const double* q_from;  // [n_edges * n_vars] device
  const double* q_to;
  size_t n_edges;
  int n_vars;
  const uint8_t* continuous;  // [n_vars] or null
  const double* factor;       // [n_vars] JointModel::getDistanceFactor, 0 for variables of mimic joints; or null (all 1)
  double max_segment;         // longest valid segment = fraction * maximum extent
  unsigned max_states_per_edge;
  unsigned* count;            // [n_edges] states checked per edge
  unsigned long long* offset; // [n_edges + 1] exclusive scan of count
};
cudaError_t launchMotionCount(const MotionArgs& a, cudaStream_t s);  // count + scan
cudaError_t launchMotionExpand(const MotionArgs& a, double* d_q, unsigned long long n_states, cudaStream_t s);
cudaError_t launchMotionReduce(const MotionArgs& a, const uint32_t* state_bits, uint32_t* edge_bits, int32_t* first_invalid,
                               cudaStream_t s);

// ---- distance field ------------------------------------------------------------------------------------
struct GridDev
{
  double om[3];  // VoxelGrid::origin_minus_
  double oo;     // VoxelGrid::oo_resolution_
  int nx, ny, nz;
  int wz;        // 32-bit occupancy words per z row
  int R;         // ceil(max_distance / resolution): propagation radius in cells
  int max_d2;    // R*R = PropagationDistanceField::max_distance_sq_
};

// A bodies::Body for bodyLatticeKernel: type = shapes::ShapeType (1 sphere, 2 cylinder, 4 box); c, R (row-major) = the pose
// the lattice is tested against; dims = { r^2 } / { r^2, half length } / { half extents }; posed: kept points are moved by
// P = [R row-major (9) | t (3)] before their voxel is looked up.
struct BodyDev
{
  int type, posed;
  double c[3], R[9], dims[3], P[12];
};

// occupancy bitmap: word ((x*ny + y)*wz + z/32), bit z%32
cudaError_t launchScatter(const double* d_xyz, size_t n, const GridDev& g, uint32_t* occ, bool set, cudaStream_t s);
// addShapeToField for a sphere: lattice axes (host-accumulated) x containsPoint -> occupancy bits
cudaError_t launchSphereShape(const double* d_axes, int nxs, int nys, int nzs, const double* c, double r2, const GridDev& g,
                              uint32_t* occ, cudaStream_t s);
// occ' = (occ - (old \\ new)) + (new \\ old): updatePointsInField on bitmaps of the same layout
cudaError_t launchBodyLattice(const double* d_axes, int nxs, int nys, int nzs, const BodyDev& body, const GridDev& g,
                              uint32_t* occ, bool set, cudaStream_t s);
cudaError_t launchUpdateOcc(uint32_t* occ, const uint32_t* old_bits, const uint32_t* new_bits, size_t n_words, cudaStream_t s);
// bodies::ConvexMesh (geometric_shapes): interior lattice points of a convex polytope given by its planes.
//   d_axes   lattice axis values (host-accumulated), d_planes [n_planes][4] = unit normal, w (inside: n.p + w - 1e-9 <= 0)
//   ipose    the body's inverse pose, rows of [R^T | -R^T t] (i_pose_ * p); post/posed as in BodyDev
struct ConvexDev
{
  int n_planes, posed;
  double ipose[12], P[12];
};
cudaError_t launchConvexLattice(const double* d_axes, int nxs, int nys, int nzs, const double* d_planes, const ConvexDev& body,
                                const GridDev& g, uint32_t* occ, bool set, cudaStream_t s);
// DistanceField::getOcTreePoints (distance_field.cpp:240-284) for leaves the caller enumerated: d_leaves [n][4] = centre, size
cudaError_t launchOctreeLeaves(const double* d_leaves, size_t n, double resolution, const GridDev& g, uint32_t* occ,
                               cudaStream_t s);
// exact truncated EDT of the occupancy bitmap -> d2 (uint16, min(d^2, max_d2)) and this field's half of the
// INTERLEAVED compact field (element 2 * voxel; `compact` arrives offset by 0 = world / 1 = self; uint8:
// min(d^2, 255), uint16: d^2).  tmp: uint16 [nx*ny*nz] scratch.  invert: the transform of the complement (the
// negative half of a signed field; pass compact = nullptr, it has no compact copy).
constexpr int kPlaneBusyStride = 32;  // bytes between two planes' flags (one L2 sector each)
// own_busy / other_busy: [nx * kPlaneBusyStride] flags per x plane (the first byte of each stride), nonzero = this / the OTHER field's compact elements of that plane are not
// all "free" (nullptr: unknown).  The x pass clears and sets own_busy, and where the other field's plane is free it writes
// whole {world, self} words instead of merging into them (no read of the interleaved field).
cudaError_t launchEdt(const GridDev& g, const uint32_t* occ, uint16_t* tmp, uint16_t* d2, void* compact,
                      int compact_bytes, int sm_count, cudaStream_t s, bool invert = false, uint8_t* own_busy = nullptr,
                      const uint8_t* other_busy = nullptr);
cudaError_t launchInjectNegD2(const GridDev& g, const int32_t* d_in, uint16_t* nd2, cudaStream_t s);
cudaError_t launchFillField(const GridDev& g, uint16_t* d2, void* compact, int compact_bytes, cudaStream_t s);
cudaError_t launchInjectD2(const GridDev& g, const int32_t* d_in, uint32_t* occ, uint16_t* d2, void* compact,
                           int compact_bytes, cudaStream_t s);
// ---- the reference's own propagation, reproduced phase-parallel (propagation.cu) -----------------------------------------
// A PropField holds PropagationDistanceField's per-voxel record (distance_square_, closest_point_, update_direction_) and its
// bucket queue for one field of a context; every call below leaves the squared distances in propDistances() (int32, one per
// voxel, voxel_grid.hpp:425 order), ready for launchInjectD2.  All calls synchronise `s` many times.
struct PropField;
struct PropStats
{
  unsigned long long rounds, hazard_rounds, entries, backward_offers, hazard_offers, queued_for_next_call, flood_pops_batches;
};
cudaError_t propCreate(const GridDev& g, bool is_signed, PropField** out);  // is_signed: propagate_negative_ (:48-58)
void propDestroy(PropField* f);
cudaError_t propReset(PropField* f, cudaStream_t s);                                          // reset() :506-524
cudaError_t propAddPoints(PropField* f, const double* d_xyz, size_t n, cudaStream_t s);      // addPointsToField :177-196
cudaError_t propRemovePoints(PropField* f, const double* d_xyz, size_t n, cudaStream_t s);   // removePointsFromField :198-219
cudaError_t propUpdatePoints(PropField* f, const double* d_old, size_t n_old, const double* d_new, size_t n_new,
                             cudaStream_t s);                                                // updatePointsInField :130-175
// readFromStream's tail (:730-758): every set bit of an occupancy bitmap (launchScatter's layout), in voxel order
cudaError_t propAddOccupancy(PropField* f, const uint32_t* d_occ, cudaStream_t s);
const int32_t* propDistances(const PropField* f);
const int32_t* propNegativeDistances(const PropField* f);  // negative_distance_square_ (signed fields), else nullptr
const PropStats& propStats(const PropField* f);
// One field's elements of the interleaved compact field <-> a dense array (field replication across contexts):
// `compact` is already offset to this field's first element; extract = to the dense array, else from it.
cudaError_t launchCompactCopy(void* compact, void* dense, size_t n_voxels, int compact_bytes, bool extract, cudaStream_t s);
// float states -> double states (b200sv_check_batch_f32)
cudaError_t launchWidenStates(const float* in, double* out, size_t n, cudaStream_t s);
}  // namespace b200sv
