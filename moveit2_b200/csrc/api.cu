// The C-ABI of libb200sv.so (include/b200sv.h): context, device tables, field management, batch entry points.
// Host logic only -- kernels live in check_kernels.cu and edt_kernels.cu.  There is no CPU compute path here:
// every entry point either runs on the CUDA device or returns an error.
#include <cuda_runtime.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200sv.h"
#include "device_model.hpp"
#include "host_model.hpp"
#include "kernels.hpp"

using namespace b200sv;

namespace
{
thread_local std::string g_create_error;

struct Field
{
  uint32_t* occ = nullptr;    // occupancy bits
  uint16_t* d2 = nullptr;     // min(d^2, max_d2)
  uint16_t* nd2 = nullptr;    // signed fields only: negative_distance_square_ (distance to the nearest free voxel)
  void* compact = nullptr;    // this field's half of the interleaved compact field (base + which elements)
  uint8_t* plane_busy = nullptr;  // [nx] nonzero: plane x of this field's compact elements may hold a non-free value
  bool dirty = false;         // occupancy changed inside a begin/end_update bracket: the EDT is owed
  bool deferred = false;      // inside b200sv_field_begin_update .. b200sv_field_end_update
  PropField* prop = nullptr;  // B200SV_PROPAGATION_REFERENCE: the reference's voxel records and bucket queue (propagation.cu)
};

template <typename T>
struct DevBuf
{
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n)
  {
    if (n <= cap)
      return cudaSuccess;
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = std::max<size_t>(n, 1024);
    cudaError_t e = cudaMalloc(&p, want * sizeof(T));
    if (e == cudaSuccess)
      cap = want;
    return e;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};
}  // namespace

struct b200sv_ctx
{
  int device = 0;
  std::mutex mu;
  mutable std::string err;
  HostRobot robot;
  HostCollisionModel cm;
  b200sv_field_cfg cfg{};
  std::vector<double> self_points;
  SelfLocalPoints self_local;      // URDF path: link-frame interior points of the links the group does not move
  // grid
  GridDev grid{};
  int compact_bytes = 1;
  size_t n_voxels = 0;
  // device tables
  std::vector<uint8_t> blob_f, blob_d, blob_fast;
  uint8_t *d_blob_f = nullptr, *d_blob_d = nullptr, *d_blob_fast = nullptr;
  bool fast_ok = false;      // the group's joints are all FIXED / REVOLUTE / PRISMATIC: fast_check.cu runs the FP32 pass
  bool fast_disabled = false;  // B200SV_FAST=0: force the generic FP32 kernel (A/B measurements)
  int state_stride_fast = 0, fast_t_rows = 0, fast_n_linkpairs = 0, fast_n_reps = 0, fast_n_reps_global = 0;
  uint8_t* d_rep_scratch = nullptr;  // hierarchical cull: per resident warp, the cluster-level parked centres
  bool fast_t_global = false;  // link transforms in an L2-resident scratch instead of shared memory (large robots)
  int fast_smem_blob_bytes = 0;  // bytes of the fast blob the kernel copies to shared memory
  std::vector<int> fast_hot_sphere;  // fast blob: Hot record -> sphere (dfce order), -1 = padding
  double fp32_eps_derived = 0.0;  // worst-case FP32 error of a posed sphere centre (m), from the model (deriveFp32Eps)
  std::vector<double> fp32_link_eps;  // the same per device link (the fast kernel's cell-face margin is per link)
  bool fp32_eps_user = false;     // b200sv_set_tuning fixed the budget
  bool fast_wide = false;        // the group takes the wide instance of the fast kernel (> 16 variables or a planar joint)
  float* d_t_scratch = nullptr;
  int n_dev_links = 0, n_spheres = 0, n_bs = 0, n_pairs = 0, n_link_pairs = 0;
  double cluster_radius = 0.2;   // max tight-sphere radius of a sphere cluster (m); clusters also hold <= 4 spheres
  int state_stride_f = 0, state_stride_d = 0;  // shared-memory words per state (float / double blob)
  std::vector<int> link_slot;      // model link -> device slot
  std::vector<int> glink_slot;     // dfce entry -> device slot (attached bodies have pseudo links)
  std::vector<double> const_T;     // model links * 16 (base state)
  int* d_link_slot = nullptr;
  double* d_const_T = nullptr;
  Field fields[2];
  uint16_t* d_tmp = nullptr;  // EDT intermediate
  double* d_sqrt_table = nullptr;  // sqrt_table_ of the reference's field (gradient kernel)
  uint8_t* h_small = nullptr;      // pinned staging of the small-batch (latency) path of b200sv_check_batch
  uint32_t* d_small = nullptr;     // its device-side result: 128 bitmap words, then the recheck counter
  std::vector<double> edge_factor_cached;  // what d_edge_aux / d_edge_cont hold (b200sv_check_motions re-uploads on change)
  std::vector<uint8_t> edge_cont_cached;
  size_t h_small_bytes = 0;
  void* d_combined = nullptr; // interleaved compact field {world, self} per voxel
  // launch config
  int sm_count = 148, smem_optin = 0;
  struct LaunchCfg
  {
    int warps = 0, blocks_per_sm = 1;
    size_t smem = 0;
  };
  LaunchCfg cfg_f, cfg_d, cfg_list;  // generic FP32 pass, exact pass over all states, exact pass over the recheck list
  LaunchCfg cfg_fast;                // fast_check.cu
  int tune_warps = 0, tune_blocks = 0;
  double fp32_eps = 4e-6;
  // streams / staging
  static constexpr int kMaxChunks = 32;
  cudaStream_t stream = nullptr, copy_stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_go = nullptr;
  // The *_device entry points run on the CALLER's stream but use the context's scratch (recheck list and counter, transform
  // scratch, EDT intermediate, the fields themselves): ev_user marks the end of the last such call, and whichever stream
  // touches the context next waits for it first (enterCtx / enterUser).
  cudaEvent_t ev_user = nullptr;
  cudaStream_t user_stream = nullptr;
  bool user_pending = false;
  cudaEvent_t ev_chunk[kMaxChunks] = { nullptr };
  DevBuf<double> d_q;
  DevBuf<float> d_qf;              // b200sv_check_batch_f32: the states as they arrive
  DevBuf<uint32_t> d_bitmap, d_list;
  DevBuf<double> d_points;
  DevBuf<double> d_fk;
  DevBuf<int32_t> d_i32;
  // batched edge validation
  DevBuf<double> d_edges, d_edge_q, d_edge_aux;
  DevBuf<unsigned> d_edge_count;
  DevBuf<unsigned long long> d_edge_offset;
  DevBuf<uint32_t> d_edge_state_bits, d_edge_bits;
  DevBuf<int32_t> d_edge_first;
  DevBuf<uint8_t> d_edge_cont;
  // per-link PosedDistanceFields of the gradient path (GroupStateRepresentation::link_distance_fields_)
  struct LinkField
  {
    uint16_t* d2 = nullptr;
    int n[3] = { 0, 0, 0 };
    double corner[3] = { 0, 0, 0 };
  };
  std::vector<LinkField> link_fields;        // [n_glinks]
  std::vector<uint8_t> lf_enabled;           // [n_glinks^2]; empty: the ACM is empty, the fields are not consulted
  DevBuf<uint8_t> d_lf_enabled;
  LinkFieldDev* d_lf = nullptr;
  bool lf_dirty = true;
  SelfLocalPoints link_points;               // URDF path: link-frame interior points of every link with geometry
  // K5 outputs
  DevBuf<double> d_g_dist, d_g_grad, d_g_closest, d_g_centers;
  DevBuf<uint8_t> d_g_types, d_g_self;
  // contacts
  DevBuf<int32_t> d_c_start, d_c_slot, d_c_count, d_c_body;
  DevBuf<double> d_c_bs;
  DevBuf<uint8_t> d_c_intra, d_c_coll;
  DevBuf<b200sv_contact_rec> d_c_out;
  unsigned* d_count = nullptr;
  unsigned* d_sink = nullptr;
  unsigned* h_count = nullptr;     // pinned: the recheck counter of the last large batch
  uint8_t* h_lat = nullptr;        // mapped pinned memory of the latency path: states, then result bytes
  uint8_t* d_lat = nullptr;        // its device-side address
  bool latency_ok = false;         // the exact blob + 4 warps of transforms fit shared memory, V <= 32
  // single states back to back: a kernel that stays resident for a short while and is fed through a mailbox (checkLinger)
  uint8_t* h_edge_lat = nullptr;    // mapped pinned memory of the few-edges path: the edges' ends, then result bytes
  uint8_t* d_edge_lat = nullptr;
  LingerMailbox* h_mail = nullptr;  // mapped pinned
  LingerMailbox* d_mail = nullptr;
  unsigned long long mail_seq = 0, linger_epoch = 0;
  bool linger_maybe_alive = false;
  std::chrono::steady_clock::time_point linger_last{}, linger_launched{};
  b200sv_stats stats{};
  bool stats_pending = false;
};

struct b200sv_multi
{
  std::vector<b200sv_ctx*> ctx;
  std::mutex mu;
  std::string err;
  b200sv_stats stats{};
};

namespace
{
int fail(b200sv_ctx* c, int code, const std::string& msg)
{
  if (c)
    c->err = msg;
  else
    g_create_error = msg;
  return code;
}

// Compute entry points call this first: a host-only context has no device and there is no CPU path to fall back to.
int needDevice(b200sv_ctx* c)
{
  if (c->device < 0)
    return fail(c, B200SV_ERR_NO_DEVICE, "host-only context (B200SV_DEVICE_NONE): no CUDA device, and no CPU fallback exists");
  return B200SV_OK;
}

#define CU(ctx, expr)                                                                                              \
  do                                                                                                               \
  {                                                                                                                \
    cudaError_t e__ = (expr);                                                                                      \
    if (e__ != cudaSuccess)                                                                                        \
      return fail(ctx, B200SV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));                      \
  } while (0)

inline int align16(int x) { return (x + 15) & ~15; }

// Every entry point that works on the context's own stream: select the device and order the stream behind the last call that
// ran on a caller's stream (it may still be using the context's scratch buffers or rewriting a field).
// A resident single-state kernel (checkLinger) sits on the context's stream: everything else asks it to leave first.  The
// kernel sees the word within a PCIe round trip; stream order does the waiting.
void quitLinger(b200sv_ctx* c)
{
  if (c->linger_maybe_alive)
  {
    *reinterpret_cast<volatile unsigned long long*>(&c->h_mail->quit) = c->linger_epoch;
    c->linger_maybe_alive = false;
  }
}

void enterCtx(b200sv_ctx* c, bool keep_linger = false)
{
  cudaSetDevice(c->device);
  if (!keep_linger)
    quitLinger(c);
  if (c->user_pending)
  {
    cudaStreamWaitEvent(c->stream, c->ev_user, 0);
    c->user_pending = false;
  }
}

// Entry points that run on the caller's stream `s`: behind the context's own stream (a field rebuild or upload may be in
// flight there) and behind the last call on a DIFFERENT caller stream.
int enterUser(b200sv_ctx* c, cudaStream_t s)
{
  cudaSetDevice(c->device);
  quitLinger(c);
  if (c->user_pending && c->user_stream != s)
    CU(c, cudaStreamWaitEvent(s, c->ev_user, 0));
  CU(c, cudaEventRecord(c->ev_go, c->stream));
  CU(c, cudaStreamWaitEvent(s, c->ev_go, 0));
  return B200SV_OK;
}

int leaveUser(b200sv_ctx* c, cudaStream_t s)
{
  CU(c, cudaEventRecord(c->ev_user, s));
  c->user_stream = s;
  c->user_pending = true;
  return B200SV_OK;
}

// Integer form of the reference predicate (collision_distance_field_types.cpp:229):
//   collide = (maximum_value > dist) && (radius - dist > tolerance),  dist = sqrt_table[d2] (unsigned field)
// with sqrt_table[i] = sqrt(double(i)) * resolution (propagation_distance_field.cpp:99-101).  dist grows with d2,
// so the predicate is true exactly for d2 < T.  T in [0, max_d2 + 1].
// Signed fields: an occupied voxel (d2 = 0) carries dist = -sqrt(neg d2) * res with neg d2 in [1, max_d2]; free voxels are as
// in the unsigned field.  The integer form still exists when the predicate gives one answer for every such neg d2 (it does
// unless tolerance exceeds radius + resolution); -1 = it does not.
int sphereThreshold(double radius, double res, double max_prop, double tol, int max_d2, bool signed_field)
{
  if (signed_field)
  {
    auto hit = [&](int nd2) {
      const double dist = sqrt(double(0)) * res - sqrt(double(nd2)) * res;
      return (max_prop > dist) && (radius - dist > tol);
    };
    if (hit(1) != hit(max_d2))
      return -1;
    if (!hit(1))
      return 0;  // radius - dist only shrinks as d2 grows
  }
  for (int i = signed_field ? 1 : 0; i <= max_d2; ++i)
  {
    const double dist = sqrt(double(i)) * res - sqrt(double(0)) * res;  // getDistance(): sqrt_table[d2] - sqrt_table[neg d2 = 0]
    if (!((max_prop > dist) && (radius - dist > tol)))
      return i;
  }
  return max_d2 + 1;
}

template <typename R>
void buildBlob(const b200sv_ctx& c, const std::vector<LinkRec<double>>& links, const std::vector<VarRec>& vars,
               const std::vector<SphereRec<double>>& spheres, const std::vector<BsRec<double>>& bs,
               const std::vector<PairRec>& pairs, const std::vector<GroupRec>& groups, const std::vector<AltRec>& alts,
               double margin, double pair_eps, double bs_eps, double tight_eps, std::vector<uint8_t>& out)
{
  ModelHeader<R> h{};
  h.n_links = (int)links.size();
  h.n_vars = (int)vars.size();
  h.n_spheres = c.n_spheres;
  h.n_spheres_padded = (int)spheres.size();
  h.n_groups = (int)groups.size();
  h.n_bs = (int)bs.size();
  h.n_pairs = (int)pairs.size();
  h.n_group_vars = (int)c.robot.group_var_index.size();
  int off = align16(sizeof(ModelHeader<R>));
  h.off_links = off;
  off = align16(off + (int)(links.size() * sizeof(LinkRec<R>)));
  h.off_vars = off;
  off = align16(off + (int)(vars.size() * sizeof(VarRec)));
  h.off_spheres = off;
  off = align16(off + (int)(spheres.size() * sizeof(SphereRec<R>)));
  h.off_bs = off;
  off = align16(off + (int)(bs.size() * sizeof(BsRec<R>)));
  h.off_pairs = off;
  off = align16(off + (int)(pairs.size() * sizeof(PairRec)));
  h.off_groups = off;
  off = align16(off + (int)(groups.size() * sizeof(GroupRec)));
  h.off_alts = off;
  h.n_alts = (int)alts.size();
  off = align16(off + (int)(alts.size() * sizeof(AltRec)));
  h.total_bytes = off;
  h.state_stride = sizeof(R) == 4 ? c.state_stride_f : c.state_stride_d;
  h.link_stride = kLinkStrideWords;
  h.nx = c.grid.nx;
  h.ny = c.grid.ny;
  h.nz = c.grid.nz;
  for (int k = 0; k < 3; ++k)
    h.om[k] = (R)c.grid.om[k];
  h.oo_res = (R)c.grid.oo;
  h.margin = (R)margin;
  h.pair_eps = (R)pair_eps;
  h.bs_eps = (R)bs_eps;
  h.tight_eps = (R)tight_eps;
  out.assign(off, 0);
  memcpy(out.data(), &h, sizeof(h));
  auto* L = reinterpret_cast<LinkRec<R>*>(out.data() + h.off_links);
  for (size_t i = 0; i < links.size(); ++i)
  {
    for (int k = 0; k < 12; ++k)
      L[i].o[k] = (R)links[i].o[k];
    for (int k = 0; k < 3; ++k)
      L[i].ax[k] = (R)links[i].ax[k];
    L[i].parent = links[i].parent;
    L[i].type = links[i].type;
    L[i].var = links[i].var;
    L[i].has_origin = links[i].has_origin;
    L[i].s0 = links[i].s0;
    L[i].s1 = links[i].s1;
  }
  if (!vars.empty())
    memcpy(out.data() + h.off_vars, vars.data(), vars.size() * sizeof(VarRec));
  auto* S = reinterpret_cast<SphereRec<R>*>(out.data() + h.off_spheres);
  for (size_t i = 0; i < spheres.size(); ++i)
  {
    S[i].x = (R)spheres[i].x;
    S[i].y = (R)spheres[i].y;
    S[i].z = (R)spheres[i].z;
    S[i].r = (R)spheres[i].r;
    S[i].slot = spheres[i].slot;
    S[i].thr = spheres[i].thr;
    S[i].thr_self = spheres[i].thr_self;
  }
  auto* B = reinterpret_cast<BsRec<R>*>(out.data() + h.off_bs);
  for (size_t i = 0; i < bs.size(); ++i)
  {
    B[i].x = (R)bs[i].x;
    B[i].y = (R)bs[i].y;
    B[i].z = (R)bs[i].z;
    B[i].r = (R)bs[i].r;
    B[i].tx = (R)bs[i].tx;
    B[i].ty = (R)bs[i].ty;
    B[i].tz = (R)bs[i].tz;
    B[i].tr = (R)bs[i].tr;
    B[i].slot = bs[i].slot;
    B[i].s0 = bs[i].s0;
    B[i].s1 = bs[i].s1;
    B[i].pad = bs[i].pad;
  }
  if (!pairs.empty())
    memcpy(out.data() + h.off_pairs, pairs.data(), pairs.size() * sizeof(PairRec));
  if (!groups.empty())
    memcpy(out.data() + h.off_groups, groups.data(), groups.size() * sizeof(GroupRec));
  if (!alts.empty())
    memcpy(out.data() + h.off_alts, alts.data(), alts.size() * sizeof(AltRec));
}

// Smallest IEEE half >= x (x > 0), as bits; +inf when x exceeds the half range.
unsigned halfBitsUp(double x)
{
  if (!(x > 0.0))
    return 0u;
  if (x > 65504.0)
    return 0x7c00u;
  int e = 0;
  frexp(x, &e);  // x = f * 2^e, f in [0.5, 1)
  int E = e - 1;  // x in [2^E, 2^(E+1))
  if (E < -14)
    E = -14;  // subnormal range shares the spacing 2^-24
  const double step = ldexp(1.0, E - 10);
  double q = ceil(x / step) * step;
  if (q >= ldexp(1.0, E + 1) && E >= -14)
  {
    ++E;
    q = ldexp(1.0, E);
  }
  if (E > 15)
    return 0x7c00u;
  const int mant = (int)llround(q / ldexp(1.0, E - 10)) - 1024;
  if (q < ldexp(1.0, -14))
    return (unsigned)llround(q / ldexp(1.0, -24));  // subnormal
  return ((unsigned)(E + 15) << 10) | (unsigned)mant;
}

// Tables of fast_check.cu (device_model.hpp, namespace fast), from the same links / spheres / clusters / pairs the
// generic blobs are built from.  All constants are evaluated in double and narrowed once.
bool buildFastBlob(b200sv_ctx& c, const std::vector<LinkRec<double>>& links_in, const std::vector<VarRec>& vars,
                   const std::vector<SphereRec<double>>& spheres_in, int n_spheres, const std::vector<BsRec<double>>& bs_in,
                   const std::vector<PairRec>& pairs, const std::vector<GroupRec>& groups, const std::vector<AltRec>& alts,
                   double margin, double pair_eps, double bs_eps, double tight_eps)
{
  c.blob_fast.clear();
  c.fast_wide = (int)c.robot.group_var_index.size() > fastCheckRegisterPrefetchVars();
  for (const auto& L : links_in)
  {
    if (L.type == PLANAR)
    {
      // a planar joint the group moves: its three variables must be plain, consecutive group variables (no mimic); one the
      // group does not move is a constant link (folded below, or a root: converted to a fixed origin)
      const VarRec &vx = vars[L.var], &vy = vars[L.var + 1], &vt = vars[L.var + 2];
      const bool moving = vx.k >= 0 || vy.k >= 0 || vt.k >= 0;
      if (moving && !(vx.k >= 0 && vy.k == vx.k + 1 && vt.k == vx.k + 2 && vx.a == 1.0 && vy.a == 1.0 && vt.a == 1.0 &&
                      vx.b == 0.0 && vy.b == 0.0 && vt.b == 0.0))
        return false;
      if (moving)
        c.fast_wide = true;
    }
    else if (L.type == FLOATING)
    {  // seven plain, consecutive group variables (x, y, z, then the quaternion x, y, z, w); a constant one is not folded: generic
      bool ok = true;
      for (int k = 0; k < 7; ++k)
        ok = ok && vars[L.var + k].k == vars[L.var].k + k && vars[L.var].k >= 0 && vars[L.var + k].a == 1.0 && vars[L.var + k].b == 0.0;
      if (!ok)
        return false;
      c.fast_wide = true;
    }
    else if (L.type != FIXED && L.type != REVOLUTE && L.type != PRISMATIC)
      return false;
  }
  if ((int)c.robot.group_var_index.size() > fastCheckMaxVars())
    return false;
  const int unroll = fastCheckUnroll();
  const double oo = c.grid.oo;
  const int nl = (int)links_in.size();
  // ---- constant links folded into their ancestors ("alias" links) -----------------------------------------------------------
  // A link behind a fixed joint, or behind a joint the group does not move (its value is the base state's), is rigidly
  // attached to its parent.  Its spheres, cluster centres and bounding sphere are re-expressed (in double) in the frame of
  // the nearest ancestor that has a transform of its own, and a moving link behind it gets that constant transform folded
  // into its joint origin: the kernel then spends no FK on such links (Panda: link8, hand, both fingers; the dual arm: half
  // of its 28 links).
  std::vector<LinkRec<double>> links = links_in;
  std::vector<SphereRec<double>> spheres = spheres_in;
  std::vector<BsRec<double>> bs = bs_in;
  struct X34
  {
    double m[12];  // rows of [R | t]
  };
  auto mul34 = [](const X34& a, const X34& b) {
    X34 o{};
    for (int r = 0; r < 3; ++r)
    {
      for (int col = 0; col < 3; ++col)
        o.m[4 * r + col] = a.m[4 * r] * b.m[col] + a.m[4 * r + 1] * b.m[4 + col] + a.m[4 * r + 2] * b.m[8 + col];
      o.m[4 * r + 3] = a.m[4 * r] * b.m[3] + a.m[4 * r + 1] * b.m[7] + a.m[4 * r + 2] * b.m[11] + a.m[4 * r + 3];
    }
    return o;
  };
  auto originOf = [](const LinkRec<double>& L) {
    X34 o{};
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 4; ++col)
        o.m[4 * r + col] = L.has_origin ? L.o[4 * r + col] : (r == col ? 1.0 : 0.0);
    return o;
  };
  std::vector<char> alias(nl, 0);
  std::vector<int> anchor(nl, -1);  // alias links: the ancestor whose transform they use
  std::vector<X34> rel(nl);         // alias links: their frame in the anchor's frame
  static const bool fold = !(getenv("B200SV_NO_FOLD") && atoi(getenv("B200SV_NO_FOLD")) != 0);
  for (int l = 0; l < nl && fold; ++l)
  {
    const LinkRec<double>& L = links_in[l];
    const int P = L.parent;  // device links are in DFS pre-order: P < l
    const bool constant = L.type == FIXED || (L.type == PLANAR ? (vars[L.var].k < 0 && vars[L.var + 1].k < 0 && vars[L.var + 2].k < 0) :
                                                                 vars[L.var].k < 0);
    if (constant && P >= 0)
    {
      // local transform origin * joint(value): Rodrigues rotation (revolute_joint_model.cpp:250-287) or a translation along
      // the axis (prismatic_joint_model.cpp:124-146)
      X34 J{};
      J.m[0] = J.m[5] = J.m[10] = 1.0;
      if (L.type != FIXED)
      {
        const double q = vars[L.var].b, x = L.ax[0], y = L.ax[1], z = L.ax[2];
        if (L.type == PLANAR)
        {  // Translation(x, y, 0) * AngleAxis(theta, UnitZ) (planar_joint_model.cpp:345-349)
          const double th = vars[L.var + 2].b, cs = cos(th), sn = sin(th);
          J.m[0] = cs; J.m[1] = -sn; J.m[4] = sn; J.m[5] = cs;
          J.m[3] = vars[L.var].b;
          J.m[7] = vars[L.var + 1].b;
        }
        else if (L.type == REVOLUTE)
        {
          const double cs = cos(q), sn = sin(q), t = 1.0 - cs;
          const double R[9] = { t * x * x + cs, t * x * y - z * sn, t * x * z + y * sn, t * x * y + z * sn, t * y * y + cs,
                                t * y * z - x * sn, t * x * z - y * sn, t * y * z + x * sn, t * z * z + cs };
          for (int r = 0; r < 3; ++r)
            for (int col = 0; col < 3; ++col)
              J.m[4 * r + col] = R[3 * r + col];
        }
        else
        {
          J.m[3] = x * q;
          J.m[7] = y * q;
          J.m[11] = z * q;
        }
      }
      const X34 local = mul34(originOf(L), J);
      alias[l] = 1;
      anchor[l] = alias[P] ? anchor[P] : P;
      rel[l] = alias[P] ? mul34(rel[P], local) : local;
    }
  }
  for (int l = 0; l < nl; ++l)
  {
    LinkRec<double>& L = links[l];
    const int P = links_in[l].parent;
    if (alias[l])
    {
      L.type = fast::kAlias;
      L.parent = anchor[l];
      L.has_origin = 1;
      for (int i = 0; i < 12; ++i)
        L.o[i] = rel[l].m[i];  // only its translation is used below (reach bound)
    }
    else if (P >= 0 && alias[P])
    {
      const X34 o = mul34(rel[P], originOf(links_in[l]));
      L.parent = anchor[P];
      L.has_origin = 1;
      for (int i = 0; i < 12; ++i)
        L.o[i] = o.m[i];
    }
  }
  auto move = [&](int slot, double& x, double& y, double& z) {
    const double* m = rel[slot].m;
    const double nx = m[0] * x + m[1] * y + m[2] * z + m[3], ny = m[4] * x + m[5] * y + m[6] * z + m[7],
                 nz = m[8] * x + m[9] * y + m[10] * z + m[11];
    x = nx;
    y = ny;
    z = nz;
  };
  for (auto& S : spheres)
    if (alias[S.slot])
      move(S.slot, S.x, S.y, S.z);
  for (auto& B : bs)
    if (alias[B.slot])
    {
      move(B.slot, B.x, B.y, B.z);
      move(B.slot, B.tx, B.ty, B.tz);
    }
  auto tslot = [&](int slot) { return alias[slot] ? anchor[slot] : slot; };  // whose transform poses this link's geometry
  // which transforms are stored: parents that are not in registers when needed, and links whose clusters are in pairs
  std::vector<char> stored(nl, 0);
  {
    int in_regs = -2;  // the link whose transform the registers hold after the previous link
    for (int l = 0; l < nl; ++l)
    {
      const int P = links[l].parent;
      if (P >= 0 && P != in_regs)
        stored[P] = 1;
      in_regs = alias[l] ? P : l;
    }
  }
  for (const auto& P : pairs)
  {
    stored[tslot(bs[P.a].slot)] = 1;
    stored[tslot(bs[P.b].slot)] = 1;
  }
  std::vector<int> store_off(nl, -1);
  int n_stored = 0;
  for (int l = 0; l < nl; ++l)
    if (stored[l])
      store_off[l] = kLinkStrideWords * n_stored++;
  c.state_stride_fast = std::max(1, n_stored) * kLinkStrideWords;
  while (c.state_stride_fast % 32 != 4)
    c.state_stride_fast += 4;
  // spheres -> Hot entries, per link padded to the unroll
  std::vector<fast::Hot> hot;
  c.fast_hot_sphere.clear();
  std::vector<fast::Link> fl(nl);
  const bool bytes8 = c.compact_bytes == 1;
  int in_regs = -2;
  for (int l = 0; l < nl; ++l)
  {
    const LinkRec<double>& L = links[l];
    fast::Link F{};
    F.type = L.type;
    F.k = -1;
    F.a = 0.0;
    F.b = 0.0;
    if (L.type != FIXED && L.type != fast::kAlias)
    {
      const VarRec& v = vars[L.var];
      F.k = v.k;
      F.a = v.a;
      F.b = v.b;
    }
    const bool root = L.parent < 0;
    F.parent_off = root ? fast::kParentRoot : (L.parent == in_regs ? fast::kParentPrev : store_off[L.parent]);
    F.store_off = alias[l] ? -1 : store_off[l];
    in_regs = alias[l] ? L.parent : l;
    // O: origin rotation (rows), scaled into the voxel frame for a root
    double O[9], ot[3];
    for (int r = 0; r < 3; ++r)
    {
      for (int col = 0; col < 3; ++col)
        O[3 * r + col] = L.has_origin ? L.o[4 * r + col] : (r == col ? 1.0 : 0.0);
      ot[r] = L.has_origin ? L.o[4 * r + 3] : 0.0;
    }
    if (root)
    {
      for (int i = 0; i < 9; ++i)
        O[i] *= oo;
      for (int r = 0; r < 3; ++r)
        ot[r] = (ot[r] - c.grid.om[r]) * oo;
    }
    const bool planar = L.type == PLANAR;  // its rotation is about the joint frame's z
    const double x = planar ? 0.0 : L.ax[0], y = planar ? 0.0 : L.ax[1], z = planar ? 1.0 : L.ax[2];
    const double K[9] = { 0, -z, y, z, 0, -x, -y, x, 0 };
    double K2[9], OK[9], OK2[9];
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col)
      {
        double s2 = 0;
        for (int k = 0; k < 3; ++k)
          s2 += K[3 * r + k] * K[3 * k + col];
        K2[3 * r + col] = s2;
      }
    for (int r = 0; r < 3; ++r)
      for (int col = 0; col < 3; ++col)
      {
        double s1 = 0, s2 = 0;
        for (int k = 0; k < 3; ++k)
        {
          s1 += O[3 * r + k] * K[3 * k + col];
          s2 += O[3 * r + k] * K2[3 * k + col];
        }
        OK[3 * r + col] = s1;
        OK2[3 * r + col] = s2;
      }
    for (int i = 0; i < 9; ++i)
    {
      F.a0[i] = (float)O[i];
      F.a1[i] = 0.f;
      F.a2[i] = 0.f;
    }
    for (int r = 0; r < 3; ++r)
      F.ot[r] = (float)ot[r];
    if (L.type == REVOLUTE || planar)
      for (int i = 0; i < 9; ++i)
      {
        F.a1[i] = (float)OK[i];
        F.a2[i] = (float)OK2[i];
      }
    else if (L.type == PRISMATIC)
      for (int r = 0; r < 3; ++r)
        F.a1[r] = (float)(O[3 * r + 0] * x + O[3 * r + 1] * y + O[3 * r + 2] * z);
    F.h0 = (int)hot.size();
    for (int k = L.s0; k < L.s1 && k < n_spheres; ++k)
    {
      c.fast_hot_sphere.push_back(k);
      const SphereRec<double>& S = spheres[k];
      fast::Hot H{};
      H.x = (float)S.x;
      H.y = (float)S.y;
      H.z = (float)S.z;
      const unsigned t = (unsigned)S.thr | ((unsigned)S.thr_self << 16);
      H.t = bytes8 ? 0x01000100u - t : t;
      hot.push_back(H);
    }
    while (((int)hot.size() - F.h0) % unroll)
    {  // padding: posed like a real sphere (at the link origin), thresholds 0: never collides
      c.fast_hot_sphere.push_back(-1);
      fast::Hot H{};
      H.t = bytes8 ? 0x01000100u : 0u;
      hot.push_back(H);
    }
    F.h1 = (int)hot.size();
    F.c0 = F.c1 = 0;
    // the link's own cell-face margin (cells): its spheres' worst-case FP32 error + the rounding of the magic-add index
    {
      const int nmax_l = std::max(c.grid.nx, std::max(c.grid.ny, c.grid.nz));
      const double e = l < (int)c.fp32_link_eps.size() ? c.fp32_link_eps[l] : c.fp32_eps;
      F.pad3 = (float)std::min(margin, e * c.grid.oo + nmax_l * ldexp(1.0, -22));
    }
    fl[l] = F;
  }
  // Level-1 cull in half precision: cluster centres are parked relative to rep_base (the first root's origin).
  // |centre - rep_base| <= reach: the longest chain of |origin translation| + prismatic travel unknown => bounded by
  // the group's joint limits is not available here, so prismatic joints count with 2 m.  Spacing of half at that
  // magnitude bounds the rounding of each coordinate; the slack covers both centres, all three axes.
  double rep_base[3] = { 0, 0, 0 };
  std::vector<double> reach(nl, 0.0);
  bool have_base = false;
  double max_reach = 0.0;
  for (int l = 0; l < nl; ++l)
  {
    const LinkRec<double>& L = links[l];
    double ot[3] = { L.has_origin ? L.o[3] : 0.0, L.has_origin ? L.o[7] : 0.0, L.has_origin ? L.o[11] : 0.0 };
    if (L.parent < 0 && !have_base)
    {
      for (int k = 0; k < 3; ++k)
        rep_base[k] = (ot[k] - c.grid.om[k]) * oo;
      have_base = true;
    }
    double r;
    if (L.parent < 0)
    {
      const double dx = (ot[0] - c.grid.om[0]) * oo - rep_base[0], dy = (ot[1] - c.grid.om[1]) * oo - rep_base[1],
                   dz = (ot[2] - c.grid.om[2]) * oo - rep_base[2];
      r = sqrt(dx * dx + dy * dy + dz * dz);
    }
    else
      r = reach[L.parent] + sqrt(ot[0] * ot[0] + ot[1] * ot[1] + ot[2] * ot[2]) * oo;
    if (L.type == PRISMATIC || L.type == PLANAR || L.type == FLOATING)
      r += 2.0 * oo;  // assumed travel; the kernel verifies it (Link::range_check) and defers to the exact pass beyond
    reach[l] = r;
    fl[l].range_check =
        (L.type == PRISMATIC || L.type == PLANAR || L.type == FLOATING || (L.parent >= 0 && fl[L.parent].range_check)) ? 1 : 0;
  }
  for (size_t i = 0; i < bs.size(); ++i)
  {
    const double cr = sqrt(bs[i].tx * bs[i].tx + bs[i].ty * bs[i].ty + bs[i].tz * bs[i].tz) * oo;
    max_reach = std::max(max_reach, reach[bs[i].slot] + cr);
  }
  int e2 = 0;
  frexp(std::max(max_reach, 1.0), &e2);                    // max_reach < 2^e2
  const double half_spacing = ldexp(1.0, e2 - 11);         // spacing of half values below 2^e2
  const double half_slack = 2.0 * sqrt(3.0) * half_spacing + 0.01;
  // Hierarchical cull for robots with many cluster pairs: a LINK-level tight sphere per body is parked next to its
  // clusters' ("rep" = a parked centre: the body's clusters, then the body itself), link pairs are culled first.
  // Measured on the dual-arm tree (profiles/r1_hier_cull_rates.txt): slower at 278 cluster pairs (0.94e9 vs 1.31e9
  // states/s: too many link pairs survive for the expansion to pay), faster from ~1 700 pairs (3.6e8 vs 3.0e8).
  bool hier = pairs.size() > 1024;
  if (const char* e = getenv("B200SV_HIER"))
    hier = atoi(e) != 0 && !pairs.empty();
  const int n_bodies = c.cm.n_glinks;
  std::vector<int> rep_of_cluster(bs.size(), -1), body_rep(n_bodies, -1);
  std::vector<double> body_tight(4 * (size_t)n_bodies, 0.0);
  int n_rep = 0;
  for (size_t i = 0; i < bs.size();)
  {
    const int body = bs[i].pad;
    size_t j = i;
    while (j < bs.size() && bs[j].pad == body)
      rep_of_cluster[j++] = n_rep++;
    if (hier)
    {
      const int s0 = bs[i].s0, s1 = bs[j - 1].s1;
      double cx = 0, cy = 0, cz = 0, rr = 0;
      for (int k = s0; k < s1; ++k)
      {
        cx += spheres[k].x; cy += spheres[k].y; cz += spheres[k].z;
      }
      cx /= (s1 - s0); cy /= (s1 - s0); cz /= (s1 - s0);
      for (int k = s0; k < s1; ++k)
      {
        const double ex = spheres[k].x - cx, ey = spheres[k].y - cy, ez = spheres[k].z - cz;
        rr = std::max(rr, sqrt(ex * ex + ey * ey + ez * ez) + spheres[k].r);
      }
      body_tight[4 * body + 0] = cx; body_tight[4 * body + 1] = cy; body_tight[4 * body + 2] = cz;
      body_tight[4 * body + 3] = rr + 1e-9;
      body_rep[body] = n_rep++;
    }
    i = j;
  }
  if (n_rep > 0xffff)
    return false;
  std::vector<fast::Cluster> cl(n_rep);
  std::vector<fast::PSphere> ps((size_t)n_rep * fast::kPSphereStride);
  for (auto& Q : ps)
    Q.r_v = -1e30f;
  auto linkRange = [&](int slot, int rep) {
    fast::Link& F = fl[slot];  // a body's reps are contiguous, and a device link carries exactly one body
    if (F.c1 == 0)
      F.c0 = rep;
    F.c0 = std::min(F.c0, rep);
    F.c1 = std::max(F.c1, rep + 1);
  };
  for (size_t i = 0; i < bs.size(); ++i)
  {
    if (bs[i].s1 - bs[i].s0 > fast::kClusterSpheres)
      return false;  // cannot happen: buildTables caps clusters at kClusterSpheres
    const int rep = rep_of_cluster[i];
    fast::Cluster C{};
    C.tx = (float)bs[i].tx;
    C.ty = (float)bs[i].ty;
    C.tz = (float)bs[i].tz;
    cl[rep] = C;
    linkRange(bs[i].slot, rep);
    for (int k = 0; k < fast::kClusterSpheres; ++k)
      if (bs[i].s0 + k < bs[i].s1)
      {
        const SphereRec<double>& S = spheres[bs[i].s0 + k];
        fast::PSphere Q{};
        Q.x = (float)S.x;
        Q.y = (float)S.y;
        Q.z = (float)S.z;
        Q.r_v = (float)(S.r * oo);
        ps[(size_t)rep * fast::kPSphereStride + k] = Q;
      }
    if (hier && (i + 1 == bs.size() || bs[i + 1].pad != bs[i].pad))
    {
      const int body = bs[i].pad;
      fast::Cluster L{};
      L.tx = (float)body_tight[4 * body + 0];
      L.ty = (float)body_tight[4 * body + 1];
      L.tz = (float)body_tight[4 * body + 2];
      cl[body_rep[body]] = L;
      linkRange(bs[i].slot, body_rep[body]);
    }
  }
  // Where a centre is parked (hierarchical cull): the link-level ones in shared memory rows 0 .. n_parked - 1, the cluster-level
  // ones in the warp's global scratch at their own index.  Without the hierarchy every centre is in shared memory at its index.
  std::vector<int> slot_of_rep(n_rep, -1);
  int n_parked = n_rep;
  if (hier)
  {
    n_parked = 0;
    for (int b = 0; b < n_bodies; ++b)
      if (body_rep[b] >= 0)
        slot_of_rep[body_rep[b]] = n_parked++;
  }
  for (int r2 = 0; r2 < n_rep; ++r2)
    cl[r2].slot = hier ? slot_of_rep[r2] : r2;
  // cluster pairs, ordered by (body a, body b) when the cull is hierarchical so a link pair's cluster pairs are contiguous
  std::vector<int> order(pairs.size());
  for (size_t p = 0; p < pairs.size(); ++p)
    order[p] = (int)p;
  if (hier)
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
      const long kx = (long)bs[pairs[x].a].pad * n_bodies + bs[pairs[x].b].pad;
      const long ky = (long)bs[pairs[y].a].pad * n_bodies + bs[pairs[y].b].pad;
      return kx < ky;
    });
  std::vector<fast::Group> fg;  // the fast kernel's cull is transposed: no per-cluster groups
  std::vector<fast::Pair> fp(pairs.size());
  std::vector<fast::Quirk> fq;
  std::vector<fast::LinkPair> flp;
  for (size_t q = 0; q < order.size(); ++q)
  {
    const size_t p = (size_t)order[q];
    const BsRec<double>&A = bs[pairs[p].a], &B = bs[pairs[p].b];
    fast::Pair P{};
    const double rt = (A.tr + B.tr + tight_eps) * oo + half_slack;
    P.rt2_h = halfBitsUp(rt * rt * (1.0 + 4.0 / 1024.0));  // + the half rounding of the squared sum itself
    const int offA = store_off[tslot(A.slot)], offB = store_off[tslot(B.slot)];
    if (offA > 0xffff || offB > 0xffff)
      return false;  // beyond the packed pair record: the generic kernel takes the group
    P.ab = (unsigned)rep_of_cluster[pairs[p].a] | ((unsigned)rep_of_cluster[pairs[p].b] << 16);
    P.offs = (unsigned)offA | ((unsigned)offB << 16);
    P.quirk = -1;
    if (!pairs[p].quirk_free)
    {
      auto record = [&](const BsRec<double>& U, const BsRec<double>& V, int next) {
        fast::Quirk Q{};
        Q.ax = (float)U.x; Q.ay = (float)U.y; Q.az = (float)U.z;
        Q.bx = (float)V.x; Q.by = (float)V.y; Q.bz = (float)V.z;
        Q.lim_v2 = (float)((U.r + V.r) * oo * oo);
        Q.eps_v2 = (float)(bs_eps * oo * oo);
        Q.a_off = offA;
        Q.b_off = offB;
        Q.next = next;
        fq.push_back(Q);
        return (int)fq.size() - 1;
      };
      // the pair's own bounding spheres head the chain; the other shapes of a several-shape body follow.  All shapes of a
      // body ride on one link (checked in buildTables), so the records of a chain share the pair's two transforms.
      int next = -1;
      for (int alt = pairs[p].alt0; alt >= 0; alt = alts[alt].next)
      {
        if (store_off[tslot(bs[alts[alt].a].slot)] != offA || store_off[tslot(bs[alts[alt].b].slot)] != offB)
          return false;
        next = record(bs[alts[alt].a], bs[alts[alt].b], next);
      }
      P.quirk = record(A, B, next);
    }
    fp[q] = P;
    if (hier)
    {
      const int ba = A.pad, bb = B.pad;
      const unsigned ab = (unsigned)slot_of_rep[body_rep[ba]] | ((unsigned)slot_of_rep[body_rep[bb]] << 16);  // shared-memory rows
      if (flp.empty() || flp.back().ab != ab || flp.back().ncp >= 16)
      {
        fast::LinkPair L{};
        const double rl = (body_tight[4 * ba + 3] + body_tight[4 * bb + 3] + tight_eps) * oo + half_slack;
        L.rt2_h = halfBitsUp(rl * rl * (1.0 + 4.0 / 1024.0));
        L.ab = ab;
        L.cp0 = (unsigned)q;
        L.ncp = 0;
        flp.push_back(L);
      }
      ++flp.back().ncp;
    }
  }
  {  // sentinel after the last link: the kernel reads link l + 1's variable index one link ahead
    fast::Link S{};
    S.k = -1;
    fl.push_back(S);
  }
  fast::Header h{};
  h.n_links = nl;
  h.n_groups = (int)fg.size();
  h.n_pairs = (int)fp.size();
  h.n_group_vars = (int)c.robot.group_var_index.size();
  int off = align16(sizeof(fast::Header));
  h.off_links = off;
  off = align16(off + (int)(fl.size() * sizeof(fast::Link)));
  h.off_hot = off;
  off = align16(off + (int)(hot.size() * sizeof(fast::Hot)));
  h.off_psphere = off;
  off = align16(off + (int)(ps.size() * sizeof(fast::PSphere)));
  h.off_clusters = off;
  off = align16(off + (int)(cl.size() * sizeof(fast::Cluster)));
  h.off_groups = off;
  off = align16(off + (int)(fg.size() * sizeof(fast::Group)));
  h.off_pairs = off;
  off = align16(off + (int)(fp.size() * sizeof(fast::Pair)));
  h.off_quirks = off;
  off = align16(off + (int)(fq.size() * sizeof(fast::Quirk)));
  h.off_linkpairs = off;
  off = align16(off + (int)(flp.size() * sizeof(fast::LinkPair)));
  h.n_linkpairs = (int)flp.size();
  c.fast_n_linkpairs = h.n_linkpairs;
  // with the hierarchical cull the pair tables stay in global memory: shared memory holds the blob up to off_pairs
  c.fast_smem_blob_bytes = h.n_linkpairs > 0 ? h.off_pairs : off;
  c.fast_n_reps = n_parked;                  // rows of the per-warp shared-memory arrays
  c.fast_n_reps_global = hier ? n_rep : 0;   // rows of the per-warp global scratch (hierarchical cull)
  h.total_bytes = off;
  h.state_stride = c.state_stride_fast;
  h.t_rows = 3 * std::max(1, n_stored);
  c.fast_t_rows = h.t_rows;
  h.n_clusters = (int)cl.size();
  h.n_parked = n_parked;
  h.nz = c.grid.nz;
  h.nynz = c.grid.ny * c.grid.nz;
  h.idx_bias = 0u - 0x4B400000u * (unsigned)(h.nynz + h.nz + 1);
  const int n[3] = { c.grid.nx, c.grid.ny, c.grid.nz };
  for (int k = 0; k < 3; ++k)
  {
    h.lo[k] = (float)(0.5 - margin);
    h.hi[k] = (float)(n[k] - 0.5 - margin);
  }
  h.margin = (float)margin;
  h.two_margin = (float)(2.0 * margin);
  for (int k = 0; k < 3; ++k)
    h.rep_base[k] = (float)rep_base[k];
  h.rep_limit = (float)ldexp(1.0, e2);
  for (int r = 0; r < 3; ++r)
    h.ident[4 * r + r] = 1.f;

  h.pair_eps_v = (float)(pair_eps * oo);
  h.tight_eps_v = (float)(tight_eps * oo);
  c.blob_fast.assign(off, 0);
  uint8_t* o = c.blob_fast.data();
  memcpy(o, &h, sizeof(h));
  memcpy(o + h.off_links, fl.data(), fl.size() * sizeof(fast::Link));
  if (!hot.empty())
  {
    memcpy(o + h.off_hot, hot.data(), hot.size() * sizeof(fast::Hot));
  }
  if (!cl.empty())
  {
    memcpy(o + h.off_clusters, cl.data(), cl.size() * sizeof(fast::Cluster));
    memcpy(o + h.off_psphere, ps.data(), ps.size() * sizeof(fast::PSphere));
  }
  if (!fg.empty())
    memcpy(o + h.off_groups, fg.data(), fg.size() * sizeof(fast::Group));
  if (!fp.empty())
    memcpy(o + h.off_pairs, fp.data(), fp.size() * sizeof(fast::Pair));
  if (!fq.empty())
    memcpy(o + h.off_quirks, fq.data(), fq.size() * sizeof(fast::Quirk));
  if (!flp.empty())
    memcpy(o + h.off_linkpairs, flp.data(), flp.size() * sizeof(fast::LinkPair));
  return true;
}

// column-major 4x4 -> row-major 3x4
void toRow34(const Mat4& m, double* o)
{
  for (int r = 0; r < 3; ++r)
    for (int col = 0; col < 4; ++col)
      o[r * 4 + col] = m[col * 4 + r];
}

int validateModel(b200sv_ctx* c)
{
  const HostRobot& r = c->robot;
  const HostCollisionModel& m = c->cm;
  if (r.n_links <= 0 || r.n_vars < 0)
    return fail(c, B200SV_ERR_INVALID, "robot has no links");
  for (int l = 0; l < r.n_links; ++l)
  {
    if (r.parent[l] >= l)
      return fail(c, B200SV_ERR_INVALID, "links must be in DFS order: parent index must be smaller than the link's");
    if (r.joint_type[l] < FIXED || r.joint_type[l] > FLOATING)
      return fail(c, B200SV_ERR_INVALID, "unknown joint type");
    const int nv = jointVarCount(r.joint_type[l]);
    if (nv && (r.first_var[l] < 0 || r.first_var[l] + nv > r.n_vars))
      return fail(c, B200SV_ERR_INVALID, "joint variable index out of range");
  }
  for (int v : r.group_var_index)
    if (v < 0 || v >= r.n_vars)
      return fail(c, B200SV_ERR_INVALID, "group variable index out of range");
  if (r.group_var_index.empty())
    return fail(c, B200SV_ERR_INVALID, "the group has no variables");
  for (size_t i = 0; i < r.mimic_dst.size(); ++i)
    if (r.mimic_dst[i] < 0 || r.mimic_dst[i] >= r.n_vars || r.mimic_src[i] < 0 || r.mimic_src[i] >= r.n_vars)
      return fail(c, B200SV_ERR_INVALID, "mimic variable index out of range");
  if ((int)m.sphere_start.size() != m.n_glinks + 1)
    return fail(c, B200SV_ERR_INVALID, "sphere_start must have n_glinks + 1 entries");
  for (int i = 0; i < m.n_glinks; ++i)
  {
    if (m.glink_index[i] < 0 || m.glink_index[i] >= r.n_links)
      return fail(c, B200SV_ERR_INVALID, "glink_index out of range");
    if (m.sphere_start[i + 1] < m.sphere_start[i])
      return fail(c, B200SV_ERR_INVALID, "sphere_start must be non-decreasing");
    // a link index that already appeared denotes an ATTACHED BODY carried by that link (dfce->attached_body_names_ follow
    // dfce->link_names_, collision_env_distance_field.cpp:798-822); the link entries themselves come in
    // getUpdatedLinkModels() order
  }
  return B200SV_OK;
}

// Worst-case FP32 error (metres) of a sphere centre posed by the fast kernel, from the model: the budget every FP32 decision
// carries (a decision that could flip within it goes to the exact pass).  First-order forward analysis of what the kernel
// computes, u = 2^-24 the unit roundoff, all magnitudes in metres (the kernel works in cells = metres * 1/res, a pure scale):
//   rotation block of link l     E_R[l] = E_R[parent] + joint term
//        every link              4 u   an entry of T_parent * M is three FMAs (partial sums <= 1) + the rounded constants
//        revolute / planar       + 4 u (M = A0 + s A1 + (1 - c) A2: two FMAs, constants) + e_trig + (|q|max + 1) u (q
//                                narrowed to float; 1 - c)
//        floating                + 24 u        quaternion normalisation and the nine products
//   origin of link l             E_p[l] = E_p[parent] + sqrt(3) E_R[parent] |t_l| + 3 u (P + sqrt(3) |t_l|) + u |t_l|
//        t_l = the joint origin's translation (+ travel for prismatic / planar / floating joints), P = the largest
//        coordinate magnitude the kernel holds (field extent + reach): three FMAs with partial sums <= P, rounded constants
//   sphere s of link l           E[s] = E_p[l] + sqrt(3) E_R[l] |offset_s| + 3 u (P + sqrt(3) |offset_s|) + u |offset_s| + u P
//        (the last term: the margin-lowered translation column is rounded once more)
// e_trig = 2^-22: sincosf's documented 2 ulp on values <= 1.  Links the host folds into an ancestor only shorten the chain.
// link_eps[l] = the largest E[s] over the link's spheres: the kernel tests cell-face proximity with a per-link margin.
double deriveFp32Eps(const b200sv_ctx& c, const std::vector<LinkRec<double>>& links, const std::vector<VarRec>& vars,
                     const std::vector<SphereRec<double>>& spheres, std::vector<double>& link_eps)
{
  const double u = ldexp(1.0, -24), r3 = sqrt(3.0), e_trig = ldexp(1.0, -22);
  const size_t nl = links.size();
  const bool have_bounds = c.robot.group_lower.size() == c.robot.group_var_index.size();
  auto qmax = [&](int var, bool angle) {
    const VarRec& v = vars[var];
    double m = fabs(v.b);
    if (v.k >= 0)
    {
      const double lo = have_bounds ? c.robot.group_lower[v.k] : (angle ? -2.0 * M_PI : -2.0);
      const double hi = have_bounds ? c.robot.group_upper[v.k] : (angle ? 2.0 * M_PI : 2.0);
      const double lim = angle ? 2.0 * M_PI : 1e3;  // unbounded (continuous / planar travel): angles wrap, travel is checked
      m = fabs(v.a) * std::min(lim, std::max(fabs(lo), fabs(hi))) + fabs(v.b);
    }
    return m;
  };
  std::vector<double> ER(nl, 0.0), EP(nl, 0.0), reach(nl, 0.0), tl(nl, 0.0), travel(nl, 0.0);
  double reach_max = 0.0;
  for (size_t l = 0; l < nl; ++l)
  {
    const LinkRec<double>& L = links[l];
    tl[l] = L.has_origin ? sqrt(L.o[3] * L.o[3] + L.o[7] * L.o[7] + L.o[11] * L.o[11]) : 0.0;
    if (L.type == PRISMATIC)
      travel[l] = qmax(L.var, false);
    else if (L.type == PLANAR)
      travel[l] = qmax(L.var, false) + qmax(L.var + 1, false);
    else if (L.type == FLOATING)
      travel[l] = qmax(L.var, false) + qmax(L.var + 1, false) + qmax(L.var + 2, false);
    travel[l] = std::min(travel[l], 4.0);  // beyond the assumed travel the kernel itself defers to the exact pass
    reach[l] = (L.parent >= 0 ? reach[L.parent] + tl[l] : 0.0) + travel[l];
    reach_max = std::max(reach_max, reach[l]);
  }
  double off_max = 0.0;
  for (const SphereRec<double>& S : spheres)
    off_max = std::max(off_max, sqrt(S.x * S.x + S.y * S.y + S.z * S.z));
  const double extent = c.cfg.resolution * std::max(c.grid.nx, std::max(c.grid.ny, c.grid.nz));
  double root_off = 0.0;  // a root's origin relative to the field corner
  for (size_t l = 0; l < nl; ++l)
    if (links[l].parent < 0)
      for (int k = 0; k < 3; ++k)
        root_off = std::max(root_off, fabs((links[l].has_origin ? links[l].o[4 * k + 3] : 0.0) - c.grid.om[k]));
  const double P = std::max(extent, root_off) + reach_max + off_max;
  double worst = 0.0;
  for (size_t l = 0; l < nl; ++l)
  {
    const LinkRec<double>& L = links[l];
    const double er_p = L.parent >= 0 ? ER[L.parent] : 0.0, ep_p = L.parent >= 0 ? EP[L.parent] : 0.0;
    double joint = 4.0 * u;
    if (L.type == REVOLUTE)
      joint += 4.0 * u + e_trig + (qmax(L.var, true) + 1.0) * u;
    else if (L.type == PLANAR)
      joint += 4.0 * u + e_trig + (qmax(L.var + 2, true) + 1.0) * u;
    else if (L.type == FLOATING)
      joint += 24.0 * u;
    const double t = tl[l] + travel[l];
    ER[l] = er_p + joint;
    EP[l] = ep_p + r3 * er_p * t + 3.0 * u * (P + r3 * t) + u * t;
  }
  link_eps.assign(nl, 0.0);
  for (size_t l = 0; l < nl; ++l)
    link_eps[l] = EP[l] + 4.0 * u * P;  // a link without spheres still poses cluster centres
  for (const SphereRec<double>& S : spheres)
  {
    if (S.slot < 0 || S.slot >= (int)nl)
      continue;
    const double off = sqrt(S.x * S.x + S.y * S.y + S.z * S.z);
    const double e = EP[S.slot] + r3 * ER[S.slot] * off + 3.0 * u * (P + r3 * off) + u * off + u * P;
    link_eps[S.slot] = std::max(link_eps[S.slot], e);
    worst = std::max(worst, e);
  }
  return worst;
}

int buildTables(b200sv_ctx* c)
{
  const HostRobot& r = c->robot;
  const HostCollisionModel& m = c->cm;
  // ---- variable map: setJointGroupPositions then updateMimicJoints(group) ----------------------------------
  std::vector<VarRec> vmap(r.n_vars);
  for (int v = 0; v < r.n_vars; ++v)
    vmap[v] = VarRec{ 0.0, r.base_positions[v], -1, 0 };
  for (size_t i = 0; i < r.group_var_index.size(); ++i)
    vmap[r.group_var_index[i]] = VarRec{ 1.0, 0.0, (int)i, 0 };
  for (size_t i = 0; i < r.mimic_dst.size(); ++i)
  {
    const VarRec s = vmap[r.mimic_src[i]];
    VarRec d;
    d.pad = 0;
    if (s.k >= 0)
    {  // factor * (a q + b) + offset; a == 1, b == 0 for a group variable, so this is factor * q + offset exactly
      d.k = s.k;
      d.a = r.mimic_factor[i] * s.a;
      d.b = r.mimic_factor[i] * s.b + r.mimic_offset[i];
    }
    else
    {
      d.k = -1;
      d.a = 0.0;
      d.b = r.mimic_factor[i] * s.b + r.mimic_offset[i];
    }
    vmap[r.mimic_dst[i]] = d;
  }
  // positions of the base state after the same update (mimic joints of the group follow their sources)
  std::vector<double> base = r.base_positions;
  for (size_t i = 0; i < r.mimic_dst.size(); ++i)
    if (vmap[r.mimic_dst[i]].k < 0)
      base[r.mimic_dst[i]] = vmap[r.mimic_dst[i]].b;
  std::vector<Mat4> T0;
  hostForwardKinematics(r, base, T0);
  c->const_T.resize((size_t)r.n_links * 16);
  for (int l = 0; l < r.n_links; ++l)
    std::copy(T0[l].begin(), T0[l].end(), c->const_T.begin() + 16 * l);
  // ---- which links move with the group ------------------------------------------------------------------------
  std::vector<char> dynamic(r.n_links, 0), needed(r.n_links, 0);
  for (int l = 0; l < r.n_links; ++l)
  {
    bool dyn = r.parent[l] >= 0 && dynamic[r.parent[l]];
    const int nv = jointVarCount(r.joint_type[l]);
    for (int k = 0; k < nv; ++k)
      dyn |= vmap[r.first_var[l] + k].k >= 0;
    dynamic[l] = dyn;
  }
  for (int i = 0; i < m.n_glinks; ++i)
    if (m.sphere_start[i + 1] > m.sphere_start[i])
      needed[m.glink_index[i]] = 1;
  c->link_slot.assign(r.n_links, -1);
  std::vector<LinkRec<double>> links;
  std::vector<VarRec> vars;
  for (int l = 0; l < r.n_links; ++l)
  {
    if (!dynamic[l] && !needed[l])
      continue;
    LinkRec<double> L{};
    Mat4 origin;
    std::copy(r.joint_origin.begin() + 16 * l, r.joint_origin.begin() + 16 * (l + 1), origin.begin());
    L.var = (int)vars.size();
    L.s0 = L.s1 = 0;
    for (int k = 0; k < 3; ++k)
      L.ax[k] = r.joint_axis[3 * l + k];
    if (!dynamic[l])
    {  // a checked link that does not move: its base-state transform, as a fixed root
      toRow34(T0[l], L.o);
      L.parent = -1;
      L.type = FIXED;
      L.has_origin = 1;
    }
    else
    {
      L.type = r.joint_type[l];
      const int nv = jointVarCount(L.type);
      for (int k = 0; k < nv; ++k)
        vars.push_back(vmap[r.first_var[l] + k]);
      const int p = r.parent[l];
      if (p >= 0 && c->link_slot[p] >= 0 && dynamic[p])
      {
        L.parent = c->link_slot[p];
        toRow34(origin, L.o);
        L.has_origin = isIdentityTransform(origin.data()) ? 0 : 1;
      }
      else
      {  // parent does not move: fold its base-state transform into the origin (same product order as the reference)
        L.parent = -1;
        Mat4 g = p >= 0 ? mulAffine(T0[p], origin) : origin;
        if (p >= 0 && isIdentityTransform(origin.data()))
          g = T0[p];
        toRow34(g, L.o);
        L.has_origin = 1;
      }
    }
    c->link_slot[l] = (int)links.size();
    links.push_back(L);
  }
  // Attached bodies (second and later entries of the same link): their spheres are given in the LINK frame, so a body
  // is a pseudo link fixed to its link with an identity origin.  Appended after the real links: device sphere order
  // stays the caller's (dfce) order.
  std::vector<int>& glink_slot = c->glink_slot;
  glink_slot.assign(m.n_glinks, -1);
  {
    std::vector<char> seen(r.n_links, 0);
    for (int i = 0; i < m.n_glinks; ++i)
    {
      const int li = m.glink_index[i];
      if (!seen[li])
      {
        seen[li] = 1;
        glink_slot[i] = c->link_slot[li];
        continue;
      }
      if (m.sphere_start[i + 1] <= m.sphere_start[i])
        continue;
      if (c->link_slot[li] < 0)
        return fail(c, B200SV_ERR_INVALID, "attached body on a link that neither moves with the group nor carries spheres");
      LinkRec<double> L{};
      L.o[0] = L.o[5] = L.o[10] = 1.0;
      L.parent = c->link_slot[li];
      L.type = FIXED;
      L.var = (int)vars.size();
      L.has_origin = 0;
      glink_slot[i] = (int)links.size();
      links.push_back(L);
    }
  }
  c->n_dev_links = (int)links.size();
  if (links.empty())
    return fail(c, B200SV_ERR_INVALID, "no link moves with the group and none carries spheres");
  // FP32: state stride = 4 (mod 32) words so the 8 lanes of a quarter-warp store to 8 distinct 16-byte bank groups;
  // FP64: 2 (mod 16) doubles, the same condition for 16-byte accesses of 8-byte words.
  c->state_stride_f = c->n_dev_links * kLinkStrideWords;
  while (c->state_stride_f % 32 != 4)
    c->state_stride_f += 4;
  c->state_stride_d = c->n_dev_links * kLinkStrideWords;
  while (c->state_stride_d % 16 != 2)
    c->state_stride_d += 2;
  // ---- spheres, bounding spheres, pairs -----------------------------------------------------------------------------
  std::vector<SphereRec<double>> spheres;
  std::vector<BsRec<double>> bs;                                 // clusters
  std::vector<std::vector<int>> clusters_of_glink(m.n_glinks);  // BsRec indices per dfce link
  int max_thr = 0;
  for (int i = 0; i < m.n_glinks; ++i)
  {
    const int a = m.sphere_start[i], b = m.sphere_start[i + 1];
    if (b <= a)
      continue;
    const int slot = glink_slot[i];
    const int first = (int)spheres.size();
    for (int k = a; k < b; ++k)
    {
      SphereRec<double> S{};
      S.x = m.sphere_xyzr[4 * k + 0];
      S.y = m.sphere_xyzr[4 * k + 1];
      S.z = m.sphere_xyzr[4 * k + 2];
      S.r = m.sphere_xyzr[4 * k + 3];
      S.slot = slot;
      S.thr = sphereThreshold(S.r, c->cfg.resolution, c->cfg.max_propagation_distance, c->cfg.collision_tolerance,
                              c->grid.max_d2, c->cfg.use_signed_distance_field != 0);
      if (S.thr < 0)
        return fail(c, B200SV_ERR_UNSUPPORTED,
                    "signed field: collision_tolerance exceeds a sphere's radius + resolution, the outcome inside obstacles "
                    "would depend on the depth");
      S.thr_self = m.self_enabled[i] ? S.thr : 0;
      max_thr = std::max(max_thr, S.thr);
      spheres.push_back(S);
    }
    links[slot].s0 = first;
    links[slot].s1 = (int)spheres.size();
    // Clusters: greedy contiguous runs of the link's spheres whose tight sphere (centroid, farthest sphere surface,
    // + 1 nm) stays within cluster_radius.  Sphere lists are usually laid out along the link, so runs are compact.
    auto tight = [&](int s0, int s1, double* out4) {
      double cx = 0, cy = 0, cz = 0;
      for (int k = s0; k < s1; ++k)
      {
        cx += spheres[k].x;
        cy += spheres[k].y;
        cz += spheres[k].z;
      }
      const double inv = 1.0 / (s1 - s0);
      cx *= inv; cy *= inv; cz *= inv;
      double rr = 0.0;
      for (int k = s0; k < s1; ++k)
      {
        const double ex = spheres[k].x - cx, ey = spheres[k].y - cy, ez = spheres[k].z - cz;
        rr = std::max(rr, sqrt(ex * ex + ey * ey + ez * ez) + spheres[k].r);
      }
      out4[0] = cx; out4[1] = cy; out4[2] = cz; out4[3] = rr + 1e-9;
    };
    int s0 = first;
    while (s0 < (int)spheres.size())
    {
      int s1 = s0 + 1;
      double t4[4], best4[4];
      tight(s0, s1, best4);
      while (s1 < (int)spheres.size() && s1 - s0 < fast::kClusterSpheres)
      {
        tight(s0, s1 + 1, t4);
        if (t4[3] > c->cluster_radius)
          break;
        std::copy(t4, t4 + 4, best4);
        ++s1;
      }
      BsRec<double> B{};
      B.x = m.bounding_sphere[4 * i + 0];
      B.y = m.bounding_sphere[4 * i + 1];
      B.z = m.bounding_sphere[4 * i + 2];
      B.r = m.bounding_sphere[4 * i + 3];
      B.tx = best4[0]; B.ty = best4[1]; B.tz = best4[2]; B.tr = best4[3];
      B.slot = slot;
      B.s0 = s0;
      B.s1 = s1;
      B.pad = i;
      clusters_of_glink[i].push_back((int)bs.size());
      bs.push_back(B);
      s0 = s1;
    }
  }
  // Shapes of one attached body (entries that share a body_id): the reference culls such a body as a whole.
  std::vector<std::vector<int>> shapes_of(m.n_glinks);
  for (int i = 0; i < m.n_glinks; ++i)
  {
    if (clusters_of_glink[i].empty())
      continue;
    shapes_of[i].push_back(i);
    if (m.body_id.empty() || m.body_id[i] < 0)
      continue;
    for (int k = 0; k < m.n_glinks; ++k)
      if (k != i && m.body_id[k] == m.body_id[i] && !clusters_of_glink[k].empty())
      {
        if (m.glink_index[k] != m.glink_index[i])
          return fail(c, B200SV_ERR_INVALID, "entries that share a body_id must be attached to the same link");
        shapes_of[i].push_back(k);
      }
  }
  std::vector<AltRec> alts;
  std::map<std::pair<int, int>, int> alt_chain_of;  // (glink i, glink j) -> head of its AltRec chain
  // cluster pairs of every enabled link pair, grouped by the first cluster
  std::vector<PairRec> pairs;
  std::vector<GroupRec> groups;
  double max_rs = 0.0;
  int n_link_pairs = 0;
  for (int i = 0; i < m.n_glinks; ++i)
  {
    for (int j = i + 1; j < m.n_glinks; ++j)
      if (!clusters_of_glink[i].empty() && !clusters_of_glink[j].empty() && m.intra_enabled[(size_t)i * m.n_glinks + j])
      {
        ++n_link_pairs;
        max_rs = std::max(max_rs, bs[clusters_of_glink[i][0]].r + bs[clusters_of_glink[j][0]].r);
      }
    for (int ca : clusters_of_glink[i])
    {
      GroupRec G{ ca, (int)pairs.size(), 0, 0 };
      for (int j = i + 1; j < m.n_glinks; ++j)
        if (!clusters_of_glink[j].empty() && m.intra_enabled[(size_t)i * m.n_glinks + j])
          for (int cb : clusters_of_glink[j])
          {
            // The reference culls a link pair unless |Ci - Cj|^2 < Ri + Rj on the links' bounding spheres
            // (types.cpp:416-417).  If our tight spheres are near, |Ci - Cj| <= (tr_a + tr_b + slack) + |Ci - ta| +
            // |Cj - tb| (rigid offsets inside each link).  When that bound, squared, is safely below Ri + Rj the
            // reference's test cannot fail, and the kernel skips evaluating it.
            const BsRec<double>&A = bs[ca], &B = bs[cb];
            const double oa = sqrt((A.x - A.tx) * (A.x - A.tx) + (A.y - A.ty) * (A.y - A.ty) + (A.z - A.tz) * (A.z - A.tz));
            const double ob = sqrt((B.x - B.tx) * (B.x - B.tx) + (B.y - B.ty) * (B.y - B.ty) + (B.z - B.tz) * (B.z - B.tz));
            const double dmax = A.tr + B.tr + oa + ob + 1e-3;
            const int quirk_free = dmax * dmax < (A.r + B.r) - 1e-3 ? 1 : 0;
            // A body of several shapes is culled as a whole by the reference: any (shape, shape) combination whose
            // bounding spheres pass lets ALL the body's spheres be tested (:455-507).  The entries here are per shape,
            // so a pair the proof above does not cover carries the other combinations as a chain.
            int alt0 = -1;
            if (!quirk_free && (shapes_of[i].size() > 1 || shapes_of[j].size() > 1))
            {
              auto it = alt_chain_of.find({ i, j });
              if (it == alt_chain_of.end())
              {
                int head = -1;
                for (int u : shapes_of[i])
                  for (int v : shapes_of[j])
                    if (u != i || v != j)
                    {
                      alts.push_back(AltRec{ clusters_of_glink[u][0], clusters_of_glink[v][0], head, 0 });
                      head = (int)alts.size() - 1;
                    }
                it = alt_chain_of.emplace(std::make_pair(i, j), head).first;
              }
              alt0 = it->second;
            }
            pairs.push_back(PairRec{ ca, cb, quirk_free, alt0 });
          }
      G.p1 = (int)pairs.size();
      if (G.p1 > G.p0)
        groups.push_back(G);
    }
  }
  c->n_link_pairs = n_link_pairs;
  c->n_spheres = (int)spheres.size();
  c->n_bs = (int)bs.size();
  c->n_pairs = (int)pairs.size();
  if (c->n_pairs >= (1 << 26))
    return fail(c, B200SV_ERR_UNSUPPORTED, "too many enabled cluster pairs for the work-queue item format");
  while (spheres.size() % B200SV_SPHERE_CHUNK)
  {  // pad to a multiple of the chunk the field loop is unrolled by: thresholds 0 => never read, never collide
    SphereRec<double> S{};
    S.slot = spheres.empty() ? 0 : spheres.back().slot;
    spheres.push_back(S);
  }
  c->compact_bytes = max_thr <= 255 ? 1 : 2;
  if (getenv("B200SV_VERBOSE"))
    fprintf(stderr, "b200sv: %d device links, %d spheres, %d clusters, %d cluster pairs in %d groups (%d link pairs)\n",
            c->n_dev_links, c->n_spheres, c->n_bs, c->n_pairs, (int)groups.size(), n_link_pairs);
  // ---- FP32 uncertainty budget (see DESIGN.md "precision") -------------------------------------------------------------
  // derived from the model (deriveFp32Eps) unless b200sv_set_tuning / B200SV_FP32_EPS fixed it
  c->fp32_eps_derived = deriveFp32Eps(*c, links, vars, spheres, c->fp32_link_eps);
  {
    const char* e = getenv("B200SV_FP32_EPS");
    if (!c->fp32_eps_user && !(e && atof(e) > 0.0))
      c->fp32_eps = c->fp32_eps_derived;
    else
    {  // a fixed budget applies to every link alike
      if (!c->fp32_eps_user)
        c->fp32_eps = atof(e);
      c->fp32_link_eps.assign(links.size(), c->fp32_eps);
    }
  }
  const double eps = c->fp32_eps;
  const int nmax = std::max(c->grid.nx, std::max(c->grid.ny, c->grid.nz));
  const double margin = eps * c->grid.oo + nmax * ldexp(1.0, -22);
  const double pair_eps = 2.0 * eps + 1e-7;
  const double bs_eps = 4.0 * eps * (sqrt(max_rs) + 0.05) + 1e-7;
  const double tight_eps = 4.0 * eps + 1e-6;
  buildBlob<float>(*c, links, vars, spheres, bs, pairs, groups, alts, margin, pair_eps, bs_eps, tight_eps, c->blob_f);
  buildBlob<double>(*c, links, vars, spheres, bs, pairs, groups, alts, 0.0, 0.0, 0.0, 1e-9, c->blob_d);
  c->fast_ok =
      buildFastBlob(*c, links, vars, spheres, c->n_spheres, bs, pairs, groups, alts, margin, pair_eps, bs_eps, tight_eps);
  return B200SV_OK;
}

// Pick (warps per CTA, CTAs per SM) for the thread-per-state check kernel: as many resident warps as shared memory
// (32 states of link transforms per warp) and registers allow.  Limits: 227 KB of shared memory per CTA / 228 KB per
// SM with 1 KB reserved per CTA, 64 K registers per SM, 512 threads per CTA (__launch_bounds__).
b200sv_ctx::LaunchCfg chooseCfg(const b200sv_ctx* c, bool fp64, bool from_list, int blob_bytes, bool fast = false)
{
  b200sv_ctx::LaunchCfg best;
  const int regs = ((fast ? fastCheckKernelRegs(c->compact_bytes, (int)c->robot.group_var_index.size()) : checkKernelRegs(fp64, from_list, c->compact_bytes)) + 7) & ~7;
  const int reg_warps = 65536 / (regs * 32);
  const long pw = fast ? fastCheckPerWarpBytes(c->state_stride_fast, c->fast_n_reps, (int)c->robot.group_var_index.size(), c->fast_n_linkpairs, c->fast_wide) :
                         checkPerWarpBytes(fp64, fp64 ? c->state_stride_d : c->state_stride_f);
  for (int b = 1; b <= 4; ++b)
  {
    if (c->tune_blocks > 0 && b != c->tune_blocks)
      continue;
    const long per_block = std::min<long>(c->smem_optin, (233472L - 1024L * b) / b) - align16(blob_bytes);
    int w = per_block > 0 ? (int)(per_block / pw) : 0;
    w = std::min(w, fast ? fastCheckMaxWarps((int)c->robot.group_var_index.size()) : 16);  // __launch_bounds__ of the kernels
    w = std::min(w, reg_warps / b);
    if (c->tune_warps > 0)
      w = std::min(w, c->tune_warps);
    if (w * b > best.warps * best.blocks_per_sm)
    {
      best.warps = w;
      best.blocks_per_sm = b;
      best.smem = align16(blob_bytes) + (size_t)pw * w;
    }
  }
  return best;
}

int uploadTables(b200sv_ctx* c)
{
  if (c->d_blob_f)
    cudaFree(c->d_blob_f);
  if (c->d_blob_d)
    cudaFree(c->d_blob_d);
  if (c->d_blob_fast)
    cudaFree(c->d_blob_fast);
  c->d_blob_f = c->d_blob_d = c->d_blob_fast = nullptr;
  c->cfg_fast = b200sv_ctx::LaunchCfg();
  if (c->fast_ok)
  {
    CU(c, cudaMalloc(&c->d_blob_fast, c->blob_fast.size()));
    CU(c, cudaMemcpy(c->d_blob_fast, c->blob_fast.data(), c->blob_fast.size(), cudaMemcpyHostToDevice));
    c->cfg_fast = chooseCfg(c, false, false, c->fast_smem_blob_bytes, true);
    c->fast_t_global = false;
    {  // large robots: with the transforms in shared memory only a few warps fit; move them to an L2-resident scratch
      const int keep = c->state_stride_fast;
      c->state_stride_fast = 0;
      const b200sv_ctx::LaunchCfg g = chooseCfg(c, false, false, c->fast_smem_blob_bytes, true);
      c->state_stride_fast = keep;
      const char* e = getenv("B200SV_T_GLOBAL");
      // measured on B200: the scratch variant is also the faster one when shared memory would fit 12 warps
      // (Panda C2: 0.546 vs 0.624 ms per million states) -- fewer shared-memory wavefronts on the L1 pipe the
      // gathers share, coalesced 512-byte stores.  B200SV_T_GLOBAL=0 keeps the transforms in shared memory.
      const bool want = e ? atoi(e) != 0 : g.warps * g.blocks_per_sm >= c->cfg_fast.warps * c->cfg_fast.blocks_per_sm;
      if (want && g.warps >= 1)
      {
        c->cfg_fast = g;
        c->fast_t_global = true;
      }
    }
    if (c->d_t_scratch)
      cudaFree(c->d_t_scratch);
    c->d_t_scratch = nullptr;
    if (c->fast_t_global)
      CU(c, cudaMalloc(&c->d_t_scratch, (size_t)c->sm_count * c->cfg_fast.blocks_per_sm * c->cfg_fast.warps *
                                            (size_t)c->fast_t_rows * 128 * sizeof(float)));
    if (c->d_rep_scratch)
      cudaFree(c->d_rep_scratch);
    c->d_rep_scratch = nullptr;
    if (c->fast_n_reps_global > 0 && c->cfg_fast.warps >= 1)
      CU(c, cudaMalloc(&c->d_rep_scratch, (size_t)c->sm_count * c->cfg_fast.blocks_per_sm * c->cfg_fast.warps *
                                              (size_t)fastCheckRepBytes(c->fast_n_reps_global)));
    if (getenv("B200SV_VERBOSE"))
      fprintf(stderr, "b200sv: fast kernel: %d words of transforms per state, %d bytes per warp, %d warps x %d CTAs per SM, blob %zu bytes\n",
              c->state_stride_fast, fastCheckPerWarpBytes(c->fast_t_global ? 0 : c->state_stride_fast, c->fast_n_reps, (int)c->robot.group_var_index.size(), c->fast_n_linkpairs, c->fast_wide),
              c->cfg_fast.warps, c->cfg_fast.blocks_per_sm, c->blob_fast.size());
    if (c->cfg_fast.warps < 1)
      c->fast_ok = false;  // the generic kernel's limits decide below whether the robot fits at all
  }
  CU(c, cudaMalloc(&c->d_blob_f, c->blob_f.size()));
  CU(c, cudaMalloc(&c->d_blob_d, c->blob_d.size()));
  CU(c, cudaMemcpy(c->d_blob_f, c->blob_f.data(), c->blob_f.size(), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(c->d_blob_d, c->blob_d.data(), c->blob_d.size(), cudaMemcpyHostToDevice));
  if (!c->d_link_slot)
  {
    CU(c, cudaMalloc(&c->d_link_slot, c->link_slot.size() * sizeof(int)));
    CU(c, cudaMalloc(&c->d_const_T, c->const_T.size() * sizeof(double)));
  }
  CU(c, cudaMemcpy(c->d_link_slot, c->link_slot.data(), c->link_slot.size() * sizeof(int), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(c->d_const_T, c->const_T.data(), c->const_T.size() * sizeof(double), cudaMemcpyHostToDevice));
  {
    const char* e = getenv("B200SV_LATENCY");
    c->latency_ok = !(e && atoi(e) == 0) && c->robot.group_var_index.size() <= 32 &&
                    (size_t)align16((int)c->blob_d.size()) + 4 * ((size_t)c->state_stride_d + 32 + 2 * (size_t)c->n_dev_links) * 8 <= (size_t)c->smem_optin;
  }
  c->cfg_f = chooseCfg(c, false, false, (int)c->blob_f.size());
  c->cfg_d = chooseCfg(c, true, false, (int)c->blob_d.size());
  c->cfg_list = chooseCfg(c, true, true, (int)c->blob_d.size());
  if (c->cfg_f.warps < 1 || c->cfg_d.warps < 1 || c->cfg_list.warps < 1)
    return fail(c, B200SV_ERR_UNSUPPORTED,
                "robot too large for the shared-memory tile: " + std::to_string(c->n_dev_links) + " moving links need " +
                    std::to_string(32 * c->state_stride_d * 8) + " bytes per warp");
  return B200SV_OK;
}

int initGrid(b200sv_ctx* c)
{
  const b200sv_field_cfg& f = c->cfg;
  if (!(f.resolution > 0.0) || !(f.max_propagation_distance > 0.0))
    return fail(c, B200SV_ERR_INVALID, "resolution and max_propagation_distance must be positive");
  // corner = origin - size / 2 (collision_env_distance_field.cpp:1786-1787); VoxelGrid::resize (voxel_grid.hpp:364-396)
  const double res = f.resolution;
  const double oo = 1.0 / res;
  GridDev& g = c->grid;
  int n[3];
  for (int k = 0; k < 3; ++k)
  {
    const double corner = f.origin[k] - 0.5 * f.size[k];
    g.om[k] = corner - 0.5 * res;
    n[k] = (int)(f.size[k] * oo);
    if (n[k] < 1)
      return fail(c, B200SV_ERR_INVALID, "field has no cells along one axis");
  }
  g.oo = oo;
  g.nx = n[0];
  g.ny = n[1];
  g.nz = n[2];
  g.wz = (g.nz + 31) / 32;
  const double rc = ceil(f.max_propagation_distance / res);  // propagation_distance_field.cpp:88
  if (rc > 180.0)  // the EDT works in 16-bit lanes: (R + 1)^2 + R^2 must stay below 2^16 (edt_kernels.cu)
    return fail(c, B200SV_ERR_UNSUPPORTED, "max_propagation_distance / resolution exceeds 180 cells");
  g.R = (int)rc;
  g.max_d2 = (int)(rc * rc);
  c->n_voxels = (size_t)g.nx * g.ny * g.nz;
  if (c->n_voxels >= (1ull << 31))
    return fail(c, B200SV_ERR_UNSUPPORTED, "more than 2^31 voxels");
  return B200SV_OK;
}

int allocFields(b200sv_ctx* c)
{
  const GridDev& g = c->grid;
  const size_t occ_bytes = (size_t)g.nx * g.ny * g.wz * 4;
  // what the check kernel gathers: [voxel] = { world, self } side by side, so one load serves both checks
  CU(c, cudaMalloc(&c->d_combined, c->n_voxels * 2 * (size_t)c->compact_bytes));
  for (int w = 0; w < 2; ++w)
  {
    Field& F = c->fields[w];
    CU(c, cudaMalloc(&F.occ, occ_bytes));
    CU(c, cudaMalloc(&F.d2, c->n_voxels * 2));
    F.compact = static_cast<uint8_t*>(c->d_combined) + (size_t)w * c->compact_bytes;
    CU(c, cudaMalloc(&F.plane_busy, (size_t)g.nx * kPlaneBusyStride));
    CU(c, cudaMemsetAsync(F.occ, 0, occ_bytes, c->stream));
    CU(c, launchFillField(g, F.d2, F.compact, c->compact_bytes, c->stream));
    CU(c, cudaMemsetAsync(F.plane_busy, 0, (size_t)g.nx * kPlaneBusyStride, c->stream));
    if (c->cfg.use_signed_distance_field)
    {  // reset(): negative_distance_square_ = 0 everywhere (propagation_distance_field.cpp:506-524)
      CU(c, cudaMalloc(&F.nd2, c->n_voxels * 2));
      CU(c, cudaMemsetAsync(F.nd2, 0, c->n_voxels * 2, c->stream));
    }
  }
  CU(c, cudaMalloc(&c->d_tmp, c->n_voxels * 2));
  return B200SV_OK;
}

// The positive transform, and for a signed field (propagate_negative_) the transform of the complement
int edtPasses(b200sv_ctx* c, Field& F, cudaStream_t s)
{
  // the x pass writes whole {world, self} words where the OTHER field's plane is all free (no merge read)
  Field& other = c->fields[&F == &c->fields[0] ? 1 : 0];
  CU(c, launchEdt(c->grid, F.occ, c->d_tmp, F.d2, F.compact, c->compact_bytes, c->sm_count, s, false, F.plane_busy,
                  other.plane_busy));
  if (F.nd2)
    CU(c, launchEdt(c->grid, F.occ, c->d_tmp, F.nd2, nullptr, c->compact_bytes, c->sm_count, s, true));
  return B200SV_OK;
}

int rebuildField(b200sv_ctx* c, int which)
{
  Field& F = c->fields[which];
  if (F.prop)
    return fail(c, B200SV_ERR_UNSUPPORTED,
                "this field reproduces the reference's propagation (b200sv_field_set_propagation): only reset, add_points, "
                "remove_points and update_points change it");
  if (F.deferred)
  {  // several occupancy changes, ONE transform (b200sv_field_end_update runs it)
    F.dirty = true;
    return B200SV_OK;
  }
  CU(c, cudaEventRecord(c->ev0, c->stream));
  if (int rc = edtPasses(c, F, c->stream))
    return rc;
  CU(c, cudaEventRecord(c->ev1, c->stream));
  CU(c, cudaEventSynchronize(c->ev1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->stats.edt_ms = ms;
  return B200SV_OK;
}

int finishCreate(b200sv_ctx* c, int device, b200sv_ctx** out)
{
  if (device == B200SV_DEVICE_NONE)
  {  // host-only context: model + tables for inspection; every compute entry point refuses
    c->device = B200SV_DEVICE_NONE;
    int rc0 = validateModel(c);
    if (rc0 == B200SV_OK)
      rc0 = initGrid(c);
    if (rc0 == B200SV_OK)
      rc0 = buildTables(c);
    if (rc0 != B200SV_OK)
    {
      g_create_error = c->err;
      delete c;
      return rc0;
    }
    *out = c;
    return B200SV_OK;
  }
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
  {
    fail(nullptr, B200SV_ERR_NO_DEVICE,
         std::string("no CUDA device (") + (e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e)) +
             "); libb200sv has no CPU fallback");
    cudaGetLastError();
    delete c;
    return B200SV_ERR_NO_DEVICE;
  }
  if (device < 0 || device >= n_dev)
  {
    fail(nullptr, B200SV_ERR_INVALID, "device ordinal out of range");
    delete c;
    return B200SV_ERR_INVALID;
  }
  c->device = device;
  int rc = B200SV_OK;
  auto bail = [&](int code) {
    g_create_error = c->err;
    b200sv_destroy(c);
    return code;
  };
  if (cudaSetDevice(device) != cudaSuccess)
    return bail(fail(c, B200SV_ERR_CUDA, "cudaSetDevice failed"));
  cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
  cudaDeviceGetAttribute(&c->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
  // profiling knobs (same meaning as b200sv_set_tuning), so a bench run under ncu can pin a configuration
  if (const char* e = getenv("B200SV_WARPS"))
    c->tune_warps = atoi(e);
  if (const char* e = getenv("B200SV_CLUSTER_R"))
    c->cluster_radius = atof(e);
  if (const char* e = getenv("B200SV_BLOCKS"))
    c->tune_blocks = atoi(e);
  if (const char* e = getenv("B200SV_FAST"))
    c->fast_disabled = atoi(e) == 0;
  if ((rc = validateModel(c)) != B200SV_OK)
    return bail(rc);
  if ((rc = initGrid(c)) != B200SV_OK)
    return bail(rc);
  if ((rc = buildTables(c)) != B200SV_OK)
    return bail(rc);
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_go, getenv("B200SV_TRACE") ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&c->ev_user, cudaEventDisableTiming) != cudaSuccess)
    return bail(fail(c, B200SV_ERR_CUDA, "stream/event creation failed"));
  for (int i = 0; i < b200sv_ctx::kMaxChunks; ++i)
    if (cudaEventCreateWithFlags(&c->ev_chunk[i], getenv("B200SV_TRACE") ? cudaEventDefault : cudaEventDisableTiming) != cudaSuccess)
      return bail(fail(c, B200SV_ERR_CUDA, "event creation failed"));
  if ((rc = uploadTables(c)) != B200SV_OK)
    return bail(rc);
  if ((rc = allocFields(c)) != B200SV_OK)
    return bail(rc);
  if (cudaMalloc(&c->d_count, sizeof(unsigned)) != cudaSuccess || cudaMalloc(&c->d_sink, sizeof(unsigned)) != cudaSuccess)
    return bail(fail(c, B200SV_ERR_CUDA, "cudaMalloc failed"));
  cudaMemsetAsync(c->d_count, 0, sizeof(unsigned), c->stream);
  if (cudaStreamSynchronize(c->stream) != cudaSuccess)
    return bail(fail(c, B200SV_ERR_CUDA, std::string("device initialisation failed: ") + cudaGetErrorString(cudaGetLastError())));
  *out = c;
  return B200SV_OK;
}

int addOrRemove(b200sv_ctx* c, int which, const double* d_xyz, size_t n, bool set, cudaStream_t s)
{
  CU(c, launchScatter(d_xyz, n, c->grid, c->fields[which].occ, set, s));
  return B200SV_OK;
}

template <typename FT>
CheckArgs<FT> baseArgs(b200sv_ctx* c, const double* d_q, size_t n, uint32_t flags, uint32_t* d_bitmap)
{
  CheckArgs<FT> a{};
  a.q = d_q;
  a.n_states = n;
  a.state_base = 0;
  a.field = c->d_combined;
  a.flags = flags & 7u;
  a.bitmap = d_bitmap;
  a.flag_list = c->d_list.p;
  a.flag_count = c->d_count;
  a.list_cap = (unsigned)std::min<size_t>(c->d_list.cap, 0xffffffffu);
  a.list = c->d_list.p;
  a.list_count = c->d_count;
  return a;
}

// fast (FP32) or exact-for-all (FP64) pass over states [base, base + n) of the batch that starts at d_q0 / d_bitmap0
template <typename FT>
int launchMain(b200sv_ctx* c, const double* d_q0, uint32_t* d_bitmap0, size_t base, size_t n, uint32_t flags, cudaStream_t s)
{
  const size_t V = c->robot.group_var_index.size();
  CheckArgs<FT> a = baseArgs<FT>(c, d_q0 + base * V, n, flags, d_bitmap0 + base / 32);
  a.state_base = base;
  const bool fp64 = (flags & B200SV_CHECK_FP64) != 0;
  if (!fp64 && c->fast_ok && !c->fast_disabled)
  {
    const b200sv_ctx::LaunchCfg& k = c->cfg_fast;
    a.model = c->d_blob_fast;
    a.model_bytes = (int)c->blob_fast.size();
    a.t_scratch = c->fast_t_global ? c->d_t_scratch : nullptr;
    a.rep_scratch = c->d_rep_scratch;
    a.hier = c->fast_n_linkpairs > 0;
    a.n_group_vars = (int)c->robot.group_var_index.size();
    a.wide = c->fast_wide ? 1 : 0;
    // persistent grid, but no more CTAs than there are tiles of 32 states to hand out (small batches: latency)
    const size_t tiles = (n + 31) / 32, ctas_needed = (tiles + k.warps - 1) / k.warps;
    const int grid = (int)std::min<size_t>((size_t)c->sm_count * k.blocks_per_sm, std::max<size_t>(1, ctas_needed));
    CU(c, launchFastCheck<FT>(a, grid, k.warps * 32, k.smem, s));
    return B200SV_OK;
  }
  const b200sv_ctx::LaunchCfg& k = fp64 ? c->cfg_d : c->cfg_f;
  a.model = fp64 ? c->d_blob_d : c->d_blob_f;
  a.model_bytes = (int)(fp64 ? c->blob_d.size() : c->blob_f.size());
  CU(c, launchCheck<FT>(a, fp64, false, c->sm_count * k.blocks_per_sm, k.warps * 32, k.smem, s));
  return B200SV_OK;
}

// exact pass over the states the fast pass could not decide; the count stays on the device, so the launch is
// unconditional
template <typename FT>
int launchRecheck(b200sv_ctx* c, const double* d_q0, uint32_t* d_bitmap0, size_t n_total, uint32_t flags, cudaStream_t s)
{
  CheckArgs<FT> a = baseArgs<FT>(c, d_q0, n_total, flags, d_bitmap0);
  a.model = c->d_blob_d;
  a.model_bytes = (int)c->blob_d.size();
  static const bool warp_per_state = !(getenv("B200SV_RECHECK_WARP") && atoi(getenv("B200SV_RECHECK_WARP")) == 0);
  if (warp_per_state && (size_t)align16(a.model_bytes) + 4 * (size_t)c->state_stride_d * 8 <= (size_t)c->smem_optin)
    CU(c, launchRecheckWarp<FT>(a, (int)std::min<size_t>((size_t)c->sm_count * 2, std::max<size_t>(1, (n_total + 3) / 4)),
                                c->state_stride_d, s));  // 4 warps per CTA, one listed state per warp: never more CTAs than states / 4
  else
    CU(c, launchCheck<FT>(a, true, true, c->sm_count * c->cfg_list.blocks_per_sm, c->cfg_list.warps * 32,
                          c->cfg_list.smem, s));
  return B200SV_OK;
}

template <typename FT>
int runCheck(b200sv_ctx* c, const double* d_q, size_t n, uint32_t flags, uint32_t* d_bitmap, cudaStream_t s)
{
  if (flags & B200SV_CHECK_FP64)
    return launchMain<FT>(c, d_q, d_bitmap, 0, n, flags, s);
  CU(c, cudaMemsetAsync(c->d_count, 0, sizeof(unsigned), s));
  int rc = launchMain<FT>(c, d_q, d_bitmap, 0, n, flags, s);
  if (rc != B200SV_OK)
    return rc;
  return launchRecheck<FT>(c, d_q, d_bitmap, n, flags, s);
}

// Host-buffer entry: the H2D copy of the states is cut into chunks on a second stream so the check of chunk i
// overlaps the copy of chunk i + 1.  The check kernel is a persistent grid in which every warp takes tiles of 32 states
// (one state per lane, ~36 us per tile on C2): a launch costs whole ROUNDS of the grid, so a chunk that is not a multiple
// of the wave (grid x warps x 32 states) pays for a round it only partly uses -- uniform 128 Ki-state chunks spent twice
// the kernel time of one launch over the batch (B200SV_TRACE=1 prints the timeline).  Chunks are therefore whole waves:
// 1, 2, 3, then 4 (or more, to stay within kMaxChunks) waves (the first copy cannot overlap anything, so it is short; later chunks amortise launch and DMA
// set-up), one wave before the end (the last kernel cannot overlap anything either), and the ragged remainder last --
// the rounds add up to those of a single launch.  Sizes are multiples of 32 so bitmap words do not straddle chunks.
std::vector<size_t> chunkPlan(size_t n, size_t wave)
{
  std::vector<size_t> plan;
  if (const char* e = getenv("B200SV_CHUNK"))
  {  // experiments: uniform chunks of this many states
    size_t chunk = std::max<size_t>(32, (size_t)atol(e) & ~(size_t)31);
    while ((n + chunk - 1) / chunk > (size_t)b200sv_ctx::kMaxChunks)
      chunk *= 2;
    for (size_t base = 0; base < n; base += chunk)
      plan.push_back(std::min(chunk, n - base));
    return plan;
  }
  wave = std::max<size_t>(32, wave & ~(size_t)31);
  const size_t full = n / wave, rest = n % wave;
  if (full == 0)
    return { n };
  const size_t tail = full >= 3 ? 1 : 0;
  size_t body = full - tail;
  const size_t mid = std::max<size_t>(4, (body + 22) / 23);  // 3 ramp chunks + at most 24 of these + 2: within kMaxChunks
  for (size_t step = 1; body > 0; ++step)
  {
    const size_t t = std::min(step <= 3 ? step : mid, body);
    plan.push_back(t * wave);
    body -= t;
  }
  if (tail)
    plan.push_back(wave);
  if (rest)
    plan.push_back(rest);
  return plan;
}

template <typename FT>
int runCheckHost(b200sv_ctx* c, const void* h_q_any, size_t n, uint32_t flags, bool f32 = false)
{
  const double* h_q = static_cast<const double*>(h_q_any);
  const float* h_qf = static_cast<const float*>(h_q_any);
  const size_t V = c->robot.group_var_index.size();
  const b200sv_ctx::LaunchCfg& lc = (flags & B200SV_CHECK_FP64) ? c->cfg_d : (c->fast_ok && !c->fast_disabled) ? c->cfg_fast : c->cfg_f;
  const std::vector<size_t> plan = chunkPlan(n, (size_t)c->sm_count * lc.blocks_per_sm * lc.warps * 32);
  const bool fp64 = (flags & B200SV_CHECK_FP64) != 0;
  // B200SV_TRACE=1: print when every chunk's copy and kernel finished (debug; adds events and a synchronisation)
  static const bool trace = getenv("B200SV_TRACE") != nullptr;
  static cudaEvent_t tr_k[b200sv_ctx::kMaxChunks + 1] = { nullptr };
  if (trace && !tr_k[0])
    for (auto& e : tr_k)
      cudaEventCreate(&e);
  if (!fp64)
    CU(c, cudaMemsetAsync(c->d_count, 0, sizeof(unsigned), c->stream));
  CU(c, cudaEventRecord(c->ev_go, c->stream));
  CU(c, cudaStreamWaitEvent(c->copy_stream, c->ev_go, 0));  // earlier work on the compute stream may still read d_q
  size_t base = 0;
  for (size_t i = 0; i < plan.size(); base += plan[i], ++i)
  {
    const size_t m = plan[i];
    if (f32)
      CU(c, cudaMemcpyAsync(c->d_qf.p + base * V, h_qf + base * V, m * V * sizeof(float), cudaMemcpyHostToDevice,
                            c->copy_stream));
    else
      CU(c, cudaMemcpyAsync(c->d_q.p + base * V, h_q + base * V, m * V * sizeof(double), cudaMemcpyHostToDevice,
                            c->copy_stream));
    CU(c, cudaEventRecord(c->ev_chunk[i], c->copy_stream));
    CU(c, cudaStreamWaitEvent(c->stream, c->ev_chunk[i], 0));
    if (f32)  // widen on the device: q = double(float), exactly what a float-holding caller would pass to the f64 entry
      CU(c, launchWidenStates(c->d_qf.p + base * V, c->d_q.p + base * V, m * V, c->stream));
    int rc = launchMain<FT>(c, c->d_q.p, c->d_bitmap.p, base, m, flags, c->stream);
    if (rc != B200SV_OK)
      return rc;
    if (trace)
      cudaEventRecord(tr_k[i], c->stream);
  }
  int rc = fp64 ? B200SV_OK : launchRecheck<FT>(c, c->d_q.p, c->d_bitmap.p, n, flags, c->stream);
  if (trace)
  {
    cudaEventRecord(tr_k[plan.size()], c->stream);
    cudaStreamSynchronize(c->stream);
    fprintf(stderr, "[b200sv trace] %s, %zu states:", f32 ? "f32" : "f64", n);
    for (size_t i = 0; i < plan.size(); ++i)
    {
      float tc = 0.f, tk = 0.f;
      cudaEventElapsedTime(&tc, c->ev_go, c->ev_chunk[i]);
      cudaEventElapsedTime(&tk, c->ev_go, tr_k[i]);
      fprintf(stderr, " [%zu: copied %.0f us, checked %.0f us]", plan[i], tc * 1e3, tk * 1e3);
    }
    float tr = 0.f;
    cudaEventElapsedTime(&tr, c->ev_go, tr_k[plan.size()]);
    fprintf(stderr, " recheck done %.0f us\n", tr * 1e3);
  }
  return rc;
}

// batched FK of device-resident states into device-resident 4x4 column-major doubles, all model links
int fkLaunch(b200sv_ctx* c, const double* d_q, size_t n, bool fp64, double* d_out, cudaStream_t s)
{
  FkArgs a{};
  a.model = fp64 ? c->d_blob_d : c->d_blob_f;
  a.model_bytes = (int)(fp64 ? c->blob_d.size() : c->blob_f.size());
  a.q = d_q;
  a.n_states = n;
  a.n_model_links = c->robot.n_links;
  a.link_slot = c->d_link_slot;
  a.const_T = c->d_const_T;
  a.out = d_out;
  const int elem = fp64 ? 8 : 4;
  const int ss = fp64 ? c->state_stride_d : c->state_stride_f;
  int warps = 4;
  while (warps > 1 && align16(a.model_bytes) + (size_t)warps * 32 * ss * elem > (size_t)c->smem_optin)
    --warps;
  const size_t smem = align16(a.model_bytes) + (size_t)warps * 32 * ss * elem;
  // as many CTAs as fit an SM at this much shared memory, times the SM count: a persistent grid over the state tiles
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(16, (size_t)(228 * 1024) / (smem + 1024)));
  const int grid = (int)std::min<size_t>((n + 32 * warps - 1) / (32 * warps), (size_t)c->sm_count * per_sm);
  CU(c, launchFk(a, fp64, grid, warps * 32, smem, s));
  return B200SV_OK;
}

int checkDevice(b200sv_ctx* c, const double* d_q, size_t n, uint32_t flags, uint32_t* d_bitmap, cudaStream_t s)
{
  if ((flags & 7u) == 0)
    return fail(c, B200SV_ERR_INVALID, "flags select no check");
  if (n >= (1ull << 32))
    return fail(c, B200SV_ERR_INVALID, "more than 2^32 - 1 states in one call");
  CU(c, c->d_list.ensure(n));
  return c->compact_bytes == 1 ? runCheck<uint8_t>(c, d_q, n, flags, d_bitmap, s) :
                                 runCheck<uint16_t>(c, d_q, n, flags, d_bitmap, s);
}
}  // namespace

// ================================================================================================================
extern "C" {

int b200sv_create(int device, const b200sv_robot* robot, const b200sv_collision_model* model,
                  const b200sv_field_cfg* field, b200sv_ctx** out)
{
  if (!robot || !model || !field || !out)
    return fail(nullptr, B200SV_ERR_INVALID, "null argument");
  *out = nullptr;
  if (robot->n_links <= 0 || robot->n_vars < 0 || robot->n_group_vars <= 0 || model->n_glinks < 0)
    return fail(nullptr, B200SV_ERR_INVALID, "empty robot or group");
  auto* c = new b200sv_ctx();
  HostRobot& r = c->robot;
  r.n_links = robot->n_links;
  r.n_vars = robot->n_vars;
  r.parent.assign(robot->parent, robot->parent + r.n_links);
  r.joint_type.assign(robot->joint_type, robot->joint_type + r.n_links);
  r.first_var.assign(robot->first_var, robot->first_var + r.n_links);
  r.joint_axis.assign(robot->joint_axis, robot->joint_axis + 3 * r.n_links);
  r.joint_origin.assign(robot->joint_origin, robot->joint_origin + 16 * r.n_links);
  r.base_positions.assign(robot->base_positions, robot->base_positions + r.n_vars);
  r.group_var_index.assign(robot->group_var_index, robot->group_var_index + robot->n_group_vars);
  if (robot->n_mimic > 0)
  {
    r.mimic_dst.assign(robot->mimic_dst, robot->mimic_dst + robot->n_mimic);
    r.mimic_src.assign(robot->mimic_src, robot->mimic_src + robot->n_mimic);
    r.mimic_factor.assign(robot->mimic_factor, robot->mimic_factor + robot->n_mimic);
    r.mimic_offset.assign(robot->mimic_offset, robot->mimic_offset + robot->n_mimic);
  }
  HostCollisionModel& m = c->cm;
  m.n_glinks = model->n_glinks;
  const int n = m.n_glinks;
  m.glink_index.assign(model->glink_index, model->glink_index + n);
  m.self_enabled.assign(model->self_enabled, model->self_enabled + n);
  m.intra_enabled.assign(model->intra_enabled, model->intra_enabled + (size_t)n * n);
  m.sphere_start.assign(model->sphere_start, model->sphere_start + n + 1);
  const int ns = n ? m.sphere_start[n] : 0;
  if (ns < 0)
  {
    delete c;
    return fail(nullptr, B200SV_ERR_INVALID, "negative sphere count");
  }
  m.sphere_xyzr.assign(model->sphere_xyzr, model->sphere_xyzr + 4 * (size_t)ns);
  m.bounding_sphere.assign(model->bounding_sphere, model->bounding_sphere + 4 * (size_t)n);
  if (model->body_id)
    m.body_id.assign(model->body_id, model->body_id + n);
  c->cfg = *field;
  return finishCreate(c, device, out);
}

int b200sv_create_from_urdf(int device, const char* urdf_xml, const char* srdf_xml, const char* group_name,
                            const b200sv_field_cfg* field, b200sv_ctx** out)
{
  if (!urdf_xml || !group_name || !field || !out)
    return fail(nullptr, B200SV_ERR_INVALID, "null argument");
  *out = nullptr;
  auto* c = new b200sv_ctx();
  std::string err;
  bool parse_error = false;
  if (!loadUrdfSrdf(urdf_xml, srdf_xml ? srdf_xml : "", group_name, field->resolution, c->robot, c->cm, c->self_points,
                    c->self_local, c->link_points, err, parse_error))
  {
    delete c;
    return fail(nullptr, parse_error ? B200SV_ERR_PARSE : B200SV_ERR_INVALID, err);
  }
  c->cfg = *field;
  int rc = finishCreate(c, device, out);
  if (rc != B200SV_OK)
    return rc;
  // self field from the non-group links (generateDistanceFieldCacheEntry :907-953)
  if (!c->self_points.empty() && c->device >= 0)
  {
    rc = b200sv_field_add_points(c, B200SV_FIELD_SELF, c->self_points.data(), c->self_points.size() / 3);
    if (rc != B200SV_OK)
    {
      g_create_error = c->err;
      b200sv_destroy(c);
      *out = nullptr;
    }
  }
  return rc;
}

void b200sv_destroy(b200sv_ctx* c)
{
  if (!c)
    return;
  if (c->device < 0)
  {
    delete c;
    return;
  }
  cudaSetDevice(c->device);
  if (c->user_pending)
    cudaEventSynchronize(c->ev_user);
  quitLinger(c);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  for (int w = 0; w < 2; ++w)
  {
    Field& F = c->fields[w];
    if (F.d2)
      cudaFree(F.d2);
    if (F.nd2)
      cudaFree(F.nd2);
    if (F.occ)
      cudaFree(F.occ);
    if (F.plane_busy)
      cudaFree(F.plane_busy);
    propDestroy(F.prop);
    F.prop = nullptr;
  }
  if (c->d_combined) cudaFree(c->d_combined);
  if (c->d_tmp) cudaFree(c->d_tmp);
  if (c->d_sqrt_table) cudaFree(c->d_sqrt_table);
  if (c->h_small) cudaFreeHost(c->h_small);
  if (c->h_count) cudaFreeHost(c->h_count);
  if (c->h_mail) cudaFreeHost(c->h_mail);
  if (c->h_edge_lat) cudaFreeHost(c->h_edge_lat);
  if (c->h_lat) cudaFreeHost(c->h_lat);
  if (c->d_small) cudaFree(c->d_small);
  if (c->d_blob_f) cudaFree(c->d_blob_f);
  if (c->d_blob_d) cudaFree(c->d_blob_d);
  if (c->d_blob_fast) cudaFree(c->d_blob_fast);
  if (c->d_t_scratch) cudaFree(c->d_t_scratch);
  if (c->d_rep_scratch) cudaFree(c->d_rep_scratch);
  c->d_edges.release(); c->d_edge_q.release(); c->d_edge_aux.release(); c->d_edge_count.release();
  c->d_edge_offset.release(); c->d_edge_state_bits.release(); c->d_edge_bits.release(); c->d_edge_first.release();
  c->d_edge_cont.release();
  c->d_g_dist.release(); c->d_g_grad.release(); c->d_g_closest.release(); c->d_g_centers.release();
  c->d_g_types.release(); c->d_g_self.release();
  for (auto& lf : c->link_fields)
    if (lf.d2) cudaFree(lf.d2);
  if (c->d_lf) cudaFree(c->d_lf);
  c->d_lf_enabled.release();
  c->d_c_start.release(); c->d_c_slot.release(); c->d_c_count.release(); c->d_c_bs.release(); c->d_c_intra.release(); c->d_c_body.release();
  c->d_c_coll.release(); c->d_c_out.release();
  if (c->d_link_slot) cudaFree(c->d_link_slot);
  if (c->d_const_T) cudaFree(c->d_const_T);
  if (c->d_count) cudaFree(c->d_count);
  if (c->d_sink) cudaFree(c->d_sink);
  c->d_q.release();
  c->d_qf.release();
  c->d_bitmap.release();
  c->d_list.release();
  c->d_points.release();
  c->d_fk.release();
  c->d_i32.release();
  for (int i = 0; i < b200sv_ctx::kMaxChunks; ++i)
    if (c->ev_chunk[i]) cudaEventDestroy(c->ev_chunk[i]);
  if (c->ev_go) cudaEventDestroy(c->ev_go);
  if (c->ev_user) cudaEventDestroy(c->ev_user);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* b200sv_last_error(const b200sv_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int b200sv_get_info(const b200sv_ctx* c, b200sv_info* info)
{
  if (!c || !info)
    return B200SV_ERR_INVALID;
  info->n_links = c->robot.n_links;
  info->n_vars = c->robot.n_vars;
  info->n_group_vars = (int)c->robot.group_var_index.size();
  info->n_glinks = c->cm.n_glinks;
  info->n_spheres = c->n_spheres;
  info->n_link_pairs = c->n_link_pairs;
  info->num_cells[0] = c->grid.nx;
  info->num_cells[1] = c->grid.ny;
  info->num_cells[2] = c->grid.nz;
  info->max_distance_sq = c->grid.max_d2;
  info->device = c->device;
  info->field_bytes_per_voxel = c->compact_bytes;
  info->sm_count = c->sm_count;
  return B200SV_OK;
}

int b200sv_get_stats(b200sv_ctx* c, b200sv_stats* stats)
{
  if (!c || !stats)
    return B200SV_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  if (c->device < 0)
  {
    *stats = c->stats;
    return B200SV_OK;
  }
  cudaSetDevice(c->device);
  if (c->stats_pending)
  {
    if (c->user_pending)
      CU(c, cudaEventSynchronize(c->ev_user));
    CU(c, cudaStreamSynchronize(c->stream));
    unsigned cnt = 0;
    CU(c, cudaMemcpy(&cnt, c->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost));
    c->stats.n_rechecked = cnt;
    c->stats_pending = false;
  }
  *stats = c->stats;
  return B200SV_OK;
}

int b200sv_get_robot_arrays(const b200sv_ctx* c, int32_t* parent, int32_t* joint_type, double* joint_axis,
                            double* joint_origin, int32_t* first_var, double* base_positions, int32_t* group_var_index)
{
  if (!c)
    return B200SV_ERR_INVALID;
  const HostRobot& r = c->robot;
  if (parent) std::copy(r.parent.begin(), r.parent.end(), parent);
  if (joint_type) std::copy(r.joint_type.begin(), r.joint_type.end(), joint_type);
  if (joint_axis) std::copy(r.joint_axis.begin(), r.joint_axis.end(), joint_axis);
  if (joint_origin) std::copy(r.joint_origin.begin(), r.joint_origin.end(), joint_origin);
  if (first_var) std::copy(r.first_var.begin(), r.first_var.end(), first_var);
  if (base_positions) std::copy(r.base_positions.begin(), r.base_positions.end(), base_positions);
  if (group_var_index) std::copy(r.group_var_index.begin(), r.group_var_index.end(), group_var_index);
  return B200SV_OK;
}

int b200sv_get_model_arrays(const b200sv_ctx* c, int32_t* glink_index, uint8_t* self_enabled, uint8_t* intra_enabled,
                            int32_t* sphere_start, double* sphere_xyzr, double* bounding_sphere)
{
  if (!c)
    return B200SV_ERR_INVALID;
  const HostCollisionModel& m = c->cm;
  if (glink_index) std::copy(m.glink_index.begin(), m.glink_index.end(), glink_index);
  if (self_enabled) std::copy(m.self_enabled.begin(), m.self_enabled.end(), self_enabled);
  if (intra_enabled) std::copy(m.intra_enabled.begin(), m.intra_enabled.end(), intra_enabled);
  if (sphere_start) std::copy(m.sphere_start.begin(), m.sphere_start.end(), sphere_start);
  if (sphere_xyzr) std::copy(m.sphere_xyzr.begin(), m.sphere_xyzr.end(), sphere_xyzr);
  if (bounding_sphere) std::copy(m.bounding_sphere.begin(), m.bounding_sphere.end(), bounding_sphere);
  return B200SV_OK;
}

int b200sv_get_link_names(const b200sv_ctx* c, char* buf, size_t cap)
{
  if (!c || !buf || cap == 0)
    return B200SV_ERR_INVALID;
  std::string s;
  for (size_t i = 0; i < c->robot.link_names.size(); ++i)
    s += (i ? "\n" : "") + c->robot.link_names[i];
  if (s.size() + 1 > cap)
    return B200SV_ERR_INVALID;
  memcpy(buf, s.c_str(), s.size() + 1);
  return B200SV_OK;
}

int b200sv_get_group_bounds(const b200sv_ctx* c, double* lower, double* upper)
{
  if (!c || !lower || !upper)
    return B200SV_ERR_INVALID;
  if (c->robot.group_lower.size() != c->robot.group_var_index.size())
    return fail(const_cast<b200sv_ctx*>(c), B200SV_ERR_UNSUPPORTED, "bounds are only known on the URDF path");
  std::copy(c->robot.group_lower.begin(), c->robot.group_lower.end(), lower);
  std::copy(c->robot.group_upper.begin(), c->robot.group_upper.end(), upper);
  return B200SV_OK;
}

int b200sv_get_sphere_thresholds(const b200sv_ctx* c, int32_t* thresholds)
{
  if (!c || !thresholds)
    return B200SV_ERR_INVALID;
  const auto* h = reinterpret_cast<const ModelHeader<double>*>(c->blob_d.data());
  const auto* S = reinterpret_cast<const SphereRec<double>*>(c->blob_d.data() + h->off_spheres);
  for (int i = 0; i < h->n_spheres; ++i)
    thresholds[i] = S[i].thr;
  return B200SV_OK;
}

int b200sv_get_self_points(const b200sv_ctx* c, double* xyz, size_t cap_points, size_t* n_out)
{
  if (!c || !n_out)
    return B200SV_ERR_INVALID;
  *n_out = c->self_points.size() / 3;
  if (xyz)
  {
    if (cap_points < *n_out)
      return B200SV_ERR_INVALID;
    std::copy(c->self_points.begin(), c->self_points.end(), xyz);
  }
  return B200SV_OK;
}

// ---- fields ------------------------------------------------------------------------------------------------------
static int checkWhich(b200sv_ctx* c, int which)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (which != B200SV_FIELD_WORLD && which != B200SV_FIELD_SELF)
    return fail(c, B200SV_ERR_INVALID, "field selector must be B200SV_FIELD_WORLD or B200SV_FIELD_SELF");
  return needDevice(c);
}

// entry points that edit the occupancy bits directly have no meaning for a field in reference-propagation mode
static int notInReferenceMode(b200sv_ctx* c, int which)
{
  if (c->fields[which].prop)
    return fail(c, B200SV_ERR_UNSUPPORTED,
                "this field reproduces the reference's propagation (b200sv_field_set_propagation): only reset, add_points, "
                "remove_points and update_points change it");
  return B200SV_OK;
}

// the propagation's distances become the field every kernel reads: occupancy bits, d2, this field's compact elements
static int publishPropagation(b200sv_ctx* c, int which, cudaStream_t s)
{
  Field& F = c->fields[which];
  CU(c, cudaMemsetAsync(F.occ, 0, (size_t)c->grid.nx * c->grid.ny * c->grid.wz * 4, s));
  CU(c, launchInjectD2(c->grid, propDistances(F.prop), F.occ, F.d2, F.compact, c->compact_bytes, s));
  CU(c, cudaMemsetAsync(F.plane_busy, 1, (size_t)c->grid.nx * kPlaneBusyStride, s));
  if (F.nd2)
    CU(c, launchInjectNegD2(c->grid, propNegativeDistances(F.prop), F.nd2, s));
  return B200SV_OK;
}

int b200sv_field_set_propagation(b200sv_ctx* c, int which, int mode)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (mode != B200SV_PROPAGATION_EXACT && mode != B200SV_PROPAGATION_REFERENCE)
    return fail(c, B200SV_ERR_INVALID, "mode must be B200SV_PROPAGATION_EXACT or B200SV_PROPAGATION_REFERENCE");
  {
    std::lock_guard<std::mutex> lk(c->mu);
    enterCtx(c);
    Field& F = c->fields[which];
    CU(c, cudaStreamSynchronize(c->stream));
    if (mode == B200SV_PROPAGATION_REFERENCE && !F.prop)
    {
      cudaError_t e = propCreate(c->grid, c->cfg.use_signed_distance_field != 0, &F.prop);
      if (e != cudaSuccess)
        return fail(c, e == cudaErrorInvalidConfiguration ? B200SV_ERR_UNSUPPORTED : B200SV_ERR_CUDA,
                    std::string("reference propagation: ") + cudaGetErrorString(e));
      F.deferred = F.dirty = false;
    }
    else if (mode == B200SV_PROPAGATION_EXACT && F.prop)
    {
      propDestroy(F.prop);
      F.prop = nullptr;
    }
  }
  rc = b200sv_field_reset(c, which);  // either way the field starts empty in its new mode ...
  if (rc != B200SV_OK || which != B200SV_FIELD_SELF)
    return rc;
  // ... except the self field, which is the context's own: the links' points the context holds are added again
  std::vector<double> pts;
  {
    std::lock_guard<std::mutex> lk(c->mu);
    pts = c->self_points;
  }
  return pts.empty() ? B200SV_OK : b200sv_field_add_points(c, which, pts.data(), pts.size() / 3);
}

int b200sv_field_propagation_stats(b200sv_ctx* c, int which, b200sv_propagation_stats* out)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!out)
    return fail(c, B200SV_ERR_INVALID, "null output");
  std::lock_guard<std::mutex> lk(c->mu);
  if (!c->fields[which].prop)
    return fail(c, B200SV_ERR_INVALID, "the field is not in reference-propagation mode");
  const PropStats& st = propStats(c->fields[which].prop);
  out->rounds = st.rounds;
  out->hazard_rounds = st.hazard_rounds;
  out->queue_entries = st.entries;
  out->backward_offers = st.backward_offers;
  out->queued_for_next_call = st.queued_for_next_call;
  return B200SV_OK;
}

int b200sv_field_reset(b200sv_ctx* c, int which)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  Field& F = c->fields[which];
  if (F.prop)
    CU(c, propReset(F.prop, c->stream));
  CU(c, cudaMemsetAsync(F.occ, 0, (size_t)c->grid.nx * c->grid.ny * c->grid.wz * 4, c->stream));
  CU(c, launchFillField(c->grid, F.d2, F.compact, c->compact_bytes, c->stream));
  CU(c, cudaMemsetAsync(F.plane_busy, 0, (size_t)c->grid.nx * kPlaneBusyStride, c->stream));
  if (F.nd2)
    CU(c, cudaMemsetAsync(F.nd2, 0, c->n_voxels * 2, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return B200SV_OK;
}

int b200sv_field_begin_update(b200sv_ctx* c, int which)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  c->fields[which].deferred = true;
  return B200SV_OK;
}

int b200sv_field_end_update(b200sv_ctx* c, int which)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  Field& F = c->fields[which];
  F.deferred = false;
  if (!F.dirty)
    return B200SV_OK;
  F.dirty = false;
  return rebuildField(c, which);
}

static int pointsHost(b200sv_ctx* c, int which, const double* xyz, size_t n, bool set)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (n && !xyz)
    return fail(c, B200SV_ERR_INVALID, "null point array");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (c->fields[which].prop)
  {  // addPointsToField / removePointsFromField as the reference runs them, call by call
    PropField* pf = c->fields[which].prop;
    CU(c, c->d_points.ensure(3 * n + 3));
    if (n)
      CU(c, cudaMemcpyAsync(c->d_points.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, cudaEventRecord(c->ev0, c->stream));
    CU(c, set ? propAddPoints(pf, c->d_points.p, n, c->stream) : propRemovePoints(pf, c->d_points.p, n, c->stream));
    if ((rc = publishPropagation(c, which, c->stream)) != B200SV_OK)
      return rc;
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    c->stats.edt_ms = ms;
    return B200SV_OK;
  }
  if (n)
  {
    CU(c, c->d_points.ensure(3 * n));
    CU(c, cudaMemcpyAsync(c->d_points.p, xyz, 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = addOrRemove(c, which, c->d_points.p, n, set, c->stream)) != B200SV_OK)
      return rc;
  }
  return rebuildField(c, which);
}

int b200sv_field_add_points(b200sv_ctx* c, int which, const double* xyz, size_t n)
{
  return pointsHost(c, which, xyz, n, true);
}

int b200sv_field_remove_points(b200sv_ctx* c, int which, const double* xyz, size_t n)
{
  return pointsHost(c, which, xyz, n, false);
}

int b200sv_field_update_points(b200sv_ctx* c, int which, const double* old_xyz, size_t n_old, const double* new_xyz, size_t n_new)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((n_old && !old_xyz) || (n_new && !new_xyz))
    return fail(c, B200SV_ERR_INVALID, "null point array");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (c->fields[which].prop)
  {
    CU(c, c->d_points.ensure(3 * (n_old + n_new) + 3));
    double* d_old = c->d_points.p;
    double* d_new = c->d_points.p + 3 * n_old;
    if (n_old)
      CU(c, cudaMemcpyAsync(d_old, old_xyz, 3 * n_old * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if (n_new)
      CU(c, cudaMemcpyAsync(d_new, new_xyz, 3 * n_new * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, propUpdatePoints(c->fields[which].prop, d_old, n_old, d_new, n_new, c->stream));
    if ((rc = publishPropagation(c, which, c->stream)) != B200SV_OK)
      return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return B200SV_OK;
  }
  const size_t n_words = (size_t)c->grid.nx * c->grid.ny * c->grid.wz;
  if (2 * n_words * 4 > c->n_voxels * 2)
    return fail(c, B200SV_ERR_UNSUPPORTED, "grid too shallow for the update scratch (nz < 16)");
  // two scratch bitmaps in the EDT's intermediate buffer: the voxels of the old and of the new point set
  uint32_t* old_bits = reinterpret_cast<uint32_t*>(c->d_tmp);
  uint32_t* new_bits = old_bits + n_words;
  CU(c, cudaMemsetAsync(old_bits, 0, 2 * n_words * 4, c->stream));
  CU(c, c->d_points.ensure(3 * std::max(n_old, n_new) + 3));
  if (n_old)
  {
    CU(c, cudaMemcpyAsync(c->d_points.p, old_xyz, 3 * n_old * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, launchScatter(c->d_points.p, n_old, c->grid, old_bits, true, c->stream));
  }
  if (n_new)
  {
    CU(c, cudaMemcpyAsync(c->d_points.p, new_xyz, 3 * n_new * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, launchScatter(c->d_points.p, n_new, c->grid, new_bits, true, c->stream));
  }
  CU(c, launchUpdateOcc(c->fields[which].occ, old_bits, new_bits, n_words, c->stream));
  return rebuildField(c, which);
}

int b200sv_field_add_sphere(b200sv_ctx* c, int which, const double* center, double radius)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  if (!center || !(radius >= 0.0))
    return fail(c, B200SV_ERR_INVALID, "null centre or negative radius");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  // findInternalPointsConvex (find_internal_points.cpp:44-54): the bounding sphere of a bodies::Sphere is the sphere;
  // start = floor((c - r - res) / res) * res, then `pt += res` while pt <= c + r + res -- accumulated, not k * res
  const double res = c->cfg.resolution;
  std::vector<double> axes;
  int cnt[3];
  for (int k = 0; k < 3; ++k)
  {
    const double e = center[k] + radius + res;
    int n = 0;
    for (double v = floor((center[k] - radius - res) / res) * res; v <= e; v += res, ++n)
      axes.push_back(v);
    cnt[k] = n;
  }
  if ((double)cnt[0] * cnt[1] * cnt[2] > 4.0e9)
    return fail(c, B200SV_ERR_INVALID, "sphere spans more than 4e9 lattice points");
  CU(c, c->d_points.ensure(axes.size() + 1));
  CU(c, cudaMemcpyAsync(c->d_points.p, axes.data(), axes.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, launchSphereShape(c->d_points.p, cnt[0], cnt[1], cnt[2], center, radius * radius, c->grid, c->fields[which].occ,
                          c->stream));
  CU(c, cudaStreamSynchronize(c->stream));  // `axes` lives on this stack frame
  return rebuildField(c, which);
}

int b200sv_field_add_shape(b200sv_ctx* c, int which, const b200sv_shape* sh, int remove)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  if (!sh)
    return fail(c, B200SV_ERR_INVALID, "null shape");
  if (sh->lattice_frame != B200SV_LATTICE_WORLD && sh->lattice_frame != B200SV_LATTICE_BODY)
    return fail(c, B200SV_ERR_INVALID, "lattice_frame must be B200SV_LATTICE_WORLD or B200SV_LATTICE_BODY");
  // bodies::{Sphere,Cylinder,Box}::updateInternalData with scale 1, padding 0 (what DistanceField::getShapePoints and
  // BodyDecomposition(shape, resolution, 0.0) use), and computeBoundingSphere: centre = pose translation, radius below
  BodyDev b{};
  b.type = sh->type;
  double radius = 0.0;
  switch (sh->type)
  {
    case B200SV_SHAPE_SPHERE:
      if (!(sh->dims[0] >= 0.0))
        return fail(c, B200SV_ERR_INVALID, "negative radius");
      b.dims[0] = sh->dims[0] * sh->dims[0];
      radius = sh->dims[0];
      break;
    case B200SV_SHAPE_CYLINDER:
    {
      if (!(sh->dims[0] >= 0.0) || !(sh->dims[1] >= 0.0))
        return fail(c, B200SV_ERR_INVALID, "negative cylinder dimension");
      const double length2 = 1.0 * sh->dims[1] / 2.0;
      b.dims[0] = sh->dims[0] * sh->dims[0];
      b.dims[1] = length2;
      radius = sqrt(length2 * length2 + b.dims[0]);
      break;
    }
    case B200SV_SHAPE_BOX:
    {
      for (int k = 0; k < 3; ++k)
      {
        if (!(sh->dims[k] >= 0.0))
          return fail(c, B200SV_ERR_INVALID, "negative box dimension");
        b.dims[k] = sh->dims[k] * (1.0 / 2.0);
      }
      radius = sqrt(b.dims[0] * b.dims[0] + b.dims[1] * b.dims[1] + b.dims[2] * b.dims[2]);
      break;
    }
    default:
      return fail(c, B200SV_ERR_UNSUPPORTED,
                  "shape type: spheres, cylinders and boxes are decomposed on the device; cones, planes and meshes go through "
                  "the caller's bodies::Body and b200sv_field_add_points, octrees through their node centres");
  }
  const bool body_frame = sh->lattice_frame == B200SV_LATTICE_BODY;
  static const double ident[9] = { 1, 0, 0, 0, 1, 0, 0, 0, 1 };
  for (int r = 0; r < 3; ++r)
  {
    for (int k = 0; k < 3; ++k)
    {
      b.R[3 * r + k] = body_frame ? ident[3 * r + k] : sh->pose[4 * r + k];
      b.P[3 * r + k] = sh->pose[4 * r + k];
    }
    b.c[r] = body_frame ? 0.0 : sh->pose[4 * r + 3];
    b.P[9 + r] = sh->pose[4 * r + 3];
  }
  b.posed = body_frame ? 1 : 0;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  // findInternalPointsConvex (find_internal_points.cpp:44-54): start = floor((c - r - res) / res) * res, then `pt += res`
  // while pt <= c + r + res -- accumulated, not k * res
  const double res = c->cfg.resolution;
  std::vector<double> axes;
  int cnt[3];
  for (int k = 0; k < 3; ++k)
  {
    const double e = b.c[k] + radius + res;
    int n = 0;
    for (double v = floor((b.c[k] - radius - res) / res) * res; v <= e; v += res, ++n)
      axes.push_back(v);
    cnt[k] = n;
  }
  if ((double)cnt[0] * cnt[1] * cnt[2] > 4.0e9)
    return fail(c, B200SV_ERR_INVALID, "shape spans more than 4e9 lattice points");
  CU(c, c->d_points.ensure(axes.size() + 1));
  CU(c, cudaMemcpyAsync(c->d_points.p, axes.data(), axes.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, launchBodyLattice(c->d_points.p, cnt[0], cnt[1], cnt[2], b, c->grid, c->fields[which].occ, remove == 0, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));  // `axes` lives on this stack frame
  return rebuildField(c, which);
}

int b200sv_field_add_convex(b200sv_ctx* c, int which, const b200sv_convex* cv, int remove)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  if (!cv || cv->n_planes < 1 || !cv->planes)
    return fail(c, B200SV_ERR_INVALID, "null convex body or no planes");
  if (cv->lattice_frame != B200SV_LATTICE_WORLD && cv->lattice_frame != B200SV_LATTICE_BODY)
    return fail(c, B200SV_ERR_INVALID, "lattice_frame must be B200SV_LATTICE_WORLD or B200SV_LATTICE_BODY");
  if (!(cv->bounding_sphere[3] >= 0.0))
    return fail(c, B200SV_ERR_INVALID, "negative bounding radius");
  const bool body_frame = cv->lattice_frame == B200SV_LATTICE_BODY;
  ConvexDev b{};
  b.n_planes = cv->n_planes;
  b.posed = body_frame ? 1 : 0;
  double R[9], t[3];
  for (int r = 0; r < 3; ++r)
  {
    for (int k = 0; k < 3; ++k)
    {
      R[3 * r + k] = cv->pose[4 * r + k];
      b.P[3 * r + k] = cv->pose[4 * r + k];
    }
    t[r] = cv->pose[4 * r + 3];
    b.P[9 + r] = cv->pose[4 * r + 3];
  }
  // i_pose_ = pose_.inverse() for an isometry: [R^T | -(R^T t)]; the identity when the lattice is laid out around the body
  double centre[3];
  for (int r = 0; r < 3; ++r)
  {
    for (int k = 0; k < 3; ++k)
      b.ipose[3 * r + k] = body_frame ? (r == k ? 1.0 : 0.0) : R[3 * k + r];
    b.ipose[9 + r] = body_frame ? 0.0 : -((R[0 + r] * t[0] + R[3 + r] * t[1]) + R[6 + r] * t[2]);
    // the lattice is laid out around the body's bounding sphere (computeBoundingSphere: pose_ * mesh centre)
    centre[r] = body_frame ? cv->bounding_sphere[r] :
                             ((R[3 * r] * cv->bounding_sphere[0] + R[3 * r + 1] * cv->bounding_sphere[1]) +
                              R[3 * r + 2] * cv->bounding_sphere[2]) + t[r];
  }
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const double res = c->cfg.resolution, radius = cv->bounding_sphere[3];
  std::vector<double> host;  // axes, then the planes
  int cnt[3];
  for (int k = 0; k < 3; ++k)
  {
    const double e = centre[k] + radius + res;
    int n = 0;
    for (double v = floor((centre[k] - radius - res) / res) * res; v <= e; v += res, ++n)
      host.push_back(v);
    cnt[k] = n;
  }
  if ((double)cnt[0] * cnt[1] * cnt[2] > 4.0e9)
    return fail(c, B200SV_ERR_INVALID, "body spans more than 4e9 lattice points");
  const size_t n_axes = host.size();
  host.insert(host.end(), cv->planes, cv->planes + 4 * (size_t)cv->n_planes);
  CU(c, c->d_points.ensure(host.size() + 1));
  CU(c, cudaMemcpyAsync(c->d_points.p, host.data(), host.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(c, launchConvexLattice(c->d_points.p, cnt[0], cnt[1], cnt[2], c->d_points.p + n_axes, b, c->grid, c->fields[which].occ,
                            remove == 0, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));  // `host` lives on this stack frame
  return rebuildField(c, which);
}

int b200sv_field_add_octree_leaves(b200sv_ctx* c, int which, const double* leaves, size_t n_leaves)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  if (n_leaves && !leaves)
    return fail(c, B200SV_ERR_INVALID, "null leaf array");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (n_leaves)
  {
    CU(c, c->d_points.ensure(4 * n_leaves));
    CU(c, cudaMemcpyAsync(c->d_points.p, leaves, 4 * n_leaves * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(c, launchOctreeLeaves(c->d_points.p, n_leaves, c->cfg.resolution, c->grid, c->fields[which].occ, c->stream));
  }
  return rebuildField(c, which);
}

int b200sv_field_add_points_device(b200sv_ctx* c, int which, const double* d_xyz, size_t n, void* cuda_stream)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  if ((rc = enterUser(c, s)) != B200SV_OK)
    return rc;
  if (c->fields[which].prop)
  {  // not asynchronous in this mode: the propagation synchronises the stream once per round
    CU(c, propAddPoints(c->fields[which].prop, d_xyz, n, s));
    if ((rc = publishPropagation(c, which, s)) != B200SV_OK)
      return rc;
    return leaveUser(c, s);
  }
  if ((rc = addOrRemove(c, which, d_xyz, n, true, s)) != B200SV_OK)
    return rc;
  if ((rc = edtPasses(c, c->fields[which], s)) != B200SV_OK)
    return rc;
  return leaveUser(c, s);
}

int b200sv_field_get_d2(b200sv_ctx* c, int which, int32_t* out)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!out)
    return fail(c, B200SV_ERR_INVALID, "null output");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  std::vector<uint16_t> h(c->n_voxels);
  CU(c, cudaMemcpyAsync(h.data(), c->fields[which].d2, c->n_voxels * 2, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < c->n_voxels; ++i)
    out[i] = h[i];
  return B200SV_OK;
}

int b200sv_field_get_neg_d2(b200sv_ctx* c, int which, int32_t* out)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!out)
    return fail(c, B200SV_ERR_INVALID, "null output");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (!c->fields[which].nd2)
  {  // an unsigned field: negative_distance_square_ stays at its reset value
    std::fill(out, out + c->n_voxels, 0);
    return B200SV_OK;
  }
  std::vector<uint16_t> h(c->n_voxels);
  CU(c, cudaMemcpyAsync(h.data(), c->fields[which].nd2, c->n_voxels * 2, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  for (size_t i = 0; i < c->n_voxels; ++i)
    out[i] = h[i];
  return B200SV_OK;
}

int b200sv_field_set_neg_d2(b200sv_ctx* c, int which, const int32_t* nd2)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!nd2)
    return fail(c, B200SV_ERR_INVALID, "null input");
  if (!c->fields[which].nd2)
    return fail(c, B200SV_ERR_INVALID, "the context's fields are unsigned (use_signed_distance_field = 0)");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  CU(c, c->d_i32.ensure(c->n_voxels));
  CU(c, cudaMemcpyAsync(c->d_i32.p, nd2, c->n_voxels * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, launchInjectNegD2(c->grid, c->d_i32.p, c->fields[which].nd2, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return B200SV_OK;
}

int b200sv_field_set_d2(b200sv_ctx* c, int which, const int32_t* d2)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if ((rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;
  if (!d2)
    return fail(c, B200SV_ERR_INVALID, "null input");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  Field& F = c->fields[which];
  CU(c, c->d_i32.ensure(c->n_voxels));
  CU(c, cudaMemcpyAsync(c->d_i32.p, d2, c->n_voxels * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaMemsetAsync(F.occ, 0, (size_t)c->grid.nx * c->grid.ny * c->grid.wz * 4, c->stream));
  CU(c, launchInjectD2(c->grid, c->d_i32.p, F.occ, F.d2, F.compact, c->compact_bytes, c->stream));
  CU(c, cudaMemsetAsync(F.plane_busy, 1, (size_t)c->grid.nx * kPlaneBusyStride, c->stream));  // injected values: any plane may be busy
  CU(c, cudaStreamSynchronize(c->stream));
  return B200SV_OK;
}

// ---- .df streams: PropagationDistanceField::writeToStream / readFromStream (propagation_distance_field.cpp:635-760) -----
// 7 ASCII header lines (resolution, size_xyz, origin_xyz = the grid CORNER), then zlib data: x-major, then y, then z packed
// 8 voxels per byte, bit zi LSB-first, 1 = occupied (distance_square_ == 0).  Distances are not stored; reading resets
// the field and rebuilds it from the occupied voxels.
int b200sv_field_write_df(b200sv_ctx* c, int which, uint8_t* out, size_t cap, size_t* n_out)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!n_out)
    return fail(c, B200SV_ERR_INVALID, "null n_out");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const GridDev& g = c->grid;
  const size_t rows = (size_t)g.nx * g.ny, zb = (size_t)(g.nz + 7) / 8;
  std::vector<uint32_t> occ(rows * g.wz);
  CU(c, cudaStreamSynchronize(c->stream));
  CU(c, cudaMemcpy(occ.data(), c->fields[which].occ, occ.size() * 4, cudaMemcpyDeviceToHost));
  std::vector<uint8_t> packed(rows * zb);
  for (size_t r = 0; r < rows; ++r)
    memcpy(packed.data() + r * zb, reinterpret_cast<const uint8_t*>(occ.data() + r * g.wz), zb);  // little-endian words
  char head[512];
  const b200sv_field_cfg& f = c->cfg;
  const int hl = snprintf(head, sizeof(head),
                          "resolution: %g\nsize_x: %g\nsize_y: %g\nsize_z: %g\norigin_x: %g\norigin_y: %g\norigin_z: %g\n",
                          f.resolution, f.size[0], f.size[1], f.size[2], f.origin[0] - 0.5 * f.size[0],
                          f.origin[1] - 0.5 * f.size[1], f.origin[2] - 0.5 * f.size[2]);
  uLongf zl = compressBound((uLong)packed.size());
  std::vector<uint8_t> z(zl);
  if (compress(z.data(), &zl, packed.data(), (uLong)packed.size()) != Z_OK)
    return fail(c, B200SV_ERR_INVALID, "zlib compress failed");
  *n_out = (size_t)hl + zl;
  if (!out || cap < *n_out)
    return out ? fail(c, B200SV_ERR_INVALID, "output buffer too small (n_out holds the size needed)") : B200SV_OK;
  memcpy(out, head, (size_t)hl);
  memcpy(out + hl, z.data(), zl);
  return B200SV_OK;
}

int b200sv_field_read_df(b200sv_ctx* c, int which, const uint8_t* bytes, size_t n_bytes)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!bytes)
    return fail(c, B200SV_ERR_INVALID, "null input");
  static const char* keys[7] = { "resolution:", "size_x:", "size_y:", "size_z:", "origin_x:", "origin_y:", "origin_z:" };
  double v[7];
  size_t pos = 0;
  for (int k = 0; k < 7; ++k)
  {
    size_t end = pos;
    while (end < n_bytes && bytes[end] != '\n')
      ++end;
    if (end >= n_bytes)
      return fail(c, B200SV_ERR_PARSE, ".df header truncated");
    const std::string line(reinterpret_cast<const char*>(bytes) + pos, end - pos);
    const size_t kl = strlen(keys[k]);
    if (line.compare(0, kl, keys[k]) != 0)
      return fail(c, B200SV_ERR_PARSE, std::string(".df header: expected ") + keys[k]);
    v[k] = atof(line.c_str() + kl);
    pos = end + 1;
  }
  // readFromStream resizes the field to the stream's geometry; a context's geometry is fixed at creation, so the stream
  // must describe the same grid
  const b200sv_field_cfg& f = c->cfg;
  const double want[7] = { f.resolution, f.size[0], f.size[1], f.size[2], f.origin[0] - 0.5 * f.size[0],
                           f.origin[1] - 0.5 * f.size[1], f.origin[2] - 0.5 * f.size[2] };
  for (int k = 0; k < 7; ++k)
    if (fabs(v[k] - want[k]) > 1e-5 * std::max(1.0, fabs(want[k])))  // the header carries 6 significant digits
      return fail(c, B200SV_ERR_INVALID, std::string(".df geometry differs from the context's field: ") + keys[k]);
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const GridDev& g = c->grid;
  const size_t rows = (size_t)g.nx * g.ny, zb = (size_t)(g.nz + 7) / 8;
  std::vector<uint8_t> packed(rows * zb);
  uLongf got = (uLongf)packed.size();
  const int zr = uncompress(packed.data(), &got, bytes + pos, (uLong)(n_bytes - pos));
  if (zr != Z_OK || got != packed.size())
    return fail(c, B200SV_ERR_PARSE, ".df payload: zlib error or wrong size for this grid");
  std::vector<uint32_t> occ(rows * g.wz, 0u);
  const uint8_t tail_mask = (g.nz & 7) ? (uint8_t)((1u << (g.nz & 7)) - 1u) : (uint8_t)0xff;
  for (size_t r = 0; r < rows; ++r)
  {
    uint8_t* dst = reinterpret_cast<uint8_t*>(occ.data() + r * g.wz);
    memcpy(dst, packed.data() + r * zb, zb);
    dst[zb - 1] &= tail_mask;  // bits beyond nz are not voxels
  }
  Field& F = c->fields[which];
  CU(c, cudaMemcpyAsync(F.occ, occ.data(), occ.size() * 4, cudaMemcpyHostToDevice, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  if (F.prop)
  {  // readFromStream: initialize() (a reset), then addNewObstacleVoxels of the occupied voxels in x, y, z order (:720-758)
    CU(c, propReset(F.prop, c->stream));
    CU(c, propAddOccupancy(F.prop, F.occ, c->stream));
    if ((rc = publishPropagation(c, which, c->stream)) != B200SV_OK)
      return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return B200SV_OK;
  }
  return rebuildField(c, which);
}

// ---- hot path --------------------------------------------------------------------------------------------------------
int b200sv_fk_batch(b200sv_ctx* c, const double* q, size_t n, int precision_bits, double* link_T)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if ((n && (!q || !link_T)) || (precision_bits != 32 && precision_bits != 64))
    return fail(c, B200SV_ERR_INVALID, "bad arguments to b200sv_fk_batch");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t V = c->robot.group_var_index.size();
  const size_t out_n = n * (size_t)c->robot.n_links * 16;
  CU(c, c->d_q.ensure(n * V));
  CU(c, c->d_fk.ensure(out_n));
  CU(c, cudaMemcpyAsync(c->d_q.p, q, n * V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (int rc = fkLaunch(c, c->d_q.p, n, precision_bits == 64, c->d_fk.p, c->stream))
    return rc;
  CU(c, cudaMemcpyAsync(link_T, c->d_fk.p, out_n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return B200SV_OK;
}

int b200sv_fk_batch_device(b200sv_ctx* c, const double* d_q, size_t n, int precision_bits, double* d_link_T, void* cuda_stream)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if ((n && (!d_q || !d_link_T)) || (precision_bits != 32 && precision_bits != 64))
    return fail(c, B200SV_ERR_INVALID, "bad arguments to b200sv_fk_batch_device");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  if (int rc = enterUser(c, s))
    return rc;
  if (int rc = fkLaunch(c, d_q, n, precision_bits == 64, d_link_T, s))
    return rc;
  return leaveUser(c, s);
}

int b200sv_sphere_centers(b200sv_ctx* c, const double* q_one, double* centers)
{
  if (!c || !q_one || !centers)
    return B200SV_ERR_INVALID;
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t V = c->robot.group_var_index.size();
  CU(c, c->d_q.ensure(V));
  CU(c, c->d_fk.ensure((size_t)c->n_spheres * 3 + 1));
  CU(c, cudaMemcpyAsync(c->d_q.p, q_one, V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  FkArgs a{};
  a.model = c->d_blob_d;
  a.model_bytes = (int)c->blob_d.size();
  a.q = c->d_q.p;
  a.n_states = 1;
  const size_t smem = align16(a.model_bytes) + (size_t)c->state_stride_d * 8;
  CU(c, launchSphereCenters(a, c->d_fk.p, smem, c->stream));
  CU(c, cudaMemcpyAsync(centers, c->d_fk.p, (size_t)c->n_spheres * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  return B200SV_OK;
}

int b200sv_check_batch_device(b200sv_ctx* c, const double* d_q, size_t n, uint32_t flags, uint8_t* d_valid_bitmap,
                              void* cuda_stream)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!d_q || !d_valid_bitmap))
    return fail(c, B200SV_ERR_INVALID, "null device pointer");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  int rc = enterUser(c, s);
  if (rc != B200SV_OK)
    return rc;
  rc = checkDevice(c, d_q, n, flags, reinterpret_cast<uint32_t*>(d_valid_bitmap), s);
  c->stats.n_states = n;
  c->stats.n_valid = 0;
  c->stats.check_ms = 0.f;
  c->stats_pending = true;
  if (rc != B200SV_OK)
    return rc;
  return leaveUser(c, s);
}

// pinned staging of the latency paths (single states, single edges): the context's own, so no copy of a small call falls
// back to the driver's pageable staging
static int ensureSmallStaging(b200sv_ctx* c, size_t bytes)
{
  if (c->h_small_bytes >= bytes)
    return B200SV_OK;
  if (c->h_small)
    cudaFreeHost(c->h_small);
  c->h_small = nullptr;
  c->h_small_bytes = 0;
  CU(c, cudaHostAlloc(&c->h_small, bytes, cudaHostAllocDefault));
  c->h_small_bytes = bytes;
  return B200SV_OK;
}

// Popcount of a validity bitmap of n states (bits beyond n are 0: the kernels write whole words from ballots of active lanes)
static uint64_t countValid(const uint8_t* valid_bitmap, size_t n)
{
  uint64_t valid = 0;
  const size_t nb = (n + 7) / 8;
  size_t i = 0;
  for (; i + 8 <= nb; i += 8)
  {
    uint64_t w;
    memcpy(&w, valid_bitmap + i, 8);
    valid += (uint64_t)__builtin_popcountll(w);
  }
  for (; i < nb; ++i)
    valid += (uint64_t)__builtin_popcount(valid_bitmap[i]);
  return valid;
}

constexpr size_t kSmallBatch = 4096;   // up to here: one stream, the context's own pinned staging, no timing events
constexpr size_t kLatencyBatch = 32;   // up to here: ONE kernel launch, no copies (latencyCheckKernel)

// Small batches: everything on one stream, states and results staged through the context's own pinned buffer.
static int checkSmall(b200sv_ctx* c, const double* q, size_t n, uint32_t flags, uint8_t* valid_bitmap)
{
  const size_t V = c->robot.group_var_index.size();
  CU(c, c->d_q.ensure(n * V));
  CU(c, c->d_list.ensure(n));
  unsigned cnt = 0;
  // Latency path (a single-state checkCollision through the plugin is a batch of 1): everything on one stream, states and
  // results staged through the context's own pinned buffer, so no copy falls back to the driver's pageable staging.
  const size_t q_bytes = kSmallBatch * V * sizeof(double), out_bytes = kSmallBatch / 8 + 16;
  if (int rc = ensureSmallStaging(c, q_bytes + out_bytes))
    return rc;
  uint8_t* h_out = c->h_small + q_bytes;
  // bitmap words and the recheck counter side by side in device memory: ONE copy back instead of two
  if (!c->d_small)
    CU(c, cudaMalloc(&c->d_small, kSmallBatch / 8 + 16));
  uint32_t* d_small_bitmap = c->d_small;
  unsigned* const count_saved = c->d_count;
  c->d_count = c->d_small + kSmallBatch / 32;
  struct Restore
  {
    b200sv_ctx* c;
    unsigned* p;
    ~Restore() { c->d_count = p; }
  } restore{ c, count_saved };
  memcpy(c->h_small, q, n * V * sizeof(double));
  const bool fp64 = (flags & B200SV_CHECK_FP64) != 0;
  if (!fp64)
    CU(c, cudaMemsetAsync(c->d_count, 0, sizeof(unsigned), c->stream));
  const auto t_small = std::chrono::steady_clock::now();  // no timing events on the latency path: three API calls fewer
  CU(c, cudaMemcpyAsync(c->d_q.p, c->h_small, n * V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int rc = c->compact_bytes == 1 ? launchMain<uint8_t>(c, c->d_q.p, d_small_bitmap, 0, n, flags, c->stream) :
                                   launchMain<uint16_t>(c, c->d_q.p, d_small_bitmap, 0, n, flags, c->stream);
  if (rc == B200SV_OK && !fp64)
    rc = c->compact_bytes == 1 ? launchRecheck<uint8_t>(c, c->d_q.p, d_small_bitmap, n, flags, c->stream) :
                                 launchRecheck<uint16_t>(c, c->d_q.p, d_small_bitmap, n, flags, c->stream);
  if (rc != B200SV_OK)
    return rc;
  CU(c, cudaMemcpyAsync(h_out, c->d_small, kSmallBatch / 8 + sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaStreamSynchronize(c->stream));
  memcpy(valid_bitmap, h_out, (n + 7) / 8);
  memcpy(&cnt, h_out + kSmallBatch / 8, sizeof(unsigned));
  const float ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_small).count();
  c->stats.n_states = n;
  c->stats.n_valid = countValid(valid_bitmap, n);
  c->stats.n_rechecked = (flags & B200SV_CHECK_FP64) ? 0 : cnt;
  c->stats.check_ms = ms;
  c->stats_pending = false;
  return B200SV_OK;
}

// The latency path: n <= 32 states (OMPL's per-state isValid through the plugin is n = 1).  One kernel launch, nothing else:
// the states travel in the kernel parameters (or mapped pinned memory), the exact FP64 check runs one warp per state, the
// result bytes land in mapped pinned memory and the host polls them -- no memcpy, no memset, no second launch, no
// stream synchronisation on the way.
static int checkLatency(b200sv_ctx* c, const double* q, size_t n, uint32_t flags, uint8_t* valid_bitmap)
{
  const size_t V = c->robot.group_var_index.size();
  const size_t q_bytes = kLatencyBatch * V * sizeof(double);
  if (!c->h_lat)
  {
    CU(c, cudaHostAlloc(&c->h_lat, q_bytes + 64, cudaHostAllocMapped));
    CU(c, cudaHostGetDevicePointer(&c->d_lat, c->h_lat, 0));
  }
  const auto t0 = std::chrono::steady_clock::now();
  volatile uint8_t* res = c->h_lat + q_bytes;
  for (size_t i = 0; i < n; ++i)
    res[i] = 0xff;
  auto fill = [&](auto& a) {
    a.model = c->d_blob_d;
    a.model_bytes = (int)c->blob_d.size();
    a.field = c->d_combined;
    a.flags = flags & 7u;
    a.n_states = (int)n;
    a.state_stride = c->state_stride_d;
    a.n_links = c->n_dev_links;
    a.use_inline = n * V <= 32 ? 1 : 0;
    a.q = reinterpret_cast<const double*>(c->d_lat);
    a.result = c->d_lat + q_bytes;
    if (a.use_inline)
      memcpy(a.q_inline, q, n * V * sizeof(double));
    else
      memcpy(c->h_lat, q, n * V * sizeof(double));
  };
  if (c->compact_bytes == 1)
  {
    LatencyArgs<uint8_t> a{};
    fill(a);
    CU(c, launchLatencyCheck<uint8_t>(a, c->stream));
  }
  else
  {
    LatencyArgs<uint16_t> a{};
    fill(a);
    CU(c, launchLatencyCheck<uint16_t>(a, c->stream));
  }
  // poll the result bytes; a kernel that cannot finish (launch failure, device error) is caught by the synchronise below
  bool done = false;
  for (long spin = 0; spin < 4000000L && !done; ++spin)
  {
    done = true;
    for (size_t i = 0; i < n; ++i)
      if (res[i] == 0xff)
      {
        done = false;
        break;
      }
  }
  if (!done)
  {
    CU(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; ++i)
      if (res[i] == 0xff)
        return fail(c, B200SV_ERR_CUDA, "latency path: the kernel finished without writing a result");
  }
  memset(valid_bitmap, 0, (n + 7) / 8);
  uint64_t valid = 0;
  for (size_t i = 0; i < n; ++i)
    if (res[i] == 1)
    {
      valid_bitmap[i >> 3] |= (uint8_t)(1u << (i & 7));
      ++valid;
    }
  c->stats.n_states = n;
  c->stats.n_valid = valid;
  c->stats.n_rechecked = n;  // every state of this path is decided by the exact FP64 arithmetic
  c->stats.check_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  c->stats_pending = false;
  return B200SV_OK;
}

// ONE state, and probably another one right after it (OMPL validates state by state): answered by a kernel that stays
// resident for kLingerUs after its last answer.  The first call launches it with the state in the launch parameters; while
// it is believed alive a call is one store into its mailbox and a poll of the answer word -- no launch, no copy.  If the
// kernel has just left (the two clocks race), the stream turns idle without an answer and the call launches a new one.
// Every other entry point asks the kernel to quit first (quitLinger in enterCtx / enterUser).
static int checkLinger(b200sv_ctx* c, const double* q, uint32_t flags, uint8_t* valid_bitmap)
{
  static const long linger_us = getenv("B200SV_LINGER_US") ? atol(getenv("B200SV_LINGER_US")) : 200;
  const size_t V = c->robot.group_var_index.size();
  if (!c->h_mail)
  {
    CU(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_mail), sizeof(LingerMailbox), cudaHostAllocMapped));
    memset(c->h_mail, 0, sizeof(LingerMailbox));
    CU(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_mail), c->h_mail, 0));
  }
  const auto t0 = std::chrono::steady_clock::now();
  LingerMailbox* mb = c->h_mail;
  const unsigned long long seq = ++c->mail_seq;
  const unsigned long long req = (seq << 3) | (flags & 7u);
  auto launch = [&]() -> int {
    auto fill = [&](auto& a) {
      a.model = c->d_blob_d;
      a.model_bytes = (int)c->blob_d.size();
      a.field = c->d_combined;
      a.state_stride = c->state_stride_d;
      a.n_links = c->n_dev_links;
      a.mail = c->d_mail;
      a.first_req = req;
      a.epoch = ++c->linger_epoch;
      a.linger_ns = (unsigned long long)linger_us * 1000ull;
      a.life_ns = 20000000ull;  // 20 ms: a kernel never outlives this, whatever the host does
      memcpy(a.q_inline, q, V * sizeof(double));
    };
    if (c->compact_bytes == 1)
    {
      LingerArgs<uint8_t> a{};
      fill(a);
      CU(c, launchLingerCheck<uint8_t>(a, c->stream));
    }
    else
    {
      LingerArgs<uint16_t> a{};
      fill(a);
      CU(c, launchLingerCheck<uint16_t>(a, c->stream));
    }
    c->linger_maybe_alive = true;
    c->linger_launched = std::chrono::steady_clock::now();
    return B200SV_OK;
  };
  // believed alive: launched, not asked to quit, and its linger window (minus a margin for the two clocks) still open
  bool posted = false;
  if (c->linger_maybe_alive &&
      std::chrono::duration_cast<std::chrono::microseconds>(t0 - c->linger_last).count() < std::max(20L, linger_us - 40) &&
      std::chrono::duration_cast<std::chrono::microseconds>(t0 - c->linger_launched).count() < 19000)
  {
    memcpy(const_cast<double*>(mb->q), q, V * sizeof(double));
    std::atomic_thread_fence(std::memory_order_release);
    *reinterpret_cast<volatile unsigned long long*>(&mb->req) = req;
    posted = true;
  }
  else
  {
    quitLinger(c);  // a kernel past its window is leaving or gone; the new one queues behind it
    if (int rc = launch())
      return rc;
  }
  const volatile unsigned long long* done = reinterpret_cast<const volatile unsigned long long*>(&mb->done);
  unsigned long long d = 0;
  for (long spin = 0;; ++spin)
  {
    d = *done;
    if ((d >> 8) == seq)
      break;
    if ((spin & 255) == 255)
    {
      const long us = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
      if (posted && us > 30 && cudaStreamQuery(c->stream) == cudaSuccess)
      {  // the kernel left before it saw the request
        d = *done;
        if ((d >> 8) == seq)
          break;
        posted = false;
        c->linger_maybe_alive = false;
        if (int rc = launch())
          return rc;
      }
      if (us > 2000000)
      {
        c->linger_maybe_alive = false;
        CU(c, cudaStreamSynchronize(c->stream));
        return fail(c, B200SV_ERR_CUDA, "single-state path: no answer from the resident kernel");
      }
    }
  }
  c->linger_last = std::chrono::steady_clock::now();
  valid_bitmap[0] = (uint8_t)(d & 1u);
  c->stats.n_states = 1;
  c->stats.n_valid = d & 1u;
  c->stats.n_rechecked = 1;  // decided by the exact FP64 arithmetic
  c->stats.check_ms = std::chrono::duration<float, std::milli>(c->linger_last - t0).count();
  c->stats_pending = false;
  return B200SV_OK;
}

// The large-batch path in two halves, so several contexts (one per GPU, multi.cu) can be in flight at once:
// enqueue = chunked H2D of the states overlapped with the kernels + D2H of the bitmap, all asynchronous; finish = wait + stats.
// The caller holds c->mu across both.
static int checkHostEnqueue(b200sv_ctx* c, const void* q, size_t n, uint32_t flags, uint8_t* valid_bitmap, bool f32 = false)
{
  cudaSetDevice(c->device);
  const size_t V = c->robot.group_var_index.size();
  const size_t words = (n + 31) / 32;
  if ((flags & 7u) == 0)
    return fail(c, B200SV_ERR_INVALID, "flags select no check");
  if (n >= (1ull << 32))
    return fail(c, B200SV_ERR_INVALID, "more than 2^32 - 1 states in one call");
  CU(c, c->d_q.ensure(n * V));
  if (f32)
    CU(c, c->d_qf.ensure(n * V));
  CU(c, c->d_bitmap.ensure(words));
  CU(c, c->d_list.ensure(n));
  if (!c->h_count)
    CU(c, cudaHostAlloc(&c->h_count, sizeof(unsigned), cudaHostAllocDefault));
  CU(c, cudaEventRecord(c->ev0, c->stream));
  int rc = c->compact_bytes == 1 ? runCheckHost<uint8_t>(c, q, n, flags, f32) : runCheckHost<uint16_t>(c, q, n, flags, f32);
  if (rc != B200SV_OK)
    return rc;
  CU(c, cudaEventRecord(c->ev1, c->stream));
  // bitmap back: whole bytes of the caller's buffer; the device words are little-endian, bit i%8 of byte i/8
  CU(c, cudaMemcpyAsync(valid_bitmap, c->d_bitmap.p, (n + 7) / 8, cudaMemcpyDeviceToHost, c->stream));
  CU(c, cudaMemcpyAsync(c->h_count, c->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, c->stream));
  return B200SV_OK;
}

static int checkHostFinish(b200sv_ctx* c, size_t n, uint32_t flags, const uint8_t* valid_bitmap)
{
  cudaSetDevice(c->device);
  CU(c, cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->stats.n_states = n;
  c->stats.n_valid = countValid(valid_bitmap, n);
  c->stats.n_rechecked = (flags & B200SV_CHECK_FP64) ? 0 : *c->h_count;
  c->stats.check_ms = ms;
  c->stats_pending = false;
  return B200SV_OK;
}

int b200sv_check_batch(b200sv_ctx* c, const double* q, size_t n, uint32_t flags, uint8_t* valid_bitmap)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!q || !valid_bitmap))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  static const bool linger_on = !(getenv("B200SV_LINGER") && atoi(getenv("B200SV_LINGER")) == 0);
  const bool one = n == 1 && linger_on && c->latency_ok && (flags & 7u) != 0 && !(flags & B200SV_CHECK_FP64) && !c->user_pending;
  enterCtx(c, one);
  if ((flags & 7u) == 0)
    return fail(c, B200SV_ERR_INVALID, "flags select no check");
  if (one)
    return checkLinger(c, q, flags, valid_bitmap);
  if (n <= kLatencyBatch && !(flags & B200SV_CHECK_FP64) && c->latency_ok)
    return checkLatency(c, q, n, flags, valid_bitmap);
  if (n <= kSmallBatch)
    return checkSmall(c, q, n, flags, valid_bitmap);
  if (int rc = checkHostEnqueue(c, q, n, flags, valid_bitmap))
    return rc;
  return checkHostFinish(c, n, flags, valid_bitmap);
}

// A few edges, ONE launch (an OMPL MotionValidator validates one edge per checkMotion, and the general path below costs two
// stream synchronisations: the batch is sized on the device).  Here the host sizes it -- the state counts are the arithmetic of
// motionCountKernel, operation for operation -- and launches one CTA per visited state (edgeLatencyKernel); the ends travel in
// the launch parameters (one edge) or are read from mapped host memory, the result bytes are polled there.
// Returns 1 when the call does not qualify (more than 1 024 states, a group of more than 32 variables): the general path takes it.
constexpr unsigned kEdgeLatencyStates = 1024;  // measured break-even with the general path: ~1 400 states (64 edges)
static int checkMotionsLatency(b200sv_ctx* c, const double* q_from, const double* q_to, size_t n_edges, double max_segment_length,
                               const uint8_t* var_continuous, const double* var_distance_factor, uint32_t max_states_per_edge,
                               uint32_t flags, uint8_t* valid_bitmap, int32_t* first_invalid, uint64_t* n_states_checked)
{
  const double kPi = 3.14159265358979323846;
  const size_t V = c->robot.group_var_index.size();
  if (V > 32 || n_edges > (size_t)kEdgeLatencyEdges)
    return 1;
  uint32_t offsets[kEdgeLatencyEdges + 1];
  offsets[0] = 0;
  for (size_t e = 0; e < n_edges; ++e)
  {
    const double *f = q_from + e * V, *t = q_to + e * V;
    double d = 0.0;
    for (size_t k = 0; k < V; ++k)
    {
      const double w = var_distance_factor ? var_distance_factor[k] : 1.0;
      if (w == 0.0)
        continue;
      double dk = fabs(f[k] - t[k]);
      if (var_continuous && var_continuous[k])
      {
        dk = fmod(dk, 2.0 * kPi);
        dk = dk > kPi ? 2.0 * kPi - dk : dk;
      }
      const volatile double prod = w * dk;  // two roundings, as __dadd_rn(d, __dmul_rn(w, dk)) on the device
      d = d + prod;
    }
    const double nd = ceil(d / max_segment_length);
    unsigned cnt = nd < 1.0 ? 1u : (nd > 4.0e9 ? 0xffffffffu : static_cast<unsigned>(nd));
    if (max_states_per_edge && cnt > max_states_per_edge)
      cnt = max_states_per_edge;
    if (cnt > kEdgeLatencyStates || offsets[e] + cnt > kEdgeLatencyStates)
      return 1;
    offsets[e + 1] = offsets[e] + cnt;
  }
  const unsigned total = offsets[n_edges];
  const size_t ends_cap = 2 * (size_t)kEdgeLatencyEdges * 32 * sizeof(double);
  if (!c->h_edge_lat)
  {
    CU(c, cudaHostAlloc(reinterpret_cast<void**>(&c->h_edge_lat), ends_cap + kEdgeLatencyStates, cudaHostAllocMapped));
    CU(c, cudaHostGetDevicePointer(reinterpret_cast<void**>(&c->d_edge_lat), c->h_edge_lat, 0));
  }
  const auto t0 = std::chrono::steady_clock::now();
  volatile uint8_t* res = c->h_edge_lat + ends_cap;
  for (unsigned i = 0; i < total; ++i)
    res[i] = 0xff;
  auto fill = [&](auto& a) {
    a.model = c->d_blob_d;
    a.model_bytes = (int)c->blob_d.size();
    a.field = c->d_combined;
    a.flags = flags & 7u;
    a.state_stride = c->state_stride_d;
    a.n_links = c->n_dev_links;
    a.n_edges = (int)n_edges;
    a.n_vars = (int)V;
    a.cont_mask = 0u;
    for (size_t k = 0; k < V; ++k)
      if (var_continuous && var_continuous[k])
        a.cont_mask |= 1u << k;
    a.use_inline = n_edges == 1 ? 1 : 0;
    a.ends = reinterpret_cast<const double*>(c->d_edge_lat);
    a.result = c->d_edge_lat + ends_cap;
    memcpy(a.offsets, offsets, (n_edges + 1) * sizeof(uint32_t));
    if (a.use_inline)
    {
      memcpy(a.ends_inline, q_from, V * sizeof(double));
      memcpy(a.ends_inline + V, q_to, V * sizeof(double));
    }
    else
    {
      memcpy(c->h_edge_lat, q_from, n_edges * V * sizeof(double));
      memcpy(c->h_edge_lat + n_edges * V * sizeof(double), q_to, n_edges * V * sizeof(double));
    }
  };
  if (c->compact_bytes == 1)
  {
    EdgeLatencyArgs<uint8_t> a{};
    fill(a);
    CU(c, launchEdgeLatencyCheck<uint8_t>(a, c->stream));
  }
  else
  {
    EdgeLatencyArgs<uint16_t> a{};
    fill(a);
    CU(c, launchEdgeLatencyCheck<uint16_t>(a, c->stream));
  }
  bool done = false;
  for (long spin = 0; spin < 4000000L && !done; ++spin)
  {
    done = true;
    for (unsigned i = total; i-- > 0;)  // CTAs finish roughly in order: the last byte turns last
      if (res[i] == 0xff)
      {
        done = false;
        break;
      }
  }
  if (!done)
  {
    CU(c, cudaStreamSynchronize(c->stream));
    for (unsigned i = 0; i < total; ++i)
      if (res[i] == 0xff)
        return fail(c, B200SV_ERR_CUDA, "few-edges path: the kernel finished without writing a result");
  }
  memset(valid_bitmap, 0, (n_edges + 7) / 8);
  for (size_t e = 0; e < n_edges; ++e)
  {
    int bad = -1;
    for (unsigned i = offsets[e]; i < offsets[e + 1] && bad < 0; ++i)
      if (res[i] == 0)
        bad = (int)(i - offsets[e]) + 1;  // j of the first invalid state, 1 .. count
    if (bad < 0)
      valid_bitmap[e >> 3] |= (uint8_t)(1u << (e & 7));
    if (first_invalid)
      first_invalid[e] = bad;
  }
  if (n_states_checked)
    *n_states_checked = total;
  c->stats.n_states = total;
  c->stats.n_valid = 0;
  c->stats.n_rechecked = total;  // every state of this path is decided by the exact FP64 arithmetic
  c->stats.check_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  c->stats_pending = false;
  return B200SV_OK;
}

int b200sv_check_motions(b200sv_ctx* c, const double* q_from, const double* q_to, size_t n_edges, double max_segment_length,
                         const uint8_t* var_continuous, const double* var_distance_factor, uint32_t max_states_per_edge,
                         uint32_t flags, uint8_t* valid_bitmap, int32_t* first_invalid, uint64_t* n_states_checked)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n_edges && (!q_from || !q_to || !valid_bitmap))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if (!(max_segment_length > 0.0))
    return fail(c, B200SV_ERR_INVALID, "max_segment_length must be positive");
  if ((flags & 7u) == 0)
    return fail(c, B200SV_ERR_INVALID, "flags select no check");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n_states_checked)
    *n_states_checked = 0;
  // JointModelGroup::distance / interpolate are per-variable and linear only for 1-DOF joints.  A planar joint measures
  // Euclidean xy distance plus a weighted angle, a floating joint translation plus quaternion arc and interpolates by
  // slerp (planar_joint_model.cpp:198-225, floating_joint_model.cpp:110-160): the edge kernels do not restate those, and
  // interpolating quaternion variables linearly would check states OMPL never visits -- refuse instead of approximating.
  for (int l = 0; l < c->robot.n_links; ++l)
    if (c->robot.joint_type[l] == PLANAR || c->robot.joint_type[l] == FLOATING)
      for (int k = 0; k < jointVarCount(c->robot.joint_type[l]); ++k)
        if (std::find(c->robot.group_var_index.begin(), c->robot.group_var_index.end(), c->robot.first_var[l] + k) !=
            c->robot.group_var_index.end())
          return fail(c, B200SV_ERR_UNSUPPORTED,
                      "b200sv_check_motions: the group has a planar or floating joint; their distance / interpolation "
                      "(Euclidean + angle, slerp) is not implemented -- validate such edges state by state");
  if (n_edges == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  static const bool edge_latency_on = !(getenv("B200SV_EDGE_LATENCY") && atoi(getenv("B200SV_EDGE_LATENCY")) == 0);
  if (edge_latency_on && n_edges <= (size_t)kEdgeLatencyEdges && c->latency_ok && !(flags & B200SV_CHECK_FP64))
  {
    const int rc = checkMotionsLatency(c, q_from, q_to, n_edges, max_segment_length, var_continuous, var_distance_factor,
                                       max_states_per_edge, flags, valid_bitmap, first_invalid, n_states_checked);
    if (rc != 1)
      return rc;
  }
  const size_t V = c->robot.group_var_index.size();
  cudaStream_t s = c->stream;
  CU(c, c->d_edges.ensure(2 * n_edges * V));
  CU(c, c->d_edge_count.ensure(n_edges));
  CU(c, c->d_edge_offset.ensure(n_edges + 1));
  CU(c, c->d_edge_aux.ensure(V));
  CU(c, c->d_edge_cont.ensure(V));
  CU(c, c->d_edge_bits.ensure((n_edges + 31) / 32));
  CU(c, c->d_edge_first.ensure(n_edges));
  // Latency path (an OMPL MotionValidator validates ONE edge per checkMotion): both ends in one copy from the context's pinned
  // staging, results back through it, no timing events; the group's metric (distance factors, continuous flags) is uploaded
  // only when it changes.
  constexpr size_t kSmallEdges = 64;
  const bool small = n_edges <= kSmallEdges;
  const size_t ends_bytes = 2 * n_edges * V * sizeof(double);
  const auto t_small = std::chrono::steady_clock::now();
  uint8_t* h_res = nullptr;  // small: [total (8) | count (4) | pad (4) | valid bits (8) | first_invalid (4 * 64)]
  if (small)
  {
    const size_t stage_bytes = 2 * kSmallEdges * V * sizeof(double);
    if (int rc = ensureSmallStaging(c, stage_bytes + 512))
      return rc;
    memcpy(c->h_small, q_from, ends_bytes / 2);
    memcpy(c->h_small + ends_bytes / 2, q_to, ends_bytes / 2);
    h_res = c->h_small + stage_bytes;
    CU(c, cudaMemcpyAsync(c->d_edges.p, c->h_small, ends_bytes, cudaMemcpyHostToDevice, s));
  }
  else
  {
    CU(c, cudaEventRecord(c->ev0, s));
    CU(c, cudaMemcpyAsync(c->d_edges.p, q_from, ends_bytes / 2, cudaMemcpyHostToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_edges.p + n_edges * V, q_to, ends_bytes / 2, cudaMemcpyHostToDevice, s));
  }
  if (var_distance_factor &&
      (c->edge_factor_cached.size() != V || memcmp(c->edge_factor_cached.data(), var_distance_factor, V * sizeof(double)) != 0))
  {
    c->edge_factor_cached.assign(var_distance_factor, var_distance_factor + V);
    CU(c, cudaMemcpyAsync(c->d_edge_aux.p, c->edge_factor_cached.data(), V * sizeof(double), cudaMemcpyHostToDevice, s));
  }
  if (var_continuous && (c->edge_cont_cached.size() != V || memcmp(c->edge_cont_cached.data(), var_continuous, V) != 0))
  {
    c->edge_cont_cached.assign(var_continuous, var_continuous + V);
    CU(c, cudaMemcpyAsync(c->d_edge_cont.p, c->edge_cont_cached.data(), V, cudaMemcpyHostToDevice, s));
  }
  MotionArgs a{};
  a.q_from = c->d_edges.p;
  a.q_to = c->d_edges.p + n_edges * V;
  a.n_edges = n_edges;
  a.n_vars = (int)V;
  a.continuous = var_continuous ? c->d_edge_cont.p : nullptr;
  a.factor = var_distance_factor ? c->d_edge_aux.p : nullptr;
  a.max_segment = max_segment_length;
  a.max_states_per_edge = max_states_per_edge;
  a.count = c->d_edge_count.p;
  a.offset = c->d_edge_offset.p;
  CU(c, launchMotionCount(a, s));
  unsigned long long total = 0;
  CU(c, cudaMemcpyAsync(small ? static_cast<void*>(h_res) : static_cast<void*>(&total), c->d_edge_offset.p + n_edges, sizeof(total),
                        cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));  // the state batch is sized by the edges' lengths
  static const bool trace = getenv("B200SV_TRACE") != nullptr;
  const auto t_counted = std::chrono::steady_clock::now();
  if (small)
    memcpy(&total, h_res, sizeof(total));
  if (total >= (1ull << 32))
    return fail(c, B200SV_ERR_INVALID, "edges expand to more than 2^32 - 1 states; lower n_edges or set max_states_per_edge");
  CU(c, c->d_edge_q.ensure((size_t)total * V));
  CU(c, c->d_edge_state_bits.ensure((size_t)(total + 31) / 32));
  CU(c, launchMotionExpand(a, c->d_edge_q.p, total, s));
  int rc = checkDevice(c, c->d_edge_q.p, (size_t)total, flags, c->d_edge_state_bits.p, s);
  if (rc != B200SV_OK)
    return rc;
  CU(c, launchMotionReduce(a, c->d_edge_state_bits.p, c->d_edge_bits.p, first_invalid ? c->d_edge_first.p : nullptr, s));
  unsigned cnt = 0;
  float ms = 0.f;
  if (small)
  {
    CU(c, cudaMemcpyAsync(h_res + 16, c->d_edge_bits.p, (n_edges + 7) / 8, cudaMemcpyDeviceToHost, s));
    if (first_invalid)
      CU(c, cudaMemcpyAsync(h_res + 24, c->d_edge_first.p, n_edges * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(h_res + 8, c->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    memcpy(valid_bitmap, h_res + 16, (n_edges + 7) / 8);
    if (first_invalid)
      memcpy(first_invalid, h_res + 24, n_edges * sizeof(int32_t));
    memcpy(&cnt, h_res + 8, sizeof(unsigned));
    ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_small).count();
  }
  else
  {
    CU(c, cudaEventRecord(c->ev1, s));
    CU(c, cudaMemcpyAsync(valid_bitmap, c->d_edge_bits.p, (n_edges + 7) / 8, cudaMemcpyDeviceToHost, s));
    if (first_invalid)
      CU(c, cudaMemcpyAsync(first_invalid, c->d_edge_first.p, n_edges * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    CU(c, cudaMemcpyAsync(&cnt, c->d_count, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    CU(c, cudaStreamSynchronize(s));
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  }
  if (trace)
    fprintf(stderr, "[b200sv trace] check_motions: %zu edges -> %llu states; edges on the device and counted after %.0f us, all "
                    "done after %.0f us (device span %.0f us)\n", n_edges, total,
            std::chrono::duration<double, std::micro>(t_counted - t_small).count(),
            std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_small).count(), ms * 1e3);
  if (n_states_checked)
    *n_states_checked = total;
  c->stats.n_states = total;
  c->stats.n_valid = 0;
  c->stats.n_rechecked = (flags & B200SV_CHECK_FP64) ? 0 : cnt;
  c->stats.check_ms = ms;
  c->stats_pending = false;
  return B200SV_OK;
}

int b200sv_collision_gradients(b200sv_ctx* c, const double* q, size_t n, double* distances, double* gradients, uint8_t* types,
                               double* closest_distance, double* sphere_locations)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!q || !distances || !gradients || !types))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t V = c->robot.group_var_index.size(), S = (size_t)c->n_spheres, L = (size_t)c->cm.n_glinks;
  cudaStream_t s = c->stream;
  CU(c, c->d_q.ensure(n * V));
  CU(c, c->d_g_dist.ensure(n * S));
  CU(c, c->d_g_grad.ensure(n * S * 3));
  CU(c, c->d_g_centers.ensure(n * S * 3));
  CU(c, c->d_g_types.ensure(n * S));
  CU(c, c->d_g_closest.ensure(n * L));
  CU(c, c->d_g_self.ensure(L));
  CU(c, cudaMemcpyAsync(c->d_q.p, q, n * V * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_g_self.p, c->cm.self_enabled.data(), L, cudaMemcpyHostToDevice, s));
  GradArgs a{};
  a.model = c->d_blob_d;
  a.model_bytes = (int)c->blob_d.size();
  a.q = c->d_q.p;
  a.n_states = n;
  a.d2_world = c->fields[0].d2;
  a.d2_self = c->fields[1].d2;
  a.nd2_world = c->fields[0].nd2;
  a.nd2_self = c->fields[1].nd2;
  a.self_enabled = c->d_g_self.p;
  a.n_glinks = (int)L;
  a.n_spheres = (int)S;
  a.state_stride = c->state_stride_d;
  if (!c->d_sqrt_table)
  {  // PropagationDistanceField::initialize (propagation_distance_field.cpp:99-101)
    std::vector<double> table((size_t)c->grid.max_d2 + 1);
    for (int k = 0; k <= c->grid.max_d2; ++k)
      table[k] = sqrt(double(k)) * c->cfg.resolution;
    CU(c, cudaMalloc(&c->d_sqrt_table, table.size() * sizeof(double)));
    CU(c, cudaMemcpy(c->d_sqrt_table, table.data(), table.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  a.sqrt_table = c->d_sqrt_table;
  // the per-link posed fields, when the caller switched them on (b200sv_set_gradient_link_mask)
  a.lf = nullptr;
  if (!c->lf_enabled.empty() && c->link_fields.size() == L)
  {
    if (c->lf_dirty)
    {
      std::vector<LinkFieldDev> host(L);
      for (size_t g = 0; g < L; ++g)
      {
        host[g].d2 = c->link_fields[g].d2;
        for (int k = 0; k < 3; ++k)
        {
          host[g].n[k] = c->link_fields[g].n[k];
          host[g].om[k] = c->link_fields[g].corner[k] - 0.5 * c->cfg.resolution;
        }
      }
      if (!c->d_lf)
        CU(c, cudaMalloc(&c->d_lf, L * sizeof(LinkFieldDev)));
      CU(c, cudaMemcpyAsync(c->d_lf, host.data(), L * sizeof(LinkFieldDev), cudaMemcpyHostToDevice, s));
      CU(c, c->d_lf_enabled.ensure(L * L));
      CU(c, cudaMemcpyAsync(c->d_lf_enabled.p, c->lf_enabled.data(), L * L, cudaMemcpyHostToDevice, s));
      CU(c, cudaStreamSynchronize(s));  // `host` lives on this stack frame
      c->lf_dirty = false;
    }
    std::vector<int32_t> slot(L);
    for (size_t g = 0; g < L; ++g)
      slot[g] = c->glink_slot[g];
    CU(c, c->d_c_start.ensure(L + 1));
    CU(c, c->d_c_slot.ensure(L));
    CU(c, cudaMemcpyAsync(c->d_c_start.p, c->cm.sphere_start.data(), (L + 1) * 4, cudaMemcpyHostToDevice, s));
    CU(c, cudaMemcpyAsync(c->d_c_slot.p, slot.data(), L * 4, cudaMemcpyHostToDevice, s));
    CU(c, cudaStreamSynchronize(s));
    a.lf = c->d_lf;
    a.lf_enabled = c->d_lf_enabled.p;
    a.sphere_start = c->d_c_start.p;
    a.glink_slot = c->d_c_slot.p;
  }
  a.resolution = c->cfg.resolution;
  // DistanceField::inv_twice_resolution_ is declared int (distance_field.hpp:614): 1 / (2 res) truncated
  a.inv_twice_resolution = (double)(int)(1.0 / (2.0 * c->cfg.resolution));
  a.max_distance = c->cfg.max_propagation_distance;  // getUninitializedDistance()
  a.max_propagation = c->cfg.max_propagation_distance;
  a.distances = c->d_g_dist.p;
  a.gradients = c->d_g_grad.p;
  a.types = c->d_g_types.p;
  a.closest = c->d_g_closest.p;
  a.centers = c->d_g_centers.p;
  int block = 64;
  while (block > 32 && align16(a.model_bytes) + (size_t)block * c->state_stride_d * 8 > (size_t)c->smem_optin)
    block -= 32;
  const size_t smem = align16(a.model_bytes) + (size_t)block * c->state_stride_d * 8;
  if (smem > (size_t)c->smem_optin)
    return fail(c, B200SV_ERR_UNSUPPORTED, "robot too large for the gradient kernel's shared-memory tile");
  CU(c, cudaEventRecord(c->ev0, s));
  CU(c, launchGradients(a, (int)((n + block - 1) / block), block, smem, s));
  CU(c, cudaEventRecord(c->ev1, s));
  CU(c, cudaMemcpyAsync(distances, a.distances, n * S * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(c, cudaMemcpyAsync(gradients, a.gradients, n * S * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(c, cudaMemcpyAsync(types, a.types, n * S, cudaMemcpyDeviceToHost, s));
  if (closest_distance)
    CU(c, cudaMemcpyAsync(closest_distance, a.closest, n * L * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (sphere_locations)
    CU(c, cudaMemcpyAsync(sphere_locations, a.centers, n * S * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, c->ev0, c->ev1);
  c->stats.n_states = n;
  c->stats.check_ms = ms;
  return B200SV_OK;
}

int b200sv_check_contacts(b200sv_ctx* c, const double* q, size_t n, uint32_t flags, uint32_t max_contacts,
                          uint32_t max_contacts_per_pair, uint8_t* collision, int32_t* contact_count, b200sv_contact* contacts)
{
  static_assert(sizeof(b200sv_contact) == sizeof(b200sv_contact_rec), "contact record layout");
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!q || !collision || !contact_count || !contacts))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if ((flags & 7u) == 0 || max_contacts == 0)
    return fail(c, B200SV_ERR_INVALID, "flags select no check, or max_contacts is 0");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t V = c->robot.group_var_index.size(), S = (size_t)c->n_spheres, L = (size_t)c->cm.n_glinks;
  cudaStream_t s = c->stream;
  std::vector<int32_t> slot(c->glink_slot.begin(), c->glink_slot.end());
  for (auto& v : slot)
    v = std::max(v, 0);  // entries without spheres are never dereferenced
  CU(c, c->d_q.ensure(n * V));
  CU(c, c->d_g_centers.ensure(n * S * 3));
  CU(c, c->d_c_start.ensure(L + 1));
  CU(c, c->d_c_slot.ensure(L));
  CU(c, c->d_c_bs.ensure(L * 4));
  CU(c, c->d_g_self.ensure(L));
  CU(c, c->d_c_intra.ensure(L * L));
  CU(c, c->d_c_coll.ensure(n));
  CU(c, c->d_c_count.ensure(n));
  CU(c, c->d_c_out.ensure(n * max_contacts));
  CU(c, cudaMemcpyAsync(c->d_q.p, q, n * V * sizeof(double), cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_c_start.p, c->cm.sphere_start.data(), (L + 1) * 4, cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_c_slot.p, slot.data(), L * 4, cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_c_bs.p, c->cm.bounding_sphere.data(), L * 4 * sizeof(double), cudaMemcpyHostToDevice, s));
  std::vector<int32_t> body(L, -1);
  if (!c->cm.body_id.empty())
    body.assign(c->cm.body_id.begin(), c->cm.body_id.end());
  CU(c, c->d_c_body.ensure(L));
  CU(c, cudaMemcpyAsync(c->d_c_body.p, body.data(), L * 4, cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_g_self.p, c->cm.self_enabled.data(), L, cudaMemcpyHostToDevice, s));
  CU(c, cudaMemcpyAsync(c->d_c_intra.p, c->cm.intra_enabled.data(), L * L, cudaMemcpyHostToDevice, s));
  CU(c, cudaMemsetAsync(c->d_c_out.p, 0xff, n * max_contacts * sizeof(b200sv_contact_rec), s));
  ContactArgs a{};
  a.model = c->d_blob_d;
  a.model_bytes = (int)c->blob_d.size();
  a.q = c->d_q.p;
  a.n_states = n;
  a.d2_world = c->fields[0].d2;
  a.d2_self = c->fields[1].d2;
  a.flags = flags & 7u;
  a.max_contacts = max_contacts;
  a.max_per_pair = max_contacts_per_pair;
  a.n_glinks = (int)L;
  a.sphere_start = c->d_c_start.p;
  a.glink_slot = c->d_c_slot.p;
  a.glink_bs = c->d_c_bs.p;
  a.body_id = c->d_c_body.p;
  a.self_enabled = c->d_g_self.p;
  a.intra = c->d_c_intra.p;
  a.centers = c->d_g_centers.p;
  a.n_spheres = c->n_spheres;
  a.state_stride = c->state_stride_d;
  a.collision = c->d_c_coll.p;
  a.contact_count = c->d_c_count.p;
  a.contacts = c->d_c_out.p;
  int block = 64;
  while (block > 32 && align16(a.model_bytes) + (size_t)block * c->state_stride_d * 8 > (size_t)c->smem_optin)
    block -= 32;
  const size_t smem = align16(a.model_bytes) + (size_t)block * c->state_stride_d * 8;
  if (smem > (size_t)c->smem_optin)
    return fail(c, B200SV_ERR_UNSUPPORTED, "robot too large for the contact kernel's shared-memory tile");
  CU(c, launchContacts(a, (int)((n + block - 1) / block), block, smem, s));
  CU(c, cudaMemcpyAsync(collision, a.collision, n, cudaMemcpyDeviceToHost, s));
  CU(c, cudaMemcpyAsync(contact_count, a.contact_count, n * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CU(c, cudaMemcpyAsync(contacts, a.contacts, n * max_contacts * sizeof(b200sv_contact_rec), cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  return B200SV_OK;
}

int b200sv_check_batch_f32(b200sv_ctx* c, const float* q, size_t n, uint32_t flags, uint8_t* valid_bitmap)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!q || !valid_bitmap))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (int rc = checkHostEnqueue(c, q, n, flags, valid_bitmap, true))
    return rc;
  return checkHostFinish(c, n, flags, valid_bitmap);
}

// ---- DistanceFieldCacheEntry updates without re-creating the context (a15) ---------------------------------------------
static int rederiveTables(b200sv_ctx* c)
{
  CU(c, cudaStreamSynchronize(c->stream));
  if (c->copy_stream)
    CU(c, cudaStreamSynchronize(c->copy_stream));
  const int compact_before = c->compact_bytes;
  int rc = buildTables(c);
  if (rc != B200SV_OK)
    return rc;
  if (c->compact_bytes != compact_before)
    return fail(c, B200SV_ERR_UNSUPPORTED, "the update changes the compact field's element size");
  c->edge_factor_cached.clear();
  c->edge_cont_cached.clear();
  return uploadTables(c);
}

int b200sv_set_acm_masks(b200sv_ctx* c, const uint8_t* self_enabled, const uint8_t* intra_enabled)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (!self_enabled || !intra_enabled)
    return fail(c, B200SV_ERR_INVALID, "null mask");
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t n = (size_t)c->cm.n_glinks;
  const std::vector<uint8_t> keep_s = c->cm.self_enabled, keep_i = c->cm.intra_enabled;
  c->cm.self_enabled.assign(self_enabled, self_enabled + n);
  c->cm.intra_enabled.assign(intra_enabled, intra_enabled + n * n);
  if (c->device < 0)
    return buildTables(c);
  enterCtx(c);
  int rc = rederiveTables(c);
  if (rc != B200SV_OK)
  {  // keep the context usable with the masks it had
    c->cm.self_enabled = keep_s;
    c->cm.intra_enabled = keep_i;
    const std::string msg = c->err;
    if (buildTables(c) == B200SV_OK)
      uploadTables(c);
    c->err = msg;
  }
  return rc;
}

int b200sv_set_base_state(b200sv_ctx* c, const double* base_positions, int rebuild_self_field)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (!base_positions)
    return fail(c, B200SV_ERR_INVALID, "null base state");
  std::lock_guard<std::mutex> lk(c->mu);
  c->robot.base_positions.assign(base_positions, base_positions + c->robot.n_vars);
  if (c->device < 0)
    return buildTables(c);
  enterCtx(c);
  int rc = rederiveTables(c);
  if (rc != B200SV_OK || !rebuild_self_field)
    return rc;
  if (c->self_local.empty() && !c->self_points.empty())
    return fail(c, B200SV_ERR_UNSUPPORTED, "the self field's link points are not known to this context");
  // the self field of the new cache entry: the non-group links' interior points at the new state (:907-953)
  poseSelfPoints(c->robot, c->robot.base_positions, c->self_local, c->self_points);
  Field& F = c->fields[B200SV_FIELD_SELF];
  const size_t n = c->self_points.size() / 3;
  if (F.prop)
  {  // a new cache entry's self field is a fresh PropagationDistanceField with the links' points added (:931-953)
    CU(c, c->d_points.ensure(3 * n + 3));
    if (n)
      CU(c, cudaMemcpyAsync(c->d_points.p, c->self_points.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    propDestroy(F.prop);  // fresh: nothing queued from the entry before
    F.prop = nullptr;
    CU(c, propCreate(c->grid, c->cfg.use_signed_distance_field != 0, &F.prop));
    CU(c, propReset(F.prop, c->stream));
    CU(c, propAddPoints(F.prop, c->d_points.p, n, c->stream));
    if ((rc = publishPropagation(c, B200SV_FIELD_SELF, c->stream)) != B200SV_OK)
      return rc;
    CU(c, cudaStreamSynchronize(c->stream));
    return B200SV_OK;
  }
  CU(c, cudaMemsetAsync(F.occ, 0, (size_t)c->grid.nx * c->grid.ny * c->grid.wz * 4, c->stream));
  if (n)
  {
    CU(c, c->d_points.ensure(3 * n));
    CU(c, cudaMemcpyAsync(c->d_points.p, c->self_points.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    if ((rc = addOrRemove(c, B200SV_FIELD_SELF, c->d_points.p, n, true, c->stream)) != B200SV_OK)
      return rc;
  }
  return rebuildField(c, B200SV_FIELD_SELF);
}

int b200sv_set_self_link_points(b200sv_ctx* c, int32_t link_index, const double* xyz_link_frame, size_t n_points)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (link_index < 0 || link_index >= c->robot.n_links || (n_points && !xyz_link_frame))
    return fail(c, B200SV_ERR_INVALID, "bad link index or null points");
  std::lock_guard<std::mutex> lk(c->mu);
  for (auto& lp : c->self_local)
    if (lp.first == link_index)
    {
      lp.second.assign(xyz_link_frame, xyz_link_frame + 3 * n_points);
      return B200SV_OK;
    }
  c->self_local.push_back({ link_index, std::vector<double>(xyz_link_frame, xyz_link_frame + 3 * n_points) });
  return B200SV_OK;
}

// ---- field replication across contexts / processes -----------------------------------------------------------------------
// layout of an exported field: [occupancy words | d2 uint16 | this field's compact elements], each part 256-byte aligned
static size_t exportParts(const b200sv_ctx* c, size_t* occ_bytes, size_t* d2_off, size_t* cmp_off)
{
  const size_t ob = (size_t)c->grid.nx * c->grid.ny * c->grid.wz * 4;
  const size_t a = (ob + 255) & ~(size_t)255, b = (c->n_voxels * 2 + 255) & ~(size_t)255;
  if (occ_bytes) *occ_bytes = ob;
  if (d2_off) *d2_off = a;
  if (cmp_off) *cmp_off = a + b;
  return a + b + ((c->n_voxels * (size_t)c->compact_bytes + 255) & ~(size_t)255);
}

int b200sv_field_export_size(const b200sv_ctx* c, size_t* n_bytes)
{
  if (!c || !n_bytes)
    return B200SV_ERR_INVALID;
  *n_bytes = exportParts(c, nullptr, nullptr, nullptr);
  return B200SV_OK;
}

static int fieldTransfer(b200sv_ctx* c, int which, uint8_t* d_buf, void* cuda_stream, bool to_buf)
{
  int rc = checkWhich(c, which);
  if (rc)
    return rc;
  if (!d_buf)
    return fail(c, B200SV_ERR_INVALID, "null device buffer");
  if (c->fields[which].nd2)
    return fail(c, B200SV_ERR_UNSUPPORTED, "signed fields are rebuilt per context, not replicated");
  if (!to_buf && (rc = notInReferenceMode(c, which)) != B200SV_OK)
    return rc;  // (exporting such a field is fine: its values replicate like any other's)
  std::lock_guard<std::mutex> lk(c->mu);
  cudaStream_t s = static_cast<cudaStream_t>(cuda_stream);
  if ((rc = enterUser(c, s)) != B200SV_OK)
    return rc;
  size_t ob, d2_off, cmp_off;
  exportParts(c, &ob, &d2_off, &cmp_off);
  Field& F = c->fields[which];
  if (to_buf)
  {
    CU(c, cudaMemcpyAsync(d_buf, F.occ, ob, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(d_buf + d2_off, F.d2, c->n_voxels * 2, cudaMemcpyDeviceToDevice, s));
  }
  else
  {
    CU(c, cudaMemcpyAsync(F.occ, d_buf, ob, cudaMemcpyDeviceToDevice, s));
    CU(c, cudaMemcpyAsync(F.d2, d_buf + d2_off, c->n_voxels * 2, cudaMemcpyDeviceToDevice, s));
  }
  // this field's elements of the interleaved compact field
  CU(c, launchCompactCopy(F.compact, d_buf + cmp_off, c->n_voxels, c->compact_bytes, to_buf, s));
  if (!to_buf)
    CU(c, cudaMemsetAsync(F.plane_busy, 1, (size_t)c->grid.nx * kPlaneBusyStride, s));
  return leaveUser(c, s);
}

int b200sv_field_export_device(b200sv_ctx* c, int which, void* d_buf, void* cuda_stream)
{
  return c ? fieldTransfer(c, which, static_cast<uint8_t*>(d_buf), cuda_stream, true) : B200SV_ERR_INVALID;
}

int b200sv_field_import_device(b200sv_ctx* c, int which, const void* d_buf, void* cuda_stream)
{
  return c ? fieldTransfer(c, which, static_cast<uint8_t*>(const_cast<void*>(d_buf)), cuda_stream, false) : B200SV_ERR_INVALID;
}

// ---- several GPUs behind one handle (states sharded, fields replicated) ---------------------------
// One context per device; a batch is cut into contiguous blocks (multiples of 32 states, so every block's bitmap is whole
// words and lands directly in the caller's buffer), all blocks are enqueued before any is waited for, and a field is built
// ONCE (on the first device) and copied to the others device-to-device (NVLink when peer access is available).

static int multiFail(b200sv_multi* m, int code, const std::string& msg)
{
  if (m)
    m->err = msg;
  else
    g_create_error = msg;
  return code;
}

static void enablePeers(const std::vector<b200sv_ctx*>& ctx)
{
  for (b200sv_ctx* a : ctx)
    for (b200sv_ctx* b : ctx)
      if (a->device != b->device)
      {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, a->device, b->device);
        if (can)
        {
          cudaSetDevice(a->device);
          if (cudaDeviceEnablePeerAccess(b->device, 0) != cudaSuccess)
            cudaGetLastError();  // already enabled
        }
      }
}

int b200sv_create_multi_from_urdf(const int* devices, int n_dev, const char* urdf_xml, const char* srdf_xml,
                                  const char* group_name, const b200sv_field_cfg* field, b200sv_multi** out)
{
  if (!devices || n_dev < 1 || n_dev > 64 || !out)
    return fail(nullptr, B200SV_ERR_INVALID, "bad device list");
  *out = nullptr;
  auto* m = new b200sv_multi();
  for (int i = 0; i < n_dev; ++i)
  {
    b200sv_ctx* c = nullptr;
    const int rc = b200sv_create_from_urdf(devices[i], urdf_xml, srdf_xml, group_name, field, &c);
    if (rc != B200SV_OK)
    {
      for (b200sv_ctx* d : m->ctx)
        b200sv_destroy(d);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
  }
  enablePeers(m->ctx);
  *out = m;
  return B200SV_OK;
}

int b200sv_create_multi(const int* devices, int n_dev, const b200sv_robot* robot, const b200sv_collision_model* model,
                        const b200sv_field_cfg* field, b200sv_multi** out)
{
  if (!devices || n_dev < 1 || n_dev > 64 || !out)
    return fail(nullptr, B200SV_ERR_INVALID, "bad device list");
  *out = nullptr;
  auto* m = new b200sv_multi();
  for (int i = 0; i < n_dev; ++i)
  {
    b200sv_ctx* c = nullptr;
    const int rc = b200sv_create(devices[i], robot, model, field, &c);
    if (rc != B200SV_OK)
    {
      for (b200sv_ctx* d : m->ctx)
        b200sv_destroy(d);
      delete m;
      return rc;
    }
    m->ctx.push_back(c);
  }
  enablePeers(m->ctx);
  *out = m;
  return B200SV_OK;
}

void b200sv_destroy_multi(b200sv_multi* m)
{
  if (!m)
    return;
  for (b200sv_ctx* c : m->ctx)
    b200sv_destroy(c);
  delete m;
}

const char* b200sv_multi_last_error(const b200sv_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }
int b200sv_multi_device_count(const b200sv_multi* m) { return m ? (int)m->ctx.size() : 0; }
b200sv_ctx* b200sv_multi_ctx(b200sv_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

// field `which` of context 0 -> every other context, device to device
int b200sv_multi_field_sync(b200sv_multi* m, int which)
{
  if (!m)
    return B200SV_ERR_INVALID;
  if (which != B200SV_FIELD_WORLD && which != B200SV_FIELD_SELF)
    return multiFail(m, B200SV_ERR_INVALID, "field selector must be B200SV_FIELD_WORLD or B200SV_FIELD_SELF");
  std::lock_guard<std::mutex> lk(m->mu);
  b200sv_ctx* src = m->ctx[0];
  std::lock_guard<std::mutex> ls(src->mu);
  enterCtx(src);
  if (cudaStreamSynchronize(src->stream) != cudaSuccess)
    return multiFail(m, B200SV_ERR_CUDA, "source stream failed");
  const size_t occ_bytes = (size_t)src->grid.nx * src->grid.ny * src->grid.wz * 4;
  for (size_t i = 1; i < m->ctx.size(); ++i)
  {
    b200sv_ctx* dst = m->ctx[i];
    std::lock_guard<std::mutex> ld(dst->mu);
    enterCtx(dst);
    Field &S = src->fields[which], &D = dst->fields[which];
    cudaError_t e = cudaMemcpyPeerAsync(D.occ, dst->device, S.occ, src->device, occ_bytes, dst->stream);
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(D.d2, dst->device, S.d2, src->device, src->n_voxels * 2, dst->stream);
    if (e == cudaSuccess && S.nd2 && D.nd2)
      e = cudaMemcpyPeerAsync(D.nd2, dst->device, S.nd2, src->device, src->n_voxels * 2, dst->stream);
    // the interleaved compact field holds BOTH fields; the other one is identical on every device already (every
    // field change goes through this function), so the whole array can be copied
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(dst->d_combined, dst->device, src->d_combined, src->device,
                              src->n_voxels * 2 * (size_t)src->compact_bytes, dst->stream);
    // both fields' plane flags travel with the array they describe
    for (int w = 0; w < 2 && e == cudaSuccess; ++w)
      e = cudaMemcpyPeerAsync(dst->fields[w].plane_busy, dst->device, src->fields[w].plane_busy, src->device,
                              (size_t)src->grid.nx * kPlaneBusyStride, dst->stream);
    if (e != cudaSuccess)
      return multiFail(m, B200SV_ERR_CUDA, std::string("peer copy of the field: ") + cudaGetErrorString(e));
  }
  for (size_t i = 1; i < m->ctx.size(); ++i)
  {
    cudaSetDevice(m->ctx[i]->device);
    if (cudaStreamSynchronize(m->ctx[i]->stream) != cudaSuccess)
      return multiFail(m, B200SV_ERR_CUDA, "peer copy of the field failed");
  }
  return B200SV_OK;
}

int b200sv_multi_field_add_points(b200sv_multi* m, int which, const double* xyz, size_t n_points)
{
  if (!m)
    return B200SV_ERR_INVALID;
  int rc = b200sv_field_add_points(m->ctx[0], which, xyz, n_points);
  if (rc != B200SV_OK)
    return multiFail(m, rc, m->ctx[0]->err);
  m->stats.edt_ms = m->ctx[0]->stats.edt_ms;
  return b200sv_multi_field_sync(m, which);
}

// The field is built on the first device and copied to the others, so only the first context runs a propagation at all:
// in B200SV_PROPAGATION_REFERENCE it keeps the reference's voxel records, the others receive the values.
int b200sv_multi_field_set_propagation(b200sv_multi* m, int which, int mode)
{
  if (!m)
    return B200SV_ERR_INVALID;
  int rc = b200sv_field_set_propagation(m->ctx[0], which, mode);
  if (rc != B200SV_OK)
    return multiFail(m, rc, m->ctx[0]->err);
  return b200sv_multi_field_sync(m, which);
}

int b200sv_multi_field_reset(b200sv_multi* m, int which)
{
  if (!m)
    return B200SV_ERR_INVALID;
  for (b200sv_ctx* c : m->ctx)
    if (int rc = b200sv_field_reset(c, which))
      return multiFail(m, rc, c->err);
  return B200SV_OK;
}

// Batched check over every device of the handle: block i of the states goes to context i.  q and valid_bitmap are host
// buffers (pinned memory makes every copy a DMA that overlaps the other devices' work).
int b200sv_check_batch_multi(b200sv_multi* m, const double* q, size_t n, uint32_t flags, uint8_t* valid_bitmap)
{
  if (!m)
    return B200SV_ERR_INVALID;
  if (n && (!q || !valid_bitmap))
    return multiFail(m, B200SV_ERR_INVALID, "null host pointer");
  if ((flags & 7u) == 0)
    return multiFail(m, B200SV_ERR_INVALID, "flags select no check");
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(m->mu);
  const size_t G = m->ctx.size(), V = m->ctx[0]->robot.group_var_index.size();
  const size_t per = ((n + G - 1) / G + 31) / 32 * 32;  // states per device: whole bitmap words
  struct Part
  {
    b200sv_ctx* c;
    size_t lo, cnt;
    bool small;
  };
  std::vector<Part> parts;
  for (size_t g = 0; g < G; ++g)
  {
    const size_t lo = std::min(g * per, n), hi = std::min(lo + per, n);
    if (hi > lo)
      parts.push_back(Part{ m->ctx[g], lo, hi - lo, hi - lo <= kSmallBatch });
  }
  int rc = B200SV_OK;
  size_t locked = 0;
  for (; locked < parts.size(); ++locked)
    parts[locked].c->mu.lock();
  // enqueue everything, then wait: the devices work concurrently
  size_t enq = 0;
  for (; enq < parts.size() && rc == B200SV_OK; ++enq)
  {
    Part& p = parts[enq];
    enterCtx(p.c);
    if (p.small)
      continue;  // latency-sized tails run in the second loop
    rc = checkHostEnqueue(p.c, q + p.lo * V, p.cnt, flags, valid_bitmap + p.lo / 8);
    if (rc != B200SV_OK)
      multiFail(m, rc, p.c->err);
  }
  uint64_t valid = 0, rechecked = 0;
  float ms = 0.f;
  for (size_t i = 0; i < parts.size(); ++i)
  {
    Part& p = parts[i];
    if (rc == B200SV_OK && i < enq)
    {
      int r = p.small ? checkSmall(p.c, q + p.lo * V, p.cnt, flags, valid_bitmap + p.lo / 8) :
                        checkHostFinish(p.c, p.cnt, flags, valid_bitmap + p.lo / 8);
      if (r != B200SV_OK)
      {
        rc = r;
        multiFail(m, rc, p.c->err);
      }
      valid += p.c->stats.n_valid;
      rechecked += p.c->stats.n_rechecked;
      ms = std::max(ms, p.c->stats.check_ms);
    }
    else if (i < enq && !p.small)
    {
      cudaSetDevice(p.c->device);
      cudaStreamSynchronize(p.c->stream);  // a failed call leaves nothing in flight
    }
  }
  for (size_t i = 0; i < locked; ++i)
    parts[i].c->mu.unlock();
  m->stats.n_states = n;
  m->stats.n_valid = valid;
  m->stats.n_rechecked = rechecked;
  m->stats.check_ms = ms;
  return rc;
}

int b200sv_multi_get_stats(b200sv_multi* m, b200sv_stats* stats)
{
  if (!m || !stats)
    return B200SV_ERR_INVALID;
  std::lock_guard<std::mutex> lk(m->mu);
  *stats = m->stats;
  return B200SV_OK;
}

// ---- per-link PosedDistanceFields (getGroupStateRepresentation :1177-1197, getSelfProximityGradients :386-411) ----------
static int linkFieldDims(const b200sv_ctx* c, const double* size, int n[3])
{
  const double oo = 1.0 / c->cfg.resolution;
  for (int k = 0; k < 3; ++k)
  {
    n[k] = (int)(size[k] * oo);  // VoxelGrid::resize (voxel_grid.hpp:384)
    if (n[k] < 1)
      return 0;
  }
  return 1;
}

int b200sv_set_link_field(b200sv_ctx* c, int32_t glink, const double* size, const double* corner, const int32_t* d2)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (glink < 0 || glink >= c->cm.n_glinks || !size || !corner)
    return fail(c, B200SV_ERR_INVALID, "bad entry index or null geometry");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (c->cfg.use_signed_distance_field)
    return fail(c, B200SV_ERR_UNSUPPORTED, "per-link fields of a signed context are not implemented");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  c->link_fields.resize((size_t)c->cm.n_glinks);
  b200sv_ctx::LinkField& F = c->link_fields[(size_t)glink];
  CU(c, cudaStreamSynchronize(c->stream));
  if (F.d2)
    cudaFree(F.d2);
  F = b200sv_ctx::LinkField();
  c->lf_dirty = true;
  if (!d2)
    return B200SV_OK;  // the link has no field (no geometry)
  int n[3];
  if (!linkFieldDims(c, size, n))
    return fail(c, B200SV_ERR_INVALID, "link field has no cells along one axis");
  const size_t nv = (size_t)n[0] * n[1] * n[2];
  CU(c, cudaMalloc(&F.d2, nv * 2));
  CU(c, c->d_i32.ensure(nv));
  CU(c, cudaMemcpyAsync(c->d_i32.p, d2, nv * 4, cudaMemcpyHostToDevice, c->stream));
  GridDev g = c->grid;
  g.nx = n[0];
  g.ny = n[1];
  g.nz = n[2];
  CU(c, launchInjectNegD2(g, c->d_i32.p, F.d2, c->stream));  // clamps to [0, max_d2], narrows to 16 bits
  CU(c, cudaStreamSynchronize(c->stream));
  for (int k = 0; k < 3; ++k)
  {
    F.n[k] = n[k];
    F.corner[k] = corner[k];
  }
  return B200SV_OK;
}

int b200sv_get_link_field(b200sv_ctx* c, int32_t glink, int32_t* dims, double* corner, int32_t* d2)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (glink < 0 || glink >= c->cm.n_glinks || !dims)
    return fail(c, B200SV_ERR_INVALID, "bad entry index or null dims");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  dims[0] = dims[1] = dims[2] = 0;
  if ((size_t)glink >= c->link_fields.size() || !c->link_fields[(size_t)glink].d2)
    return B200SV_OK;
  const b200sv_ctx::LinkField& F = c->link_fields[(size_t)glink];
  for (int k = 0; k < 3; ++k)
  {
    dims[k] = F.n[k];
    if (corner)
      corner[k] = F.corner[k];
  }
  if (d2)
  {
    const size_t nv = (size_t)F.n[0] * F.n[1] * F.n[2];
    std::vector<uint16_t> h(nv);
    CU(c, cudaMemcpyAsync(h.data(), F.d2, nv * 2, cudaMemcpyDeviceToHost, c->stream));
    CU(c, cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < nv; ++i)
      d2[i] = h[i];
  }
  return B200SV_OK;
}

int b200sv_build_link_fields(b200sv_ctx* c)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (c->link_points.empty())
    return fail(c, B200SV_ERR_UNSUPPORTED, "the links' interior points are known on the URDF path only: use b200sv_set_link_field");
  if (c->cfg.use_signed_distance_field)
    return fail(c, B200SV_ERR_UNSUPPORTED, "per-link fields of a signed context are not implemented");
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const int L = c->cm.n_glinks;
  c->link_fields.resize((size_t)L);
  c->lf_dirty = true;
  CU(c, cudaStreamSynchronize(c->stream));
  for (int g = 0; g < L; ++g)
  {
    b200sv_ctx::LinkField& F = c->link_fields[(size_t)g];
    if (F.d2)
      cudaFree(F.d2);
    F = b200sv_ctx::LinkField();
    if (c->cm.sphere_start[g + 1] <= c->cm.sphere_start[g])
      continue;
    const std::vector<double>* pts = nullptr;
    for (const auto& lp : c->link_points)
      if (lp.first == c->cm.glink_index[g])
        pts = &lp.second;
    if (!pts)
      continue;
    // size = the bounding sphere's diameter, corner = its centre - size / 2, in the link frame (:1177-1188)
    const double* bs = &c->cm.bounding_sphere[4 * (size_t)g];
    const double diameter = 2 * bs[3];
    const double size[3] = { diameter, diameter, diameter };
    int n[3];
    if (!linkFieldDims(c, size, n))
      continue;
    GridDev gl = c->grid;
    gl.nx = n[0];
    gl.ny = n[1];
    gl.nz = n[2];
    gl.wz = (gl.nz + 31) / 32;
    for (int k = 0; k < 3; ++k)
    {
      F.corner[k] = bs[k] - 0.5 * size[k];
      gl.om[k] = F.corner[k] - 0.5 * c->cfg.resolution;
      F.n[k] = n[k];
    }
    const size_t nv = (size_t)n[0] * n[1] * n[2], occ_words = (size_t)n[0] * n[1] * gl.wz, np = pts->size() / 3;
    uint32_t* occ = nullptr;
    uint16_t* tmp = nullptr;
    CU(c, cudaMalloc(&F.d2, nv * 2));
    CU(c, cudaMalloc(&occ, occ_words * 4));
    CU(c, cudaMalloc(&tmp, nv * 2));
    CU(c, cudaMemsetAsync(occ, 0, occ_words * 4, c->stream));
    cudaError_t e = cudaSuccess;
    if (np)
    {
      e = c->d_points.ensure(3 * np);
      if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->d_points.p, pts->data(), 3 * np * sizeof(double), cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess)
        e = launchScatter(c->d_points.p, np, gl, occ, true, c->stream);
    }
    if (e == cudaSuccess)
      e = launchEdt(gl, occ, tmp, F.d2, nullptr, c->compact_bytes, c->sm_count, c->stream);
    if (e == cudaSuccess)
      e = cudaStreamSynchronize(c->stream);
    cudaFree(occ);
    cudaFree(tmp);
    if (e != cudaSuccess)
      return fail(c, B200SV_ERR_CUDA, std::string("building a link field: ") + cudaGetErrorString(e));
  }
  return B200SV_OK;
}

int b200sv_set_gradient_link_mask(b200sv_ctx* c, const uint8_t* enabled)
{
  if (!c)
    return B200SV_ERR_INVALID;
  std::lock_guard<std::mutex> lk(c->mu);
  const size_t L = (size_t)c->cm.n_glinks;
  if (enabled)
    c->lf_enabled.assign(enabled, enabled + L * L);
  else
    c->lf_enabled.clear();
  c->lf_dirty = true;
  return B200SV_OK;
}

int b200sv_get_fp32_budget(const b200sv_ctx* c, double* eps_used_m, double* eps_derived_m)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (eps_used_m)
    *eps_used_m = c->fp32_eps;
  if (eps_derived_m)
    *eps_derived_m = c->fp32_eps_derived;
  return B200SV_OK;
}

int b200sv_debug_fast_centres(b200sv_ctx* c, const double* q, size_t n, double* centres)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (n && (!q || !centres))
    return fail(c, B200SV_ERR_INVALID, "null host pointer");
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  if (!c->fast_ok)
    return fail(c, B200SV_ERR_UNSUPPORTED, "the group does not take the fast FP32 kernel");
  if (n == 0)
    return B200SV_OK;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t V = c->robot.group_var_index.size();
  const auto* h = reinterpret_cast<const fast::Header*>(c->blob_fast.data());
  const size_t n_hot = c->fast_hot_sphere.size(), S = (size_t)c->n_spheres;
  DevBuf<float> scratch, out;
  CU(c, c->d_q.ensure(n * V));
  CU(c, scratch.ensure(n * (size_t)std::max(h->state_stride, 1)));
  CU(c, out.ensure(n * n_hot * 3));
  CU(c, cudaMemcpyAsync(c->d_q.p, q, n * V * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  cudaError_t e = launchFastCentres(c->d_blob_fast, (int)c->blob_fast.size(), c->d_q.p, n, (int)V, c->fast_wide, scratch.p, out.p,
                                    (int)n_hot, c->stream);
  std::vector<float> host(n * n_hot * 3);
  if (e == cudaSuccess)
    e = cudaMemcpyAsync(host.data(), out.p, host.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  scratch.release();
  out.release();
  if (e != cudaSuccess)
    return fail(c, B200SV_ERR_CUDA, std::string("fast centres probe: ") + cudaGetErrorString(e));
  // cells -> metres: p = om + f * res (the inverse of VoxelGrid::getCellFromLocation's affine map, in double)
  const double res = c->cfg.resolution;
  for (size_t i = 0; i < n; ++i)
    for (size_t k = 0; k < n_hot; ++k)
    {
      const int sp = c->fast_hot_sphere[k];
      if (sp < 0 || (size_t)sp >= S)
        continue;
      for (int a = 0; a < 3; ++a)
        centres[(i * S + sp) * 3 + a] = c->grid.om[a] + (double)host[(i * n_hot + k) * 3 + a] * res;
    }
  return B200SV_OK;
}

int b200sv_gather_roofline(b200sv_ctx* c, size_t n_reads, int iters, double* gbs)
{
  if (!c || !gbs || iters < 1)
    return B200SV_ERR_INVALID;
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  const size_t field_bytes = c->n_voxels * 2 * (size_t)c->compact_bytes;
  const unsigned long long n_sectors = field_bytes / 32;
  if (n_sectors == 0)
    return fail(c, B200SV_ERR_INVALID, "field smaller than one sector");
  const int block = 256, per_thread = 64;
  const int grid = (int)std::min<size_t>((n_reads + (size_t)block * per_thread - 1) / ((size_t)block * per_thread), 1u << 20);
  const uint8_t* f = static_cast<const uint8_t*>(c->d_combined);
  CU(c, launchGather(f, n_sectors, per_thread, grid, block, c->d_sink, c->stream));  // warm-up
  float best = 1e30f;
  for (int i = 0; i < iters; ++i)
  {
    CU(c, cudaEventRecord(c->ev0, c->stream));
    CU(c, launchGather(f, n_sectors, per_thread, grid, block, c->d_sink, c->stream));
    CU(c, cudaEventRecord(c->ev1, c->stream));
    CU(c, cudaEventSynchronize(c->ev1));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev0, c->ev1);
    best = std::min(best, ms);
  }
  const double reads = (double)grid * block * per_thread;
  *gbs = reads * 32.0 / (best * 1e-3) / 1e9;
  return B200SV_OK;
}

int b200sv_set_tuning(b200sv_ctx* c, double fp32_eps_m, int warps_per_block, int blocks_per_sm)
{
  if (!c)
    return B200SV_ERR_INVALID;
  if (needDevice(c))
    return B200SV_ERR_NO_DEVICE;
  std::lock_guard<std::mutex> lk(c->mu);
  enterCtx(c);
  if (fp32_eps_m > 0.0)
  {
    c->fp32_eps = fp32_eps_m;
    c->fp32_eps_user = true;
  }
  if (warps_per_block > 0)
    c->tune_warps = warps_per_block;
  if (blocks_per_sm > 0)
    c->tune_blocks = blocks_per_sm;
  CU(c, cudaStreamSynchronize(c->stream));
  int rc = buildTables(c);
  if (rc != B200SV_OK)
    return rc;
  return uploadTables(c);
}

}  // extern "C"
