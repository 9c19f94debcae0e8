// K1 + K2, FAST pass: batched forward kinematics fused with the sphere-decomposition checks, FP32, thread-per-state.
//
// Same reference functions as check_kernels.cu (RobotState::updateLinkTransforms, PosedBodySphereDecomposition::
// updatePose, getCollisionSphereCollision, getEnvironmentCollisions / getSelfCollisions / getIntraGroupCollisions:
// robot_state/src/robot_state.cpp:793-848, collision_distance_field/src/collision_distance_field_types.cpp:211-236,
// 388-418, collision_env_distance_field.cpp:274-353, 430-642, 1561-1643), restructured around the instruction
// count per state instead of around the reference's data structures:
//
//   * Everything is computed in VOXEL coordinates: the host folds VoxelGrid::getCellFromLocation's affine map
//     (voxel_grid.hpp:522) into the root transforms, so a posed sphere centre is its fractional cell coordinate and no
//     per-sphere conversion is left.
//   * A joint's transform is ONE affine combination of host-built constants (origin * Rodrigues terms) followed by
//     ONE 3x4 product with the parent, and the link loop keeps the current transform in registers: the spheres of a
//     link are posed from registers right after its FK step (no shared-memory round trip, no second sphere phase).
//   * The cell index comes from a round-down float add (no F2I), both cells `margin` either side of the centre
//     are indexed, and a sphere whose two indices differ (its FP32 centre is within `margin` of a cell face) is parked
//     in a one-entry per-lane slot and resolved by reading BOTH cells; nothing else branches.  Centres are clamped to
//     the field's 1-cell border shell, which the field builders keep at "free": DistanceField::getDistanceGradient
//     (distance_field.cpp:82) reports every such cell, and everything outside, as max_distance.
//   * Gathers are software pipelined: the loads of one sphere chunk are consumed while the next chunk is posed.
//   * Link transforms also go to shared memory (3x4 rows of 16 bytes) for the intra-group pair phase, which is the
//     queue-compacted three-level test of check_kernels.cu in voxel units.
//
// A state whose boolean could depend on FP32 rounding is appended to the recheck list; the exact pass
// (checkKernel<double, ., true>, the reference's arithmetic operation by operation) decides it.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <type_traits>

#include "device_model.hpp"
#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;
constexpr int kU = B200SV_FAST_UNROLL;  // Hot ranges are padded per link to a multiple of this (the tail chunk)
#ifndef B200SV_FAST_THREADS
#define B200SV_FAST_THREADS 512
#endif
#ifndef B200SV_FAST_MINBLOCKS
#define B200SV_FAST_MINBLOCKS 1
#endif
#ifndef B200SV_MAXVARS
#define B200SV_MAXVARS 16
#endif
#ifndef B200SV_FAST_THREADS_SMALL
#define B200SV_FAST_THREADS_SMALL 192
#endif
#ifndef B200SV_FAST_MINBLOCKS_SMALL
#define B200SV_FAST_MINBLOCKS_SMALL 3
#endif
constexpr int kMaxVars = B200SV_MAXVARS;             // group variables the register prefetch covers (the host checks)
constexpr int kP = 4;                   // gathers in flight per lane = spheres of the main chunk
constexpr int kQueueCap = 256;          // (state, pair) work items per warp; flushed when fewer than 32 slots remain
constexpr float kMagic = 12582912.f;    // 1.5 * 2^23: x + kMagic rounded down has floor(x) in its low mantissa bits
constexpr unsigned kGuard = 0x01000100u;

struct FModel
{
  const fast::Header* h;
  const fast::Link* links;
  const fast::Hot* hot;
  const fast::PSphere* ps;
  const fast::Cluster* cl;
  const fast::Group* groups;
  const fast::Pair* pairs;
  const fast::Quirk* quirks;
  const uint4* linkpairs;
};

// Field element = { world, self } of one voxel.  8-bit fields: both compares in one add (per 16-bit lane
// 0x100 - thr + d has bit 8 set iff d >= thr).  16-bit fields: plain compares.
template <typename FT> struct Cmp;
template <> struct Cmp<uint8_t>
{
  using Elem = uint16_t;
  static constexpr unsigned kNever = kGuard;
  static __device__ __forceinline__ unsigned mask(unsigned t, unsigned fmask) { return (t & fmask) | (kGuard & ~fmask); }
  // nonzero iff either field's distance is below its threshold
  static __device__ __forceinline__ unsigned hit(unsigned v, unsigned t)
  {
    return ((__byte_perm(v, 0u, 0x3120) + t) & kGuard) ^ kGuard;
  }
};
template <> struct Cmp<uint16_t>
{
  using Elem = uint32_t;
  static constexpr unsigned kNever = 0u;
  static __device__ __forceinline__ unsigned mask(unsigned t, unsigned fmask) { return t & fmask; }
  static __device__ __forceinline__ unsigned hit(unsigned v, unsigned t)
  {
    return static_cast<unsigned>(((v & 0xffffu) < (t & 0xffffu)) | ((v >> 16) < (t >> 16)));
  }
};

__device__ __forceinline__ void load12(const float* p, float* o)
{
  const float4* v = reinterpret_cast<const float4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i)
  {
    const float4 t = v[i];
    o[4 * i + 0] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void store12(float* p, const float* o)
{
  float4* v = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    v[i] = make_float4(o[4 * i + 0], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}
// Where the per-state link transforms of a warp's tile live.  kGlobal = false: shared memory, one block of
// `stride` words per state (3x4 rows of 16 bytes).  kGlobal = true (robots whose transforms would leave shared memory
// for only a few warps): a per-warp region of an L2-resident scratch buffer, laid out [link row][state] so the
// 32 lanes of the link loop store and load 512 contiguous bytes.
template <bool kGlobal>
struct TStore
{
  float* base;
  int stride;
  __device__ __forceinline__ void load(int off, int state, float* o) const
  {
    if (kGlobal)
    {
      const float4* v = reinterpret_cast<const float4*>(base) + (off >> 2) * 32 + state;
#pragma unroll
      for (int i = 0; i < 3; ++i)
      {
        const float4 t = v[i * 32];
        o[4 * i + 0] = t.x; o[4 * i + 1] = t.y; o[4 * i + 2] = t.z; o[4 * i + 3] = t.w;
      }
    }
    else
      load12(base + state * stride + off, o);
  }
  __device__ __forceinline__ void store(int off, int state, const float* o) const
  {
    if (kGlobal)
    {
      float4* v = reinterpret_cast<float4*>(base) + (off >> 2) * 32 + state;
#pragma unroll
      for (int i = 0; i < 3; ++i)
        v[i * 32] = make_float4(o[4 * i + 0], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
    else
      store12(base + state * stride + off, o);
  }
};

__device__ __forceinline__ void pose(const float* T, float x, float y, float z, float& ox, float& oy, float& oz)
{
  ox = fmaf(T[0], x, fmaf(T[1], y, fmaf(T[2], z, T[3])));
  oy = fmaf(T[4], x, fmaf(T[5], y, fmaf(T[6], z, T[7])));
  oz = fmaf(T[8], x, fmaf(T[9], y, fmaf(T[10], z, T[11])));
}

// kHier (robots with thousands of cluster pairs): the pair, quirk and link-pair tables -- tens of KB, read with one index per
// lane -- stay in global memory (L1 / L2 resident); shared memory holds the head of the blob only, which is what decides how
// many warps fit next to it.
template <bool kHier>
__device__ __forceinline__ FModel loadFast(const uint8_t* __restrict__ g_model, int model_bytes, uint8_t* smem)
{
  const int4* src = reinterpret_cast<const int4*>(g_model);
  int4* dst = reinterpret_cast<int4*>(smem);
  const int copy_bytes = kHier ? reinterpret_cast<const fast::Header*>(g_model)->off_pairs : model_bytes;
  for (int i = threadIdx.x; i < copy_bytes / 16; i += blockDim.x)
    dst[i] = src[i];
  __syncthreads();
  FModel m;
  m.h = reinterpret_cast<const fast::Header*>(smem);
  m.links = reinterpret_cast<const fast::Link*>(smem + m.h->off_links);
  m.hot = reinterpret_cast<const fast::Hot*>(smem + m.h->off_hot);
  m.ps = reinterpret_cast<const fast::PSphere*>(smem + m.h->off_psphere);
  m.cl = reinterpret_cast<const fast::Cluster*>(smem + m.h->off_clusters);
  m.groups = reinterpret_cast<const fast::Group*>(smem + m.h->off_groups);
  const uint8_t* tail = kHier ? g_model : smem;
  m.pairs = reinterpret_cast<const fast::Pair*>(tail + m.h->off_pairs);
  m.quirks = reinterpret_cast<const fast::Quirk*>(tail + m.h->off_quirks);
  m.linkpairs = reinterpret_cast<const uint4*>(tail + m.h->off_linkpairs);
  return m;
}

__device__ __forceinline__ float sqrtApprox(float x)
{
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// One (state, cluster pair) item that survived the level-1 cull: [the reference's link-level cull when the host could
// not prove it away,] then every sphere of cluster a against every sphere of cluster b, dist < r1 + r2
// (getIntraGroupCollisions, collision_env_distance_field.cpp:571-638).  Clusters are padded to kClusterSpheres
// entries whose radius is -1e30, so the 4 x 4 block is branch-free.  T: the state's transform block.
template <bool kGlobal>
__device__ __forceinline__ void pairItem(const FModel& m, const TStore<kGlobal>& ts, int state, unsigned p, bool& hit, bool& amb)
{
  constexpr int kC = fast::kClusterSpheres;
  const fast::Header& h = *m.h;
  const uint4 pr = reinterpret_cast<const uint4*>(m.pairs)[p];  // rt2, a | b << 16, a_off | b_off << 16, quirk
  const int ca = pr.y & 0xffffu, cb = pr.y >> 16, a_off = pr.z & 0xffffu, b_off = pr.z >> 16;
  const int4 p1 = make_int4(0, static_cast<int>(pr.w), 0, 0);
  float Ta[12], Tb[12];
  ts.load(a_off, state, Ta);
  ts.load(b_off, state, Tb);
  bool cert = true;
  if (p1.y >= 0)
  {
    // the cull passes if ANY record of the chain does: one record for a pair of links, one per (shape, shape)
    // combination when an attached body of several shapes is involved (collision_env_distance_field.cpp:455-507; the
    // shapes of a body ride on one link, so every record is posed with this pair's two transforms)
    bool maybe = false;
    cert = false;
    for (int qi = p1.y; qi >= 0 && !cert; )
    {
      const fast::Quirk Q = m.quirks[qi];
      float ax, ay, az, bx, by, bz;
      pose(Ta, Q.ax, Q.ay, Q.az, ax, ay, az);
      pose(Tb, Q.bx, Q.by, Q.bz, bx, by, bz);
      const float ex = ax - bx, ey = ay - by, ez = az - bz;
      const float d2 = ex * ex + ey * ey + ez * ez;
      maybe |= d2 < Q.lim_v2 + Q.eps_v2;
      cert = d2 < Q.lim_v2 - Q.eps_v2;
      qi = Q.next;
    }
    if (!maybe)
      return;
  }
  float bx[kC], by[kC], bz[kC], br[kC];
#pragma unroll
  for (int j = 0; j < kC; ++j)
  {
    const float4 sp = reinterpret_cast<const float4*>(m.ps)[fast::kPSphereStride * cb + j];
    pose(Tb, sp.x, sp.y, sp.z, bx[j], by[j], bz[j]);
    br[j] = sp.w;
  }
  bool any_cert = false, any_maybe = false;
  const float eps = h.pair_eps_v;
#pragma unroll
  for (int i = 0; i < kC; ++i)
  {
    const float4 sp = reinterpret_cast<const float4*>(m.ps)[fast::kPSphereStride * ca + i];
    float ax, ay, az;
    pose(Ta, sp.x, sp.y, sp.z, ax, ay, az);
#pragma unroll
    for (int j = 0; j < kC; ++j)
    {
      const float gx = ax - bx[j], gy = ay - by[j], gz = az - bz[j];
      // gradient.norm() - r1 - r2 < 0 (collision_env_distance_field.cpp:577-583) with the FP32 uncertainty either side
      const float gap = (sqrtApprox(fmaf(gx, gx, fmaf(gy, gy, gz * gz))) - sp.w) - br[j];
      any_cert |= gap < -eps;
      any_maybe |= gap < eps;
    }
  }
  hit = cert && any_cert;
  amb = any_maybe && !hit;
}

// Posed cluster centres are parked as three half arrays x, y, z of [cluster][state], cluster stride 34 halves = 17 words:
// the 32 lanes of the level-1 cull (one pair each, same two states) then load from 32 distinct banks, and two adjacent
// states share one 32-bit word (the cull works on half2 = two states at a time).
constexpr int kRepStride = 34;
__device__ __host__ inline int repBytes(int n_clusters) { return (3 * n_clusters * kRepStride * 2 + 15) & ~15; }
// the hierarchical cull's cluster-level centres in global scratch: [cluster][32 states] of { x | y << 16, z } halves
__device__ __host__ inline int globalRepBytes(int n_clusters) { return n_clusters * 32 * 8; }

// The work queue of the pair phase.  The wide instance (no TMA staging) also keeps the tile's states there during the FK
// phase: the two are never live together.
__device__ __host__ inline int queueBytes(int n_group_vars, bool wide)
{
  const int stage = wide ? 32 * n_group_vars * 8 : 0;
  return stage > kQueueCap * 4 ? stage : kQueueCap * 4;
}
// The staged states of the other instances: two buffers of one tile each (32 * V doubles), filled by 1-D bulk copies
// (cp.async.bulk, the TMA engine) that complete on the two mbarriers behind them.
__device__ __host__ inline int stageOneBytes(int n_group_vars) { return (32 * n_group_vars * 8 + 15) & ~15; }
// Small groups (an arm) double-buffer: the next tile arrives while the current one is worked on.  Larger groups come with
// robots whose parked cluster centres already take 20 KB of shared memory per warp and run at 7-8 warps per SM: they keep
// ONE stage buffer, refilled after the FK phase has read the tile's last joint value (the copy overlaps the pair phase).
__device__ __host__ inline int stageCount(int n_group_vars) { return n_group_vars <= 8 ? 2 : 1; }
__device__ __host__ inline int stageBytes(int n_group_vars, bool wide)
{
  return wide ? 0 : stageCount(n_group_vars) * stageOneBytes(n_group_vars) + 16;
}

constexpr int kQcCap = 512;   // hierarchical cull: (state, cluster pair) candidates of one batch of link-pair survivors

__device__ __host__ inline int fastPerWarpBytes(int state_stride, int n_clusters, int n_group_vars, int n_linkpairs, bool wide)
{
  return 32 * state_stride * 4 + repBytes(n_clusters) + queueBytes(n_group_vars, wide) + stageBytes(n_group_vars, wide) + 16 +
         (n_linkpairs > 0 ? (kQueueCap + kQcCap) * 4 : 0);
}

// ---- 1-D bulk copies global -> shared memory through the TMA engine, completion on an mbarrier (sm_90+ PTX) ----------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
  // one thread: announce the bytes, then start the copy; the barrier's phase completes when they have landed
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dst)),
               "l"(src), "r"(bytes), "r"(smemAddr(bar))
               : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, unsigned parity)
{
  unsigned ok = 0;
  const uint32_t a = smemAddr(bar);
  do
  {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!ok);
}

// One link of the FK chain: T (the parent's transform on entry) becomes T_parent * [M | mt], M = A0 + sin q A1 + (1 - cos q) A2
// from the link's host-built constants (see the file header).  Shared by the check kernel and fastCentresKernel (the FP32
// error-budget probe), so both run the same arithmetic by construction.
template <int kV>
__device__ __forceinline__ void fastLinkStep(const int4& m0, const int4* Li, const float4* Lf, const float4& o, double q_cur,
                                             const double* qrow, float* T)
{
  if (m0.x != fast::kAlias)  // an alias link is posed by its anchor's transform as it is (host-folded geometry)
  {
    const double2 ab = reinterpret_cast<const double2*>(Li)[2];
    const float4 c0 = Lf[4], c1 = Lf[5], c2 = Lf[6];  // a0[0..8], a1[0..2]
    float M[9] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, c2.x };
    float mt[3] = { o.x, o.y, o.z };
    if (m0.x != FIXED)
    {
      // setJointGroupPositions / updateMimicJoints in double, then narrowed (k < 0: a = 0, the value is b)
      const float qf = static_cast<float>(ab.x * q_cur + ab.y);
      if (m0.x == REVOLUTE)
      {
        const float4 c3 = Lf[7], c4 = Lf[8], c5 = Lf[9], c6 = Lf[10];  // a1[3..8], a2[0..8]
        const float A1[9] = { c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w, c4.x, c4.y };
        const float A2[9] = { c4.z, c4.w, c5.x, c5.y, c5.z, c5.w, c6.x, c6.y, c6.z };
        float s, c;
        sincosf(qf, &s, &c);
        const float t = 1.f - c;
#pragma unroll
        for (int i = 0; i < 9; ++i)
          M[i] = fmaf(t, A2[i], fmaf(s, A1[i], M[i]));
      }
      else if (kV == 0 && m0.x == PLANAR)
      {  // variables k, k + 1, k + 2 of the group state = x, y, theta (the host checked a = 1, b = 0 for all three);
         // [O | ot] * [Rz(theta) | (x, y, 0)] = [A0 + s A1 + t A2 (axis z) | ot + x O(:,0) + y O(:,1)]
        const float4 c3 = Lf[7], c4 = Lf[8], c5 = Lf[9], c6 = Lf[10];
        const float A1[9] = { c2.y, c2.z, c2.w, c3.x, c3.y, c3.z, c3.w, c4.x, c4.y };
        const float A2[9] = { c4.z, c4.w, c5.x, c5.y, c5.z, c5.w, c6.x, c6.y, c6.z };
        const float px = static_cast<float>(qrow[m0.y]), py = static_cast<float>(qrow[m0.y + 1]);
        const float th = static_cast<float>(qrow[m0.y + 2]);
        mt[0] = fmaf(px, M[0], fmaf(py, M[1], mt[0]));
        mt[1] = fmaf(px, M[3], fmaf(py, M[4], mt[1]));
        mt[2] = fmaf(px, M[6], fmaf(py, M[7], mt[2]));
        float s, c;
        sincosf(th, &s, &c);
        const float t = 1.f - c;
#pragma unroll
        for (int i = 0; i < 9; ++i)
          M[i] = fmaf(t, A2[i], fmaf(s, A1[i], M[i]));
      }
      else if (kV == 0 && m0.x == FLOATING)
      {  // variables k .. k + 6 = translation, then the quaternion x, y, z, w, normalised by the joint
         // (floating_joint_model.cpp:231-236); [O | ot] * [R(q) | p] = [O R(q) | ot + O p]
        const float px = static_cast<float>(qrow[m0.y]), py = static_cast<float>(qrow[m0.y + 1]),
                    pz = static_cast<float>(qrow[m0.y + 2]);
        float x = static_cast<float>(qrow[m0.y + 3]), y = static_cast<float>(qrow[m0.y + 4]),
              z = static_cast<float>(qrow[m0.y + 5]), w = static_cast<float>(qrow[m0.y + 6]);
        const float inv = 1.f / sqrtf(fmaf(x, x, fmaf(y, y, fmaf(z, z, w * w))));
        x *= inv; y *= inv; z *= inv; w *= inv;
        const float tx = 2.f * x, ty = 2.f * y, tz = 2.f * z;
        const float twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y,
                    tyz = tz * y, tzz = tz * z;
        const float Rq[9] = { 1.f - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1.f - (txx + tzz), tyz - twx,
                              txz - twy, tyz + twx, 1.f - (txx + tyy) };
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
          const float o0 = M[3 * r], o1 = M[3 * r + 1], o2 = M[3 * r + 2];
          mt[r] = fmaf(o0, px, fmaf(o1, py, fmaf(o2, pz, mt[r])));
          M[3 * r + 0] = fmaf(o0, Rq[0], fmaf(o1, Rq[3], o2 * Rq[6]));
          M[3 * r + 1] = fmaf(o0, Rq[1], fmaf(o1, Rq[4], o2 * Rq[7]));
          M[3 * r + 2] = fmaf(o0, Rq[2], fmaf(o1, Rq[5], o2 * Rq[8]));
        }
      }
      else
      {  // PRISMATIC
        mt[0] = fmaf(qf, c2.y, mt[0]);
        mt[1] = fmaf(qf, c2.z, mt[1]);
        mt[2] = fmaf(qf, c2.w, mt[2]);
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
      const float p0 = T[4 * r + 0], p1 = T[4 * r + 1], p2 = T[4 * r + 2];
      T[4 * r + 0] = fmaf(p0, M[0], fmaf(p1, M[3], p2 * M[6]));
      T[4 * r + 1] = fmaf(p0, M[1], fmaf(p1, M[4], p2 * M[7]));
      T[4 * r + 2] = fmaf(p0, M[2], fmaf(p1, M[5], p2 * M[8]));
      T[4 * r + 3] = fmaf(p0, mt[0], fmaf(p1, mt[1], fmaf(p2, mt[2], T[4 * r + 3])));
    }
  }
}

// Dynamic shared memory: [fast blob][per warp: transforms | cluster centres | work queue | hit word, amb word]
// kV: group variables the register prefetch covers.  Groups of at most kSmallVars variables (an arm) get their own
// instance: 16 prefetch registers fewer, and launch bounds that fit 3 CTAs x 6 warps per SM at 96 registers instead of
// 16 warps at 128 (+4.7 % on C2, measured; the 16-variable instance loses with the tighter bounds).  The larger instance
// runs as ONE CTA of up to 16 warps per SM (512 threads): the model blob is staged once instead of twice, which is what let
// the 280-sphere dual arm reach 16 warps (+4.5 % over 2 x 7; C3 +1.6 %, the 18-variable whole body +4.4 %).
// kV == 0 is the WIDE instance: groups with more than 16 variables or with a planar joint (a mobile base: the dual arm's
// whole_body group has 18).  No register prefetch -- the next tile's states are pulled into L2 one tile ahead and read at the
// start of their tile -- and the link loop knows PLANAR joints (Translation(x, y, 0) * Rz(theta), planar_joint_model.cpp:345-349).
constexpr int kSmallVars = 8;
constexpr int kWideMaxVars = 32;
constexpr int fastThreads(int kV) { return (kV > 0 && kV <= kSmallVars) ? B200SV_FAST_THREADS_SMALL : B200SV_FAST_THREADS; }
constexpr int fastMinBlocks(int kV) { return (kV > 0 && kV <= kSmallVars) ? B200SV_FAST_MINBLOCKS_SMALL : B200SV_FAST_MINBLOCKS; }

template <typename FT, bool kTG, bool kHier, int kV>
__global__ void __launch_bounds__(fastThreads(kV), fastMinBlocks(kV)) fastCheckKernel(CheckArgs<FT> a)
{
  using Elem = typename Cmp<FT>::Elem;
  extern __shared__ __align__(16) uint8_t smem[];
  // the exact pass that follows (recheckWarpKernel, launched with programmatic stream serialization) may become resident as
  // this kernel's CTAs drain: it loads its model into shared memory, then waits (griddepcontrol.wait) for this kernel's end
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const FModel m = loadFast<kHier>(a.model, a.model_bytes, smem);
  const fast::Header& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int t_words = kTG ? 0 : h.state_stride;  // shared-memory words of transforms per state
  constexpr bool kWide = kV == 0;
  uint8_t* wbase = smem + (kHier ? h.off_pairs : h.total_bytes) +
                   static_cast<size_t>(warp) *
                       fastPerWarpBytes(t_words, h.n_parked, h.n_group_vars, kHier ? h.n_linkpairs : 0, kWide);
  float* Ts = reinterpret_cast<float*>(wbase);
  TStore<kTG> ts;
  ts.base = kTG ? a.t_scratch + (static_cast<size_t>(blockIdx.x) * warps + warp) * (static_cast<size_t>(h.t_rows) * 128) : Ts;
  ts.stride = h.state_stride;
  __half* rep_x = reinterpret_cast<__half*>(Ts + 32 * t_words);  // [cluster][kRepStride]
  __half* rep_y = rep_x + h.n_parked * kRepStride;
  __half* rep_z = rep_y + h.n_parked * kRepStride;
  // hierarchical cull: only the link-level centres are in shared memory (level 0 reads them for every link pair and state);
  // the cluster-level centres, read once per level-1 candidate, live in this warp's global scratch (L1 / L2) -- they were
  // 14 of the 26 KB of shared memory a warp of the 280-sphere dual arm took, i.e. 8 instead of 14 warps per SM
  uint2* grep = kHier ? reinterpret_cast<uint2*>(a.rep_scratch + (static_cast<size_t>(blockIdx.x) * warps + warp) *
                                                                    static_cast<size_t>(globalRepBytes(h.n_clusters)))
                      : nullptr;  // [cluster][state] = { x | y << 16, z }: ONE 8-byte load per centre and candidate
  uint32_t* queue = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(rep_x) + repBytes(h.n_parked));
  // the tile's states [state][var]: the wide instance stages them in the queue region (FK phase only), the others in two
  // buffers behind it that the TMA engine fills one tile ahead
  uint8_t* stage_base = reinterpret_cast<uint8_t*>(queue) + queueBytes(h.n_group_vars, kWide);
  const int stage_one = stageOneBytes(h.n_group_vars);
  constexpr bool kDouble = kV > 0 && kV <= kSmallVars;  // = stageCount(V) == 2 for the groups this instance takes
  uint64_t* mbar = reinterpret_cast<uint64_t*>(stage_base + (kDouble ? 2 : 1) * stage_one);  // [2], non-wide only
  unsigned* words = reinterpret_cast<unsigned*>(stage_base + stageBytes(h.n_group_vars, kWide));  // [0] hit, [1] amb
  uint32_t* q0 = words + 4;        // hierarchical cull only: (state, link pair) survivors of level 0
  uint32_t* qc = q0 + kQueueCap;   //                         (state, cluster pair) candidates
  const bool do_fields = (a.flags & 3u) != 0, do_intra = (a.flags & 4u) != 0;
  const unsigned fmask = ((a.flags & 1u) ? 0x0000ffffu : 0u) | ((a.flags & 2u) ? 0xffff0000u : 0u);
  const Elem* __restrict__ field = static_cast<const Elem*>(a.field);
  const unsigned lt_mask = (1u << lane) - 1u;
  const int nz = h.nz, nynz = h.nynz;
  const unsigned idx_bias = h.idx_bias;
  float two_m = h.two_margin;  // per link: twice the link's own cell-face margin (set in the link loop)

  const unsigned long long n = a.n_states;
  const unsigned long long n_tiles = (n + 31) / 32;
  const int V = h.n_group_vars;
  // The states of a tile are ONE contiguous run of 32 * V doubles.  One 1-D bulk copy (cp.async.bulk: the TMA engine, no
  // registers, no LSU instructions) brings a whole tile into this warp's second stage buffer while the current tile is
  // worked on; it completes on an mbarrier the warp polls when it gets there.  They stream from HBM exactly once.
  // Only whole tiles whose bytes are a multiple of 16 at a 16-byte aligned address go that way; the last, ragged tile
  // (and every tile of a misaligned batch) is read with plain loads when its turn comes.
  const unsigned tile_bytes = static_cast<unsigned>(32 * V * 8);
  const bool tma_ok = !kWide && (reinterpret_cast<uintptr_t>(a.q) & 15u) == 0;
  auto tileByTma = [&](unsigned long long t) { return tma_ok && (t + 1) * 32 <= n; };
  auto fetchTile = [&](unsigned long long t, int buf) {
    const double* __restrict__ src = a.q + t * 32 * V;
    if constexpr (!kWide)
    {
      if (tileByTma(t) && lane == 0)
      {
        // the buffer was read with ordinary loads a tile ago: order those before the async proxy's writes
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        bulkLoad(stage_base + buf * stage_one, src, tile_bytes, mbar + buf);
      }
    }
    else
    {  // wide instance: one 128-byte line per lane and pass, into L2
      const unsigned long long left = n - t * 32;
      const int avail = static_cast<int>(left < 32 ? left : 32) * V;
      for (int e = lane * 16; e < avail; e += 32 * 16)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(src + e));
    }
  };
  if constexpr (!kWide)
  {
    if (lane == 0)
    {
      mbarInit(mbar, 1);
      mbarInit(mbar + 1, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  const unsigned long long tile0 = static_cast<unsigned long long>(blockIdx.x) * warps + warp;
  const unsigned long long tile_step = static_cast<unsigned long long>(gridDim.x) * warps;
  if (tile0 < n_tiles)
    fetchTile(tile0, 0);
  unsigned it = 0;  // tiles this warp has started: buffer = it & 1, barrier parity = (it >> 1) & 1
  for (unsigned long long tile = tile0; tile < n_tiles; tile += tile_step, ++it)
  {
    const int buf = kDouble ? static_cast<int>(it & 1u) : 0;
    double* stage = reinterpret_cast<double*>(kWide ? reinterpret_cast<uint8_t*>(queue) : stage_base + buf * stage_one);
    if ((kDouble || kWide) && tile + tile_step < n_tiles)
      fetchTile(tile + tile_step, buf ^ 1);  // the other buffer's last reader was the previous tile (done: __syncwarp below)
    if (!kWide && tileByTma(tile))
      mbarWait(mbar + buf, kDouble ? (it >> 1) & 1u : it & 1u);
    else
    {
      const double* __restrict__ src = a.q + tile * 32 * V;
      const unsigned long long left = n - tile * 32;
      const int avail = static_cast<int>(left < 32 ? left : 32) * V;
      for (int e = lane; e < 32 * V; e += 32)
        stage[e] = e < avail ? __ldg(src + e) : 0.0;
      __syncwarp();
    }
    const unsigned long long first = tile * 32;
    const int cnt = static_cast<int>(n - first < 32 ? n - first : 32);
    const bool active = lane < cnt;
    const double* qrow = stage + (active ? lane : 0) * V;
    unsigned hacc = 0u, aacc = 0u;  // nonzero: certainly colliding / FP32 cannot decide
    unsigned pv[kP], pt[kP];      // gathers in flight: field element, thresholds
#pragma unroll
    for (int u = 0; u < kP; ++u)
    {
      pv[u] = 0u;
      pt[u] = Cmp<FT>::kNever;
    }
    int slot_a = -1, slot_b = 0;  // parked sphere: the two candidate cells and its thresholds
    unsigned slot_t = 0u;
    auto resolveSlot = [&]() {
      const unsigned va = __ldg(field + slot_a), vb = __ldg(field + slot_b);
      const bool ha = Cmp<FT>::hit(va, slot_t) != 0u, hb = Cmp<FT>::hit(vb, slot_t) != 0u;
      hacc |= static_cast<unsigned>(ha & hb);
      aacc |= static_cast<unsigned>(ha ^ hb);
      slot_a = -1;
    };
    // N consecutive spheres of the current link (transform Tl in registers, translation lowered by margin): pose,
    // clamp, both candidate cells, gather; branch-free except for the rare "centre within margin of a cell face".
    auto chunk = [&](int k0, const float* Tl, float tmx, float tmy, float tmz, auto n_tag) {
      constexpr int N = decltype(n_tag)::value;
      unsigned idx[N], tt[N], dd[N];
      unsigned rmask = 0u;
#pragma unroll
      for (int u = 0; u < N; ++u)
      {
        const float4 hr = reinterpret_cast<const float4*>(m.hot)[k0 + u];  // uniform address: one broadcast read
        float fx = fmaf(Tl[0], hr.x, fmaf(Tl[1], hr.y, fmaf(Tl[2], hr.z, tmx)));
        float fy = fmaf(Tl[4], hr.x, fmaf(Tl[5], hr.y, fmaf(Tl[6], hr.z, tmy)));
        float fz = fmaf(Tl[8], hr.x, fmaf(Tl[9], hr.y, fmaf(Tl[10], hr.z, tmz)));
        fx = fminf(fmaxf(fx, h.lo[0]), h.hi[0]);
        fy = fminf(fmaxf(fy, h.lo[1]), h.hi[1]);
        fz = fminf(fmaxf(fz, h.lo[2]), h.hi[2]);
        const int ax = __float_as_int(__fadd_rd(fx, kMagic)), bx = __float_as_int(__fadd_rd(fx + two_m, kMagic));
        const int ay = __float_as_int(__fadd_rd(fy, kMagic)), by = __float_as_int(__fadd_rd(fy + two_m, kMagic));
        const int az = __float_as_int(__fadd_rd(fz, kMagic)), bz = __float_as_int(__fadd_rd(fz + two_m, kMagic));
        const unsigned ia = static_cast<unsigned>(ax * nynz + (ay * nz + az));  // VoxelGrid::ref (voxel_grid.hpp:425)
        const unsigned ib = static_cast<unsigned>(bx * nynz + (by * nz + bz));
        const unsigned t = Cmp<FT>::mask(__float_as_uint(hr.w), fmask);
        const bool r = ia != ib;  // within `margin` of a cell face
        idx[u] = ia + idx_bias;
        dd[u] = ib - ia;
        tt[u] = r ? Cmp<FT>::kNever : t;
        rmask |= (r && t != Cmp<FT>::kNever) ? (1u << u) : 0u;
      }
#pragma unroll
      for (int u = 0; u < N; ++u)
      {
        // the gathers of the previous chunk are consumed as late as possible: not before this chunk's indices exist
        asm volatile("" : "+r"(pv[u]) : "r"(idx[u]));
        hacc |= Cmp<FT>::hit(pv[u], pt[u]);
      }
#pragma unroll
      for (int u = 0; u < N; ++u)
      {
        // nothing to learn for padding / parked spheres, nor for a state that already collides: no L1 wavefront
        pv[u] = (tt[u] != Cmp<FT>::kNever && hacc == 0u) ? __ldg(field + idx[u]) : 0u;
        pt[u] = tt[u];
      }
      if (rmask)
      {  // rare: the outcome is certain only if both candidate cells agree; park the sphere, read both cells later
#pragma unroll
        for (int u = 0; u < N; ++u)
          if ((rmask >> u) & 1u)
          {
            const unsigned d = dd[u];
            if (d == 1u || d == static_cast<unsigned>(nz) || d == static_cast<unsigned>(nynz))
            {
              if (slot_a >= 0)
                resolveSlot();
              slot_a = static_cast<int>(idx[u]);
              slot_b = static_cast<int>(idx[u] + d);
              slot_t = Cmp<FT>::mask(m.hot[k0 + u].t, fmask);
            }
            else
              aacc = 1u;  // near an edge or corner of the cell: leave it to the exact pass
          }
      }
    };
    float T[12];
    double q_next;
    {
      const int k_first = reinterpret_cast<const int*>(m.links)[1];
      q_next = k_first >= 0 ? qrow[k_first] : 0.0;
    }
    // ---- FK, and the field checks of every link as soon as its transform exists --------------------------------------
    for (int l = 0; l < h.n_links; ++l)
    {
      const int4* Li = reinterpret_cast<const int4*>(m.links + l);
      const float4* Lf = reinterpret_cast<const float4*>(m.links + l);
      const int4 m0 = Li[0];  // type, k, parent_off, store_off
      const int4 m1 = Li[1];  // h0, h1, c0, c1
      const float4 o = Lf[3];
      const int k_next = reinterpret_cast<const int*>(Li + 11)[1];  // the table ends with a sentinel record
      const double q_cur = q_next;
      q_next = k_next >= 0 ? qrow[k_next] : 0.0;  // the NEXT link's joint value: in flight during this link's work
      // T = T_parent * [M | mt]; a root's "parent" is the identity block of the header (one code path, no merge)
      if (m0.z == fast::kParentRoot)
        load12(h.ident, T);
      else if (m0.z != fast::kParentPrev)
        ts.load(m0.z, lane, T);
      fastLinkStep<kV>(m0, Li, Lf, o, q_cur, qrow, T);
      if (m0.w >= 0)
        ts.store(m0.w, lane, T);
      if (do_intra && m1.z < m1.w)
      {  // tight-sphere centres of the link's clusters, relative to rep_base and narrowed to half: the level-1 cull of
         // the pair phase is conservative, its slack (host) covers the rounding
        const float tbx = T[3] - h.rep_base[0], tby = T[7] - h.rep_base[1], tbz = T[11] - h.rep_base[2];
        for (int c = m1.z; c < m1.w; ++c)
        {
          const float4 cc = reinterpret_cast<const float4*>(m.cl)[c];
          const float x = fmaf(T[0], cc.x, fmaf(T[1], cc.y, fmaf(T[2], cc.z, tbx)));
          const float y = fmaf(T[4], cc.x, fmaf(T[5], cc.y, fmaf(T[6], cc.z, tby)));
          const float z = fmaf(T[8], cc.x, fmaf(T[9], cc.y, fmaf(T[10], cc.z, tbz)));
          if (__float_as_int(o.w) != 0 && fmaxf(fmaxf(fabsf(x), fabsf(y)), fabsf(z)) >= h.rep_limit)
            aacc = 1u;  // beyond the travel the host assumed for prismatic joints: the exact pass decides
          if constexpr (kHier)
          {
            const int slot = __float_as_int(cc.w);  // Cluster::slot
            if (slot >= 0)
            {
              rep_x[slot * kRepStride + lane] = __float2half_rn(x);
              rep_y[slot * kRepStride + lane] = __float2half_rn(y);
              rep_z[slot * kRepStride + lane] = __float2half_rn(z);
            }
            else
              grep[c * 32 + lane] = make_uint2(static_cast<uint32_t>(__half_as_ushort(__float2half_rn(x))) |
                                                   (static_cast<uint32_t>(__half_as_ushort(__float2half_rn(y))) << 16),
                                               static_cast<uint32_t>(__half_as_ushort(__float2half_rn(z))));
          }
          else
          {
            rep_x[c * kRepStride + lane] = __float2half_rn(x);
            rep_y[c * kRepStride + lane] = __float2half_rn(y);
            rep_z[c * kRepStride + lane] = __float2half_rn(z);
          }
        }
      }
      if (!do_fields)
        continue;
      // the centre minus `margin` per axis, from a translation column lowered by margin (clamp bounds likewise)
      // (the margin is the LINK's: the worst-case FP32 error of its spheres, from the host; the clamp bounds are the
      //  global margin's, which is the largest)
      const float ml = Lf[10].w;
      two_m = ml + ml;
      const float tmx = T[3] - ml, tmy = T[7] - ml, tmz = T[11] - ml;
      int k0 = m1.x;
      for (; k0 + 4 <= m1.y; k0 += 4)
        chunk(k0, T, tmx, tmy, tmz, std::integral_constant<int, 4>());
      if (k0 < m1.y)
        chunk(k0, T, tmx, tmy, tmz, std::integral_constant<int, 2>());  // Hot ranges are padded to a multiple of 2
    }
#pragma unroll
    for (int u = 0; u < kP; ++u)
      hacc |= Cmp<FT>::hit(pv[u], pt[u]);
    if (slot_a >= 0)
      resolveSlot();
    if constexpr (!kDouble && !kWide)
    {  // single stage buffer: every lane has read its last joint value; the next tile's copy runs under the pair phase
      __syncwarp();
      if (tile + tile_step < n_tiles)
        fetchTile(tile + tile_step, 0);
    }
    bool hit = hacc != 0u, amb = aacc != 0u;
    // ---- intra-group sphere pairs ---------------------------------------------------------------------------------------
    if (do_intra && h.n_pairs > 0 && __any_sync(kFull, active && !hit))
    {
      if (lane == 0)
        words[0] = words[1] = 0u;
      __syncwarp();  // the queue consumers read other lanes' transforms
      int qn = 0;
      const bool mine = active && !hit;
      auto flush = [&]() {
        __syncwarp();
        for (int i = lane; i < qn; i += 32)
        {
          const unsigned item = queue[i];
          const int s = item & 31;
          bool ihit = false, iamb = false;
          pairItem<kTG>(m, ts, s, item >> 5, ihit, iamb);
          if (ihit)
            atomicOr(&words[0], 1u << s);
          if (iamb)
            atomicOr(&words[1], 1u << s);
        }
        __syncwarp();
        qn = 0;
      };
      const unsigned mine_mask = __ballot_sync(kFull, mine);
      if constexpr (!kHier)
      {
        // level 1, TRANSPOSED: lane = cluster pair, loop over the warp's live states.  Conservative cull on the tight
        // spheres (every sphere of a cluster lies inside its tight sphere), in half precision on the parked centres.
        for (int p0 = 0; p0 < h.n_pairs; p0 += 32)
        {
          const int p = p0 + lane;
          const bool have = p < h.n_pairs;
          const uint4 pc = reinterpret_cast<const uint4*>(m.pairs)[have ? p : 0];  // rt2 (half bits), a | b << 16, ...
          const int ca = pc.y & 0xffffu, cb = pc.y >> 16;
          const __half rt2 = __ushort_as_half(static_cast<unsigned short>(pc.x));
          const __half2 rt2_2 = __half2half2(rt2);
          const uint32_t* ax = reinterpret_cast<const uint32_t*>(rep_x + ca * kRepStride);
          const uint32_t* bx = reinterpret_cast<const uint32_t*>(rep_x + cb * kRepStride);
          const uint32_t* ay = reinterpret_cast<const uint32_t*>(rep_y + ca * kRepStride);
          const uint32_t* by = reinterpret_cast<const uint32_t*>(rep_y + cb * kRepStride);
          const uint32_t* az = reinterpret_cast<const uint32_t*>(rep_z + ca * kRepStride);
          const uint32_t* bz = reinterpret_cast<const uint32_t*>(rep_z + cb * kRepStride);
          for (int c8 = 0; c8 < 4; ++c8)
          {  // chunks of 8 states (at most 256 new items, the queue's capacity), two states per half2 lane pair
            if (((mine_mask >> (8 * c8)) & 0xffu) == 0u)
              continue;
            unsigned live_states = 0u;
#pragma unroll
            for (int j = 0; j < 4; ++j)
            {
              const int w = c8 * 4 + j;  // word = states 2w, 2w + 1
              const uint32_t wax = ax[w], wbx = bx[w], way = ay[w], wby = by[w], waz = az[w], wbz = bz[w];
              const __half2 dx = __hsub2(*reinterpret_cast<const __half2*>(&wax), *reinterpret_cast<const __half2*>(&wbx));
              const __half2 dy = __hsub2(*reinterpret_cast<const __half2*>(&way), *reinterpret_cast<const __half2*>(&wby));
              const __half2 dz = __hsub2(*reinterpret_cast<const __half2*>(&waz), *reinterpret_cast<const __half2*>(&wbz));
              const __half2 dd = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
              const unsigned lt = __hlt2_mask(dd, rt2_2);  // 0xffff per half where dd < rt2
              live_states |= ((lt & 1u) | ((lt >> 15) & 2u)) << (2 * w);
            }
            live_states &= have ? mine_mask : 0u;
            // append (state, pair) items: exclusive scan of the per-lane counts, then every lane writes its own
            const int cnt = __popc(live_states);
            int pos = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
              const int t = __shfl_up_sync(kFull, pos, d);
              if (lane >= d)
                pos += t;
            }
            const int total = __shfl_sync(kFull, pos, 31);
            if (qn + total > kQueueCap)
              flush();
            pos += qn - cnt;
            while (live_states)
            {
              const int s = __ffs(live_states) - 1;
              live_states &= live_states - 1;
              queue[pos++] = static_cast<uint32_t>(s | (p << 5));
            }
            qn += total;
          }
        }
      }
      else
      {
        // TRANSPOSED cull: lane = pair (of clusters, or of links), loop over the warp's live states, two per half2.
        // Conservative test on tight spheres (every sphere of a cluster / link lies inside its tight sphere), in half
        // precision on the parked centres.  Survivors (state, pair) are appended to `dst`; `drain` empties it when full.
        auto transposedCull = [&](const uint4* __restrict__ table, int n_entries, uint32_t* dst, int& dn, auto&& drain) {
          for (int p0 = 0; p0 < n_entries; p0 += 32)
          {
            const int p = p0 + lane;
            const bool have = p < n_entries;
            const uint4 pc = table[have ? p : 0];  // rt2 (half bits), a | b << 16, ...
            const int ca = pc.y & 0xffffu, cb = pc.y >> 16;
            const __half2 rt2_2 = __half2half2(__ushort_as_half(static_cast<unsigned short>(pc.x)));
            const uint32_t* ax = reinterpret_cast<const uint32_t*>(rep_x + ca * kRepStride);
            const uint32_t* bx = reinterpret_cast<const uint32_t*>(rep_x + cb * kRepStride);
            const uint32_t* ay = reinterpret_cast<const uint32_t*>(rep_y + ca * kRepStride);
            const uint32_t* by = reinterpret_cast<const uint32_t*>(rep_y + cb * kRepStride);
            const uint32_t* az = reinterpret_cast<const uint32_t*>(rep_z + ca * kRepStride);
            const uint32_t* bz = reinterpret_cast<const uint32_t*>(rep_z + cb * kRepStride);
            for (int c8 = 0; c8 < 4; ++c8)
            {  // chunks of 8 states (at most 256 new entries, the queue's capacity), two states per half2 lane pair
              if (((mine_mask >> (8 * c8)) & 0xffu) == 0u)
                continue;
              unsigned live_states = 0u;
#pragma unroll
              for (int j = 0; j < 4; ++j)
              {
                const int w = c8 * 4 + j;  // word = states 2w, 2w + 1
                const uint32_t wax = ax[w], wbx = bx[w], way = ay[w], wby = by[w], waz = az[w], wbz = bz[w];
                const __half2 dx = __hsub2(*reinterpret_cast<const __half2*>(&wax), *reinterpret_cast<const __half2*>(&wbx));
                const __half2 dy = __hsub2(*reinterpret_cast<const __half2*>(&way), *reinterpret_cast<const __half2*>(&wby));
                const __half2 dz = __hsub2(*reinterpret_cast<const __half2*>(&waz), *reinterpret_cast<const __half2*>(&wbz));
                const __half2 dd = __hfma2(dz, dz, __hfma2(dy, dy, __hmul2(dx, dx)));
                const unsigned lt = __hlt2_mask(dd, rt2_2);  // 0xffff per half where dd < rt2
                live_states |= ((lt & 1u) | ((lt >> 15) & 2u)) << (2 * w);
              }
              live_states &= have ? mine_mask : 0u;
              // append (state, pair): exclusive scan of the per-lane counts, then every lane writes its own
              const int cnt = __popc(live_states);
              int pos = cnt;
#pragma unroll
              for (int d = 1; d < 32; d <<= 1)
              {
                const int t = __shfl_up_sync(kFull, pos, d);
                if (lane >= d)
                  pos += t;
              }
              const int total = __shfl_sync(kFull, pos, 31);
              if (dn + total > kQueueCap)
                drain();
              pos += dn - cnt;
              while (live_states)
              {
                const int s = __ffs(live_states) - 1;
                live_states &= live_states - 1;
                dst[pos++] = static_cast<uint32_t>(s | (p << 5));
              }
              dn += total;
            }
          }
        };
        // Many cluster pairs (a robot at PR2 scale has thousands): level 0 culls LINK pairs the same way, its survivors
        // (state, link pair) are expanded to (state, cluster pair) candidates, and each candidate is tested by one lane.
        int q0n = 0, qcn = 0;
        auto drainC = [&]() {
          __syncwarp();
          for (int i0 = 0; i0 < qcn; i0 += 32)
          {
            const int i = i0 + lane;
            const bool have = i < qcn;
            const unsigned cand = have ? qc[i] : 0u;
            const int s = cand & 31;
            const uint4 pc = reinterpret_cast<const uint4*>(m.pairs)[cand >> 5];
            const int ca = pc.y & 0xffffu, cb = pc.y >> 16;
            const uint2 ga = grep[ca * 32 + s], gb = grep[cb * 32 + s];
            const __half2 dxy = __hsub2(*reinterpret_cast<const __half2*>(&ga.x), *reinterpret_cast<const __half2*>(&gb.x));
            const __half dx = __low2half(dxy), dy = __high2half(dxy);
            const __half dz = __hsub(__ushort_as_half(static_cast<unsigned short>(ga.y)),
                                     __ushort_as_half(static_cast<unsigned short>(gb.y)));
            const __half dd = __hfma(dz, dz, __hfma(dy, dy, __hmul(dx, dx)));
            const bool live = have && __hlt(dd, __ushort_as_half(static_cast<unsigned short>(pc.x)));
            const unsigned mk = __ballot_sync(kFull, live);
            if (qn + __popc(mk) > kQueueCap)
              flush();
            if (live)
              queue[qn + __popc(mk & lt_mask)] = cand;
            qn += __popc(mk);
          }
          __syncwarp();
          qcn = 0;
        };
        auto drain0 = [&]() {
          __syncwarp();
          for (int i0 = 0; i0 < q0n; i0 += 32)
          {
            const int i = i0 + lane;
            const bool have = i < q0n;
            const unsigned item = have ? q0[i] : 0u;
            const uint4 lp = reinterpret_cast<const uint4*>(m.linkpairs)[item >> 5];  // rt2, la | lb << 16, cp0, ncp
            const int ncp = have ? static_cast<int>(lp.w) : 0;
            int pos = ncp;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
            {
              const int t = __shfl_up_sync(kFull, pos, d);
              if (lane >= d)
                pos += t;
            }
            const int total = __shfl_sync(kFull, pos, 31);  // <= 32 * 16: the host splits link pairs at 16 cluster pairs
            if (qcn + total > kQcCap)
              drainC();
            pos += qcn - ncp;
            for (int k = 0; k < ncp; ++k)
              qc[pos + k] = (item & 31u) | ((lp.z + k) << 5);
            qcn += total;
          }
          __syncwarp();
          q0n = 0;
        };
        transposedCull(reinterpret_cast<const uint4*>(m.linkpairs), h.n_linkpairs, q0, q0n, drain0);
        drain0();
        drainC();
      }
      if (qn > 0)
        flush();
      hit |= ((words[0] >> lane) & 1u) != 0;
      amb |= ((words[1] >> lane) & 1u) != 0;
    }
    __syncwarp();  // transforms and words are rewritten by the next tile
    const bool valid = active && !hit && !amb;
    const bool flagged = active && amb && !hit;  // FP32 could not decide: the exact pass will
    const unsigned valid_word = __ballot_sync(kFull, valid);
    const unsigned flag_word = __ballot_sync(kFull, flagged);
    if (lane == 0)
      a.bitmap[tile] = valid_word;
    if (flag_word)
    {
      unsigned base = 0;
      if (lane == 0)
        base = atomicAdd(a.flag_count, static_cast<unsigned>(__popc(flag_word)));
      base = __shfl_sync(kFull, base, 0);
      if (flagged)
      {
        const unsigned rank = __popc(flag_word & lt_mask);
        if (base + rank < a.list_cap)
          a.flag_list[base + rank] = static_cast<uint32_t>(a.state_base + first + lane);
      }
    }
  }
}

// The FP32 error-budget probe: the fast kernel's FK (fastLinkStep, the same function) and sphere posing for one state per
// thread, no checks; writes every Hot record's posed centre in cells (margin removed).  Transforms live in a global scratch
// laid out like the shared-memory form ([state][state_stride words]).
template <int kV>
__global__ void fastCentresKernel(const uint8_t* __restrict__ model, int model_bytes, const double* __restrict__ q,
                                  unsigned long long n, float* scratch, float* __restrict__ out, int n_hot)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const FModel m = loadFast<false>(model, model_bytes, smem);
  const fast::Header& h = *m.h;
  const unsigned long long state = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (state >= n)
    return;
  TStore<false> ts;
  ts.base = scratch + state * static_cast<unsigned long long>(h.state_stride);
  ts.stride = 0;
  const double* qrow = q + state * h.n_group_vars;
  float T[12];
  for (int l = 0; l < h.n_links; ++l)
  {
    const int4* Li = reinterpret_cast<const int4*>(m.links + l);
    const float4* Lf = reinterpret_cast<const float4*>(m.links + l);
    const int4 m0 = Li[0];
    const int4 m1 = Li[1];
    const float4 o = Lf[3];
    const double q_cur = m0.y >= 0 ? qrow[m0.y] : 0.0;
    if (m0.z == fast::kParentRoot)
      load12(h.ident, T);
    else if (m0.z != fast::kParentPrev)
      ts.load(m0.z, 0, T);
    fastLinkStep<kV>(m0, Li, Lf, o, q_cur, qrow, T);
    if (m0.w >= 0)
      ts.store(m0.w, 0, T);
    const float ml = Lf[10].w;
    const float tmx = T[3] - ml, tmy = T[7] - ml, tmz = T[11] - ml;
    for (int k = m1.x; k < m1.y; ++k)
    {
      const float4 hr = reinterpret_cast<const float4*>(m.hot)[k];
      float* dst = out + (state * n_hot + k) * 3;
      dst[0] = fmaf(T[0], hr.x, fmaf(T[1], hr.y, fmaf(T[2], hr.z, tmx))) + ml;
      dst[1] = fmaf(T[4], hr.x, fmaf(T[5], hr.y, fmaf(T[6], hr.z, tmy))) + ml;
      dst[2] = fmaf(T[8], hr.x, fmaf(T[9], hr.y, fmaf(T[10], hr.z, tmz))) + ml;
    }
  }
}

struct Grant
{
  size_t granted[16] = { 0 };
};
template <typename K>
cudaError_t ensure(K kernel, size_t smem_bytes, Grant& g)
{
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (smem_bytes <= g.granted[dev])
    return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e == cudaSuccess)
    g.granted[dev] = smem_bytes;
  return e;
}
}  // namespace

int fastCheckPerWarpBytes(int state_stride, int n_clusters, int n_group_vars, int n_linkpairs, bool wide)
{
  return fastPerWarpBytes(state_stride, n_clusters, n_group_vars, n_linkpairs, wide);
}
int fastCheckRepBytes(int n_rows) { return globalRepBytes(n_rows); }
int fastCheckMaxVars() { return kWideMaxVars; }
int fastCheckRegisterPrefetchVars() { return kMaxVars; }  // beyond this (or with a planar joint) the wide instance runs
int fastCheckMaxWarps(int n_group_vars) { return fastThreads(n_group_vars <= kSmallVars ? kSmallVars : kMaxVars) / 32; }
int fastCheckUnroll() { return kU; }

int fastCheckKernelRegs(int field_bytes, int n_group_vars)
{
  cudaFuncAttributes fa{};
  int regs = 0;
  auto take = [&](auto kernel) {
    const cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
    regs = std::max(regs, e == cudaSuccess ? fa.numRegs : 128);
  };
  auto all = [&](auto ft, auto kv) {
    using FT = decltype(ft);
    constexpr int kV = decltype(kv)::value;
    take(fastCheckKernel<FT, false, false, kV>);
    take(fastCheckKernel<FT, true, false, kV>);
    take(fastCheckKernel<FT, false, true, kV>);
    take(fastCheckKernel<FT, true, true, kV>);
  };
  const bool small = n_group_vars <= kSmallVars;
  if (field_bytes == 1)
    small ? all(uint8_t(), std::integral_constant<int, kSmallVars>()) : all(uint8_t(), std::integral_constant<int, kMaxVars>());
  else
    small ? all(uint16_t(), std::integral_constant<int, kSmallVars>()) : all(uint16_t(), std::integral_constant<int, kMaxVars>());
  if (!small)
  {  // the wide instance shares the launch configuration of the 16-variable one
    if (field_bytes == 1)
      all(uint8_t(), std::integral_constant<int, 0>());
    else
      all(uint16_t(), std::integral_constant<int, 0>());
  }
  return regs;
}

cudaError_t launchFastCentres(const uint8_t* model, int model_bytes, const double* q, unsigned long long n, int n_group_vars,
                              bool wide, float* scratch, float* out, int n_hot, cudaStream_t stream)
{
  static Grant g[3];
  const size_t smem = (static_cast<size_t>(model_bytes) + 15) & ~static_cast<size_t>(15);
  const int block = 128;
  const unsigned grid = static_cast<unsigned>((n + block - 1) / block);
  auto go = [&](auto kernel, Grant& gr) {
    const cudaError_t e = ensure(kernel, smem, gr);
    if (e != cudaSuccess)
      return e;
    kernel<<<grid, block, smem, stream>>>(model, model_bytes, q, n, scratch, out, n_hot);
    return cudaGetLastError();
  };
  if (wide)
    return go(fastCentresKernel<0>, g[2]);
  return n_group_vars <= kSmallVars ? go(fastCentresKernel<kSmallVars>, g[0]) : go(fastCentresKernel<kMaxVars>, g[1]);
}

template <typename FT>
cudaError_t launchFastCheck(const CheckArgs<FT>& a, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static Grant g[12];
  auto go = [&](auto kernel, Grant& gr) {
    const cudaError_t e = ensure(kernel, smem_bytes, gr);
    if (e != cudaSuccess)
      return e;
    kernel<<<grid, block, smem_bytes, stream>>>(a);
    return cudaGetLastError();
  };
  auto pick = [&](auto kv, Grant* gr) {
    constexpr int kV = decltype(kv)::value;
    if (a.hier)
      return a.t_scratch ? go(fastCheckKernel<FT, true, true, kV>, gr[3]) : go(fastCheckKernel<FT, false, true, kV>, gr[2]);
    return a.t_scratch ? go(fastCheckKernel<FT, true, false, kV>, gr[1]) : go(fastCheckKernel<FT, false, false, kV>, gr[0]);
  };
  if (a.wide)
    return pick(std::integral_constant<int, 0>(), g + 8);
  return a.n_group_vars <= kSmallVars ? pick(std::integral_constant<int, kSmallVars>(), g) :
                                        pick(std::integral_constant<int, kMaxVars>(), g + 4);
}
template cudaError_t launchFastCheck<uint8_t>(const CheckArgs<uint8_t>&, int, int, size_t, cudaStream_t);
template cudaError_t launchFastCheck<uint16_t>(const CheckArgs<uint16_t>&, int, int, size_t, cudaStream_t);
}  // namespace b200sv
