// K1 + K2: batched forward kinematics and sphere-decomposition collision checks, fused.
//
// Replaces, per state (reference paths relative to moveit_core/):
//   RobotState::setJointGroupPositions / updateMimicJoints / updateLinkTransformsInternal
//                                                  robot_state/src/robot_state.cpp:571-584, 210-220, 793-848
//   {Revolute,Prismatic,Planar,Floating,Fixed}JointModel::computeTransform
//                                                  robot_model/src/{revolute,prismatic,planar,floating,fixed}_joint_model.cpp
//   PosedBodySphereDecomposition::updatePose       collision_distance_field/src/collision_distance_field_types.cpp:388-395
//   getCollisionSphereCollision + DistanceField::getDistanceGradient (centre voxel only: the gradient never
//   influences the boolean)                        types.cpp:211-236, distance_field/src/distance_field.cpp:73-97
//   getEnvironmentCollisions / getSelfCollisions / getIntraGroupCollisions + doBoundingSpheresIntersect
//                                                  collision_env_distance_field.cpp:1561-1643, 274-353, 430-642; types.cpp:408-418
//
// Mapping (B200, sm_100a): persistent CTAs; every warp owns tiles of 32 consecutive states and works THREAD-PER-STATE,
// so each instruction serves 32 states (the first, warp-per-state version spent ~900 warp instructions per state on
// per-state scalar overhead; see profiles/r1_check_v2_tile8.txt).
//   phase 1  FK: lane s walks the device links in index order (parents first); link transforms go to shared memory as
//            3x4 row-major rows of 16 bytes.  State stride = 4 (mod 32) words: conflict-free 128-bit accesses.
//   phase 2  field checks: uniform loop over the spheres in chunks of 4; each lane poses the sphere for ITS state,
//            converts the centre to a voxel and gathers one byte from the world field and one from the self field
//            (L2-resident compact d^2): 8 independent gathers in flight per lane; compare against the sphere's integer
//            threshold.  Lanes whose state already collides stop gathering.
//   phase 3  intra-group pairs: uniform loop over the enabled link pairs; each lane applies the reference's
//            bounding-sphere cull plus a conservative tight-sphere cull for its state; surviving (state, pair) items
//            are compacted into a per-warp queue and processed one item per lane: spheres of either link against
//            the other link's tight sphere, survivors sphere against sphere.
// Precision: the fast pass is FP32 (template R = float).  A sphere whose voxel coordinate is within `margin` of a
// cell face, or a pair whose distance is within eps of its threshold, is evaluated for BOTH outcomes; if the
// state's boolean could depend on that, the state index is appended to a recheck list and the exact pass
// (R = double, same code) recomputes it.  The FP64 pass restates the reference's double arithmetic
// operation by operation.  The conservative culls only ever skip pairs that cannot collide, so they change no result.
#include <cuda_runtime.h>

#include <cstdlib>

#include <cstdint>

#include "device_model.hpp"
#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

// One element of the interleaved compact field holds both fields of a voxel: world in the low half, self in the high.
template <typename FT> struct Combined;
template <> struct Combined<uint8_t> { using type = uint16_t; };
template <> struct Combined<uint16_t> { using type = uint32_t; };
template <typename FT>
__device__ __forceinline__ void gatherBoth(const typename Combined<FT>::type* __restrict__ field, int idx, int& dw, int& ds)
{
  const unsigned v = static_cast<unsigned>(__ldg(field + idx));
  dw = static_cast<int>(v & ((1u << (8 * sizeof(FT))) - 1u));
  ds = static_cast<int>(v >> (8 * sizeof(FT)));
}

template <typename R>
struct Smem
{
  const ModelHeader<R>* h;
  const LinkRec<R>* links;
  const VarRec* vars;
  const SphereRec<R>* spheres;
  const BsRec<R>* bs;
  const PairRec* pairs;
  const GroupRec* groups;
  const AltRec* alts;
};

__device__ __forceinline__ void sincosR(float x, float* s, float* c) { sincosf(x, s, c); }
__device__ __forceinline__ void sincosR(double x, double* s, double* c) { sincos(x, s, c); }
__device__ __forceinline__ float sqrtR(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrtR(double x) { return sqrt(x); }
__device__ __forceinline__ float floorR(float x) { return floorf(x); }
__device__ __forceinline__ double floorR(double x) { return floor(x); }

// 12 consecutive, 16-byte aligned values <-> registers with 128-bit shared-memory accesses
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
__device__ __forceinline__ void load12(const double* p, double* o)
{
  const double2* v = reinterpret_cast<const double2*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i)
  {
    const double2 t = v[i];
    o[2 * i + 0] = t.x; o[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store12(float* p, const float* o)
{
  float4* v = reinterpret_cast<float4*>(p);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    v[i] = make_float4(o[4 * i + 0], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
}
__device__ __forceinline__ void store12(double* p, const double* o)
{
  double2* v = reinterpret_cast<double2*>(p);
#pragma unroll
  for (int i = 0; i < 6; ++i)
    v[i] = make_double2(o[2 * i + 0], o[2 * i + 1]);
}


// A(3x4) * B(3x4) with implicit (0,0,0,1) rows; accumulation order k = 0..3 as Eigen's fixed-size product.
template <typename R>
__device__ __forceinline__ void mulAffine(const R* a, const R* b, R* o)
{
#pragma unroll
  for (int r = 0; r < 3; ++r)
  {
#pragma unroll
    for (int c = 0; c < 4; ++c)
    {
      R s = a[r * 4 + 0] * b[0 * 4 + c];
      s += a[r * 4 + 1] * b[1 * 4 + c];
      s += a[r * 4 + 2] * b[2 * 4 + c];
      if (c == 3)
        s += a[r * 4 + 3];
      o[r * 4 + c] = s;
    }
  }
}

// rot(3x3 of A) * Rj, translation of A kept: equals A * [Rj 0; 0 1] exactly (the dropped terms are x*0 and t*1).
template <typename R>
__device__ __forceinline__ void mulRot(const R* a, const R* rj, R* o)
{
#pragma unroll
  for (int r = 0; r < 3; ++r)
  {
#pragma unroll
    for (int c = 0; c < 3; ++c)
    {
      R s = a[r * 4 + 0] * rj[0 * 3 + c];
      s += a[r * 4 + 1] * rj[1 * 3 + c];
      s += a[r * 4 + 2] * rj[2 * 3 + c];
      o[r * 4 + c] = s;
    }
    o[r * 4 + 3] = a[r * 4 + 3];
  }
}

template <typename R>
__device__ __forceinline__ R varValue(const VarRec& v, const double* __restrict__ qrow)
{
  // setJointGroupPositions / updateMimicJoints: f * q + o in double, as the reference does, then narrowed
  double x = v.b;
  if (v.k >= 0)
    x = v.a * qrow[v.k] + v.b;
  return static_cast<R>(x);
}

// Phase 1: FK of one state into shared memory.  T: this state's transform block (link stride kLinkStrideWords).
// sc (optional): sin / cos of every revolute / planar link's angle, [2 l] and [2 l + 1], computed beforehand (the latency
// kernel lets 32 lanes do that in parallel: the libm-grade FP64 sincos is most of a link's serial time).
template <typename R>
__device__ __forceinline__ void fkState(const Smem<R>& m, const double* __restrict__ qrow, R* T, const R* sc = nullptr)
{
  const int n_links = m.h->n_links;
  for (int l = 0; l < n_links; ++l)
  {
    const LinkRec<R>& L = m.links[l];
    R A[12], Lo[12];
    load12(L.o, Lo);
    if (L.parent >= 0)
    {
      R P[12];
      load12(T + L.parent * kLinkStrideWords, P);
      if (L.has_origin)
        mulAffine<R>(P, Lo, A);
      else
      {
#pragma unroll
        for (int i = 0; i < 12; ++i)
          A[i] = P[i];
      }
    }
    else
    {
#pragma unroll
      for (int i = 0; i < 12; ++i)
        A[i] = Lo[i];
    }
    R O[12];
    switch (L.type)
    {
      case REVOLUTE:
      {  // revolute_joint_model.cpp:250-287
        R s, c;
        if (sc)
        {
          s = sc[2 * l];
          c = sc[2 * l + 1];
        }
        else
          sincosR(varValue<R>(m.vars[L.var], qrow), &s, &c);
        const R x = L.ax[0], y = L.ax[1], z = L.ax[2];
        const R t = R(1) - c;
        const R txy = t * (x * y), txz = t * (x * z), tyz = t * (y * z);
        const R zs = z * s, ys = y * s, xs = x * s;
        R rj[9];
        rj[0] = t * (x * x) + c; rj[1] = txy - zs; rj[2] = txz + ys;
        rj[3] = txy + zs; rj[4] = t * (y * y) + c; rj[5] = tyz - xs;
        rj[6] = txz - ys; rj[7] = tyz + xs; rj[8] = t * (z * z) + c;
        mulRot<R>(A, rj, O);
        break;
      }
      case PRISMATIC:
      {  // prismatic_joint_model.cpp:124-146: identity rotation, translation axis * q
        const R q = varValue<R>(m.vars[L.var], qrow);
        const R tx = L.ax[0] * q, ty = L.ax[1] * q, tz = L.ax[2] * q;
#pragma unroll
        for (int r = 0; r < 3; ++r)
        {
          O[r * 4 + 0] = A[r * 4 + 0];
          O[r * 4 + 1] = A[r * 4 + 1];
          O[r * 4 + 2] = A[r * 4 + 2];
          R s2 = A[r * 4 + 0] * tx;
          s2 += A[r * 4 + 1] * ty;
          s2 += A[r * 4 + 2] * tz;
          O[r * 4 + 3] = s2 + A[r * 4 + 3];
        }
        break;
      }
      case PLANAR:
      {  // planar_joint_model.cpp:345-349: Translation(x, y, 0) * AngleAxis(theta, UnitZ)
        const R px = varValue<R>(m.vars[L.var], qrow), py = varValue<R>(m.vars[L.var + 1], qrow);
        R s, c;
        if (sc)
        {
          s = sc[2 * l];
          c = sc[2 * l + 1];
        }
        else
          sincosR(varValue<R>(m.vars[L.var + 2], qrow), &s, &c);
        R J[12] = { c, -s, R(0), px, s, c, R(0), py, R(0), R(0), (R(1) - c) + c, R(0) };
        mulAffine<R>(A, J, O);
        break;
      }
      case FLOATING:
      {  // floating_joint_model.cpp:231-236: Translation * Quaternion(w, x, y, z).normalized()
        const R px = varValue<R>(m.vars[L.var], qrow), py = varValue<R>(m.vars[L.var + 1], qrow),
                pz = varValue<R>(m.vars[L.var + 2], qrow);
        R x = varValue<R>(m.vars[L.var + 3], qrow), y = varValue<R>(m.vars[L.var + 4], qrow),
          z = varValue<R>(m.vars[L.var + 5], qrow), w = varValue<R>(m.vars[L.var + 6], qrow);
        const R n = sqrtR(x * x + y * y + z * z + w * w);
        x /= n; y /= n; z /= n; w /= n;
        const R tx = R(2) * x, ty = R(2) * y, tz = R(2) * z;
        const R twx = tx * w, twy = ty * w, twz = tz * w;
        const R txx = tx * x, txy = ty * x, txz = tz * x;
        const R tyy = ty * y, tyz = tz * y, tzz = tz * z;
        R J[12] = { R(1) - (tyy + tzz), txy - twz, txz + twy, px,
                    txy + twz, R(1) - (txx + tzz), tyz - twx, py,
                    txz - twy, tyz + twx, R(1) - (txx + tyy), pz };
        mulAffine<R>(A, J, O);
        break;
      }
      default:
#pragma unroll
        for (int i = 0; i < 12; ++i)
          O[i] = A[i];
        break;
    }
    store12(T + l * kLinkStrideWords, O);
  }
}

// Isometry3d * Vector3d from a 3x4 row-major transform held in registers: linear * p + translation
template <typename R>
__device__ __forceinline__ void posePoint(const R* T, R x, R y, R z, R& ox, R& oy, R& oz)
{
  R s = T[0] * x;
  s += T[1] * y;
  s += T[2] * z;
  ox = s + T[3];
  s = T[4] * x;
  s += T[5] * y;
  s += T[6] * z;
  oy = s + T[7];
  s = T[8] * x;
  s += T[9] * y;
  s += T[10] * z;
  oz = s + T[11];
}

template <typename R>
__device__ __forceinline__ Smem<R> loadModel(const uint8_t* __restrict__ g_model, int model_bytes, uint8_t* smem)
{
  // cooperative 16-byte copy of the model blob into shared memory
  const int4* src = reinterpret_cast<const int4*>(g_model);
  int4* dst = reinterpret_cast<int4*>(smem);
  for (int i = threadIdx.x; i < model_bytes / 16; i += blockDim.x)
    dst[i] = src[i];
  __syncthreads();
  Smem<R> m;
  m.h = reinterpret_cast<const ModelHeader<R>*>(smem);
  m.links = reinterpret_cast<const LinkRec<R>*>(smem + m.h->off_links);
  m.vars = reinterpret_cast<const VarRec*>(smem + m.h->off_vars);
  m.spheres = reinterpret_cast<const SphereRec<R>*>(smem + m.h->off_spheres);
  m.bs = reinterpret_cast<const BsRec<R>*>(smem + m.h->off_bs);
  m.pairs = reinterpret_cast<const PairRec*>(smem + m.h->off_pairs);
  m.groups = reinterpret_cast<const GroupRec*>(smem + m.h->off_groups);
  m.alts = reinterpret_cast<const AltRec*>(smem + m.h->off_alts);
  return m;
}

// FP32 only, rare: the centre of sphere k is within `margin` of a cell face.  Evaluate every candidate cell.
template <typename R, typename FT>
__device__ __noinline__ void riskySphere(const Smem<R>& m, const R* T, int k, bool do_world, bool do_self,
                                         const typename Combined<FT>::type* __restrict__ field, bool& hit, bool& amb)
{
  const ModelHeader<R>& h = *m.h;
  const SphereRec<R> sp = m.spheres[k];
  R Tl[12];
  load12(T + sp.slot * kLinkStrideWords, Tl);
  R cx, cy, cz;
  posePoint<R>(Tl, sp.x, sp.y, sp.z, cx, cy, cz);
  const R fx = (cx - h.om[0]) * h.oo_res, fy = (cy - h.om[1]) * h.oo_res, fz = (cz - h.om[2]) * h.oo_res;
  const int gx = static_cast<int>(floorR(fx - h.margin)), hx = static_cast<int>(floorR(fx + h.margin));
  const int gy = static_cast<int>(floorR(fy - h.margin)), hy = static_cast<int>(floorR(fy + h.margin));
  const int gz = static_cast<int>(floorR(fz - h.margin)), hz = static_cast<int>(floorR(fz + h.margin));
  const int thr_w = do_world ? sp.thr : 0, thr_s = do_self ? sp.thr_self : 0;
  bool any = false, all = true;
  for (int ix = gx; ix <= hx; ++ix)
    for (int iy = gy; iy <= hy; ++iy)
      for (int iz = gz; iz <= hz; ++iz)
      {
        bool c = false;
        if (ix >= 1 && iy >= 1 && iz >= 1 && ix < h.nx - 1 && iy < h.ny - 1 && iz < h.nz - 1)
        {
          int dw, ds;
          gatherBoth<FT>(field, (ix * h.ny + iy) * h.nz + iz, dw, ds);
          c = (dw < thr_w) | (ds < thr_s);  // a threshold of 0 disables that field
        }
        any |= c;
        all &= c;
      }
  hit |= all;
  amb |= (any && !all);
}

// Phase 2 for kU consecutive spheres of this lane's state: pose, voxel, gather (all loads first), compare.
template <typename R, typename FT, int kU>
__device__ __forceinline__ void sphereChunk(const Smem<R>& m, const R* T, int k0, bool do_world, bool do_self,
                                            const typename Combined<FT>::type* __restrict__ field, bool live_lane,
                                            bool& hit, bool& amb)
{
  constexpr bool kExact = sizeof(R) == 8;
  const ModelHeader<R>& h = *m.h;
  int idx[kU], tw[kU], ts[kU];
  unsigned risky = 0;
#pragma unroll
  for (int u = 0; u < kU; ++u)
  {
    const int k = k0 + u;  // the table is padded to a multiple of kU with thr = 0 entries
    {
      const SphereRec<R> sp = m.spheres[k];  // uniform address: one broadcast read for the warp
      tw[u] = do_world ? sp.thr : 0;
      ts[u] = do_self ? sp.thr_self : 0;
      R Tl[12];
      load12(T + sp.slot * kLinkStrideWords, Tl);
      R cx, cy, cz;
      posePoint<R>(Tl, sp.x, sp.y, sp.z, cx, cy, cz);
      // VoxelGrid::getCellFromLocation (voxel_grid.hpp:522): floor((loc - origin_minus) * oo_resolution)
      const R fx = (cx - h.om[0]) * h.oo_res, fy = (cy - h.om[1]) * h.oo_res, fz = (cz - h.om[2]) * h.oo_res;
      int gx, gy, gz;
      bool r = false;
      if constexpr (kExact)
      {
        gx = static_cast<int>(floorR(fx));
        gy = static_cast<int>(floorR(fy));
        gz = static_cast<int>(floorR(fz));
      }
      else
      {
        const R ffx = floorR(fx), ffy = floorR(fy), ffz = floorR(fz);
        gx = static_cast<int>(ffx);
        gy = static_cast<int>(ffy);
        gz = static_cast<int>(ffz);
        // within `margin` of a cell face on any axis <=> floor(f - margin) != floor(f + margin)
        const R lo = fminf(fminf(fx - ffx, fy - ffy), fz - ffz), hi = fmaxf(fmaxf(fx - ffx, fy - ffy), fz - ffz);
        r = (lo < h.margin) | (hi >= R(1) - h.margin);
      }
      // DistanceField::getDistanceGradient bounds (distance_field.cpp:82): a 1-cell margin is "out of bounds", which
      // reads as max_distance and therefore never collides (types.cpp:229).
      const bool inb = (static_cast<unsigned>(gx - 1) < static_cast<unsigned>(h.nx - 2)) &
                       (static_cast<unsigned>(gy - 1) < static_cast<unsigned>(h.ny - 2)) &
                       (static_cast<unsigned>(gz - 1) < static_cast<unsigned>(h.nz - 2));
      const bool want = ((tw[u] | ts[u]) != 0) & live_lane;
      if (want & r)
        risky |= 1u << u;
      // VoxelGrid::ref (voxel_grid.hpp:425); < 2^31 voxels (host checks).  Lanes with nothing to read gather voxel 0
      // with both thresholds 0: no branch around the load.
      const bool ok = want & inb & !r;
      idx[u] = ok ? (gx * h.ny + gy) * h.nz + gz : 0;
      tw[u] = ok ? tw[u] : 0;
      ts[u] = ok ? ts[u] : 0;
    }
  }
  int dw[kU], ds[kU];
#pragma unroll
  for (int u = 0; u < kU; ++u)
  {
    gatherBoth<FT>(field, idx[u], dw[u], ds[u]);  // ONE gather: both fields of the voxel sit side by side
  }
#pragma unroll
  for (int u = 0; u < kU; ++u)
    hit |= (dw[u] < tw[u]) | (ds[u] < ts[u]);
  if (!kExact && risky)
  {
#pragma unroll
    for (int u = 0; u < kU; ++u)
      if ((risky >> u) & 1u)
        riskySphere<R, FT>(m, T, k0 + u, do_world, do_self, field, hit, amb);
  }
}

// Phase 3, level 1 for this lane's state and one cluster pair (a, b): our conservative tight-sphere cull first, and
// only when that says "near" the reference's own link-level cull with its quirk.  (ax, ay, az): posed tight centre
// of cluster a, Ta: its link transform (both shared by the whole group of pairs).
template <typename R>
__device__ __forceinline__ void pairCull(const Smem<R>& m, const R* T, const BsRec<R>& A, const R* Ta, R ax, R ay, R az,
                                         const BsRec<R>& B, const PairRec& pr, bool& live, bool& certain)
{
  constexpr bool kExact = sizeof(R) == 8;
  const ModelHeader<R>& h = *m.h;
  R Tb[12];
  load12(T + B.slot * kLinkStrideWords, Tb);
  R bx, by, bz;
  // conservative cull: every sphere of a cluster lies inside its tight sphere, so separated tight spheres (with
  // slack for rounding) mean no sphere pair of the two clusters can touch
  posePoint<R>(Tb, B.tx, B.ty, B.tz, bx, by, bz);
  const R fx = ax - bx, fy = ay - by, fz = az - bz;
  const R rt = A.tr + B.tr + h.tight_eps;
  live = certain = false;
  if ((fx * fx + fy * fy + fz * fz) < rt * rt)
  {
    if (pr.quirk_free)
    {  // the host proved: whenever these two tight spheres are this close, the reference's cull below passes
      live = certain = true;
      return;
    }
    // doBoundingSpheresIntersect (types.cpp:408-418) on the LINKS' bounding spheres, quirk kept: SQUARED centre
    // distance against the plain radius sum
    auto boundsMeet = [&](const R* Tp, const BsRec<R>& P, const R* Tq, const BsRec<R>& Q) {
      R px, py, pz, qx, qy, qz;
      posePoint<R>(Tp, P.x, P.y, P.z, px, py, pz);
      posePoint<R>(Tq, Q.x, Q.y, Q.z, qx, qy, qz);
      const R ex = px - qx, ey = py - qy, ez = pz - qz;
      const R d2 = ex * ex + ey * ey + ez * ez;
      const R rs = P.r + Q.r;
      if (kExact)
      {
        live |= d2 < rs;
        certain |= d2 < rs;
      }
      else
      {
        live |= d2 < rs + h.bs_eps;
        certain |= d2 < rs - h.bs_eps;
      }
    };
    boundsMeet(Ta, A, Tb, B);
    // a body of several shapes: any (shape, shape) combination passing lets the whole body through
    // (collision_env_distance_field.cpp:455-507); rare, and only reached by tight spheres that are already near
    for (int alt = pr.alt0; alt >= 0 && !certain; alt = m.alts[alt].next)
    {
      const BsRec<R>&P = m.bs[m.alts[alt].a], &Q = m.bs[m.alts[alt].b];
      R Tp[12], Tq[12];
      load12(T + P.slot * kLinkStrideWords, Tp);
      load12(T + Q.slot * kLinkStrideWords, Tq);
      boundsMeet(Tp, P, Tq, Q);
    }
  }
}

// Phase 3, levels 2 and 3 for one (state, cluster pair) item: spheres of either cluster against the other cluster's
// tight sphere, then the survivors sphere against sphere (getIntraGroupCollisions,
// collision_env_distance_field.cpp:571-638).  Clusters hold at most 64 spheres (host guarantees it).
template <typename R>
__device__ __forceinline__ void pairItem(const Smem<R>& m, const R* T, const PairRec& pr, bool cert, bool& hit, bool& amb)
{
  constexpr bool kExact = sizeof(R) == 8;
  const ModelHeader<R>& h = *m.h;
  const BsRec<R>& A = m.bs[pr.a];
  const BsRec<R>& B = m.bs[pr.b];
  const int na = A.s1 - A.s0, nb = B.s1 - B.s0;
  R Ta[12], Tb[12];
  load12(T + A.slot * kLinkStrideWords, Ta);
  load12(T + B.slot * kLinkStrideWords, Tb);
  unsigned long long ma = ~0ull >> (64 - na), mb = ~0ull >> (64 - nb);
  if (na * nb > 4)
  {
    // level 2 pays off only when there is more than a handful of sphere pairs
    R tax, tay, taz, tbx, tby, tbz;
    posePoint<R>(Ta, A.tx, A.ty, A.tz, tax, tay, taz);
    posePoint<R>(Tb, B.tx, B.ty, B.tz, tbx, tby, tbz);
    ma = 0ull;
    for (int i = 0; i < na; ++i)
    {
      const SphereRec<R>& sp = m.spheres[A.s0 + i];
      R x, y, z;
      posePoint<R>(Ta, sp.x, sp.y, sp.z, x, y, z);
      const R gx = x - tbx, gy = y - tby, gz = z - tbz;
      const R lim = sp.r + B.tr + h.tight_eps;
      if ((gx * gx + gy * gy + gz * gz) < lim * lim)
        ma |= 1ull << i;
    }
    if (ma == 0ull)
      return;
    mb = 0ull;
    for (int j = 0; j < nb; ++j)
    {
      const SphereRec<R>& sp = m.spheres[B.s0 + j];
      R x, y, z;
      posePoint<R>(Tb, sp.x, sp.y, sp.z, x, y, z);
      const R gx = x - tax, gy = y - tay, gz = z - taz;
      const R lim = sp.r + A.tr + h.tight_eps;
      if ((gx * gx + gy * gy + gz * gz) < lim * lim)
        mb |= 1ull << j;
    }
    if (mb == 0ull)
      return;
  }
  while (ma)
  {
    const int i = __ffsll(static_cast<long long>(ma)) - 1;
    ma &= ma - 1;
    const SphereRec<R>& sa = m.spheres[A.s0 + i];
    R axp, ayp, azp;
    posePoint<R>(Ta, sa.x, sa.y, sa.z, axp, ayp, azp);
    unsigned long long mj = mb;
    while (mj)
    {
      const int j = __ffsll(static_cast<long long>(mj)) - 1;
      mj &= mj - 1;
      const SphereRec<R>& sb = m.spheres[B.s0 + j];
      R bxp, byp, bzp;
      posePoint<R>(Tb, sb.x, sb.y, sb.z, bxp, byp, bzp);
      const R gx2 = axp - bxp, gy2 = ayp - byp, gz2 = azp - bzp;
      const R dd = gx2 * gx2 + gy2 * gy2 + gz2 * gz2;
      const R rsum = sa.r + sb.r;
      if (kExact)
        hit |= sqrtR(dd) < rsum;  // gradient.norm() < r1 + r2 (collision_env_distance_field.cpp:577-583)
      else
      {
        // squared form of dist < rsum -+ eps (both sides positive); the 1e-7 in pair_eps covers its rounding
        const R lo = rsum - h.pair_eps, hi = rsum + h.pair_eps;
        const bool c_cert = (lo > R(0)) && (dd < lo * lo);
        const bool c_maybe = dd < hi * hi;
        hit |= (cert && c_cert);
        amb |= (c_maybe && !(cert && c_cert));
      }
    }
  }
}

constexpr int kSphereChunk = B200SV_SPHERE_CHUNK;  // spheres per unrolled pass of the field loop (host pads the table)
constexpr int kQueueCap = 512;  // (state, pair) work items per warp; flushed when fewer than 32 slots remain

// Per-warp shared memory in bytes (16-byte aligned): transforms of 32 states | work queue | result words
template <typename R>
__device__ __host__ inline int perWarpBytes(int state_stride)
{
  return 32 * state_stride * static_cast<int>(sizeof(R)) + kQueueCap * 4 + 16;
}

// Dynamic shared memory: [model blob][per warp: Ts | queue | hit word, amb word]
template <typename R, typename FT, bool kFromList>
__global__ void __launch_bounds__(512, 1) checkKernel(CheckArgs<FT> a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<R> m = loadModel<R>(a.model, a.model_bytes, smem);
  const ModelHeader<R>& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  uint8_t* wbase = smem + h.total_bytes + static_cast<size_t>(warp) * perWarpBytes<R>(h.state_stride);
  R* Ts = reinterpret_cast<R*>(wbase);
  uint32_t* queue = reinterpret_cast<uint32_t*>(wbase + 32 * h.state_stride * sizeof(R));
  unsigned* words = reinterpret_cast<unsigned*>(queue + kQueueCap);  // [0] hit, [1] amb
  const bool do_world = (a.flags & 1u) != 0, do_self = (a.flags & 2u) != 0, do_intra = (a.flags & 4u) != 0;
  const typename Combined<FT>::type* __restrict__ field = static_cast<const typename Combined<FT>::type*>(a.field);
  const unsigned lt_mask = (1u << lane) - 1u;

  unsigned long long n = a.n_states;
  if (kFromList)
    n = *a.list_count < a.list_cap ? *a.list_count : a.list_cap;
  const unsigned long long n_tiles = (n + 31) / 32;
  for (unsigned long long tile = static_cast<unsigned long long>(blockIdx.x) * warps + warp; tile < n_tiles;
       tile += static_cast<unsigned long long>(gridDim.x) * warps)
  {
    const unsigned long long first = tile * 32;
    const int cnt = static_cast<int>(n - first < 32 ? n - first : 32);
    const bool active = lane < cnt;
    unsigned long long my_state = first + (active ? lane : 0);  // idle lanes shadow lane 0; their result is dropped
    if (kFromList)
      my_state = a.list[my_state];
    R* T = Ts + lane * h.state_stride;
    // ---- phase 1: FK -------------------------------------------------------------------------------------------
    fkState<R>(m, a.q + my_state * h.n_group_vars, T);
    // ---- phase 2: world / self field ----------------------------------------------------------------------------
    bool hit = false, amb = false;
    if (do_world | do_self)
      for (int k0 = 0; k0 < h.n_spheres_padded; k0 += kSphereChunk)
        sphereChunk<R, FT, kSphereChunk>(m, T, k0, do_world, do_self, field, active && !hit, hit, amb);
    // ---- phase 3: intra-group sphere pairs ----------------------------------------------------------------------
    if (do_intra && h.n_pairs > 0 && __any_sync(kFull, active && !hit))
    {
      if (lane == 0)
        words[0] = words[1] = 0u;
      __syncwarp();  // the queue consumers read other lanes' transforms
      int qn = 0;
      const bool mine = active && !hit;
      // one (state, cluster pair) item per lane: levels 2 and 3; results OR-ed into the warp's hit / amb words
      auto flush = [&]() {
        __syncwarp();
        for (int i = lane; i < qn; i += 32)
        {
          const unsigned item = queue[i];
          const int s = item & 31;
          bool ihit = false, iamb = false;
          pairItem<R>(m, Ts + s * h.state_stride, m.pairs[item >> 6], (item & 32u) != 0, ihit, iamb);
          if (ihit)
            atomicOr(&words[0], 1u << s);
          if (iamb)
            atomicOr(&words[1], 1u << s);
        }
        __syncwarp();
        qn = 0;
      };
      for (int g = 0; g < h.n_groups; ++g)
      {
        const GroupRec G = m.groups[g];
        const BsRec<R>& A = m.bs[G.a];
        R Ta[12], ax, ay, az;
        load12(T + A.slot * kLinkStrideWords, Ta);
        posePoint<R>(Ta, A.tx, A.ty, A.tz, ax, ay, az);
        for (int p = G.p0; p < G.p1; ++p)
        {
          bool live = false, certain = false;
          if (mine)
          {
            const PairRec pr = m.pairs[p];
            pairCull<R>(m, T, A, Ta, ax, ay, az, m.bs[pr.b], pr, live, certain);
          }
          const unsigned mk = __ballot_sync(kFull, live);
          if (live)
            queue[qn + __popc(mk & lt_mask)] = static_cast<uint32_t>(lane | (certain ? 32 : 0) | (p << 6));
          qn += __popc(mk);
          if (qn > kQueueCap - 32)
            flush();
        }
      }
      if (qn > 0)
        flush();
      hit |= ((words[0] >> lane) & 1u) != 0;
      amb |= ((words[1] >> lane) & 1u) != 0;
    }
    __syncwarp();  // transforms and words are rewritten by the next tile
    const bool valid = active && !hit && !amb;
    const bool flagged = active && amb && !hit;  // FP32 could not decide: the exact pass will
    if (kFromList)
    {
      // patch the bits of the rechecked states (the fast pass left them 0)
      if (valid)
        atomicOr(a.bitmap + (my_state >> 5), 1u << (my_state & 31));
    }
    else
    {
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
}

// FK only: one warp per 32 states, results written as 4x4 column-major doubles for EVERY model link.
template <typename R>
__global__ void __launch_bounds__(128) fkKernel(FkArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<R> m = loadModel<R>(a.model, a.model_bytes, smem);
  const ModelHeader<R>& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  R* Ts = reinterpret_cast<R*>(smem + h.total_bytes) + static_cast<size_t>(warp) * 32 * h.state_stride;
  const unsigned long long n_tiles = (a.n_states + 31) / 32;
  for (unsigned long long tile = static_cast<unsigned long long>(blockIdx.x) * warps + warp; tile < n_tiles;
       tile += static_cast<unsigned long long>(gridDim.x) * warps)
  {
    const unsigned long long first = tile * 32;
    const int cnt = static_cast<int>(a.n_states - first < 32 ? a.n_states - first : 32);
    if (lane < cnt)
      fkState<R>(m, a.q + (first + lane) * h.n_group_vars, Ts + lane * h.state_stride);
    __syncwarp();
    // write-out: lanes stride over (state, model link, element)
    // (a state's L 4x4 matrices are 128 L contiguous bytes: every warp store is 256 contiguous bytes; no index division)
    // A lane's element of the 4x4 matrix is fixed (rem = lane + 32 k: e = lane & 15), so what the element is -- a constant, or a
    // word of the state's transform block -- is decided once per pair of links, outside the loop over the tile's states:
    // the inner loop is one shared-memory load, one conversion, one store.
    const int per_state = a.n_model_links * 16;
    const int e = lane & 15, c = e >> 2, r = e & 3;
    double* dst0 = a.out + first * per_state;
    for (int rem = lane; rem < per_state; rem += 32)
    {
      const int l = rem >> 4;
      const int slot = __ldg(a.link_slot + l);
      if (r == 3 || slot < 0)
      {
        const double v = r == 3 ? (c == 3 ? 1.0 : 0.0) : __ldg(a.const_T + l * 16 + e);
#pragma unroll 8
        for (int s = 0; s < cnt; ++s)
          dst0[static_cast<size_t>(s) * per_state + rem] = v;
      }
      else
      {
        const R* src = Ts + slot * kLinkStrideWords + r * 4 + c;
#pragma unroll 8
        for (int s = 0; s < cnt; ++s)
          dst0[static_cast<size_t>(s) * per_state + rem] = static_cast<double>(src[s * h.state_stride]);
      }
    }
    __syncwarp();
  }
}

// Posed sphere centres of one state, FP64.
__global__ void sphereCentersKernel(FkArgs a, double* centers)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  double* Ts = reinterpret_cast<double*>(smem + h.total_bytes);
  if (threadIdx.x == 0)
    fkState<double>(m, a.q, Ts);
  __syncthreads();
  for (int k = threadIdx.x; k < h.n_spheres; k += blockDim.x)
  {
    const SphereRec<double> sp = m.spheres[k];
    double x, y, z;
    posePoint<double>(Ts + sp.slot * kLinkStrideWords, sp.x, sp.y, sp.z, x, y, z);
    centers[k * 3 + 0] = x;
    centers[k * 3 + 1] = y;
    centers[k * 3 + 2] = z;
  }
}

// Random 32-byte-sector gather: every thread reads `per_thread` pseudo-random sectors of the field footprint.  The inner
// loop is one 32-bit xorshift step (3 shifts + 3 xors), one multiply-shift to map the draw onto [0, n_sectors) (no division)
// and the load, so the measured rate is the memory system's, not the address generator's (ncu: profiles/r2_gather_*).
__global__ void gatherKernel(const uint8_t* __restrict__ field, unsigned long long n_sectors, int per_thread,
                             unsigned* sink)
{
  unsigned x = (blockIdx.x * blockDim.x + threadIdx.x) * 0x9E3779B9u + 0x7F4A7C15u;
  x |= 1u;
  const unsigned ns = static_cast<unsigned>(n_sectors);  // the launcher keeps the footprint below 2^32 sectors
  unsigned acc = 0;
#pragma unroll 8
  for (int i = 0; i < per_thread; ++i)
  {
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 5;
    const unsigned sec = __umulhi(x, ns);  // uniform on [0, ns)
    acc += __ldg(field + static_cast<size_t>(sec) * 32 + (x & 31u));
  }
  if (acc == 0xffffffffu)
    *sink = acc;
}

// K5: CollisionEnvDistanceField::getCollisionGradients (collision_env_distance_field.cpp:1517-1538) for n states, in the
// reference's order and arithmetic (FP64): self field (type SELF, :355-428 -- the dfce distance_field_ part; the per-link
// PosedDistanceFields are consulted by the reference only when an ACM is present, :386, which CHOMP's optimisation loop
// does not pass, chomp_optimizer.cpp:890), every sphere pair of every enabled link pair (INTRA, :644-709), world field
// (ENVIRONMENT, :1645-1681).  getCollisionSphereGradients: collision_distance_field_types.cpp:149-209 with
// subtract_radii = false; DistanceField::getDistanceGradient: distance_field.cpp:73-97 (1-cell margin, central
// differences times inv_twice_resolution_); PropagationDistanceField::getDistance = sqrt(d2) * resolution.
// One thread per state; sphere centres live in shared memory ([sphere][xyz] per thread, 24 B each).
__device__ __forceinline__ double cellDistance(const GradArgs& a, const uint16_t* __restrict__ d2,
                                               const uint16_t* __restrict__ nd2, size_t idx)
{
  // getDistance (propagation_distance_field.hpp:604-607): sqrt_table_[d2] - sqrt_table_[negative d2], the table being
  // sqrt(double(i)) * resolution (propagation_distance_field.cpp:99-101), built on the host with the same expression; the
  // negative part is 0 in unsigned fields
  const double pos = __ldg(a.sqrt_table + __ldg(d2 + idx));
  return nd2 ? __dsub_rn(pos, __ldg(a.sqrt_table + __ldg(nd2 + idx))) : pos;
}

__device__ __forceinline__ void fieldGradients(const GradArgs& a, const uint16_t* __restrict__ d2,
                                               const uint16_t* __restrict__ nd2, const Smem<double>& m,
                                               const BsRec<double>& C, const double* cen, int type, double* dist_out,
                                               double* grad_out, uint8_t* type_out, double& closest)
{
  const ModelHeader<double>& h = *m.h;
  for (int k = C.s0; k < C.s1; ++k)
  {
    const double x = cen[3 * k], y = cen[3 * k + 1], z = cen[3 * k + 2];
    const int gx = static_cast<int>(floor((x - h.om[0]) * h.oo_res)), gy = static_cast<int>(floor((y - h.om[1]) * h.oo_res)),
              gz = static_cast<int>(floor((z - h.om[2]) * h.oo_res));
    double dist = a.max_distance, g0 = 0.0, g1 = 0.0, g2 = 0.0;
    if (!(gx < 1 || gy < 1 || gz < 1 || gx >= h.nx - 1 || gy >= h.ny - 1 || gz >= h.nz - 1))
    {
      const size_t sy = static_cast<size_t>(h.nz), sx = sy * h.ny;
      const size_t i0 = (static_cast<size_t>(gx) * h.ny + gy) * h.nz + gz;
      g0 = (cellDistance(a, d2, nd2, i0 + sx) - cellDistance(a, d2, nd2, i0 - sx)) * a.inv_twice_resolution;
      g1 = (cellDistance(a, d2, nd2, i0 + sy) - cellDistance(a, d2, nd2, i0 - sy)) * a.inv_twice_resolution;
      g2 = (cellDistance(a, d2, nd2, i0 + 1) - cellDistance(a, d2, nd2, i0 - 1)) * a.inv_twice_resolution;
      dist = cellDistance(a, d2, nd2, i0);
    }
    if (dist < a.max_propagation)
    {
      if (dist < closest)
        closest = dist;
      if (dist < dist_out[k])
      {
        type_out[k] = static_cast<uint8_t>(type);
        dist_out[k] = dist;
        grad_out[3 * k] = g0;
        grad_out[3 * k + 1] = g1;
        grad_out[3 * k + 2] = g2;
      }
    }
  }
}

// PosedDistanceField::getCollisionSphereGradients (collision_distance_field_types.cpp:87-147) on link j's field for the
// spheres [s0, s1) of one entry, with PosedDistanceField::getDistanceGradient (types.hpp:163-176) exactly as written:
// rel_pos = pose.linear().transpose() * p (the translation is not subtracted), grad = pose * g (the translation is added),
// so an out-of-bounds query returns the link's translation as its gradient, `grad.norm() > 0` (:104) holds for every link
// away from the origin and the loop ends there.  Tj: rows of [R | t] of link j.  Explicit round-to-nearest operations in
// the reference's (Eigen's) order, no FMA.
__device__ __forceinline__ void posedFieldGradients(const GradArgs& a, const LinkFieldDev& F, const double* Tj, int s0, int s1,
                                                    const double* cen, double* dist_out, double* grad_out, uint8_t* type_out,
                                                    double& closest, double oo_res)
{
  for (int k = s0; k < s1; ++k)
  {
    const double p0 = cen[3 * k], p1 = cen[3 * k + 1], p2 = cen[3 * k + 2];
    int g[3];
    bool inb = true;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax)
    {
      const double rel = __dadd_rn(__dadd_rn(__dmul_rn(Tj[0 + ax], p0), __dmul_rn(Tj[4 + ax], p1)), __dmul_rn(Tj[8 + ax], p2));
      g[ax] = static_cast<int>(floor((rel - F.om[ax]) * oo_res));
      inb = inb && !(g[ax] < 1 || g[ax] >= F.n[ax] - 1);
    }
    double dist = a.max_distance, gl0 = 0.0, gl1 = 0.0, gl2 = 0.0;
    if (inb)
    {
      const size_t sy = static_cast<size_t>(F.n[2]), sx = sy * F.n[1];
      const size_t i0 = (static_cast<size_t>(g[0]) * F.n[1] + g[1]) * F.n[2] + g[2];
      gl0 = (cellDistance(a, F.d2, nullptr, i0 + sx) - cellDistance(a, F.d2, nullptr, i0 - sx)) * a.inv_twice_resolution;
      gl1 = (cellDistance(a, F.d2, nullptr, i0 + sy) - cellDistance(a, F.d2, nullptr, i0 - sy)) * a.inv_twice_resolution;
      gl2 = (cellDistance(a, F.d2, nullptr, i0 + 1) - cellDistance(a, F.d2, nullptr, i0 - 1)) * a.inv_twice_resolution;
      dist = cellDistance(a, F.d2, nullptr, i0);
    }
    double gr[3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
      gr[r] = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(Tj[4 * r], gl0), __dmul_rn(Tj[4 * r + 1], gl1)), __dmul_rn(Tj[4 * r + 2], gl2)),
                        Tj[4 * r + 3]);
    if (!inb && sqrt(__dadd_rn(__dadd_rn(__dmul_rn(gr[0], gr[0]), __dmul_rn(gr[1], gr[1])), __dmul_rn(gr[2], gr[2]))) > 0.0)
      return;  // "out of bounds": the reference returns here, the entry's remaining spheres are not looked at
    if (dist < a.max_propagation)
    {
      if (dist < closest)
        closest = dist;
      if (dist < dist_out[k])
      {
        type_out[k] = 1;  // SELF
        dist_out[k] = dist;
        grad_out[3 * k] = gr[0];
        grad_out[3 * k + 1] = gr[1];
        grad_out[3 * k + 2] = gr[2];
      }
    }
  }
}

// kLocal: the per-state working arrays (distances, gradients, types, centres, closest) live in LOCAL memory -- which the
// hardware interleaves by lane, so the warp's accesses are coalesced and L1/L2 resident -- and are copied to the caller's
// state-major arrays once at the end.  Working on the state-major arrays directly (a warp's 32 rows are S * 8 bytes apart)
// cost a 32-byte sector per 8-byte access in the pair loops.
constexpr int kGradMaxSpheres = 96, kGradMaxLinks = 64, kGradMaxStride = 352;  // stride: 28 links x 12 words + padding

template <bool kLocal>
__global__ void __launch_bounds__(128) gradientKernel(GradArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  // the state's link transforms: local memory as well (shared memory held the kernel at 8 warps per SM)
  double Tloc[kLocal ? kGradMaxStride : 1];
  double* T = kLocal ? Tloc : reinterpret_cast<double*>(smem + h.total_bytes) + static_cast<size_t>(threadIdx.x) * h.state_stride;
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i >= a.n_states)
    return;
  const int S = h.n_spheres;
  double Dl[kLocal ? kGradMaxSpheres : 1], Gl[kLocal ? 3 * kGradMaxSpheres : 1], Cl[kLocal ? 3 * kGradMaxSpheres : 1],
      Ll[kLocal ? kGradMaxLinks : 1];
  uint8_t Tl[kLocal ? kGradMaxSpheres : 1];
  double* D = kLocal ? Dl : a.distances + i * S;
  double* G = kLocal ? Gl : a.gradients + i * S * 3;
  uint8_t* Ty = kLocal ? Tl : a.types + i * S;
  double* cen = kLocal ? Cl : a.centers + i * S * 3;  // posed sphere centres (GradientInfo::sphere_locations)
  double* closest = kLocal ? Ll : a.closest + i * a.n_glinks;
  fkState<double>(m, a.q + i * h.n_group_vars, T);
  for (int k = 0; k < S; ++k)
  {
    const SphereRec<double>& sp = m.spheres[k];
    posePoint<double>(T + sp.slot * kLinkStrideWords, sp.x, sp.y, sp.z, cen[3 * k], cen[3 * k + 1], cen[3 * k + 2]);
    D[k] = 1.7976931348623157e308;  // DBL_MAX
    Ty[k] = 0;
    G[3 * k] = G[3 * k + 1] = G[3 * k + 2] = 0.0;
  }
  // closest_distance per dfce link; links are runs of clusters with the same glink id (BsRec::pad)
  for (int g = 0; g < a.n_glinks; ++g)
    closest[g] = 1.7976931348623157e308;
  // getSelfProximityGradients (:355-428): per entry, the other links' posed fields first (only with a non-empty ACM), then
  // the cache entry's self field.  Entries are independent of each other, so doing all the posed fields first keeps the order
  // that matters (within an entry).
  if (a.lf != nullptr)
    for (int g = 0; g < a.n_glinks; ++g)
    {
      const int s0 = a.sphere_start[g], s1 = a.sphere_start[g + 1];
      if (s1 <= s0 || !a.self_enabled[g])
        continue;
      for (int j = 0; j < a.n_glinks; ++j)
      {
        if (!a.lf_enabled[g * a.n_glinks + j] || a.lf[j].d2 == nullptr)
          continue;
        posedFieldGradients(a, a.lf[j], T + a.glink_slot[j] * kLinkStrideWords, s0, s1, cen, D, G, Ty, closest[g], h.oo_res);
      }
    }
  for (int c = 0; c < h.n_bs; ++c)
  {  // self field
    const BsRec<double>& C = m.bs[c];
    if (a.self_enabled[C.pad])
      fieldGradients(a, a.d2_self, a.nd2_self, m, C, cen, 1, D, G, Ty, closest[C.pad]);
  }
  for (int p = 0; p < h.n_pairs; ++p)
  {  // intra-group: every sphere pair of the two clusters (all cluster pairs of a link pair = all its sphere pairs)
    const BsRec<double>&A = m.bs[m.pairs[p].a], &B = m.bs[m.pairs[p].b];
    for (int k = A.s0; k < A.s1; ++k)
      for (int l = B.s0; l < B.s1; ++l)
      {
        const double gx = cen[3 * k] - cen[3 * l], gy = cen[3 * k + 1] - cen[3 * l + 1], gz = cen[3 * k + 2] - cen[3 * l + 2];
        const double dd = __dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz));
        const double far = fmax(D[k], D[l]);
        if (dd > far * far * (1.0 + 1e-9))
          continue;  // sqrt(dd) exceeds both recorded distances: neither update can happen (DBL_MAX squares to +inf)
        const double dist = sqrt(dd);
        if (dist < D[k])
        {
          D[k] = dist;
          G[3 * k] = gx; G[3 * k + 1] = gy; G[3 * k + 2] = gz;
          Ty[k] = 2;
        }
        if (dist < D[l])
        {
          D[l] = dist;
          G[3 * l] = -gx; G[3 * l + 1] = -gy; G[3 * l + 2] = -gz;
          Ty[l] = 2;
        }
      }
  }
  for (int c = 0; c < h.n_bs; ++c)  // world field
    fieldGradients(a, a.d2_world, a.nd2_world, m, m.bs[c], cen, 3, D, G, Ty, closest[m.bs[c].pad]);
  if (kLocal)
  {
    for (int k = 0; k < S; ++k)
    {
      a.distances[i * S + k] = D[k];
      a.types[i * S + k] = Ty[k];
    }
    for (int k = 0; k < 3 * S; ++k)
    {
      a.gradients[i * S * 3 + k] = G[k];
      a.centers[i * S * 3 + k] = cen[k];
    }
    for (int g = 0; g < a.n_glinks; ++g)
      a.closest[i * a.n_glinks + g] = closest[g];
  }
}

// CollisionEnvDistanceField::checkCollision with req.contacts = true (collision_env_distance_field.cpp:1389-1411), n states:
// getSelfCollisions (:274-353) -> getIntraGroupCollisions (:430-642) -> getEnvironmentCollisions (:1561-1643), each pass
// ending the sequence once contact_count >= max_contacts; the contact-listing getCollisionSphereCollision
// (collision_distance_field_types.cpp:238-271) with num_coll = min(max_contacts_per_pair, max_contacts - contact_count).
// FP64, the reference's order of links, pairs and spheres: WHICH contacts are reported depends on it.  The field
// predicate is the integer threshold form (d2 < thr, exact: see sphereThreshold in api.cu).
__device__ __forceinline__ bool sphereHitsField(const ModelHeader<double>& h, const uint16_t* __restrict__ d2, const double* c,
                                                int thr)
{
  const int gx = static_cast<int>(floor((c[0] - h.om[0]) * h.oo_res)), gy = static_cast<int>(floor((c[1] - h.om[1]) * h.oo_res)),
            gz = static_cast<int>(floor((c[2] - h.om[2]) * h.oo_res));
  if (gx < 1 || gy < 1 || gz < 1 || gx >= h.nx - 1 || gy >= h.ny - 1 || gz >= h.nz - 1)
    return false;  // reads max_distance: never collides (distance_field.cpp:82, types.cpp:256)
  return static_cast<int>(__ldg(d2 + (static_cast<size_t>(gx) * h.ny + gy) * h.nz + gz)) < thr;
}

// kLocal: link transforms and posed centres in lane-interleaved local memory (see gradientKernel)
template <bool kLocal>
__global__ void __launch_bounds__(128) contactKernel(ContactArgs a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  double Tloc[kLocal ? kGradMaxStride : 1], Cl[kLocal ? 3 * kGradMaxSpheres : 1];
  double* T = kLocal ? Tloc : reinterpret_cast<double*>(smem + h.total_bytes) + static_cast<size_t>(threadIdx.x) * h.state_stride;
  const unsigned long long st = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (st >= a.n_states)
    return;
  const int S = h.n_spheres, L = a.n_glinks;
  double* cen = kLocal ? Cl : a.centers + st * S * 3;
  b200sv_contact_rec* out = a.contacts + st * a.max_contacts;
  fkState<double>(m, a.q + st * h.n_group_vars, T);
  for (int k = 0; k < S; ++k)
  {
    const SphereRec<double>& sp = m.spheres[k];
    posePoint<double>(T + sp.slot * kLinkStrideWords, sp.x, sp.y, sp.z, cen[3 * k], cen[3 * k + 1], cen[3 * k + 2]);
  }
  unsigned count = 0;
  bool coll_any = false, done = false;
  auto fieldPass = [&](const uint16_t* __restrict__ d2, bool self_pass) {
    for (int i = 0; i < L && !done; ++i)
    {
      const int s0 = a.sphere_start[i], s1 = a.sphere_start[i + 1];
      if (s1 <= s0 || (self_pass && !a.self_enabled[i]))
        continue;
      const unsigned left = a.max_contacts - count;
      const unsigned num_coll = a.max_per_pair < left ? a.max_per_pair : left;
      unsigned nc = 0;
      bool coll = false;
      for (int k = s0; k < s1; ++k)
      {
        // the world threshold is the sphere's threshold; the self pass uses the same predicate (self_enabled checked above)
        if (sphereHitsField(h, d2, cen + 3 * k, m.spheres[k].thr))
        {
          coll = true;
          if (num_coll == 0)
            break;
          b200sv_contact_rec& o = out[count + nc];
          o.body_1 = i;
          o.body_2 = self_pass ? -1 : -2;
          o.sphere_1 = k - s0;
          o.sphere_2 = -1;
          o.pos[0] = cen[3 * k]; o.pos[1] = cen[3 * k + 1]; o.pos[2] = cen[3 * k + 2];
          if (++nc >= num_coll)
            break;
        }
      }
      if (coll)
      {
        coll_any = true;
        count += nc;
        if (count >= a.max_contacts)
          done = true;
      }
    }
    if (count >= a.max_contacts)
      done = true;
  };
  if ((a.flags & 2u) && !done)
    fieldPass(a.d2_self, true);
  if ((a.flags & 4u) && !done)
  {
    for (int i = 0; i < L && !done; ++i)
    {
      if (a.sphere_start[i + 1] <= a.sphere_start[i])
        continue;
      for (int j = i + 1; j < L && !done; ++j)
      {
        if (a.sphere_start[j + 1] <= a.sphere_start[j] || !a.intra[static_cast<size_t>(i) * L + j])
          continue;
        // doBoundingSpheresIntersect, quirk kept: squared centre distance against the plain radius sum (types.cpp:416-417)
        auto boundsMeet = [&](int u, int v) {
          double cu[3], cv[3];
          posePoint<double>(T + a.glink_slot[u] * kLinkStrideWords, a.glink_bs[4 * u], a.glink_bs[4 * u + 1], a.glink_bs[4 * u + 2],
                            cu[0], cu[1], cu[2]);
          posePoint<double>(T + a.glink_slot[v] * kLinkStrideWords, a.glink_bs[4 * v], a.glink_bs[4 * v + 1], a.glink_bs[4 * v + 2],
                            cv[0], cv[1], cv[2]);
          const double ex = cu[0] - cv[0], ey = cu[1] - cv[1], ez = cu[2] - cv[2];
          return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez)) <
                 (a.glink_bs[4 * u + 3] + a.glink_bs[4 * v + 3]);
        };
        bool meet = boundsMeet(i, j);
        // a body of several shapes passes as a whole when ANY of its shapes does (collision_env_distance_field.cpp:455-507)
        const int bi = a.body_id[i], bj = a.body_id[j];
        if (!meet && (bi >= 0 || bj >= 0))
          for (int u = 0; u < L && !meet; ++u)
          {
            if ((u != i && (bi < 0 || a.body_id[u] != bi)) || a.sphere_start[u + 1] <= a.sphere_start[u])
              continue;
            for (int v = 0; v < L && !meet; ++v)
              if ((v == j || (bj >= 0 && a.body_id[v] == bj)) && a.sphere_start[v + 1] > a.sphere_start[v])
                meet = boundsMeet(u, v);
          }
        if (!meet)
          continue;
        int num_pair = 0;
        for (int k = a.sphere_start[i]; k < a.sphere_start[i + 1] && num_pair < static_cast<int>(a.max_per_pair) && !done; ++k)
          for (int l = a.sphere_start[j]; l < a.sphere_start[j + 1] && num_pair < static_cast<int>(a.max_per_pair); ++l)
          {
            const double gx = cen[3 * k] - cen[3 * l], gy = cen[3 * k + 1] - cen[3 * l + 1], gz = cen[3 * k + 2] - cen[3 * l + 2];
            const double dist = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(gx, gx), __dmul_rn(gy, gy)), __dmul_rn(gz, gz)));
            if (dist < m.spheres[k].r + m.spheres[l].r)
            {
              coll_any = true;
              b200sv_contact_rec& o = out[count++];
              o.body_1 = i;
              o.body_2 = j;
              o.sphere_1 = k - a.sphere_start[i];
              o.sphere_2 = l - a.sphere_start[j];
              o.pos[0] = cen[3 * k]; o.pos[1] = cen[3 * k + 1]; o.pos[2] = cen[3 * k + 2];
              ++num_pair;
              if (count >= a.max_contacts)
              {
                done = true;
                break;
              }
            }
          }
      }
    }
  }
  if ((a.flags & 1u) && !done)
    fieldPass(a.d2_world, false);
  a.collision[st] = coll_any ? 1 : 0;
  a.contact_count[st] = static_cast<int32_t>(count);
}

// Exact pass over the recheck list, ONE WARP PER STATE: the list is short (about one state in 7 000), so what matters is
// the latency of a single state, not throughput.  Lane 0 runs the FP64 forward kinematics into shared memory, then the
// 32 lanes share the spheres and the cluster pairs of that one state.  Same device functions (same arithmetic) as
// checkKernel<double, ., true>.
constexpr int kRecheckScLinks = 64;
template <typename FT>
__global__ void __launch_bounds__(128) recheckWarpKernel(CheckArgs<FT> a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  double* T = reinterpret_cast<double*>(smem + h.total_bytes) + static_cast<size_t>(warp) * h.state_stride;
  // sines and cosines of the state's joints, one link per lane, for the chain lane 0 walks (robots of up to kRecheckScLinks
  // links; beyond that lane 0 computes them on its way, as before)
  double* sc = reinterpret_cast<double*>(smem + h.total_bytes) + static_cast<size_t>(warps) * h.state_stride +
               static_cast<size_t>(warp) * 2 * kRecheckScLinks;
  const bool par_sc = h.n_links <= kRecheckScLinks;
  const bool do_world = (a.flags & 1u) != 0, do_self = (a.flags & 2u) != 0, do_intra = (a.flags & 4u) != 0;
  const typename Combined<FT>::type* __restrict__ field = static_cast<const typename Combined<FT>::type*>(a.field);
  // launched with programmatic stream serialization: everything above ran while the kernel before this one (the FP32 pass that
  // fills the list) was draining; nothing it wrote is read before this point.  A no-op for an ordinary launch.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const unsigned n = *a.list_count < a.list_cap ? *a.list_count : a.list_cap;
  for (unsigned i = blockIdx.x * warps + warp; i < n; i += gridDim.x * warps)
  {
    const unsigned long long state = a.list[i];
    const double* qrow = a.q + state * h.n_group_vars;
    if (par_sc)
    {
      for (int l = lane; l < h.n_links; l += 32)
      {
        const LinkRec<double>& L = m.links[l];
        if (L.type == REVOLUTE || L.type == PLANAR)
          sincos(varValue<double>(m.vars[L.var + (L.type == PLANAR ? 2 : 0)], qrow), &sc[2 * l], &sc[2 * l + 1]);
      }
      __syncwarp();
    }
    if (lane == 0)
      fkState<double>(m, qrow, T, par_sc ? sc : nullptr);
    __syncwarp();
    bool hit = false, amb = false;
    if (do_world | do_self)
      for (int k = lane; k < h.n_spheres_padded; k += 32)
        sphereChunk<double, FT, 1>(m, T, k, do_world, do_self, field, true, hit, amb);
    if (do_intra && !__any_sync(kFull, hit))
      for (int p = lane; p < h.n_pairs; p += 32)
      {
        const PairRec pr = m.pairs[p];
        const BsRec<double>& A = m.bs[pr.a];
        double Ta[12], ax, ay, az;
        load12(T + A.slot * kLinkStrideWords, Ta);
        posePoint<double>(Ta, A.tx, A.ty, A.tz, ax, ay, az);
        bool live = false, certain = false;
        pairCull<double>(m, T, A, Ta, ax, ay, az, m.bs[pr.b], pr, live, certain);
        if (live)
          pairItem<double>(m, T, pr, certain, hit, amb);
      }
    const bool any = __any_sync(kFull, hit);
    if (lane == 0 && !any)
      atomicOr(a.bitmap + (state >> 5), 1u << (state & 31));
    __syncwarp();  // T is rewritten for the next list entry
  }
}

// The latency path of b200sv_check_batch (a single-state checkCollision through the plugin is a batch of 1): ONE launch and no
// copies.  The states arrive inside the kernel parameters (or, beyond 32 doubles, are read from mapped pinned host memory); one
// CTA per state runs the exact FP64 check -- the arithmetic every FP32 decision is referred to anyway -- and lane 0 stores the
// state's result byte (1 = valid, 0 = in collision) straight into mapped host memory, where the host is polling for it.
// One state by one CTA of 128 threads, exact FP64: qs holds the state; returns nonzero when the state is in collision.
// What is left after the launch is a chain of dependent latencies (sin / cos -> FK chain -> field gathers -> pair tests),
// so the four warps shorten every link of that chain they can: the joints' sines and cosines in parallel, the spheres in
// ONE pass of 128 lanes, the cluster pairs in two.
template <typename FT>
__device__ __forceinline__ int oneStateByCta(const Smem<double>& m, double* T, const double* qs, double* sc, const void* field_v,
                                             uint32_t flags)
{
  const ModelHeader<double>& h = *m.h;
  const bool do_world = (flags & 1u) != 0, do_self = (flags & 2u) != 0, do_intra = (flags & 4u) != 0;
  const typename Combined<FT>::type* __restrict__ field = static_cast<const typename Combined<FT>::type*>(field_v);
  const int tid = threadIdx.x;
  for (int l = tid; l < h.n_links; l += blockDim.x)
  {
    const LinkRec<double>& L = m.links[l];
    if (L.type == REVOLUTE || L.type == PLANAR)
      sincos(varValue<double>(m.vars[L.var + (L.type == PLANAR ? 2 : 0)], qs), &sc[2 * l], &sc[2 * l + 1]);
  }
  __syncthreads();
  if (tid == 0)
    fkState<double>(m, qs, T, sc);
  __syncthreads();
  bool hit = false, amb = false;
  if (do_world | do_self)
    for (int k = tid; k < h.n_spheres_padded; k += blockDim.x)
      sphereChunk<double, FT, 1>(m, T, k, do_world, do_self, field, true, hit, amb);
  if (do_intra && !__syncthreads_or(hit ? 1 : 0))
    for (int p = tid; p < h.n_pairs; p += blockDim.x)
    {
      const PairRec pr = m.pairs[p];
      const BsRec<double>& A = m.bs[pr.a];
      double Ta[12], ax, ay, az;
      load12(T + A.slot * kLinkStrideWords, Ta);
      posePoint<double>(Ta, A.tx, A.ty, A.tz, ax, ay, az);
      bool live = false, certain = false;
      pairCull<double>(m, T, A, Ta, ax, ay, az, m.bs[pr.b], pr, live, certain);
      if (live)
        pairItem<double>(m, T, pr, certain, hit, amb);
    }
  return __syncthreads_or(hit ? 1 : 0);
}

template <typename FT>
__global__ void __launch_bounds__(128) latencyCheckKernel(LatencyArgs<FT> a)
{
  // ONE STATE PER CTA
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  double* T = reinterpret_cast<double*>(smem + h.total_bytes);
  double* qs = T + a.state_stride;  // the state
  double* sc = qs + 32;             // sin, cos per link
  const int state = blockIdx.x;
  const int tid = threadIdx.x;
  if (tid < h.n_group_vars)
    qs[tid] = a.use_inline ? a.q_inline[state * h.n_group_vars + tid] : a.q[state * h.n_group_vars + tid];
  __syncthreads();
  const int any = oneStateByCta<FT>(m, T, qs, sc, a.field, a.flags);
  if (tid == 0)
  {
    a.result[state] = any ? 0 : 1;
    __threadfence_system();
  }
}

// One CTA per visited state of a few edges: state j / c of edge e by JointModelGroup::interpolate (joint_model_group.cpp:474-486;
// continuous revolute joints along the shorter arc, revolute_joint_model.cpp:138-171) -- the arithmetic of motionExpandKernel,
// roundings included -- then the exact check.
template <typename FT>
__global__ void __launch_bounds__(128) edgeLatencyKernel(EdgeLatencyArgs<FT> a)
{
  constexpr double kPi = 3.14159265358979323846;
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  double* T = reinterpret_cast<double*>(smem + h.total_bytes);
  double* qs = T + a.state_stride;
  double* sc = qs + 32;
  const int tid = threadIdx.x;
  const unsigned s = blockIdx.x;
  int e = 0;
  while (e + 1 < a.n_edges && a.offsets[e + 1] <= s)
    ++e;
  const unsigned c = a.offsets[e + 1] - a.offsets[e], j = s - a.offsets[e] + 1u;  // 1 .. c
  if (tid < a.n_vars)
  {
    const double f = a.use_inline ? a.ends_inline[tid] : a.ends[static_cast<size_t>(e) * a.n_vars + tid];
    const double t = a.use_inline ? a.ends_inline[a.n_vars + tid]
                                  : a.ends[(static_cast<size_t>(a.n_edges) + e) * a.n_vars + tid];
    double v = t;  // j == c: s2 itself
    if (j != c)
    {
      const double tt = static_cast<double>(j) / static_cast<double>(c);
      double diff = __dsub_rn(t, f);
      if (((a.cont_mask >> tid) & 1u) && fabs(diff) > kPi)
      {
        diff = diff > 0.0 ? __dsub_rn(2.0 * kPi, diff) : __dsub_rn(-2.0 * kPi, diff);
        v = __dsub_rn(f, __dmul_rn(diff, tt));
        if (v > kPi)
          v = __dsub_rn(v, 2.0 * kPi);
        else if (v < -kPi)
          v = __dadd_rn(v, 2.0 * kPi);
      }
      else
        v = __dadd_rn(f, __dmul_rn(diff, tt));
    }
    qs[tid] = v;
  }
  __syncthreads();
  const int any = oneStateByCta<FT>(m, T, qs, sc, a.field, a.flags);
  if (tid == 0)
  {
    a.result[s] = any ? 0 : 1;
    __threadfence_system();
  }
}

__device__ __forceinline__ unsigned long long globalTimerNs()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// One CTA that answers its first request (state in the parameters), then keeps polling the mailbox for the next one until
// nothing has arrived for linger_ns, the host asks it to quit, or life_ns have passed since the launch.  The model blob is
// staged into shared memory once.  Thread 0 polls (one 16-byte read of {req, quit} over PCIe per round); a new request's
// state is then read by V lanes at once (one more round trip), and the answer is ONE 8-byte store the host is polling for.
template <typename FT>
__global__ void __launch_bounds__(128) lingerCheckKernel(LingerArgs<FT> a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  __shared__ unsigned long long s_req;
  const Smem<double> m = loadModel<double>(a.model, a.model_bytes, smem);
  const ModelHeader<double>& h = *m.h;
  double* T = reinterpret_cast<double*>(smem + h.total_bytes);
  double* qs = T + a.state_stride;
  double* sc = qs + 32;
  const int tid = threadIdx.x;
  const unsigned long long t_start = globalTimerNs();
  unsigned long long t_last = t_start, served = a.first_req >> 3;
  unsigned long long req = a.first_req;
  if (tid < h.n_group_vars)
    qs[tid] = a.q_inline[tid];
  __syncthreads();
  for (;;)
  {
    const int any = oneStateByCta<FT>(m, T, qs, sc, a.field, static_cast<uint32_t>(req & 7u));
    if (tid == 0)
    {
      *reinterpret_cast<volatile unsigned long long*>(&a.mail->done) = ((req >> 3) << 8) | (any ? 0ull : 1ull);
      __threadfence_system();
      served = req >> 3;
      t_last = globalTimerNs();
      // wait for the next request
      unsigned long long next = 0ull;
      for (;;)
      {
        const unsigned long long now = globalTimerNs();
        if (now - t_start > a.life_ns)
          break;  // whatever the host does: the host stops posting to a kernel this old and launches a new one
        unsigned long long r, qt;
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(r), "=l"(qt) : "l"(&a.mail->req) : "memory");
        if ((r >> 3) > served)
        {
          next = r;
          break;
        }
        if (qt == a.epoch || now - t_last > a.linger_ns)
          break;
      }
      s_req = next;
    }
    __syncthreads();
    req = s_req;
    if (req == 0ull)
      break;
    if (tid < h.n_group_vars)
      qs[tid] = *reinterpret_cast<const volatile double*>(&a.mail->q[tid]);
    __syncthreads();
  }
}

// cudaFuncSetAttribute is a host-side driver call: do it when the requirement grows, not on every launch.
struct SmemGrant
{
  size_t granted[16] = { 0 };
};

template <typename K>
cudaError_t ensureSmem(K kernel, size_t smem_bytes, SmemGrant& g)
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

template <typename R, typename FT, bool kList>
cudaError_t launchOne(const CheckArgs<FT>& a, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant g;
  cudaError_t e = ensureSmem(checkKernel<R, FT, kList>, smem_bytes, g);
  if (e != cudaSuccess)
    return e;
  checkKernel<R, FT, kList><<<grid, block, smem_bytes, stream>>>(a);
  return cudaGetLastError();
}

template <typename R, typename FT, bool kList>
int regsOf()
{
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, checkKernel<R, FT, kList>) != cudaSuccess)
    return 128;
  return fa.numRegs;
}
}  // namespace

template <typename FT>
cudaError_t launchRecheckWarp(const CheckArgs<FT>& a, int grid, int state_stride, cudaStream_t stream)
{
  static SmemGrant g;
  const int block = 128;
  const size_t smem = ((static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(block / 32) * (state_stride + 2 * kRecheckScLinks) * sizeof(double);
  cudaError_t e = ensureSmem(recheckWarpKernel<FT>, smem, g);
  if (e != cudaSuccess)
    return e;
  static const bool pdl = !(getenv("B200SV_RECHECK_PDL") && atoi(getenv("B200SV_RECHECK_PDL")) == 0);
  if (!pdl)
  {
    recheckWarpKernel<FT><<<grid, block, smem, stream>>>(a);
    return cudaGetLastError();
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(static_cast<unsigned>(grid));
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, recheckWarpKernel<FT>, a);
}
template cudaError_t launchRecheckWarp<uint8_t>(const CheckArgs<uint8_t>&, int, int, cudaStream_t);
template cudaError_t launchRecheckWarp<uint16_t>(const CheckArgs<uint16_t>&, int, int, cudaStream_t);

template <typename FT>
cudaError_t launchLatencyCheck(const LatencyArgs<FT>& a, cudaStream_t stream)
{
  static SmemGrant g;
  const int block = 128;
  const size_t smem = ((static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(a.state_stride + 32 + 2 * a.n_links) * sizeof(double);
  cudaError_t e = ensureSmem(latencyCheckKernel<FT>, smem, g);
  if (e != cudaSuccess)
    return e;
  latencyCheckKernel<FT><<<a.n_states, block, smem, stream>>>(a);
  return cudaGetLastError();
}
template cudaError_t launchLatencyCheck<uint8_t>(const LatencyArgs<uint8_t>&, cudaStream_t);
template cudaError_t launchLatencyCheck<uint16_t>(const LatencyArgs<uint16_t>&, cudaStream_t);

template <typename FT>
cudaError_t launchEdgeLatencyCheck(const EdgeLatencyArgs<FT>& a, cudaStream_t stream)
{
  static SmemGrant g;
  const size_t smem = ((static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(a.state_stride + 32 + 2 * a.n_links) * sizeof(double);
  cudaError_t e = ensureSmem(edgeLatencyKernel<FT>, smem, g);
  if (e != cudaSuccess)
    return e;
  edgeLatencyKernel<FT><<<a.offsets[a.n_edges], 128, smem, stream>>>(a);
  return cudaGetLastError();
}
template cudaError_t launchEdgeLatencyCheck<uint8_t>(const EdgeLatencyArgs<uint8_t>&, cudaStream_t);
template cudaError_t launchEdgeLatencyCheck<uint16_t>(const EdgeLatencyArgs<uint16_t>&, cudaStream_t);

template <typename FT>
cudaError_t launchLingerCheck(const LingerArgs<FT>& a, cudaStream_t stream)
{
  static SmemGrant g;
  const size_t smem = ((static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(a.state_stride + 32 + 2 * a.n_links) * sizeof(double);
  cudaError_t e = ensureSmem(lingerCheckKernel<FT>, smem, g);
  if (e != cudaSuccess)
    return e;
  lingerCheckKernel<FT><<<1, 128, smem, stream>>>(a);
  return cudaGetLastError();
}
template cudaError_t launchLingerCheck<uint8_t>(const LingerArgs<uint8_t>&, cudaStream_t);
template cudaError_t launchLingerCheck<uint16_t>(const LingerArgs<uint16_t>&, cudaStream_t);

int checkPerWarpBytes(bool fp64, int state_stride)
{
  return fp64 ? perWarpBytes<double>(state_stride) : perWarpBytes<float>(state_stride);
}

template <typename FT>
cudaError_t launchCheck(const CheckArgs<FT>& a, bool fp64, bool from_list, int grid, int block, size_t smem_bytes,
                        cudaStream_t stream)
{
  if (fp64 && from_list)
    return launchOne<double, FT, true>(a, grid, block, smem_bytes, stream);
  if (fp64)
    return launchOne<double, FT, false>(a, grid, block, smem_bytes, stream);
  return launchOne<float, FT, false>(a, grid, block, smem_bytes, stream);
}
template cudaError_t launchCheck<uint8_t>(const CheckArgs<uint8_t>&, bool, bool, int, int, size_t, cudaStream_t);
template cudaError_t launchCheck<uint16_t>(const CheckArgs<uint16_t>&, bool, bool, int, int, size_t, cudaStream_t);

int checkKernelRegs(bool fp64, bool from_list, int field_bytes)
{
  if (field_bytes == 1)
    return fp64 ? (from_list ? regsOf<double, uint8_t, true>() : regsOf<double, uint8_t, false>()) :
                  regsOf<float, uint8_t, false>();
  return fp64 ? (from_list ? regsOf<double, uint16_t, true>() : regsOf<double, uint16_t, false>()) :
                regsOf<float, uint16_t, false>();
}

cudaError_t launchFk(const FkArgs& a, bool fp64, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant gd, gf;
  cudaError_t e;
  if (fp64)
  {
    if ((e = ensureSmem(fkKernel<double>, smem_bytes, gd)) != cudaSuccess)
      return e;
    fkKernel<double><<<grid, block, smem_bytes, stream>>>(a);
  }
  else
  {
    if ((e = ensureSmem(fkKernel<float>, smem_bytes, gf)) != cudaSuccess)
      return e;
    fkKernel<float><<<grid, block, smem_bytes, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launchSphereCenters(const FkArgs& a, double* d_centers, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant g;
  cudaError_t e = ensureSmem(sphereCentersKernel, smem_bytes, g);
  if (e != cudaSuccess)
    return e;
  sphereCentersKernel<<<1, 128, smem_bytes, stream>>>(a, d_centers);
  return cudaGetLastError();
}

cudaError_t launchGradients(const GradArgs& a, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant gl, gg;
  if (a.n_spheres <= kGradMaxSpheres && a.n_glinks <= kGradMaxLinks && a.state_stride <= kGradMaxStride)
  {
    // everything per state in local memory: shared memory holds the model blob only, blocks of 128 threads
    const size_t blob = (static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15);
    cudaError_t e = ensureSmem(gradientKernel<true>, blob, gl);
    if (e != cudaSuccess)
      return e;
    const unsigned long long n = a.n_states;
    gradientKernel<true><<<static_cast<unsigned>((n + 127) / 128), 128, blob, stream>>>(a);
  }
  else
  {
    cudaError_t e = ensureSmem(gradientKernel<false>, smem_bytes, gg);
    if (e != cudaSuccess)
      return e;
    gradientKernel<false><<<grid, block, smem_bytes, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launchContacts(const ContactArgs& a, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant gl, gg;
  if (a.n_spheres <= kGradMaxSpheres && a.state_stride <= kGradMaxStride)
  {
    const size_t blob = (static_cast<size_t>(a.model_bytes) + 15) & ~static_cast<size_t>(15);
    cudaError_t e = ensureSmem(contactKernel<true>, blob, gl);
    if (e != cudaSuccess)
      return e;
    const unsigned long long n = a.n_states;
    contactKernel<true><<<static_cast<unsigned>((n + 127) / 128), 128, blob, stream>>>(a);
    return cudaGetLastError();
  }
  cudaError_t e = ensureSmem(contactKernel<false>, smem_bytes, gg);
  if (e != cudaSuccess)
    return e;
  contactKernel<false><<<grid, block, smem_bytes, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launchGather(const uint8_t* field, unsigned long long n_sectors, int per_thread, int grid, int block,
                         unsigned* sink, cudaStream_t stream)
{
  gatherKernel<<<grid, block, 0, stream>>>(field, n_sectors, per_thread, sink);
  return cudaGetLastError();
}
}  // namespace b200sv
