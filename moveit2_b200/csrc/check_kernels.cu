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
// Mapping (B200, sm_100a): persistent CTAs.  Each warp owns a tile of kTile consecutive states.
//   phase 1  thread-per-state FK: lane s (< kTile) walks the device links in index order (parents first); link
//            transforms go to shared memory as 3x4 row-major rows of 16 bytes (one 128-bit access per row).
//   phase 2  warp-per-state checks: lanes own spheres (two per lane per pass, for memory-level parallelism); each
//            lane poses its sphere, converts the centre to a voxel and gathers ONE byte from the world field and one
//            from the self field (L2-resident compact d^2), compares against the sphere's integer threshold; early
//            exit is warp-uniform.  Intra-group: lanes test link pairs (the reference's bounding-sphere cull plus a
//            conservative tight-sphere cull), then for each surviving link pair the spheres of either link are
//            tested against the other link's tight sphere, and only the survivors meet sphere against sphere.
// Precision: the fast pass is FP32 (template R = float).  A sphere whose voxel coordinate is within `margin` of a
// cell face, or a pair whose distance is within eps of its threshold, is evaluated for BOTH outcomes; if the
// state's boolean could depend on that, the state index is appended to a recheck list and the exact pass
// (R = double, same code) recomputes it.  The FP64 pass restates the reference's double arithmetic
// operation by operation.  The conservative culls only ever skip pairs that cannot collide, so they change no result.
#include <cuda_runtime.h>

#include <cstdint>

#include "device_model.hpp"
#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr unsigned kFull = 0xffffffffu;

template <typename R>
struct Smem
{
  const ModelHeader<R>* h;
  const LinkRec<R>* links;
  const VarRec* vars;
  const SphereRec<R>* spheres;
  const BsRec<R>* bs;
  const PairRec* pairs;
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
__device__ __forceinline__ void load4(const float* p, float* o)
{
  const float4 t = *reinterpret_cast<const float4*>(p);
  o[0] = t.x; o[1] = t.y; o[2] = t.z; o[3] = t.w;
}
__device__ __forceinline__ void load4(const double* p, double* o)
{
  const double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
  o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
}
__device__ __forceinline__ void store4(float* p, float a, float b, float c, float d)
{
  *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void store4(double* p, double a, double b, double c, double d)
{
  reinterpret_cast<double2*>(p)[0] = make_double2(a, b);
  reinterpret_cast<double2*>(p)[1] = make_double2(c, d);
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
template <typename R>
__device__ __forceinline__ void fkState(const Smem<R>& m, const double* __restrict__ qrow, R* T)
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
        const R q = varValue<R>(m.vars[L.var], qrow);
        R s, c;
        sincosR(q, &s, &c);
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
        const R th = varValue<R>(m.vars[L.var + 2], qrow);
        R s, c;
        sincosR(th, &s, &c);
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
  return m;
}

// What one lane knows about one of its spheres after posing it: where to read the fields, and whether FP32
// rounding leaves the voxel undetermined.
struct Probe
{
  int idx;         // voxel index of the certain cell, -1: nothing to read (no thresholds / outside the 1-cell margin)
  int thr_w, thr_s;
  bool risky;      // FP32 only: the centre is within `margin` of a cell face
  int gx, gy, gz, hx, hy, hz;
};

template <typename R>
__device__ __forceinline__ Probe probeSphere(const Smem<R>& m, const R* T, R* cent, int k, bool do_world, bool do_self)
{
  constexpr bool kExact = sizeof(R) == 8;
  const ModelHeader<R>& h = *m.h;
  Probe p;
  p.idx = -1;
  p.risky = false;
  p.thr_w = p.thr_s = 0;
  p.gx = p.gy = p.gz = p.hx = p.hy = p.hz = 0;
  if (k >= h.n_spheres)
    return p;
  const SphereRec<R> sp = m.spheres[k];
  R Tl[12];
  load12(T + sp.slot * kLinkStrideWords, Tl);
  R cx, cy, cz;
  posePoint<R>(Tl, sp.x, sp.y, sp.z, cx, cy, cz);
  store4(cent + k * 4, cx, cy, cz, sp.r);
  p.thr_w = do_world ? sp.thr : 0;
  p.thr_s = do_self ? sp.thr_self : 0;
  if ((p.thr_w | p.thr_s) == 0)
    return p;
  // VoxelGrid::getCellFromLocation (voxel_grid.hpp:522): floor((loc - origin_minus) * oo_resolution)
  const R fx = (cx - h.om[0]) * h.oo_res, fy = (cy - h.om[1]) * h.oo_res, fz = (cz - h.om[2]) * h.oo_res;
  if (kExact)
  {
    p.gx = p.hx = static_cast<int>(floorR(fx));
    p.gy = p.hy = static_cast<int>(floorR(fy));
    p.gz = p.hz = static_cast<int>(floorR(fz));
  }
  else
  {
    p.gx = static_cast<int>(floorR(fx - h.margin));
    p.hx = static_cast<int>(floorR(fx + h.margin));
    p.gy = static_cast<int>(floorR(fy - h.margin));
    p.hy = static_cast<int>(floorR(fy + h.margin));
    p.gz = static_cast<int>(floorR(fz - h.margin));
    p.hz = static_cast<int>(floorR(fz + h.margin));
    p.risky = (p.gx != p.hx) | (p.gy != p.hy) | (p.gz != p.hz);
  }
  // DistanceField::getDistanceGradient bounds (distance_field.cpp:82): a 1-cell margin is "out of bounds", which
  // reads as max_distance and therefore never collides (types.cpp:229).
  if (!p.risky && p.gx >= 1 && p.gy >= 1 && p.gz >= 1 && p.gx < h.nx - 1 && p.gy < h.ny - 1 && p.gz < h.nz - 1)
    p.idx = (p.gx * h.ny + p.gy) * h.nz + p.gz;  // VoxelGrid::ref (voxel_grid.hpp:425); < 2^31 voxels (host checks)
  return p;
}

// FP32 only, rare: evaluate every candidate cell of a sphere whose voxel is uncertain.
template <typename R, typename FT>
__device__ __noinline__ void riskySphere(const ModelHeader<R>& h, const Probe& p, const FT* __restrict__ world,
                                         const FT* __restrict__ self, bool& hit, bool& amb)
{
  bool any = false, all = true;
  for (int ix = p.gx; ix <= p.hx; ++ix)
    for (int iy = p.gy; iy <= p.hy; ++iy)
      for (int iz = p.gz; iz <= p.hz; ++iz)
      {
        bool c = false;
        if (ix >= 1 && iy >= 1 && iz >= 1 && ix < h.nx - 1 && iy < h.ny - 1 && iz < h.nz - 1)
        {
          const int idx = (ix * h.ny + iy) * h.nz + iz;
          const int dw = p.thr_w ? static_cast<int>(__ldg(world + idx)) : 0x7fffffff;
          const int ds = p.thr_s ? static_cast<int>(__ldg(self + idx)) : 0x7fffffff;
          c = (dw < p.thr_w) | (ds < p.thr_s);
        }
        any |= c;
        all &= c;
      }
  hit |= all;
  amb |= (any && !all);
}

// One state, warp-cooperative.  Returns (via uniform bools) whether it certainly collides / is undecidable in R.
template <typename R, typename FT>
__device__ __forceinline__ void checkState(const Smem<R>& m, const R* T, R* cent, R* bsc, const FT* __restrict__ world,
                                           const FT* __restrict__ self, uint32_t flags, int lane, bool& hit_out,
                                           bool& amb_out)
{
  constexpr bool kExact = sizeof(R) == 8;
  const ModelHeader<R>& h = *m.h;
  const bool do_world = (flags & 1u) != 0, do_self = (flags & 2u) != 0, do_intra = (flags & 4u) != 0;
  bool hit = false, amb = false;
  for (int base = 0; base < h.n_spheres; base += 64)
  {
    // two spheres per lane: both poses, then all four gathers in flight, then the compares
    const Probe p0 = probeSphere<R>(m, T, cent, base + lane, do_world, do_self);
    const Probe p1 = probeSphere<R>(m, T, cent, base + 32 + lane, do_world, do_self);
    int dw0 = 0x7fffffff, ds0 = 0x7fffffff, dw1 = 0x7fffffff, ds1 = 0x7fffffff;
    if (p0.idx >= 0)
    {
      if (p0.thr_w) dw0 = static_cast<int>(__ldg(world + p0.idx));
      if (p0.thr_s) ds0 = static_cast<int>(__ldg(self + p0.idx));
    }
    if (p1.idx >= 0)
    {
      if (p1.thr_w) dw1 = static_cast<int>(__ldg(world + p1.idx));
      if (p1.thr_s) ds1 = static_cast<int>(__ldg(self + p1.idx));
    }
    hit |= (dw0 < p0.thr_w) | (ds0 < p0.thr_s) | (dw1 < p1.thr_w) | (ds1 < p1.thr_s);
    if (!kExact)
    {
      if (p0.risky) riskySphere<R, FT>(h, p0, world, self, hit, amb);
      if (p1.risky) riskySphere<R, FT>(h, p1, world, self, hit, amb);
    }
  }
  hit = __any_sync(kFull, hit);
  if (!hit && do_intra && h.n_pairs > 0)
  {
    // Per link with spheres: the reference's bounding-sphere centre (PosedBodySphereDecomposition::updatePose,
    // types.cpp:391) and the centre of our tight sphere; radii alongside so pair tests read one place.
    for (int b = lane; b < h.n_bs; b += 32)
    {
      const BsRec<R> r = m.bs[b];
      R Tl[12];
      load12(T + r.slot * kLinkStrideWords, Tl);
      R x, y, z;
      posePoint<R>(Tl, r.x, r.y, r.z, x, y, z);
      store4(bsc + b * 8, x, y, z, r.r);
      posePoint<R>(Tl, r.tx, r.ty, r.tz, x, y, z);
      store4(bsc + b * 8 + 4, x, y, z, r.tr);
    }
    __syncwarp();
    bool phit = false;
    for (int p0 = 0; p0 < h.n_pairs; p0 += 32)
    {
      const int p = p0 + lane;
      bool live = false, live_certain = false;
      if (p < h.n_pairs)
      {
        const PairRec pr = m.pairs[p];
        R ca[4], cb[4], ta[4], tb[4];
        load4(bsc + pr.a * 8, ca);
        load4(bsc + pr.b * 8, cb);
        load4(bsc + pr.a * 8 + 4, ta);
        load4(bsc + pr.b * 8 + 4, tb);
        const R ex = ca[0] - cb[0], ey = ca[1] - cb[1], ez = ca[2] - cb[2];
        const R d2 = ex * ex + ey * ey + ez * ez;
        // reference quirk kept: SQUARED centre distance against the plain radius sum (types.cpp:416-417)
        const R rs = ca[3] + cb[3];
        // conservative cull: every sphere of a link lies inside its tight sphere, so separated tight spheres
        // (with slack for rounding) mean no sphere pair of the two links can touch
        const R fx = ta[0] - tb[0], fy = ta[1] - tb[1], fz = ta[2] - tb[2];
        const R t2 = fx * fx + fy * fy + fz * fz;
        const R rt = ta[3] + tb[3] + h.tight_eps;
        const bool near = t2 < rt * rt;
        if (kExact)
          live = live_certain = near && (d2 < rs);
        else
        {
          live = near && (d2 < rs + h.bs_eps);
          live_certain = d2 < rs - h.bs_eps;
        }
      }
      unsigned mask = __ballot_sync(kFull, live);
      const unsigned cmask = __ballot_sync(kFull, live_certain);
      while (mask)
      {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const bool cert = (cmask >> bit) & 1u;
        const PairRec pr = m.pairs[p0 + bit];
        if (pr.na + pr.nb <= 32)
        {
          // level 2: lanes [0, na) hold the spheres of link a and test them against b's tight sphere, lanes
          // [na, na + nb) hold b's spheres and test against a's.  Only survivors can be part of a touching pair.
          R me[4] = { R(0), R(0), R(0), R(0) };
          bool inside = false;
          if (lane < pr.na + pr.nb)
          {
            const bool is_a = lane < pr.na;
            load4(cent + (is_a ? pr.sa0 + lane : pr.sb0 + lane - pr.na) * 4, me);
            R ot[4];
            load4(bsc + (is_a ? pr.b : pr.a) * 8 + 4, ot);
            const R gx = me[0] - ot[0], gy = me[1] - ot[1], gz = me[2] - ot[2];
            const R lim = me[3] + ot[3] + h.tight_eps;
            inside = (gx * gx + gy * gy + gz * gz) < lim * lim;
          }
          const unsigned mk = __ballot_sync(kFull, inside);
          unsigned ma = mk & ((1u << pr.na) - 1u);
          const bool is_b = inside && lane >= pr.na;
          if ((mk >> pr.na) == 0u)
            ma = 0u;
          while (ma)
          {
            const int ia = __ffs(ma) - 1;
            ma &= ma - 1;
            R pa[4];
            load4(cent + (pr.sa0 + ia) * 4, pa);  // broadcast read
            if (is_b)
            {
              const R gx2 = pa[0] - me[0], gy2 = pa[1] - me[1], gz2 = pa[2] - me[2];
              const R dd = gx2 * gx2 + gy2 * gy2 + gz2 * gz2;
              const R rsum = pa[3] + me[3];
              if (kExact)
                phit |= sqrtR(dd) < rsum;  // gradient.norm() < r1 + r2 (collision_env_distance_field.cpp:577-583)
              else
              {
                // squared form of dist < rsum -+ eps (both sides positive); the 1e-7 in pair_eps covers its rounding
                const R lo = rsum - h.pair_eps, hi = rsum + h.pair_eps;
                const bool c_cert = (lo > R(0)) && (dd < lo * lo);
                const bool c_maybe = dd < hi * hi;
                phit |= (cert && c_cert);
                amb |= (c_maybe && !(cert && c_cert));
              }
            }
          }
        }
        else
        {
          // links with more than 32 spheres between them: plain enumeration of the na * nb sphere pairs
          for (int t = lane; t < pr.tot; t += 32)
          {
            const int ia = static_cast<int>((static_cast<unsigned>(t) * pr.magic) >> 16);  // t / nb
            R pa[4], pb[4];
            load4(cent + (pr.sa0 + ia) * 4, pa);
            load4(cent + (pr.sb0 + (t - ia * pr.nb)) * 4, pb);
            const R gx2 = pa[0] - pb[0], gy2 = pa[1] - pb[1], gz2 = pa[2] - pb[2];
            const R dd = gx2 * gx2 + gy2 * gy2 + gz2 * gz2;
            const R rsum = pa[3] + pb[3];
            if (kExact)
              phit |= sqrtR(dd) < rsum;
            else
            {
              const R lo = rsum - h.pair_eps, hi = rsum + h.pair_eps;
              const bool c_cert = (lo > R(0)) && (dd < lo * lo);
              const bool c_maybe = dd < hi * hi;
              phit |= (cert && c_cert);
              amb |= (c_maybe && !(cert && c_cert));
            }
          }
        }
      }
    }
    hit = __any_sync(kFull, phit);
  }
  amb = __any_sync(kFull, amb);
  hit_out = hit;
  amb_out = amb && !hit;
}

// Per-warp shared memory, in units of R (all 16-byte aligned): sphere centres + radius | link bounding / tight
// centres + radii | transforms of kTile states
template <typename R>
__device__ __forceinline__ int perWarpWords(const ModelHeader<R>& h, int tile)
{
  return 4 * h.n_spheres + 8 * h.n_bs + tile * h.state_stride;
}

// Dynamic shared memory: [model blob][per warp: cent | bsc | Ts]
template <typename R, typename FT, bool kFromList, int kTile>
__global__ void __launch_bounds__(512, 1) checkKernel(CheckArgs<FT> a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<R> m = loadModel<R>(a.model, a.model_bytes, smem);
  const ModelHeader<R>& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  R* cent = reinterpret_cast<R*>(smem + h.total_bytes) + static_cast<size_t>(warp) * perWarpWords<R>(h, kTile);
  R* bsc = cent + 4 * h.n_spheres;
  R* Ts = bsc + 8 * h.n_bs;

  unsigned long long n = a.n_states;
  if (kFromList)
    n = *a.list_count < a.list_cap ? *a.list_count : a.list_cap;
  // kFromList runs with kTile = 1: the recheck list is short, and one state per warp keeps its latency at one FK +
  // one check instead of a tile of them in a row.
  const unsigned long long n_tiles = (n + kTile - 1) / kTile;
  for (unsigned long long tile = static_cast<unsigned long long>(blockIdx.x) * warps + warp; tile < n_tiles;
       tile += static_cast<unsigned long long>(gridDim.x) * warps)
  {
    const unsigned long long first = tile * kTile;
    const int cnt = static_cast<int>(n - first < kTile ? n - first : kTile);
    unsigned long long my_state = first + lane;
    if (kFromList && lane < cnt)
      my_state = a.list[first + lane];
    if (lane < cnt)
      fkState<R>(m, a.q + my_state * h.n_group_vars, Ts + lane * h.state_stride);
    __syncwarp();
    uint32_t valid_word = 0, flag_word = 0;
    for (int s = 0; s < cnt; ++s)
    {
      bool hit, amb;
      checkState<R, FT>(m, Ts + s * h.state_stride, cent, bsc, a.world, a.self, a.flags, lane, hit, amb);
      __syncwarp();
      if (amb)
        flag_word |= 1u << s;
      else if (!hit)
        valid_word |= 1u << s;
    }
    if (kFromList)
    {
      // patch the bits of the rechecked states (the fast pass left them 0)
      if (lane < cnt && ((valid_word >> lane) & 1u))
        atomicOr(a.bitmap + (my_state >> 5), 1u << (my_state & 31));
    }
    else
    {
      if (lane == 0)
      {
        if (kTile == 32)
          a.bitmap[tile] = valid_word;
        else if (kTile == 16)
          reinterpret_cast<uint16_t*>(a.bitmap)[tile] = static_cast<uint16_t>(valid_word);
        else
          reinterpret_cast<uint8_t*>(a.bitmap)[tile] = static_cast<uint8_t>(valid_word);
      }
      if (flag_word)
      {
        const int c = __popc(flag_word);
        unsigned base = 0;
        if (lane == 0)
          base = atomicAdd(a.flag_count, static_cast<unsigned>(c));
        base = __shfl_sync(kFull, base, 0);
        if ((flag_word >> lane) & 1u)
        {
          const unsigned rank = __popc(flag_word & ((1u << lane) - 1u));
          if (base + rank < a.list_cap)
            a.flag_list[base + rank] = static_cast<uint32_t>(first + lane);
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
    const int per_state = a.n_model_links * 16;
    for (int i = lane; i < cnt * per_state; i += 32)
    {
      const int s = i / per_state, rem = i - s * per_state;
      const int l = rem >> 4, e = rem & 15, c = e >> 2, r = e & 3;
      double v;
      const int slot = a.link_slot[l];
      if (r == 3)
        v = c == 3 ? 1.0 : 0.0;
      else if (slot >= 0)
        v = static_cast<double>(Ts[s * h.state_stride + slot * kLinkStrideWords + r * 4 + c]);
      else
        v = a.const_T[l * 16 + e];
      a.out[(first + s) * per_state + rem] = v;
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

// Random 32-byte-sector gather: every thread reads `per_thread` pseudo-random sectors of the field footprint.
__global__ void gatherKernel(const uint8_t* __restrict__ field, unsigned long long n_sectors, int per_thread,
                             unsigned* sink)
{
  unsigned long long x = (static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ull +
                         0xD1B54A32D192ED03ull;
  unsigned acc = 0;
#pragma unroll 4
  for (int i = 0; i < per_thread; ++i)
  {
    x ^= x >> 12;
    x ^= x << 25;
    x ^= x >> 27;
    const unsigned long long r = x * 0x2545F4914F6CDD1Dull;
    const unsigned long long sec = (r >> 11) % n_sectors;
    acc += __ldg(field + sec * 32 + (r & 31));
  }
  if (acc == 0xffffffffu)
    *sink = acc;
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

template <typename R, typename FT, bool kList, int kTile>
cudaError_t launchOne(const CheckArgs<FT>& a, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  static SmemGrant g;
  cudaError_t e = ensureSmem(checkKernel<R, FT, kList, kTile>, smem_bytes, g);
  if (e != cudaSuccess)
    return e;
  checkKernel<R, FT, kList, kTile><<<grid, block, smem_bytes, stream>>>(a);
  return cudaGetLastError();
}

template <typename R, typename FT, bool kList, int kTile>
int regsOf()
{
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, checkKernel<R, FT, kList, kTile>) != cudaSuccess)
    return 128;
  return fa.numRegs;
}
}  // namespace

template <typename FT>
cudaError_t launchCheck(const CheckArgs<FT>& a, bool fp64, bool from_list, int tile, int grid, int block,
                        size_t smem_bytes, cudaStream_t stream)
{
  if (fp64 && from_list)
    return launchOne<double, FT, true, 1>(a, grid, block, smem_bytes, stream);
  if (fp64)
    return tile == 32 ? launchOne<double, FT, false, 32>(a, grid, block, smem_bytes, stream) :
           tile == 16 ? launchOne<double, FT, false, 16>(a, grid, block, smem_bytes, stream) :
                        launchOne<double, FT, false, 8>(a, grid, block, smem_bytes, stream);
  return tile == 32 ? launchOne<float, FT, false, 32>(a, grid, block, smem_bytes, stream) :
         tile == 16 ? launchOne<float, FT, false, 16>(a, grid, block, smem_bytes, stream) :
                      launchOne<float, FT, false, 8>(a, grid, block, smem_bytes, stream);
}
template cudaError_t launchCheck<uint8_t>(const CheckArgs<uint8_t>&, bool, bool, int, int, int, size_t, cudaStream_t);
template cudaError_t launchCheck<uint16_t>(const CheckArgs<uint16_t>&, bool, bool, int, int, int, size_t, cudaStream_t);

int checkKernelRegs(bool fp64, bool from_list, int tile, int field_bytes)
{
  if (field_bytes == 1)
  {
    if (fp64 && from_list) return regsOf<double, uint8_t, true, 1>();
    if (fp64) return tile == 32 ? regsOf<double, uint8_t, false, 32>() : tile == 16 ? regsOf<double, uint8_t, false, 16>() : regsOf<double, uint8_t, false, 8>();
    return tile == 32 ? regsOf<float, uint8_t, false, 32>() : tile == 16 ? regsOf<float, uint8_t, false, 16>() : regsOf<float, uint8_t, false, 8>();
  }
  if (fp64 && from_list) return regsOf<double, uint16_t, true, 1>();
  if (fp64) return tile == 32 ? regsOf<double, uint16_t, false, 32>() : tile == 16 ? regsOf<double, uint16_t, false, 16>() : regsOf<double, uint16_t, false, 8>();
  return tile == 32 ? regsOf<float, uint16_t, false, 32>() : tile == 16 ? regsOf<float, uint16_t, false, 16>() : regsOf<float, uint16_t, false, 8>();
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

cudaError_t launchGather(const uint8_t* field, unsigned long long n_sectors, int per_thread, int grid, int block,
                         unsigned* sink, cudaStream_t stream)
{
  gatherKernel<<<grid, block, 0, stream>>>(field, n_sectors, per_thread, sink);
  return cudaGetLastError();
}
}  // namespace b200sv
