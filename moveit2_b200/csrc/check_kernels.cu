// K1 + K2: batched forward kinematics and sphere-decomposition collision checks, fused.
//
// Replaces, per state (reference paths relative to moveit_core/):
//   RobotState::setJointGroupPositions / updateMimicJoints / updateLinkTransformsInternal
//                                                  robot_state/src/robot_state.cpp:571-584, 210-220, 793-848
//   {Revolute,Prismatic,Planar,Floating,Fixed}JointModel::computeTransform
//                                                  robot_model/src/*_joint_model.cpp
//   PosedBodySphereDecomposition::updatePose       collision_distance_field/src/collision_distance_field_types.cpp:388-395
//   getCollisionSphereCollision + DistanceField::getDistanceGradient (centre voxel only: the gradient never
//   influences the boolean)                        types.cpp:211-236, distance_field/src/distance_field.cpp:73-97
//   getEnvironmentCollisions / getSelfCollisions / getIntraGroupCollisions + doBoundingSpheresIntersect
//                                                  collision_env_distance_field.cpp:1561-1643, 274-353, 430-642; types.cpp:408-418
//
// Mapping (B200, sm_100a): persistent CTAs, one per SM slot.  Each warp owns a tile of 32 consecutive states.
//   phase 1  thread-per-state FK: lane s walks the device links in index order (parents first); link transforms
//            go to shared memory as 3x4 row-major, state stride odd, link stride 13 => conflict-free for both the
//            per-lane writes of phase 1 and the per-link reads of phase 2.
//   phase 2  warp-per-state checks: lanes own spheres; each lane poses its sphere, converts the centre to a voxel
//            and gathers ONE byte from the world field and one from the self field (L2-resident compact d^2),
//            compares against the sphere's integer threshold; early exit is warp-uniform.  Intra-group: lanes
//            test link pairs' bounding spheres, the surviving pairs are expanded into sphere pairs over lanes.
// Precision: the fast pass is FP32 (template R = float).  A sphere whose voxel coordinate is within `margin` of a
// cell face, or a pair whose distance is within eps of its threshold, is evaluated for BOTH outcomes; if the
// state's boolean could depend on that, the state index is appended to a recheck list and the exact pass
// (R = double, same code) recomputes it.  The FP64 pass restates the reference's double arithmetic
// operation by operation.
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

template <typename R>
__device__ __forceinline__ void sincosR(R x, R* s, R* c);
template <>
__device__ __forceinline__ void sincosR<float>(float x, float* s, float* c)
{
  sincosf(x, s, c);
}
template <>
__device__ __forceinline__ void sincosR<double>(double x, double* s, double* c)
{
  sincos(x, s, c);
}
template <typename R>
__device__ __forceinline__ R sqrtR(R x);
template <>
__device__ __forceinline__ float sqrtR<float>(float x)
{
  return sqrtf(x);
}
template <>
__device__ __forceinline__ double sqrtR<double>(double x)
{
  return sqrt(x);
}
template <typename R>
__device__ __forceinline__ R floorR(R x);
template <>
__device__ __forceinline__ float floorR<float>(float x)
{
  return floorf(x);
}
template <>
__device__ __forceinline__ double floorR<double>(double x)
{
  return floor(x);
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
    R A[12];
    if (L.parent >= 0)
    {
      R P[12];
      const R* src = T + L.parent * kLinkStrideWords;
#pragma unroll
      for (int i = 0; i < 12; ++i)
        P[i] = src[i];
      if (L.has_origin)
        mulAffine<R>(P, L.o, A);
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
        A[i] = L.o[i];
    }
    R O[12];
    switch (L.type)
    {
      case REVOLUTE:
      {  // revolute_joint_model.cpp:250-287
        const R q = varValue<R>(m.vars[L.var], qrow);
        R s, c;
        sincosR<R>(q, &s, &c);
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
        sincosR<R>(th, &s, &c);
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
        const R n = sqrtR<R>(x * x + y * y + z * z + w * w);
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
    R* dst = T + l * kLinkStrideWords;
#pragma unroll
    for (int i = 0; i < 12; ++i)
      dst[i] = O[i];
  }
}

template <typename R>
__device__ __forceinline__ void posePoint(const R* T, R x, R y, R z, R& ox, R& oy, R& oz)
{  // Isometry3d * Vector3d: linear * p + translation
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

template <typename FT>
__device__ __forceinline__ int ldField(const FT* __restrict__ f, long idx)
{
  return static_cast<int>(__ldg(f + idx));
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
  const int nx = h.nx, ny = h.ny, nz = h.nz;
  for (int base = 0; base < h.n_spheres; base += 32)
  {
    const int k = base + lane;
    if (k < h.n_spheres)
    {
      const SphereRec<R> sp = m.spheres[k];
      R cx, cy, cz;
      posePoint<R>(T + sp.slot * kLinkStrideWords, sp.x, sp.y, sp.z, cx, cy, cz);
      cent[k * 4 + 0] = cx;
      cent[k * 4 + 1] = cy;
      cent[k * 4 + 2] = cz;
      cent[k * 4 + 3] = sp.r;
      const int thr_w = do_world ? sp.thr : 0;
      const int thr_s = do_self ? sp.thr_self : 0;
      if (thr_w | thr_s)
      {
        // VoxelGrid::getCellFromLocation (voxel_grid.hpp:522): floor((loc - origin_minus) * oo_resolution)
        const R fx = (cx - h.om[0]) * h.oo_res, fy = (cy - h.om[1]) * h.oo_res, fz = (cz - h.om[2]) * h.oo_res;
        int gx, gy, gz, hx, hy, hz;
        if (kExact)
        {
          gx = hx = static_cast<int>(floorR<R>(fx));
          gy = hy = static_cast<int>(floorR<R>(fy));
          gz = hz = static_cast<int>(floorR<R>(fz));
        }
        else
        {
          gx = static_cast<int>(floorR<R>(fx - h.margin));
          hx = static_cast<int>(floorR<R>(fx + h.margin));
          gy = static_cast<int>(floorR<R>(fy - h.margin));
          hy = static_cast<int>(floorR<R>(fy + h.margin));
          gz = static_cast<int>(floorR<R>(fz - h.margin));
          hz = static_cast<int>(floorR<R>(fz + h.margin));
        }
        if (kExact || (gx == hx && gy == hy && gz == hz))
        {
          // DistanceField::getDistanceGradient bounds (distance_field.cpp:82): a 1-cell margin is "out of bounds",
          // which reads as max_distance and therefore never collides (types.cpp:229).
          if (gx >= 1 && gy >= 1 && gz >= 1 && gx < nx - 1 && gy < ny - 1 && gz < nz - 1)
          {
            const long idx = (static_cast<long>(gx) * ny + gy) * nz + gz;  // VoxelGrid::ref (voxel_grid.hpp:425)
            const int dw = thr_w ? ldField<FT>(world, idx) : 0x7fffffff;
            const int ds = thr_s ? ldField<FT>(self, idx) : 0x7fffffff;
            hit |= (dw < thr_w) | (ds < thr_s);
          }
        }
        else
        {
          // FP32 only: the centre is within `margin` of a cell face.  Evaluate every candidate cell.
          bool any = false, all = true;
          for (int ix = gx; ix <= hx; ++ix)
            for (int iy = gy; iy <= hy; ++iy)
              for (int iz = gz; iz <= hz; ++iz)
              {
                bool c = false;
                if (ix >= 1 && iy >= 1 && iz >= 1 && ix < nx - 1 && iy < ny - 1 && iz < nz - 1)
                {
                  const long idx = (static_cast<long>(ix) * ny + iy) * nz + iz;
                  const int dw = thr_w ? ldField<FT>(world, idx) : 0x7fffffff;
                  const int ds = thr_s ? ldField<FT>(self, idx) : 0x7fffffff;
                  c = (dw < thr_w) | (ds < thr_s);
                }
                any |= c;
                all &= c;
              }
          hit |= all;
          amb |= (any && !all);
        }
      }
    }
  }
  hit = __any_sync(kFull, hit);
  if (!hit && do_intra && h.n_pairs > 0)
  {
    // bounding-sphere centres of the links (PosedBodySphereDecomposition::updatePose, types.cpp:391) and the
    // centres of our tight spheres
    for (int b = lane; b < h.n_bs; b += 32)
    {
      const BsRec<R> r = m.bs[b];
      const R* Tl = T + r.slot * kLinkStrideWords;
      R x, y, z;
      posePoint<R>(Tl, r.x, r.y, r.z, x, y, z);
      bsc[b * 6 + 0] = x;
      bsc[b * 6 + 1] = y;
      bsc[b * 6 + 2] = z;
      posePoint<R>(Tl, r.tx, r.ty, r.tz, x, y, z);
      bsc[b * 6 + 3] = x;
      bsc[b * 6 + 4] = y;
      bsc[b * 6 + 5] = z;
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
        const R* ca = bsc + pr.a * 6;
        const R* cb = bsc + pr.b * 6;
        const R ex = ca[0] - cb[0], ey = ca[1] - cb[1], ez = ca[2] - cb[2];
        const R d2 = ex * ex + ey * ey + ez * ez;
        // reference quirk kept: SQUARED centre distance against the plain radius sum (types.cpp:416-417)
        const R rs = m.bs[pr.a].r + m.bs[pr.b].r;
        // conservative cull: every sphere of a link lies inside its tight sphere, so separated tight spheres
        // (with slack for rounding) mean no sphere pair of the two links can touch
        const R fx = ca[3] - cb[3], fy = ca[4] - cb[4], fz = ca[5] - cb[5];
        const R t2 = fx * fx + fy * fy + fz * fz;
        const R rt = m.bs[pr.a].tr + m.bs[pr.b].tr + h.tight_eps;
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
        for (int t = lane; t < pr.tot; t += 32)
        {
          const int ia = static_cast<int>((static_cast<unsigned>(t) * pr.magic) >> 16);  // t / nb (range checked on the host)
          const R* pa = cent + (pr.sa0 + ia) * 4;
          const R* pb = cent + (pr.sb0 + (t - ia * pr.nb)) * 4;
          const R gx2 = pa[0] - pb[0], gy2 = pa[1] - pb[1], gz2 = pa[2] - pb[2];
          const R dd = gx2 * gx2 + gy2 * gy2 + gz2 * gz2;
          const R rsum = pa[3] + pb[3];
          if (kExact)
            phit |= sqrtR<R>(dd) < rsum;  // gradient.norm() < r1 + r2 (collision_env_distance_field.cpp:577-583)
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
    hit = __any_sync(kFull, phit);
  }
  amb = __any_sync(kFull, amb);
  hit_out = hit;
  amb_out = amb && !hit;
}

// Dynamic shared memory: [model blob][per warp: transforms 32 * state_stride | sphere centres | bounding centres]
template <typename R, typename FT, bool kFromList>
__global__ void __launch_bounds__(512, 1) checkKernel(CheckArgs<FT> a)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const Smem<R> m = loadModel<R>(a.model, a.model_bytes, smem);
  const ModelHeader<R>& h = *m.h;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const int per_warp = 32 * h.state_stride + 4 * h.n_spheres + 6 * h.n_bs + 4;
  R* wbase = reinterpret_cast<R*>(smem + ((h.total_bytes + 15) & ~15)) + static_cast<size_t>(warp) * per_warp;
  R* cent = wbase;
  R* bsc = cent + 4 * h.n_spheres;
  R* Ts = bsc + 6 * h.n_bs;

  unsigned long long n = a.n_states;
  if (kFromList)
    n = *a.list_count < a.list_cap ? *a.list_count : a.list_cap;
  // The recheck list is short: one state per warp keeps its latency at one FK + one check instead of 32 in a row.
  constexpr int kTile = kFromList ? 1 : 32;
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
        a.bitmap[tile] = valid_word;
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
  R* Ts = reinterpret_cast<R*>(smem + ((h.total_bytes + 15) & ~15)) + static_cast<size_t>(warp) * 32 * h.state_stride;
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
  double* Ts = reinterpret_cast<double*>(smem + ((h.total_bytes + 15) & ~15));
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
}  // namespace

// cudaFuncSetAttribute is a host-side driver call: do it when the requirement grows, not on every launch.
template <typename K>
cudaError_t ensureSmem(K kernel, size_t smem_bytes, size_t& granted)
{
  if (smem_bytes <= granted)
    return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
  if (e == cudaSuccess)
    granted = smem_bytes;
  return e;
}

template <typename FT>
cudaError_t launchCheck(const CheckArgs<FT>& a, bool fp64, bool from_list, int grid, int block, size_t smem_bytes,
                        cudaStream_t stream)
{
  static size_t g_dl = 0, g_d = 0, g_f = 0;  // per FT instantiation
  cudaError_t e;
  if (fp64 && from_list)
  {
    if ((e = ensureSmem(checkKernel<double, FT, true>, smem_bytes, g_dl)) != cudaSuccess)
      return e;
    checkKernel<double, FT, true><<<grid, block, smem_bytes, stream>>>(a);
  }
  else if (fp64)
  {
    if ((e = ensureSmem(checkKernel<double, FT, false>, smem_bytes, g_d)) != cudaSuccess)
      return e;
    checkKernel<double, FT, false><<<grid, block, smem_bytes, stream>>>(a);
  }
  else
  {
    if ((e = ensureSmem(checkKernel<float, FT, false>, smem_bytes, g_f)) != cudaSuccess)
      return e;
    checkKernel<float, FT, false><<<grid, block, smem_bytes, stream>>>(a);
  }
  return cudaGetLastError();
}
template cudaError_t launchCheck<uint8_t>(const CheckArgs<uint8_t>&, bool, bool, int, int, size_t, cudaStream_t);
template cudaError_t launchCheck<uint16_t>(const CheckArgs<uint16_t>&, bool, bool, int, int, size_t, cudaStream_t);

cudaError_t launchFk(const FkArgs& a, bool fp64, int grid, int block, size_t smem_bytes, cudaStream_t stream)
{
  cudaError_t e;
  if (fp64)
  {
    e = cudaFuncSetAttribute(fkKernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess)
      return e;
    fkKernel<double><<<grid, block, smem_bytes, stream>>>(a);
  }
  else
  {
    e = cudaFuncSetAttribute(fkKernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess)
      return e;
    fkKernel<float><<<grid, block, smem_bytes, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launchSphereCenters(const FkArgs& a, double* d_centers, size_t smem_bytes, cudaStream_t stream)
{
  cudaError_t e = cudaFuncSetAttribute(sphereCentersKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
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
