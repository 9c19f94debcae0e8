// K3 + K4: point cloud -> occupancy bitmap -> exact truncated squared Euclidean distance transform.
//
// Replaces (reference: moveit_core/distance_field/src/propagation_distance_field.cpp)
//   addPointsToField :177-196 (worldToGrid per point, voxel_grid.hpp:508-560)           -> scatterKernel
//   addNewObstacleVoxels :221-249 + propagatePositive :390-444 (serial bucket queue)     -> edtZYKernel + edtXKernel
//   reset :506-524                                                                        -> fillKernel
// Output semantics: d2[v] = min(true squared cell distance to the nearest occupied voxel, max_distance_sq_) --
// what the reference's propagation computes whenever it is exact (it is an approximation that never
// under-estimates and is insertion-order dependent in sparse scenes).
//
// Separable, because the transform is truncated at R = ceil(max_distance / resolution) cells:
//   z pass  nearest occupied voxel along z within +-R, straight from the occupancy BITS (ffs/clz on 32-bit words)
//   y pass  min over |dy| <= R of g_z(y+dy)^2 + dy^2, scanning outward and stopping as soon as dy^2 >= best
//   x pass  same along x on the 16-bit intermediate
// The z and y passes are fused per x-plane in shared memory (one CTA per plane slab); only the 2-byte
// intermediate and the outputs touch HBM.  Integer arithmetic throughout: results are exact.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "kernels.hpp"

namespace b200sv
{
namespace
{
// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may become
// resident while the kernel before it in the stream drains -- its CTAs run their prologue and then block in depWait()
// until that kernel has completed and its writes are visible.  Nothing of the context's global state may be read OR
// written before depWait().  depLaunch() lets the NEXT kernel's CTAs in once every CTA of this one has called it.
// Both are no-ops for a kernel launched the ordinary way.
__device__ __forceinline__ void depWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void depLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__global__ void __launch_bounds__(256) scatterKernel(const double* __restrict__ xyz, size_t n, GridDev g,
                                                     uint32_t* __restrict__ occ, int set)
{
  // A warp takes 32 consecutive points = 96 consecutive doubles: three coalesced 256-byte loads into shared memory, then
  // every lane picks up its own point (stride 3 doubles: conflict-free).  Reading xyz[3 i + k] straight from global memory
  // touched every sector three times.
  __shared__ double sh[8][96];
  depLaunch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t n_tiles = (n + 31) / 32;
  const size_t stride = static_cast<size_t>(gridDim.x) * 8;
  // The tile after this one is already on its way from HBM while this one is turned into cells and atomics: the REDs of a
  // warp occupy the load/store pipe for ~40 cycles each (32 spread addresses), and without the prefetch every iteration
  // also waited a full DRAM latency for its points (ncu, round 1: 77 % of the stall samples sat on this load).
  auto loadTile = [&](size_t t, double* v) {
    const double* src = xyz + 96 * t;
    const size_t left = 3 * n - 96 * t;
    const int avail = left < 96 ? static_cast<int>(left) : 96;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      v[j] = (j * 32 + lane < avail) ? __ldcs(src + j * 32 + lane) : 0.0;  // streamed: read exactly once
  };
  double cur[3], nxt[3] = { 0.0, 0.0, 0.0 };
  size_t t = blockIdx.x * static_cast<size_t>(8) + warp;
  if (t < n_tiles)
    loadTile(t, cur);
  for (; t < n_tiles; t += stride)
  {
    if (t + stride < n_tiles)
      loadTile(t + stride, nxt);
    const size_t left = 3 * n - 96 * t;
    const int avail = left < 96 ? static_cast<int>(left) : 96;
#pragma unroll
    for (int j = 0; j < 3; ++j)
      sh[warp][j * 32 + lane] = cur[j];
    __syncwarp();
    const bool have = 3 * lane < avail;
    const double px = sh[warp][3 * lane], py = sh[warp][3 * lane + 1], pz = sh[warp][3 * lane + 2];
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 3; ++j)
      cur[j] = nxt[j];
    if (!have)
      continue;
    // VoxelGrid::getCellFromLocation (voxel_grid.hpp:522), in double exactly as the reference
    const int x = static_cast<int>(floor((px - g.om[0]) * g.oo));
    const int y = static_cast<int>(floor((py - g.om[1]) * g.oo));
    const int z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;  // isCellValid (voxel_grid.hpp:405-408)
    uint32_t* w = occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5);
    if (set)
      atomicOr(w, 1u << (z & 31));
    else
      atomicAnd(w, ~(1u << (z & 31)));
  }
}

// ---- 1-D bulk copies global -> shared memory through the TMA engine, completion on an mbarrier (sm_90+ PTX) ----------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void bulkLoad(void* dst, const void* src, unsigned bytes, uint64_t* bar)
{
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

// scatterKernel with the point tiles staged by the TMA engine: every warp owns kScatterStages buffers of one tile (32 points =
// 768 bytes) and keeps kScatterStages - 1 bulk copies in flight, so the atomics of a tile never wait for its points (the
// register prefetch of scatterKernel covers one tile: ncu still showed long_scoreboard as its first stall).  Whole tiles of
// a 16-byte aligned cloud only; the ragged last tile is read with plain loads by the warp whose turn it is.
constexpr int kScatterStages = 4;
__global__ void __launch_bounds__(256) scatterKernelTma(const double* __restrict__ xyz, size_t n, GridDev g,
                                                        uint32_t* __restrict__ occ, int set)
{
  __shared__ __align__(16) double sh[8][kScatterStages][96];
  __shared__ uint64_t bars[8][kScatterStages];
  depLaunch();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t n_tiles = (n + 31) / 32, n_full = n / 32;
  const size_t stride = static_cast<size_t>(gridDim.x) * 8;
  const size_t t0 = blockIdx.x * static_cast<size_t>(8) + warp;
  if (lane == 0)
  {
    for (int b = 0; b < kScatterStages; ++b)
      mbarInit(&bars[warp][b], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto fetch = [&](size_t it) {  // lane 0 only
    const size_t t = t0 + it * stride;
    if (t < n_full)
      bulkLoad(sh[warp][it % kScatterStages], xyz + 96 * t, 768u, &bars[warp][it % kScatterStages]);
  };
  if (lane == 0)
    for (int it = 0; it < kScatterStages - 1; ++it)
      fetch(it);
  size_t it = 0;
  for (size_t t = t0; t < n_tiles; t += stride, ++it)
  {
    const int b = static_cast<int>(it % kScatterStages);
    if (lane == 0)
    {
      // the buffer about to be refilled was read with ordinary loads an iteration ago: order those before the async writes
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      fetch(it + kScatterStages - 1);
    }
    double px, py, pz;
    bool have = true;
    if (t < n_full)
    {
      mbarWait(&bars[warp][b], static_cast<unsigned>((it / kScatterStages) & 1));
      px = sh[warp][b][3 * lane];
      py = sh[warp][b][3 * lane + 1];
      pz = sh[warp][b][3 * lane + 2];
    }
    else
    {
      const size_t i = 32 * t + lane;
      have = i < n;
      px = have ? xyz[3 * i] : 0.0;
      py = have ? xyz[3 * i + 1] : 0.0;
      pz = have ? xyz[3 * i + 2] : 0.0;
    }
    __syncwarp();
    if (!have)
      continue;
    const int x = static_cast<int>(floor((px - g.om[0]) * g.oo));
    const int y = static_cast<int>(floor((py - g.om[1]) * g.oo));
    const int z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;
    uint32_t* w = occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5);
    if (set)
      atomicOr(w, 1u << (z & 31));
    else
      atomicAnd(w, ~(1u << (z & 31)));
  }
}

// i / d by multiply-high with magic = ceil(2^32 / d); d = 1 has no 32-bit magic (2^32): the launcher passes 0 for it
__device__ __forceinline__ unsigned divByMagic(unsigned i, unsigned magic) { return magic ? __umulhi(i, magic) : i; }

// Tables of the z pass (see edtZYKernel): 256 + 2 (R + 2) rows of 8 uint16 squared distances, a function of R alone.
__global__ void edtLutKernel(int R, uint4* __restrict__ lut)
{
  const int cap = R + 1;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < 256 + 2 * (cap + 1); t += gridDim.x * blockDim.x)
  {
    uint32_t v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j)
    {
      int d;
      if (t < 256)
      {  // nearest set bit of pattern t to bit j
        const uint32_t lo = static_cast<uint32_t>(t) & ((2u << j) - 1u), hi = static_cast<uint32_t>(t) >> j;
        d = min(lo ? j - 31 + __clz(lo) : cap, hi ? __ffs(hi) - 1 : cap);
      }
      else if (t < 256 + cap + 1)
        d = (t - 256) + j;        // below[dL]
      else
        d = (t - 256 - cap - 1) + 7 - j;  // above[dR]
      d = min(d, cap);
      v[j] = static_cast<uint32_t>(d * d);
    }
    lut[t] = make_uint4(v[0] | (v[1] << 16), v[2] | (v[3] << 16), v[4] | (v[5] << 16), v[6] | (v[7] << 16));
  }
}

// One CTA per (x plane, y chunk).  Shared memory: occupancy words of rows [y0-R, y1+R), then their SQUARED z distances
// as uint16 (row stride nzp = nz rounded up to 8: a thread takes 8 consecutive z with one 128-bit access).  Squares,
// because the y pass then is two packed 16-bit instructions per candidate row and voxel pair: VIMNMX.U16x2 (min of the
// rows k below and above) and VIADDMNMX.U16x2 (min(best, m + k^2)), both native on sm_100a.  "No occupied voxel within
// R" is (R+1)^2: above max_d2, and small enough that adding k^2 <= R^2 stays below 2^16 (the launcher checks R <= 180).
// Intermediate values above max_d2 mean "nothing within range"; the x pass treats them so.
// `invert`: transform of the COMPLEMENT (distance to the nearest free voxel of the grid = negative_distance_square_ of the
// signed field, propagateNegative :446-504); cells outside the grid are not free.
__global__ void __launch_bounds__(1024) edtZYKernel(GridDev g, const uint32_t* __restrict__ occ, uint16_t* __restrict__ out,
                                                  int chunk, int n_chunks, unsigned nz8_magic, int invert, uint8_t* own_busy,
                                                  const uint4* __restrict__ lut)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const int x = blockIdx.x / n_chunks, ci = blockIdx.x - x * n_chunks;
  if (own_busy != nullptr && ci == 0 && threadIdx.x == 0)
    own_busy[static_cast<size_t>(x) * kPlaneBusyStride] = 0;  // the x pass (next launch) raises it again where the plane is not free
  const int y0 = ci * chunk, y1 = min(g.ny, y0 + chunk);
  const int ya = y0 - g.R, rows = (y1 - y0) + 2 * g.R;  // haloed row range [ya, ya + rows)
  const int nzp = (g.nz + 7) & ~7;
  const int wz = g.wz;
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem);
  uint16_t* gz = reinterpret_cast<uint16_t*>(smem + ((static_cast<size_t>(rows) * wz * 4 + 15) & ~static_cast<size_t>(15)));
  // tables of the z pass (see below), behind the slab: 256 + 2 (R + 2) rows of 8 uint16
  uint4* lut_within = reinterpret_cast<uint4*>(gz + static_cast<size_t>(chunk + 2 * g.R) * nzp);
  uint4* lut_below = lut_within + 256;
  uint4* lut_above = lut_below + (g.R + 2);
  uint32_t* ends = reinterpret_cast<uint32_t*>(lut_above + (g.R + 2));  // [rows * wz]: see the z pass, step (a)
  // the tables depend on R only: built once per (device, R) by edtLutKernel, copied in here (5 KB, L2-resident)
  for (int t = threadIdx.x; t < 256 + 2 * (g.R + 2); t += blockDim.x)
    lut_within[t] = lut[t];
  {
    const uint32_t* src = occ + (static_cast<size_t>(x) * g.ny + ya) * wz;  // rows are contiguous in the bitmap
    const int lo = max(0, -ya) * wz, hi = min(rows, g.ny - ya) * wz;
    if (!invert)
      for (int i = threadIdx.x; i < rows * wz; i += blockDim.x)
        bits[i] = (i >= lo && i < hi) ? src[i] : 0u;
    else
    {
      const uint32_t last = (g.nz & 31) ? (1u << (g.nz & 31)) - 1u : 0xffffffffu;  // bits of the last word inside the grid
      for (int i = threadIdx.x; i < rows * wz; i += blockDim.x)
        bits[i] = (i >= lo && i < hi) ? (~src[i] & ((i % wz == wz - 1) ? last : 0xffffffffu)) : 0u;
    }
  }
  __syncthreads();
  // z pass, one thread per occupancy BYTE (8 voxels), by table: squaring is monotonic, so the squared distance of a voxel
  // is the min of three squared candidates, each one 128-bit table row of 8 uint16 --
  //   within[byte pattern]   nearest set bit inside the byte
  //   below[dL]              nearest set bit below the byte, dL = its distance from the byte's bit 0 (clz on the bits below)
  //   above[dR]              nearest set bit above the byte, dR = its distance from the byte's bit 7 (ffs on the bits above)
  // -- i.e. 3 LDS.128 + 8 VIMNMX.U16x2 + one STS.128 per 8 voxels instead of two 32-step sweeps per word.
  const int cap = g.R + 1;
  // (a) per word: distance of its bit 0 / bit 31 to the nearest set bit in the words below / above (capped)
  for (int i = threadIdx.x; i < rows * wz; i += blockDim.x)
  {
    const int r = i / wz, w = i - r * wz;
    const uint32_t* row = bits + r * wz;
    int dl = cap, dr = cap;
    for (int ww = w - 1, off = 1; ww >= 0 && off <= g.R; --ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        dl = min(cap, off + __clz(v));
        break;
      }
    }
    for (int ww = w + 1, off = 1; ww < wz && off <= g.R; ++ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        dr = min(cap, off + __ffs(v) - 1);
        break;
      }
    }
    ends[i] = static_cast<uint32_t>(dl) | (static_cast<uint32_t>(dr) << 16);
  }
  __syncthreads();
  // (b) per byte
  {
    const int nzb = nzp >> 3;
    for (int i = threadIdx.x; i < rows * nzb; i += blockDim.x)
    {
      const int r = static_cast<int>(divByMagic(static_cast<unsigned>(i), nz8_magic));  // i / nzb
      const int kb = i - r * nzb;
      const int w = kb >> 2, o = (kb & 3) * 8;
      const uint32_t cur = bits[r * wz + w], e = ends[r * wz + w];
      const uint32_t low = o ? (cur & ((1u << o) - 1u)) : 0u, high = o < 24 ? (cur >> (o + 8)) : 0u;
      const int dl = low ? o - 31 + __clz(low) : o + static_cast<int>(e & 0xffffu);
      const int dr = high ? __ffs(high) : 24 - o + static_cast<int>(e >> 16);
      const uint4 a = lut_within[(cur >> o) & 0xffu], b = lut_below[min(dl, cap)], c = lut_above[min(dr, cap)];
      uint4 q;
      q.x = __vminu2(__vminu2(a.x, b.x), c.x);
      q.y = __vminu2(__vminu2(a.y, b.y), c.y);
      q.z = __vminu2(__vminu2(a.z, b.z), c.z);
      q.w = __vminu2(__vminu2(a.w, b.w), c.w);
      reinterpret_cast<uint4*>(gz + r * nzp)[kb] = q;
      // entries at z >= nz (occupancy bits there are 0) are never read as centres, only as padding
    }
  }
  __syncthreads();
  const int R = g.R;
  const int nz8 = nzp >> 3;
  const bool vec_out = (g.nz & 7) == 0;  // then every (x, y) row of `out` starts 16-byte aligned
  const uint4* gz4 = reinterpret_cast<const uint4*>(gz);
  for (int i = threadIdx.x; i < (y1 - y0) * nz8; i += blockDim.x)
  {
    const int yy = static_cast<int>(divByMagic(static_cast<unsigned>(i), nz8_magic));  // i / nz8
    const int zc = i - yy * nz8;
    const uint4* ctr = gz4 + (yy + R) * nz8 + zc;  // row yy + R of the haloed block
    uint4 best = *ctr;
    auto rowPair = [&](int k) {
      const uint32_t k2p = static_cast<uint32_t>(k * k) * 0x10001u;
      const uint4 a = ctr[-k * nz8], b = ctr[k * nz8];
      best.x = __viaddmin_u16x2(__vminu2(a.x, b.x), k2p, best.x);
      best.y = __viaddmin_u16x2(__vminu2(a.y, b.y), k2p, best.y);
      best.z = __viaddmin_u16x2(__vminu2(a.z, b.z), k2p, best.z);
      best.w = __viaddmin_u16x2(__vminu2(a.w, b.w), k2p, best.w);
    };
    for (int k = 1; k <= R; k += 2)
    {  // one termination test per two candidate distances (a candidate too many changes nothing: it is a min)
      const uint32_t w2 = __vmaxu2(__vmaxu2(best.x, best.y), __vmaxu2(best.z, best.w));
      if (static_cast<uint32_t>(k * k) >= max(w2 & 0xffffu, w2 >> 16))
        break;  // every further candidate is at least k^2 away
      rowPair(k);
      if (k < R)
        rowPair(k + 1);
    }
    uint16_t* dst = out + (static_cast<size_t>(x) * g.ny + (y0 + yy)) * g.nz + zc * 8;
    if (vec_out)
      *reinterpret_cast<uint4*>(dst) = best;
    else
    {
      const uint32_t bw[4] = { best.x, best.y, best.z, best.w };
      for (int j = 0; j < 8 && zc * 8 + j < g.nz; ++j)
        dst[j] = static_cast<uint16_t>(bw[j >> 1] >> (16 * (j & 1)));
    }
  }
}

// The z + y passes tiled (x plane) x (y chunk) x (z range): with the whole y extent of a plane in one CTA (the usual case)
// no halo row is transformed twice -- the first version cut a plane into y chunks of 4R rows and ran the z pass on 2R halo
// rows per chunk, 1.5x the z work -- and a z range of a few 8-voxel items keeps the slab small enough for four 512-thread
// CTAs per SM.  The z pass of an item needs the occupancy words within R of it, nothing else of the row.
//   shared memory: bits [rows_in][nw] | ends [rows_in][nwo] | gz [chunk + 2R][zi] x 16 bytes | the z pass's tables
// rows of gz outside the grid are filled with "nothing within R" instead of being computed from zero bits.
struct ZYPlan
{
  int chunk, n_chunks, n_zs;        // y rows per CTA, CTAs per plane along y, z ranges per row
  int zi_max, nw_max, nwo_max, rows_in_max;  // layout maxima over all CTAs
};

__device__ __forceinline__ unsigned smallMagic(unsigned d) { return d <= 1u ? 0u : 0xffffffffu / d + 1u; }  // i / d for i < 2^16

__global__ void __launch_bounds__(512, 4) edtZYKernel2(GridDev g, const uint32_t* __restrict__ occ, uint16_t* __restrict__ out,
                                                       ZYPlan p, int invert, uint8_t* own_busy, const uint4* __restrict__ lut)
{
  extern __shared__ __align__(16) uint8_t smem[];
  depLaunch();
  int b = blockIdx.x;
  const int zs = b % p.n_zs;
  b /= p.n_zs;
  const int ci = b % p.n_chunks, x = b / p.n_chunks;
  const int R = g.R, cap = R + 1;
  const int nz8 = (g.nz + 7) >> 3;
  const int i0 = (zs * nz8) / p.n_zs, i1 = ((zs + 1) * nz8) / p.n_zs, zi = i1 - i0;  // this CTA's items [i0, i1) of a row
  const int y0 = ci * p.chunk, y1 = min(g.ny, y0 + p.chunk), ny_c = y1 - y0;
  const int ya = y0 - R, rows_g = ny_c + 2 * R;                     // gz row j <-> y = ya + j
  const int ra = max(0, ya), rb = min(g.ny, y1 + R), n_in = rb - ra;  // the rows of it inside the grid
  const int w_lo = max(0, (8 * i0 - R) >> 5), w_hi = min(g.wz - 1, (8 * i1 - 1 + R) >> 5), nw = w_hi - w_lo + 1;
  const int wo_lo = (8 * i0) >> 5, wo_hi = (8 * i1 - 1) >> 5, nwo = wo_hi - wo_lo + 1;  // words the items live in
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem);
  uint32_t* ends = bits + ((p.rows_in_max * p.nw_max + 3) & ~3);
  uint4* gz4 = reinterpret_cast<uint4*>(ends + ((p.rows_in_max * p.nwo_max + 3) & ~3));
  uint4* lut_within = gz4 + static_cast<size_t>(p.chunk + 2 * R) * p.zi_max;
  const uint4* lut_below = lut_within + 256;
  const uint4* lut_above = lut_below + (R + 2);
  const unsigned m_nw = smallMagic(nw), m_nwo = smallMagic(nwo), m_zi = smallMagic(zi);
  // prologue: nothing here depends on the kernels before this one (the tables are a function of R alone)
  for (int t = threadIdx.x; t < 256 + 2 * (R + 2); t += blockDim.x)
    lut_within[t] = lut[t];
  {
    const uint32_t capv = static_cast<uint32_t>(cap * cap) * 0x10001u;
    const uint4 capq = make_uint4(capv, capv, capv, capv);
    const int below = ra - ya, above = rows_g - (rb - ya);  // rows of gz before / behind the grid
    for (int i = threadIdx.x; i < (below + above) * zi; i += blockDim.x)
    {
      const int r = static_cast<int>(divByMagic(static_cast<unsigned>(i), m_zi)), kb = i - r * zi;
      const int j = r < below ? r : (rb - ya) + (r - below);
      gz4[j * zi + kb] = capq;
    }
  }
  depWait();
  if (own_busy != nullptr && ci == 0 && zs == 0 && threadIdx.x == 0)
    own_busy[static_cast<size_t>(x) * kPlaneBusyStride] = 0;  // the x pass (next launch) raises it again where the plane is not free
  {
    const uint32_t* src = occ + (static_cast<size_t>(x) * g.ny + ra) * g.wz + w_lo;
    const uint32_t last = (g.nz & 31) ? (1u << (g.nz & 31)) - 1u : 0xffffffffu;  // bits of a row's last word inside the grid
    for (int i = threadIdx.x; i < n_in * nw; i += blockDim.x)
    {
      const int r = static_cast<int>(divByMagic(static_cast<unsigned>(i), m_nw)), w = i - r * nw;
      uint32_t v = src[r * g.wz + w];
      if (invert)
        v = ~v & ((w_lo + w == g.wz - 1) ? last : 0xffffffffu);
      bits[i] = v;
    }
  }
  __syncthreads();
  // (a) per word the items live in: distance of its bit 0 / bit 31 to the nearest set bit in the words below / above
  for (int i = threadIdx.x; i < n_in * nwo; i += blockDim.x)
  {
    const int r = static_cast<int>(divByMagic(static_cast<unsigned>(i), m_nwo)), wq = i - r * nwo;
    const int w = wo_lo - w_lo + wq;
    const uint32_t* row = bits + r * nw;
    int dl = cap, dr = cap;
    for (int ww = w - 1, off = 1; ww >= 0 && off <= R; --ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        dl = min(cap, off + __clz(v));
        break;
      }
    }
    for (int ww = w + 1, off = 1; ww < nw && off <= R; ++ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        dr = min(cap, off + __ffs(v) - 1);
        break;
      }
    }
    ends[i] = static_cast<uint32_t>(dl) | (static_cast<uint32_t>(dr) << 16);
  }
  __syncthreads();
  // (b) per item (occupancy byte): three table rows, as in edtZYKernel
  {
    uint4* gzin = gz4 + (ra - ya) * zi;
    for (int i = threadIdx.x; i < n_in * zi; i += blockDim.x)
    {
      const int r = static_cast<int>(divByMagic(static_cast<unsigned>(i), m_zi)), kb = i - r * zi;
      const int item = i0 + kb, w = item >> 2, o = (item & 3) * 8;
      const uint32_t cur = bits[r * nw + (w - w_lo)], e = ends[r * nwo + (w - wo_lo)];
      const uint32_t low = o ? (cur & ((1u << o) - 1u)) : 0u, high = o < 24 ? (cur >> (o + 8)) : 0u;
      const int dl = low ? o - 31 + __clz(low) : o + static_cast<int>(e & 0xffffu);
      const int dr = high ? __ffs(high) : 24 - o + static_cast<int>(e >> 16);
      const uint4 a = lut_within[(cur >> o) & 0xffu], bb = lut_below[min(dl, cap)], c = lut_above[min(dr, cap)];
      uint4 q;
      q.x = __vminu2(__vminu2(a.x, bb.x), c.x);
      q.y = __vminu2(__vminu2(a.y, bb.y), c.y);
      q.z = __vminu2(__vminu2(a.z, bb.z), c.z);
      q.w = __vminu2(__vminu2(a.w, bb.w), c.w);
      gzin[i] = q;
    }
  }
  __syncthreads();
  // y pass: 8 voxels per thread, outward scan, one termination test per two candidate distances
  const bool vec_out = (g.nz & 7) == 0;  // then every (x, y) row of `out` starts 16-byte aligned
  uint16_t* out_x = out + static_cast<size_t>(x) * g.ny * g.nz;
  for (int i = threadIdx.x; i < ny_c * zi; i += blockDim.x)
  {
    const int yy = static_cast<int>(divByMagic(static_cast<unsigned>(i), m_zi)), zc = i - yy * zi;
    const uint4* pa = gz4 + R * zi + i;  // row yy + R of the haloed block, item zc
    const uint4* pb = pa;
    uint4 best = *pa;
    uint32_t k2p = 0x10001u, inc = 3u * 0x10001u;  // k^2 in both halves; (k + 1)^2 - k^2
    auto rowPair = [&]() {
      pa -= zi;
      pb += zi;
      const uint4 a = *pa, bb = *pb;
      best.x = __viaddmin_u16x2(__vminu2(a.x, bb.x), k2p, best.x);
      best.y = __viaddmin_u16x2(__vminu2(a.y, bb.y), k2p, best.y);
      best.z = __viaddmin_u16x2(__vminu2(a.z, bb.z), k2p, best.z);
      best.w = __viaddmin_u16x2(__vminu2(a.w, bb.w), k2p, best.w);
      k2p += inc;
      inc += 2u * 0x10001u;
    };
    for (int k = 1; k <= R; k += 2)
    {
      const uint32_t w2 = __vmaxu2(__vmaxu2(best.x, best.y), __vmaxu2(best.z, best.w));
      if (__vmaxu2(w2, k2p) == k2p)
        break;  // k^2 >= every best: every further candidate is at least k^2 away
      rowPair();
      if (k < R)
        rowPair();
    }
    const int off = (y0 + yy) * g.nz + (i0 + zc) * 8;
    if (vec_out)
      *reinterpret_cast<uint4*>(out_x + off) = best;
    else
    {
      const uint32_t bw[4] = { best.x, best.y, best.z, best.w };
      for (int j = 0; j < 8 && (i0 + zc) * 8 + j < g.nz; ++j)
        out_x[off + j] = static_cast<uint16_t>(bw[j >> 1] >> (16 * (j & 1)));
    }
  }
}

// voxel i = (x * ny + y) * nz + z lies in the outermost cell layer of the grid
__device__ __forceinline__ bool onShell(const GridDev& g, size_t i)
{
  const size_t row = i / g.nz;
  const int z = static_cast<int>(i - row * g.nz);
  const int x = static_cast<int>(row / g.ny), y = static_cast<int>(row - static_cast<size_t>(x) * g.ny);
  return x == 0 || x == g.nx - 1 || y == 0 || y == g.ny - 1 || z == 0 || z == g.nz - 1;
}

// Raise plane x's flag if any thread of the CTA holds a non-free element: ONE store per CTA (a CTA works on one plane),
// and the flags sit kPlaneBusyStride bytes apart -- thousands of threads share each flag, and same-address (or same-line)
// traffic serialises in L2.  Called by every thread of the CTA that has not exited.
__device__ __forceinline__ void raiseBusy(uint8_t* own_busy, int x, bool mine)
{
  if (own_busy == nullptr)
    return;  // uniform
  if (__syncthreads_or(mine ? 1 : 0) != 0 && threadIdx.x == 0)
    own_busy[static_cast<size_t>(x) * kPlaneBusyStride] = 1;
}

// x pass, 8 consecutive z per thread (16-byte loads / stores); requires nz % 8 == 0.  The compact field is
// interleaved {world, self} per voxel: this field owns every other element, so its 8 elements are merged into the
// 16 (uint8) or 32 (uint16) bytes they share with the other field (read-modify-write; fields are built one at a time
// on one stream).
template <typename FT>
__global__ void __launch_bounds__(256) edtXKernelVec(GridDev g, const uint16_t* __restrict__ in, uint16_t* __restrict__ d2,
                                                    FT* compact, unsigned nz8_magic, uint8_t* own_busy,
                                                    const uint8_t* __restrict__ other_busy)
{
  // grid: x = tiles of the (y, z/8) plane, y = x planes
  depLaunch();
  const int nz8 = g.nz >> 3;
  const unsigned plane8 = static_cast<unsigned>(g.ny) * nz8;
  const unsigned p = blockIdx.x * blockDim.x + threadIdx.x;
  depWait();  // the intermediate comes from the z + y kernel right before this one (programmatic dependent launch)
  if (p >= plane8)
    return;
  const int y = static_cast<int>(divByMagic(p, nz8_magic));  // p / nz8
  const int zc = static_cast<int>(p) - y * nz8;
  const bool shell_y = y == 0 || y == g.ny - 1;
  const int R = g.R, max_d2 = g.max_d2;
  const uint4* in4 = reinterpret_cast<const uint4*>(in);
  // packed 16-bit lanes throughout; anything above max_d2 means "nothing within range".  The z+y pass writes at most
  // (R+1)^2 + R^2, so the sums below stay far from 2^16; the centre is capped to max_d2 + 1 so the scan ends at k = R.
  const uint32_t cap1 = static_cast<uint32_t>(max_d2 + 1) * 0x10001u;
  const uint32_t capm = static_cast<uint32_t>(max_d2) * 0x10001u;
  for (int x = blockIdx.y; x < g.nx; x += gridDim.y)
  {
    const size_t i = static_cast<size_t>(x) * plane8 + p;
    const uint4 c = in4[i];
    uint32_t bw[4] = { __vminu2(c.x, cap1), __vminu2(c.y, cap1), __vminu2(c.z, cap1), __vminu2(c.w, cap1) };
    for (int k = 1; k <= R; ++k)
    {
      const uint32_t k2 = static_cast<uint32_t>(k * k);
      const uint32_t w2 = __vmaxu2(__vmaxu2(bw[0], bw[1]), __vmaxu2(bw[2], bw[3]));
      if (k2 >= max(w2 & 0xffffu, w2 >> 16))
        break;
      uint4 a = make_uint4(cap1, cap1, cap1, cap1), b = a;
      if (x - k >= 0)
        a = in4[i - static_cast<size_t>(k) * plane8];
      if (x + k < g.nx)
        b = in4[i + static_cast<size_t>(k) * plane8];
      const uint32_t k2p = k2 | (k2 << 16);
      bw[0] = __viaddmin_u16x2(__vminu2(a.x, b.x), k2p, bw[0]);
      bw[1] = __viaddmin_u16x2(__vminu2(a.y, b.y), k2p, bw[1]);
      bw[2] = __viaddmin_u16x2(__vminu2(a.z, b.z), k2p, bw[2]);
      bw[3] = __viaddmin_u16x2(__vminu2(a.w, b.w), k2p, bw[3]);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
      bw[j] = __vminu2(bw[j], capm);
    reinterpret_cast<uint4*>(d2)[i] = make_uint4(bw[0], bw[1], bw[2], bw[3]);
    if (compact == nullptr)
      continue;  // the negative half of a signed field has no compact copy
    // The 1-cell border shell of the COMPACT field is kept at "free": DistanceField::getDistanceGradient
    // (distance_field.cpp:82) answers max_distance for those cells, and the fast check kernel clamps to them.
    if (shell_y || x == 0 || x == g.nx - 1)
      bw[0] = bw[1] = bw[2] = bw[3] = capm;
    else
    {
      if (zc == 0)
        bw[0] = (bw[0] & 0xffff0000u) | static_cast<uint32_t>(max_d2);
      if (zc == nz8 - 1)
        bw[3] = (bw[3] & 0x0000ffffu) | (static_cast<uint32_t>(max_d2) << 16);
    }
    // our 8 elements of the interleaved {world, self} compact field: element 2 * (8 i + j) counted from `compact`
    // (already offset by 0 = world / 1 = self).  Where the other field's plane is known to be all "free" the whole words
    // are written; elsewhere the other field's elements in between are kept (read-modify-write).
    const bool second = (reinterpret_cast<uintptr_t>(compact) & (2 * sizeof(FT) - 1)) != 0;
    const bool other_free = other_busy != nullptr && other_busy[static_cast<size_t>(x) * kPlaneBusyStride] == 0;  // uniform over the CTA
    if (sizeof(FT) == 1)
    {
      const uint32_t sh = second ? 8u : 0u, osh = 8u - sh, keep = ~(0x00ff00ffu << sh);
      const uint32_t freeb = static_cast<uint32_t>(min(max_d2, 255)) * 0x00010001u;
      uint4* q = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(compact) - (second ? 1 : 0)) + i;
      const uint4 own = make_uint4(__vminu2(bw[0], 0x00ff00ffu), __vminu2(bw[1], 0x00ff00ffu), __vminu2(bw[2], 0x00ff00ffu),
                                   __vminu2(bw[3], 0x00ff00ffu));
      raiseBusy(own_busy, x, ((own.x ^ freeb) | (own.y ^ freeb) | (own.z ^ freeb) | (own.w ^ freeb)) != 0u);
      uint4 v;  // voxel pair per 32-bit word: bytes { world, self, world, self }
      if (other_free)
        v = make_uint4(freeb << osh, freeb << osh, freeb << osh, freeb << osh);
      else
      {
        v = *q;
        v.x &= keep; v.y &= keep; v.z &= keep; v.w &= keep;
      }
      v.x |= own.x << sh;
      v.y |= own.y << sh;
      v.z |= own.z << sh;
      v.w |= own.w << sh;
      *q = v;
    }
    else
    {
      const uint32_t freew = static_cast<uint32_t>(max_d2) * 0x00010001u;
      raiseBusy(own_busy, x, ((bw[0] ^ freew) | (bw[1] ^ freew) | (bw[2] ^ freew) | (bw[3] ^ freew)) != 0u);
      if (other_free)
      {  // voxel per 32-bit word: { world, self }
        const uint32_t fo = static_cast<uint32_t>(max_d2) << (second ? 0 : 16);
        uint32_t w[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          w[j] = (((bw[j >> 1] >> (16 * (j & 1))) & 0xffffu) << (second ? 16 : 0)) | fo;
        uint4* q = reinterpret_cast<uint4*>(reinterpret_cast<uint16_t*>(compact) - (second ? 1 : 0)) + 2 * i;
        q[0] = make_uint4(w[0], w[1], w[2], w[3]);
        q[1] = make_uint4(w[4], w[5], w[6], w[7]);
      }
      else
      {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          compact[2 * (8 * i + j)] = static_cast<FT>((bw[j >> 1] >> (16 * (j & 1))) & 0xffffu);
      }
    }
  }
}

template <typename FT>
__global__ void __launch_bounds__(256) edtXKernel(GridDev g, const uint16_t* __restrict__ in, uint16_t* __restrict__ d2,
                                                 FT* compact)
{
  const size_t plane = static_cast<size_t>(g.ny) * g.nz;
  const size_t total = plane * g.nx;
  const int R = g.R, max_d2 = g.max_d2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int x = static_cast<int>(i / plane);
    int best = in[i];  // anything above max_d2 means "nothing within range": it can only lose
    const int limit = min(best, max_d2 + 1);
    best = limit;
    for (int k = 1; k <= R && k * k < best; ++k)
    {
      const int k2 = k * k;
      if (x - k >= 0)
        best = min(best, static_cast<int>(in[i - k * plane]) + k2);
      if (x + k < g.nx)
        best = min(best, static_cast<int>(in[i + k * plane]) + k2);
    }
    best = min(best, max_d2);
    d2[i] = static_cast<uint16_t>(best);
    // interleaved compact field: element 2 * i + which (the caller passes compact already offset by `which`); its
    // 1-cell border shell is kept at "free" (see edtXKernelVec)
    if (compact == nullptr)
      continue;
    if (onShell(g, i))
      best = max_d2;
    compact[2 * i] = static_cast<FT>(sizeof(FT) == 1 ? min(best, 255) : best);
  }
}

template <typename FT>
__global__ void fillKernel(size_t total, int max_d2, uint16_t* d2, FT* compact)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    d2[i] = static_cast<uint16_t>(max_d2);
    compact[2 * i] = static_cast<FT>(sizeof(FT) == 1 ? min(max_d2, 255) : max_d2);
  }
}

template <typename FT>
__global__ void injectKernel(GridDev g, const int32_t* __restrict__ in, uint32_t* occ, uint16_t* d2, FT* compact)
{
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int v = min(max(in[i], 0), g.max_d2);
    d2[i] = static_cast<uint16_t>(v);
    const int vc = onShell(g, i) ? g.max_d2 : v;
    compact[2 * i] = static_cast<FT>(sizeof(FT) == 1 ? min(vc, 255) : vc);
    if (v == 0)
    {
      const int z = static_cast<int>(i % g.nz);
      const size_t row = i / g.nz;
      atomicOr(occ + row * g.wz + (z >> 5), 1u << (z & 31));
    }
  }
}

__global__ void injectNegKernel(size_t total, int max_d2, const int32_t* __restrict__ in, uint16_t* __restrict__ nd2)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    nd2[i] = static_cast<uint16_t>(min(max(in[i], 0), max_d2));
}

// DistanceField::addShapeToField for a SPHERE (distance_field.cpp:208-238): findInternalPointsConvex
// (find_internal_points.cpp:39-64) visits the lattice xs x ys x zs -- the axis values come from the host, accumulated
// `pt += resolution` exactly as the reference's loops do -- and keeps the points bodies::Sphere::containsPoint accepts,
// |c - p|^2 <= r^2 (the predicate test_big.df pins); each kept point marks its voxel (addPointsToField).
__global__ void sphereShapeKernel(const double* __restrict__ axes, int nxs, int nys, int nzs, double cx, double cy, double cz,
                                  double r2, GridDev g, uint32_t* __restrict__ occ)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int iz = static_cast<int>(i % nzs);
    const size_t t = i / nzs;
    const int iy = static_cast<int>(t % nys), ix = static_cast<int>(t / nys);
    const double px = axes[ix], py = axes[nxs + iy], pz = axes[nxs + nys + iz];
    const double dx = cx - px, dy = cy - py, dz = cz - pz;
    if (!(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)) <= r2))
      continue;
    const int x = static_cast<int>(floor((px - g.om[0]) * g.oo)), y = static_cast<int>(floor((py - g.om[1]) * g.oo)),
              z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;
    atomicOr(occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5), 1u << (z & 31));
  }
}

// Interior lattice points of a bodies::Body (findInternalPointsConvex, find_internal_points.cpp:39-64) on the device: the axis
// values come from the host, accumulated `pt += resolution` as the reference's loops do; containsPoint per body type as
// geometric_shapes' bodies.cpp has it (restated from the published source: un-vendored, parity unpinned) in explicit
// round-to-nearest double operations, left to right, no FMA (the order the numpy restatement in oracle/ uses):
//   Sphere    |c - p|^2 <= r^2
//   Box       |R^T (p - c)| <= half extents, per axis
//   Cylinder  |(p - c) . nH| <= half length and ((p - c) . nB1)^2 <= r^2 and ((p - c) . nB2)^2 <= r^2 - ((p - c) . nB1)^2
// A kept point is then moved by `post` (PosedBodyPointDecomposition::updatePose, types.cpp:388-406: the lattice was laid out
// in the body frame) or left where it is (DistanceField::addShapeToField: the lattice is in the world frame), and its voxel is
// marked (addPointsToField) or cleared (removePointsFromField).
__device__ __forceinline__ double dot3rn(double ax, double ay, double az, double bx, double by, double bz)
{
  return __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
}

__global__ void bodyLatticeKernel(const double* __restrict__ axes, int nxs, int nys, int nzs, BodyDev b, GridDev g,
                                  uint32_t* __restrict__ occ, int set)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int iz = static_cast<int>(i % nzs);
    const size_t t = i / nzs;
    const int iy = static_cast<int>(t % nys), ix = static_cast<int>(t / nys);
    double px = axes[ix], py = axes[nxs + iy], pz = axes[nxs + nys + iz];
    bool inside;
    if (b.type == 1)
    {  // (center_ - p).squaredNorm() <= radius2_
      const double dx = __dsub_rn(b.c[0], px), dy = __dsub_rn(b.c[1], py), dz = __dsub_rn(b.c[2], pz);
      inside = dot3rn(dx, dy, dz, dx, dy, dz) <= b.dims[0];
    }
    else
    {
      const double vx = __dsub_rn(px, b.c[0]), vy = __dsub_rn(py, b.c[1]), vz = __dsub_rn(pz, b.c[2]);
      // columns of the body's rotation: R[r][k] = b.R[3 * r + k]
      const double a0 = dot3rn(vx, vy, vz, b.R[0], b.R[3], b.R[6]);
      const double a1 = dot3rn(vx, vy, vz, b.R[1], b.R[4], b.R[7]);
      const double a2 = dot3rn(vx, vy, vz, b.R[2], b.R[5], b.R[8]);
      if (b.type == 4)
        inside = fabs(a0) <= b.dims[0] && fabs(a1) <= b.dims[1] && fabs(a2) <= b.dims[2];
      else
      {  // cylinder: dims = { radius^2, half length }
        const double remaining = __dsub_rn(b.dims[0], __dmul_rn(a0, a0));
        inside = !(fabs(a2) > b.dims[1]) && !(remaining < 0.0) && __dmul_rn(a1, a1) <= remaining;
      }
    }
    if (!inside)
      continue;
    if (b.posed)
    {  // trans * p: linear() * p + translation()
      const double qx = __dadd_rn(dot3rn(b.P[0], b.P[1], b.P[2], px, py, pz), b.P[9]);
      const double qy = __dadd_rn(dot3rn(b.P[3], b.P[4], b.P[5], px, py, pz), b.P[10]);
      const double qz = __dadd_rn(dot3rn(b.P[6], b.P[7], b.P[8], px, py, pz), b.P[11]);
      px = qx;
      py = qy;
      pz = qz;
    }
    const int x = static_cast<int>(floor((px - g.om[0]) * g.oo)), y = static_cast<int>(floor((py - g.om[1]) * g.oo)),
              z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;
    uint32_t* w = occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5);
    if (set)
      atomicOr(w, 1u << (z & 31));
    else
      atomicAnd(w, ~(1u << (z & 31)));
  }
}

// bodies::ConvexMesh::containsPoint as geometric_shapes' published bodies.cpp has it (un-vendored in the reference: parity
// unpinned): the point goes to the body's base frame (i_pose_ * p) and must lie on the inner side of every plane of the hull,
// plane_normal . p + w - 1e-9 <= 0 (isPointInsidePlanes; w is the caller's, recomputed from a scaled + padded vertex as the
// library does).  The bounding-box pre-test of the library is implied by the planes of a closed hull.  Explicit round-to-
// nearest double operations, left to right, no FMA (the order the numpy restatement in oracle/ uses).
__global__ void convexLatticeKernel(const double* __restrict__ axes, int nxs, int nys, int nzs,
                                    const double* __restrict__ planes, ConvexDev b, GridDev g, uint32_t* __restrict__ occ, int set)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int iz = static_cast<int>(i % nzs);
    const size_t t = i / nzs;
    const int iy = static_cast<int>(t % nys), ix = static_cast<int>(t / nys);
    double px = axes[ix], py = axes[nxs + iy], pz = axes[nxs + nys + iz];
    const double bx = __dadd_rn(dot3rn(b.ipose[0], b.ipose[1], b.ipose[2], px, py, pz), b.ipose[9]);
    const double by = __dadd_rn(dot3rn(b.ipose[3], b.ipose[4], b.ipose[5], px, py, pz), b.ipose[10]);
    const double bz = __dadd_rn(dot3rn(b.ipose[6], b.ipose[7], b.ipose[8], px, py, pz), b.ipose[11]);
    bool inside = true;
    for (int k = 0; k < b.n_planes && inside; ++k)
    {
      const double dist = __dsub_rn(__dadd_rn(dot3rn(__ldg(planes + 4 * k), __ldg(planes + 4 * k + 1), __ldg(planes + 4 * k + 2),
                                                     bx, by, bz),
                                              __ldg(planes + 4 * k + 3)),
                                    1e-9);
      inside = !(dist > 0.0);
    }
    if (!inside)
      continue;
    if (b.posed)
    {
      const double qx = __dadd_rn(dot3rn(b.P[0], b.P[1], b.P[2], px, py, pz), b.P[9]);
      const double qy = __dadd_rn(dot3rn(b.P[3], b.P[4], b.P[5], px, py, pz), b.P[10]);
      const double qz = __dadd_rn(dot3rn(b.P[6], b.P[7], b.P[8], px, py, pz), b.P[11]);
      px = qx;
      py = qy;
      pz = qz;
    }
    const int x = static_cast<int>(floor((px - g.om[0]) * g.oo)), y = static_cast<int>(floor((py - g.om[1]) * g.oo)),
              z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;
    uint32_t* w = occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5);
    if (set)
      atomicOr(w, 1u << (z & 31));
    else
      atomicAnd(w, ~(1u << (z & 31)));
  }
}

// DistanceField::getOcTreePoints (distance_field.cpp:240-284), the part after octomap's leaf iteration (which stays with the
// caller: octomap is its dependency): an occupied leaf no larger than a cell contributes its centre; a larger one the lattice
// x = cx - c, x <= cx + c, x += resolution (accumulated, as the reference's loops) with c = ceil(size / resolution) *
// resolution / 2, on all three axes.  Every point then marks its voxel (addPointsToField).  One thread per leaf.
__global__ void octreeLeavesKernel(const double* __restrict__ leaves, size_t n, double res, GridDev g,
                                   uint32_t* __restrict__ occ)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const double cx = leaves[4 * i], cy = leaves[4 * i + 1], cz = leaves[4 * i + 2], size = leaves[4 * i + 3];
    auto mark = [&](double px, double py, double pz) {
      const int x = static_cast<int>(floor((px - g.om[0]) * g.oo)), y = static_cast<int>(floor((py - g.om[1]) * g.oo)),
                z = static_cast<int>(floor((pz - g.om[2]) * g.oo));
      if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
        return;
      atomicOr(occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5), 1u << (z & 31));
    };
    if (size <= res)
    {
      mark(cx, cy, cz);
      continue;
    }
    const double c = __ddiv_rn(__dmul_rn(ceil(__ddiv_rn(size, res)), res), 2.0);
    const double ex = __dadd_rn(cx, c), ey = __dadd_rn(cy, c), ez = __dadd_rn(cz, c);
    for (double x = __dsub_rn(cx, c); x <= ex; x = __dadd_rn(x, res))
      for (double y = __dsub_rn(cy, c); y <= ey; y = __dadd_rn(y, res))
        for (double z = __dsub_rn(cz, c); z <= ez; z = __dadd_rn(z, res))
          mark(x, y, z);
  }
}

// updatePointsInField (propagation_distance_field.cpp:130-175): occ' = (occ - (old \ new)) + (new \ old), word by word
__global__ void updateOccKernel(uint32_t* __restrict__ occ, const uint32_t* __restrict__ old_bits,
                                const uint32_t* __restrict__ new_bits, size_t n_words)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n_words;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const uint32_t a = old_bits[i], b = new_bits[i];
    occ[i] = (occ[i] & ~(a & ~b)) | (b & ~a);
  }
}

template <typename FT>
__global__ void compactCopyKernel(FT* compact, FT* dense, size_t n, int extract)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    if (extract)
      dense[i] = compact[2 * i];
    else
      compact[2 * i] = dense[i];
  }
}

__global__ void widenStatesKernel(const float* __restrict__ in, double* __restrict__ out, size_t n)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = static_cast<double>(in[i]);
}

int gridFor(size_t n, int block, int cap)
{
  size_t b = (n + block - 1) / block;
  if (b < 1)
    b = 1;
  return static_cast<int>(b > static_cast<size_t>(cap) ? cap : b);
}
}  // namespace

cudaError_t launchScatter(const double* d_xyz, size_t n, const GridDev& g, uint32_t* occ, bool set, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  static const int variant = getenv("B200SV_SCATTER") ? atoi(getenv("B200SV_SCATTER")) : 0;
  static const int grid_env = getenv("B200SV_SCATTER_GRID") ? atoi(getenv("B200SV_SCATTER_GRID")) : 0;
  const int grid = gridFor(n, 256, grid_env > 0 ? grid_env : 148 * 16);
  if (variant == 1 && (reinterpret_cast<uintptr_t>(d_xyz) & 15u) == 0)
    scatterKernelTma<<<grid, 256, 0, s>>>(d_xyz, n, g, occ, set ? 1 : 0);
  else
    scatterKernel<<<grid, 256, 0, s>>>(d_xyz, n, g, occ, set ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchSphereShape(const double* d_axes, int nxs, int nys, int nzs, const double* c, double r2, const GridDev& g,
                              uint32_t* occ, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  if (total == 0)
    return cudaSuccess;
  sphereShapeKernel<<<gridFor(total, 256, 148 * 16), 256, 0, s>>>(d_axes, nxs, nys, nzs, c[0], c[1], c[2], r2, g, occ);
  return cudaGetLastError();
}

cudaError_t launchBodyLattice(const double* d_axes, int nxs, int nys, int nzs, const BodyDev& body, const GridDev& g,
                              uint32_t* occ, bool set, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  if (total == 0)
    return cudaSuccess;
  bodyLatticeKernel<<<gridFor(total, 256, 148 * 16), 256, 0, s>>>(d_axes, nxs, nys, nzs, body, g, occ, set ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchConvexLattice(const double* d_axes, int nxs, int nys, int nzs, const double* d_planes, const ConvexDev& body,
                                const GridDev& g, uint32_t* occ, bool set, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(nxs) * nys * nzs;
  if (total == 0)
    return cudaSuccess;
  convexLatticeKernel<<<gridFor(total, 256, 148 * 16), 256, 0, s>>>(d_axes, nxs, nys, nzs, d_planes, body, g, occ, set ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchOctreeLeaves(const double* d_leaves, size_t n, double resolution, const GridDev& g, uint32_t* occ,
                               cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  octreeLeavesKernel<<<gridFor(n, 128, 148 * 16), 128, 0, s>>>(d_leaves, n, resolution, g, occ);
  return cudaGetLastError();
}

cudaError_t launchUpdateOcc(uint32_t* occ, const uint32_t* old_bits, const uint32_t* new_bits, size_t n_words, cudaStream_t s)
{
  updateOccKernel<<<gridFor(n_words, 256, 148 * 8), 256, 0, s>>>(occ, old_bits, new_bits, n_words);
  return cudaGetLastError();
}

// The z pass's tables, one per (device, R), built on first use and kept for the life of the process (a few KB each).
const uint4* edtLut(int dev, int R, cudaStream_t s)
{
  struct Entry { int dev, R; uint4* d; };
  static std::mutex mu;
  static std::vector<Entry> cache;
  std::lock_guard<std::mutex> lk(mu);
  for (const Entry& e : cache)
    if (e.dev == dev && e.R == R)
      return e.d;
  uint4* d = nullptr;
  if (cudaMalloc(&d, (256 + 2 * static_cast<size_t>(R + 2)) * sizeof(uint4)) != cudaSuccess)
    return nullptr;
  edtLutKernel<<<2, 256, 0, s>>>(R, d);
  // other streams (other contexts on this device) may read the table right after this returns
  if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(s) != cudaSuccess)
  {
    cudaFree(d);
    return nullptr;
  }
  cache.push_back(Entry{ dev, R, d });
  return d;
}

// Tiling of the z + y kernel: y rows per CTA (the whole extent when it fits: no halo rows) and z ranges per row, chosen by a
// small cost model -- thread-instructions per CTA from the measured per-item costs of the phases, CTAs per SM from shared
// memory (at most four of 512 threads), whole waves of the grid.  B200SV_EDT_CHUNK / B200SV_EDT_ZS override.
namespace
{
size_t zyPlanBytes(const GridDev& g, ZYPlan& p)
{
  const int nz8 = (g.nz + 7) / 8;
  p.n_chunks = (g.ny + p.chunk - 1) / p.chunk;
  p.zi_max = (nz8 + p.n_zs - 1) / p.n_zs;
  p.rows_in_max = std::min(g.ny, p.chunk + 2 * g.R);
  p.nw_max = std::min(g.wz, (8 * p.zi_max + 2 * g.R + 30) / 32 + 1);
  p.nwo_max = std::min(g.wz, (8 * p.zi_max + 30) / 32 + 1);
  const size_t lut = (256 + 2 * static_cast<size_t>(g.R + 2)) * 16;
  return static_cast<size_t>((p.rows_in_max * p.nw_max + 3) & ~3) * 4 + static_cast<size_t>((p.rows_in_max * p.nwo_max + 3) & ~3) * 4 +
         static_cast<size_t>(p.chunk + 2 * g.R) * p.zi_max * 16 + lut;
}

ZYPlan planZY(const GridDev& g, int sm_count, size_t* smem_out)
{
  static const int chunk_env = getenv("B200SV_EDT_CHUNK") ? atoi(getenv("B200SV_EDT_CHUNK")) : 0;
  static const int zs_env = getenv("B200SV_EDT_ZS") ? atoi(getenv("B200SV_EDT_ZS")) : 0;
  const int nz8 = (g.nz + 7) / 8;
  ZYPlan best{};
  double best_cost = 0.0;
  size_t best_bytes = 0;
  for (int parts = 1; parts <= g.ny; ++parts)
  {
    int chunk = (g.ny + parts - 1) / parts;
    if (chunk_env > 0)
      chunk = std::min(chunk_env, g.ny);
    if (parts > 1 && chunk < std::min(g.ny, std::max(2 * g.R, 16)) && best.chunk != 0)
      break;  // halo rows would outnumber the rows they serve
    for (int n_zs = 1; n_zs <= nz8; ++n_zs)
    {
      if (zs_env > 0 && n_zs != std::min(zs_env, nz8))
        continue;
      ZYPlan p{};
      p.chunk = chunk;
      p.n_zs = n_zs;
      const size_t bytes = zyPlanBytes(g, p);
      if (bytes > 200 * 1024)
        continue;
      const int per_sm = static_cast<int>(std::min<size_t>(4, (227 * 1024) / (bytes + 1024)));
      const long n_ctas = static_cast<long>(g.nx) * p.n_chunks * n_zs;
      const long waves = (n_ctas + static_cast<long>(sm_count) * per_sm - 1) / (static_cast<long>(sm_count) * per_sm);
      // thread-instructions of one CTA: z pass 59 per item, y pass 120, word ends 62 per word, loads, fixed costs
      const double rows_in = std::min(g.ny, chunk + 2 * g.R), zi = static_cast<double>(nz8) / n_zs;
      const double work = rows_in * zi * 59 + chunk * zi * 120 + rows_in * p.nwo_max * 62 + rows_in * p.nw_max * 10 + 6000;
      // CTAs on an SM share its issue slots; fewer than 32 warps per SM hide latency badly
      const double cost = static_cast<double>(waves) * per_sm * work * (per_sm >= 2 ? 1.0 : 1.5);
      if (best.chunk == 0 || cost < best_cost)
      {
        best = p;
        best_cost = cost;
        best_bytes = bytes;
      }
    }
    if (chunk_env > 0)
      break;
  }
  *smem_out = best_bytes;
  return best;
}
}  // namespace

cudaError_t launchEdt(const GridDev& g, const uint32_t* occ, uint16_t* tmp, uint16_t* d2, void* compact,
                      int compact_bytes, int sm_count, cudaStream_t s, bool invert, uint8_t* own_busy,
                      const uint8_t* other_busy)
{
  if (g.R > 180)
    return cudaErrorInvalidConfiguration;  // (R+1)^2 + R^2 must fit 16 bits
  const int nzp = (g.nz + 7) & ~7;
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  const unsigned nz8 = static_cast<unsigned>(nzp / 8);
  const unsigned nz8_magic = static_cast<unsigned>(((1ull << 32) + nz8 - 1) / nz8);  // i / nz8 = umulhi(i, magic), i < 2^16
  if (compact == nullptr)
    own_busy = nullptr;
  const uint4* lut = edtLut(dev, g.R, s);
  if (lut == nullptr)
    return cudaErrorMemoryAllocation;
  static const bool use_pdl = !(getenv("B200SV_EDT_PDL") && atoi(getenv("B200SV_EDT_PDL")) == 0);
  static const bool zy_v1 = getenv("B200SV_EDT_ZY") && atoi(getenv("B200SV_EDT_ZY")) == 1;
  cudaLaunchAttribute pdl_attr[1];
  pdl_attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  pdl_attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaError_t e = cudaSuccess;
  if (!zy_v1)
  {
    size_t smem = 0;
    const ZYPlan plan = planZY(g, sm_count, &smem);
    if (plan.chunk == 0)
      return cudaErrorInvalidConfiguration;  // not even one row block fits: grid too deep for this kernel
    static size_t granted2[16] = { 0 };  // the attribute is per device
    if (smem > granted2[dev])
    {
      e = cudaFuncSetAttribute(edtZYKernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess)
        return e;
      granted2[dev] = smem;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(static_cast<unsigned>(g.nx) * plan.n_chunks * plan.n_zs);
    cfg.blockDim = dim3(512);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, edtZYKernel2, g, occ, tmp, plan, invert ? 1 : 0, own_busy, lut);
    if (e != cudaSuccess)
      return e;
  }
  else
  {
  // y chunks: small enough for several CTAs per SM (a haloed slab = occupancy words + squared z distances in shared
  // memory), large enough that the 2R halo rows stay a modest overhead
  const size_t per_row = static_cast<size_t>(g.wz) * 8 + static_cast<size_t>(nzp) * 2;  // bits, word ends, squared z distances
  const long max_rows = static_cast<long>((200 * 1024 - (256 + 2 * (g.R + 2)) * 16) / per_row);
  if (max_rows < 2L * g.R + 1)
    return cudaErrorInvalidConfiguration;  // a haloed row block does not fit: grid too deep for this kernel
  static const int chunk_env = getenv("B200SV_EDT_CHUNK") ? atoi(getenv("B200SV_EDT_CHUNK")) : 0;
  int chunk = chunk_env > 0 ? chunk_env : std::min<long>(g.ny, std::max<long>(4L * g.R, 64));
  chunk = static_cast<int>(std::max<long>(1, std::min<long>(std::min<long>(chunk, g.ny), max_rows - 2L * g.R)));
  const int n_chunks = (g.ny + chunk - 1) / chunk;
  const size_t lut_bytes = (256 + 2 * static_cast<size_t>(g.R + 2)) * 16 + static_cast<size_t>(chunk + 2 * g.R) * g.wz * 4;
  const size_t smem = ((static_cast<size_t>(chunk + 2 * g.R) * g.wz * 4 + 15) & ~static_cast<size_t>(15)) +
                      static_cast<size_t>(chunk + 2 * g.R) * nzp * 2 + lut_bytes;
  static size_t granted[16] = { 0 };  // the attribute is per device
  if (smem > granted[dev])
  {
    e = cudaFuncSetAttribute(edtZYKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess)
      return e;
    granted[dev] = smem;
  }
  static const int zy_threads = getenv("B200SV_EDT_THREADS") ? atoi(getenv("B200SV_EDT_THREADS")) : 512;
  edtZYKernel<<<g.nx * n_chunks, zy_threads, smem, s>>>(g, occ, tmp, chunk, n_chunks, nz8_magic, invert ? 1 : 0, own_busy, lut);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  }
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  const unsigned long long plane8 = static_cast<unsigned long long>(g.ny) * nz8;
  if (g.nz % 8 == 0 && plane8 * nz8 < (1ull << 32))  // the second condition keeps p / nz8 = umulhi(p, magic) exact
  {
    const dim3 grid(static_cast<unsigned>((plane8 + 255) / 256), static_cast<unsigned>(std::min(g.nx, 65535)));
    static const bool no_rmw_skip = getenv("B200SV_EDT_RMW") != nullptr;  // A/B: always merge, as before
    if (no_rmw_skip)
      other_busy = nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(256);
    cfg.stream = s;
    cfg.attrs = pdl_attr;
    cfg.numAttrs = use_pdl ? 1 : 0;
    const uint16_t* tmp_in = tmp;
    if (compact_bytes == 1)
      return cudaLaunchKernelEx(&cfg, edtXKernelVec<uint8_t>, g, tmp_in, d2, static_cast<uint8_t*>(compact), nz8_magic, own_busy,
                                other_busy);
    return cudaLaunchKernelEx(&cfg, edtXKernelVec<uint16_t>, g, tmp_in, d2, static_cast<uint16_t*>(compact), nz8_magic, own_busy,
                              other_busy);
  }
  if (own_busy != nullptr && compact != nullptr)
  {  // the scalar x pass below does not track the planes: everything may be busy
    e = cudaMemsetAsync(own_busy, 1, static_cast<size_t>(g.nx) * kPlaneBusyStride, s);
    if (e != cudaSuccess)
      return e;
  }
  const int grid = gridFor(total, 256, sm_count * 32);
  if (compact_bytes == 1)
    edtXKernel<uint8_t><<<grid, 256, 0, s>>>(g, tmp, d2, static_cast<uint8_t*>(compact));
  else
    edtXKernel<uint16_t><<<grid, 256, 0, s>>>(g, tmp, d2, static_cast<uint16_t*>(compact));
  return cudaGetLastError();
}

cudaError_t launchCompactCopy(void* compact, void* dense, size_t n_voxels, int compact_bytes, bool extract, cudaStream_t s)
{
  const int grid = gridFor(n_voxels, 256, 148 * 16);
  if (compact_bytes == 1)
    compactCopyKernel<uint8_t><<<grid, 256, 0, s>>>(static_cast<uint8_t*>(compact), static_cast<uint8_t*>(dense), n_voxels,
                                                    extract ? 1 : 0);
  else
    compactCopyKernel<uint16_t><<<grid, 256, 0, s>>>(static_cast<uint16_t*>(compact), static_cast<uint16_t*>(dense), n_voxels,
                                                     extract ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchWidenStates(const float* in, double* out, size_t n, cudaStream_t s)
{
  if (n == 0)
    return cudaSuccess;
  widenStatesKernel<<<gridFor(n, 256, 148 * 16), 256, 0, s>>>(in, out, n);
  return cudaGetLastError();
}

cudaError_t launchFillField(const GridDev& g, uint16_t* d2, void* compact, int compact_bytes, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  const int grid = gridFor(total, 256, 148 * 16);
  if (compact_bytes == 1)
    fillKernel<uint8_t><<<grid, 256, 0, s>>>(total, g.max_d2, d2, static_cast<uint8_t*>(compact));
  else
    fillKernel<uint16_t><<<grid, 256, 0, s>>>(total, g.max_d2, d2, static_cast<uint16_t*>(compact));
  return cudaGetLastError();
}

cudaError_t launchInjectNegD2(const GridDev& g, const int32_t* d_in, uint16_t* nd2, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  injectNegKernel<<<gridFor(total, 256, 148 * 16), 256, 0, s>>>(total, g.max_d2, d_in, nd2);
  return cudaGetLastError();
}

cudaError_t launchInjectD2(const GridDev& g, const int32_t* d_in, uint32_t* occ, uint16_t* d2, void* compact,
                           int compact_bytes, cudaStream_t s)
{
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  const int grid = gridFor(total, 256, 148 * 16);
  if (compact_bytes == 1)
    injectKernel<uint8_t><<<grid, 256, 0, s>>>(g, d_in, occ, d2, static_cast<uint8_t*>(compact));
  else
    injectKernel<uint16_t><<<grid, 256, 0, s>>>(g, d_in, occ, d2, static_cast<uint16_t*>(compact));
  return cudaGetLastError();
}
}  // namespace b200sv
