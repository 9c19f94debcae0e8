// The reference's OWN propagation, reproduced on the device: PropagationDistanceField's bucket-queue wavefront
// (distance_field/src/propagation_distance_field.cpp:177-444) evaluated in data-parallel phases whose result is the serial
// queue's, voxel for voxel -- including what makes that result differ from the exact transform of edt_kernels.cu
// (closest points only travel along the queue's direction-limited neighbourhoods, ties go to whoever was queued first,
// entries pushed behind the bucket being drained are dropped or stay queued for the NEXT call).
//
// Serial semantics restated (propagatePositive :390-444).  Bucket i is a FIFO of voxel locations.  Draining it, every entry
// reads its voxel's CURRENT (closest point, update direction), walks the neighbourhood of that direction (26 cells in bucket 0,
// the non-opposing axis steps from bucket 1 on) and OFFERS each neighbour the distance to its own closest point; an offer is
// accepted iff it is strictly below the neighbour's distance at that moment, and an accepted offer appends the neighbour to
// bucket[new distance].  Number the offers of one bucket key = position * S + slot (S = 26 or 6).  Then
//   * an offer is accepted  <=>  it is below the target's distance before the bucket AND below every offer to the same
//     target with a smaller key (accepted or not: a rejected offer was not below something smaller still);
//   * the target's final state is the accepted offer of smallest distance; its queue entries are all accepted offers, in
//     key order within each destination bucket
// PROVIDED no entry's own state changes while its bucket is drained.  That can happen (an offer may go "backwards" to a
// voxel whose own entry sits later in the same bucket): such a HAZARD is detected, the bucket is cut right before the
// earliest position a hazard touches, the part before the cut is applied -- it is unaffected, by induction over positions --
// and the rest is redone as another ROUND with the updated states.  A round is therefore: stamp the entries' positions,
// compute all offers and their acceptance (read-only), apply them up to the cut, sort the accepted offers by destination
// bucket (stable: key order is kept) straight into the queue pool, un-stamp.  Big rounds are driven by the host (a handful of
// kernels, cub's radix sort for the partition, one synchronisation per round); buckets whose round has a few thousand offers
// at most are drained by ONE resident CTA that walks buckets and rounds by itself (residentKernel below).
//
// The removal flood (removeObstacleVoxels :303-388) is a depth-first walk whose visiting order decides the order of the
// bucket-0 entries it queues, i.e. later ties: it runs as ONE warp (27 lanes = the 27 neighbours of the voxel on top of
// the stack), exactly in the reference's order -- over one pre-classified byte per voxel, since what the walk does with a voxel
// depends only on the state before the walk; its writes are applied afterwards, in parallel.
//
// Signed fields (propagate_negative_) keep a second half of records and a second queue; the same code runs on either half.
//
// This mode exists for parity with the reference's numbers; it costs milliseconds where the exact transform costs
// a tenth of one.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr uint32_t kNone = 0xffffffffu;
constexpr uint16_t kNoKey = 0xffffu;
constexpr int kInitialDirection = 13;  // getDirectionNumber(0, 0, 0)

// neighbourhoods_[n][direction] (:526-582), the slot of a step in them, their sizes
struct NeighbourTables
{
  int8_t nb[2][27][26][3];
  int8_t cnt[2][27];
  int8_t slot[2][27][27];  // [n][source direction][direction number of the step] -> index in nb, or -1
};
__constant__ NeighbourTables c_nb;

NeighbourTables makeTables()
{
  NeighbourTables t;
  std::memset(&t, 0, sizeof(t));
  for (int n = 0; n < 2; ++n)
    for (int dx = -1; dx <= 1; ++dx)
      for (int dy = -1; dy <= 1; ++dy)
        for (int dz = -1; dz <= 1; ++dz)
        {
          const int dn = (dx + 1) * 9 + (dy + 1) * 3 + dz + 1;
          for (int k = 0; k < 27; ++k)
            t.slot[n][dn][k] = -1;
          int cnt = 0;
          for (int tx = -1; tx <= 1; ++tx)
            for (int ty = -1; ty <= 1; ++ty)
              for (int tz = -1; tz <= 1; ++tz)
              {
                if (tx == 0 && ty == 0 && tz == 0)
                  continue;
                if (n >= 1)
                {
                  if (std::abs(tx) + std::abs(ty) + std::abs(tz) != 1)
                    continue;
                  if (dx * tx < 0 || dy * ty < 0 || dz * tz < 0)
                    continue;
                }
                t.nb[n][dn][cnt][0] = static_cast<int8_t>(tx);
                t.nb[n][dn][cnt][1] = static_cast<int8_t>(ty);
                t.nb[n][dn][cnt][2] = static_cast<int8_t>(tz);
                t.slot[n][dn][(tx + 1) * 9 + (ty + 1) * 3 + tz + 1] = static_cast<int8_t>(cnt);
                ++cnt;
              }
          t.cnt[n][dn] = static_cast<int8_t>(cnt);
        }
  return t;
}

// closest point as three 16-bit fields holding coordinate + 1 (0 = the reference's UNINITIALIZED -1)
__host__ __device__ inline uint64_t packPoint(int x, int y, int z)
{
  return (static_cast<uint64_t>(x + 1) << 32) | (static_cast<uint64_t>(y + 1) << 16) | static_cast<uint64_t>(z + 1);
}
constexpr uint64_t kPointMask = 0xffffffffffffull;
__device__ inline void unpackPoint(uint64_t p, int& x, int& y, int& z)
{
  x = static_cast<int>((p >> 32) & 0xffffu) - 1;
  y = static_cast<int>((p >> 16) & 0xffffu) - 1;
  z = static_cast<int>(p & 0xffffu) - 1;
}

struct Dims
{
  int nx, ny, nz;
};
__device__ inline void coordsOf(const Dims& g, uint32_t v, int& x, int& y, int& z)
{
  const uint32_t plane = static_cast<uint32_t>(g.ny) * g.nz;
  x = static_cast<int>(v / plane);
  const uint32_t r = v - static_cast<uint32_t>(x) * plane;
  y = static_cast<int>(r / g.nz);
  z = static_cast<int>(r - static_cast<uint32_t>(y) * g.nz);
}
__device__ inline bool validCell(const Dims& g, int x, int y, int z)
{
  return x >= 0 && y >= 0 && z >= 0 && x < g.nx && y < g.ny && z < g.nz;
}
__device__ inline uint32_t refOf(const Dims& g, int x, int y, int z)
{
  return (static_cast<uint32_t>(x) * g.ny + y) * g.nz + z;
}
__device__ inline long long sq3(int ax, int ay, int az)
{
  return static_cast<long long>(ax) * ax + static_cast<long long>(ay) * ay + static_cast<long long>(az) * az;
}

struct State
{
  int32_t* dist;     // distance_square_
  uint64_t* cp;      // closest_point_
  uint8_t* dir;      // update_direction_
  uint32_t* first;   // position of a voxel's first entry in the round being processed (kNone: none)
  uint32_t* lastp1;  // position + 1 of its last entry (0: none)
};

__global__ void resetKernel(State s, size_t n, int max_d2, Dims g, int negative)
{  // reset() :506-524 and the voxel constructor (propagation_distance_field.hpp:589-598): the negative half starts at
   // distance 0 with every voxel its own closest free voxel
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    if (negative)
    {
      int x, y, z;
      coordsOf(g, static_cast<uint32_t>(i), x, y, z);
      s.dist[i] = 0;
      s.cp[i] = packPoint(x, y, z);
    }
    else
    {
      s.dist[i] = max_d2;
      s.cp[i] = 0;
    }
    s.dir[i] = 0;
    s.first[i] = kNone;
    s.lastp1[i] = 0;
  }
}

struct Seg
{
  unsigned long long start;
  uint32_t len, out;
};
__global__ void gatherKernel(const uint32_t* __restrict__ pool, const Seg* __restrict__ segs, int n_segs, uint32_t* __restrict__ L)
{
  for (int k = blockIdx.y; k < n_segs; k += gridDim.y)
  {
    const Seg sg = segs[k];
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < sg.len; j += gridDim.x * blockDim.x)
      L[sg.out + j] = pool[sg.start + j];
  }
}

// round control block on the device
struct RoundCtl
{
  uint32_t cut;        // earliest position a hazard touches (init: the round's end)
  uint32_t hazards;    // accepted offers that touched a later entry of their own round
  uint32_t backward;   // accepted offers to a bucket at or below the one being drained
  uint32_t pad;
};

// positions of the round's entries; block 0 also resets the round's control block and segment table (read by the kernels
// behind this one)
__global__ void stampKernel(const uint32_t* __restrict__ L, uint32_t a, uint32_t b, State s, RoundCtl* ctl, uint32_t* starts,
                            uint32_t* ends, int n_buckets)
{
  if (blockIdx.x == 0)
  {
    if (threadIdx.x == 0)
      ctl->cut = b;
    for (int j = threadIdx.x; j < n_buckets; j += blockDim.x)
    {
      starts[j] = 0;
      ends[j] = 0;
    }
  }
  const uint32_t p = a + blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= b)
    return;
  const uint32_t v = L[p];
  atomicMin(&s.first[v], p);
  atomicMax(&s.lastp1[v], p + 1);
}
// every offer of entries [a, b) of bucket i: rec_key = its distance if it is accepted (kNoKey otherwise), rec_val = the target,
// rec_cp = the closest point it carries | the direction number of the step << 48.  Reads the state, writes none of it.
// One offer: record idx of the round [a, b) of bucket `bucket` (entry a + idx / S, slot idx % S).
template <int S>
__device__ __forceinline__ void offerBody(size_t idx, const uint32_t* __restrict__ L, uint32_t a, Dims g, int max_d2, const State& s,
                                          uint16_t* __restrict__ rec_key, uint32_t* __restrict__ rec_val,
                                          uint64_t* __restrict__ rec_cp, uint32_t* cut, uint32_t* hazards)
{
  constexpr int D = S == 26 ? 0 : 1;
  const uint32_t p = a + static_cast<uint32_t>(idx / S);
  const int slot = static_cast<int>(idx % S);
  uint16_t key_out = kNoKey;
  const uint32_t v = L[p];
  const int dv = __ldcg(s.dir + v);
  if (dv <= 26 && slot < c_nb.cnt[D][dv])
  {
    const int ex = c_nb.nb[D][dv][slot][0], ey = c_nb.nb[D][dv][slot][1], ez = c_nb.nb[D][dv][slot][2];
    int vx, vy, vz;
    coordsOf(g, v, vx, vy, vz);
    const int tx = vx + ex, ty = vy + ey, tz = vz + ez;
    if (validCell(g, tx, ty, tz))
    {
      const uint64_t cpv = __ldcg(s.cp + v);
      int cx, cy, cz;
      unpackPoint(cpv, cx, cy, cz);
      const long long nd = sq3(cx - tx, cy - ty, cz - tz);
      const uint32_t t = refOf(g, tx, ty, tz);
      if (nd <= max_d2 && nd < __ldcg(s.dist + t))
      {
        bool acc = true;
        const unsigned long long key = static_cast<unsigned long long>(p) * S + slot;
        // the other entries of this round that make an offer to t: all of t's neighbours u = t - step
        for (int k = 0; k < 27 && acc; ++k)
        {
          const int sx = k / 9 - 1, sy = (k / 3) % 3 - 1, sz = k % 3 - 1;
          if (k == 13 || (D == 1 && abs(sx) + abs(sy) + abs(sz) != 1))
            continue;
          const int ux = tx - sx, uy = ty - sy, uz = tz - sz;
          if (!validCell(g, ux, uy, uz))
            continue;
          const uint32_t u = refOf(g, ux, uy, uz);
          const uint32_t fp = __ldcg(s.first + u);
          if (fp == kNone)
            continue;
          const int du = __ldcg(s.dir + u);
          if (du > 26)
            continue;
          const int slot_u = c_nb.slot[D][du][k];
          if (slot_u < 0)
            continue;
          const unsigned long long key_u = static_cast<unsigned long long>(fp) * S + slot_u;
          if (key_u >= key)
            continue;
          int ux2, uy2, uz2;
          unpackPoint(__ldcg(s.cp + u), ux2, uy2, uz2);
          const long long nd_u = sq3(ux2 - tx, uy2 - ty, uz2 - tz);
          if (nd_u <= max_d2 && nd_u <= nd)
            acc = false;
        }
        if (acc)
        {
          key_out = static_cast<uint16_t>(nd);
          rec_val[idx] = t;
          rec_cp[idx] = (cpv & kPointMask) | (static_cast<uint64_t>((ex + 1) * 9 + (ey + 1) * 3 + ez + 1) << 48);
          const uint32_t lp1 = __ldcg(s.lastp1 + t);
          if (lp1 > p + 1)
          {  // t still has an entry behind this one: its state would change under it
            const uint32_t ft = __ldcg(s.first + t);
            atomicMin(cut, ft > p ? ft : p + 1);
            atomicAdd(hazards, 1u);
          }
        }
      }
    }
  }
  rec_key[idx] = key_out;
}


template <int S>
__global__ void __launch_bounds__(256) offerKernel(const uint32_t* __restrict__ L, uint32_t a, uint32_t b, int bucket, Dims g,
                                                   int max_d2, State s, uint16_t* __restrict__ rec_key,
                                                   uint32_t* __restrict__ rec_val, uint64_t* __restrict__ rec_cp, RoundCtl* ctl)
{
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= static_cast<size_t>(b - a) * S)
    return;
  offerBody<S>(idx, L, a, g, max_d2, s, rec_key, rec_val, rec_cp, &ctl->cut, &ctl->hazards);
}

template <int S>
__global__ void applyMinKernel(uint32_t a, size_t n_rec, const uint16_t* __restrict__ rec_key, const uint32_t* __restrict__ rec_val,
                               State s, const RoundCtl* ctl)
{
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= n_rec)
    return;
  const uint16_t k = rec_key[idx];
  if (k == kNoKey || a + static_cast<uint32_t>(idx / S) >= ctl->cut)
    return;
  atomicMin(&s.dist[rec_val[idx]], static_cast<int32_t>(k));
}

// the accepted offer of smallest distance is the target's state; keys become the sort keys of the queue entries
template <int S>
__global__ void applyStateKernel(uint32_t a, size_t n_rec, int bucket, uint16_t* __restrict__ rec_key,
                                 const uint32_t* __restrict__ rec_val, const uint64_t* __restrict__ rec_cp, State s, RoundCtl* ctl)
{
  const size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (idx >= n_rec)
    return;
  const uint16_t k = rec_key[idx];
  if (k == kNoKey)
    return;
  if (a + static_cast<uint32_t>(idx / S) >= ctl->cut)
  {
    rec_key[idx] = kNoKey;  // behind the cut: redone in the next round
    return;
  }
  const uint32_t t = rec_val[idx];
  if (s.dist[t] == static_cast<int32_t>(k))
  {
    const uint64_t r = rec_cp[idx];
    s.cp[t] = r & kPointMask;
    s.dir[t] = static_cast<uint8_t>(r >> 48);
  }
  if (static_cast<int>(k) <= bucket)
    atomicAdd(&ctl->backward, 1u);
  if (static_cast<int>(k) == bucket)
    rec_key[idx] = kNoKey;  // pushed behind the end captured before the loop (:398), cleared with the bucket (:441)
}

// runs of equal keys in the sorted records -> [start, end) per destination bucket
// (L != nullptr: the first b - a threads also clear their entries' stamps -- the round is over)
__global__ void runsKernel(const uint16_t* __restrict__ keys, size_t n, uint32_t* __restrict__ starts, uint32_t* __restrict__ ends,
                           const uint32_t* __restrict__ L = nullptr, uint32_t a = 0, uint32_t b = 0, State s = State{})
{
  const size_t j = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (L != nullptr && j < b - a)
  {
    const uint32_t v = L[a + j];
    s.first[v] = kNone;
    s.lastp1[v] = 0;
  }
  if (j >= n)
    return;
  const uint16_t k = keys[j];
  if (k == kNoKey)
    return;
  if (j == 0 || keys[j - 1] != k)
    starts[k] = static_cast<uint32_t>(j);
  if (j + 1 == n || keys[j + 1] != k)
    ends[k] = static_cast<uint32_t>(j + 1);
}

__global__ void roundInitKernel(RoundCtl* ctl, uint32_t cut, uint32_t* starts, uint32_t* ends, int n_buckets)
{
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0)
    ctl->cut = cut;
  if (j < n_buckets)
  {
    starts[j] = 0;
    ends[j] = 0;
  }
}

// ---- small buckets: the whole loop on the device -------------------------------------------------------------------------
// A round of a few thousand entries is all latency: ~7 launches and a host synchronisation for microseconds of work.  The
// resident kernel is ONE CTA that walks buckets and rounds by itself -- the same phases with __syncthreads between them, the
// queue as a device-side segment table, the stable partition by destination bucket done in shared memory (every warp owns a
// contiguous range of the round's records: per-warp histograms, a scan over (destination, warp), match_any ranks inside a
// warp) -- and hands back to the host at the first bucket that is too big for it (or when the pool or the table is full):
// always BEFORE it touches that bucket, so the host path can take over there.
#ifndef B200SV_RESIDENT_RECORDS
#define B200SV_RESIDENT_RECORDS 4096
#endif
constexpr int kMaxSeg = 256;            // segments per bucket in the device table ([kMaxSeg][n_buckets], column-major)
constexpr int kResidentThreads = 1024;
constexpr int kResidentWarps = kResidentThreads / 32;
constexpr unsigned kResidentRecords = B200SV_RESIDENT_RECORDS;  // records per round it accepts: beyond a few iterations of the CTA per phase ONE SM has too few loads in flight and the host-driven kernels win
constexpr int kResidentMaxBuckets = 1600;     // 32 histograms of that many bins fit shared memory
enum { kResDone = 0, kResTooBig = 1, kResNeedPool = 2, kResSegFull = 3, kResFatal = 4 };
struct ResidentCtl
{
  int status, bucket;
  unsigned long long pool_off, rounds, hazard_rounds, entries;
  unsigned hazards, backward, max_cnt, pad;
};
struct DevQueue
{
  uint32_t* start;  // [kMaxSeg][nb]
  uint32_t* len;    // [kMaxSeg][nb]
  uint32_t* cnt;    // [nb]
};

__global__ void __launch_bounds__(kResidentThreads, 1)
residentKernel(State st, Dims g, int max_d2, uint32_t* __restrict__ pool, unsigned long long pool_cap, DevQueue q, int i0,
               uint32_t* __restrict__ L, uint16_t* __restrict__ rec_key, uint32_t* __restrict__ rec_val,
               uint64_t* __restrict__ rec_cp, ResidentCtl* ctl)
{
  extern __shared__ uint32_t rsm[];
  const int nb = max_d2 + 1;
  uint32_t* hist = rsm;                       // [kResidentWarps][nb]: counts, then running offsets
  uint32_t* tot = hist + kResidentWarps * nb;  // [nb] kept records per destination
  uint32_t* base = tot + nb;                  // [nb] exclusive scan of tot
  uint32_t* sg_start = base + nb;             // [kMaxSeg] the segments of the bucket being drained
  uint32_t* sg_out = sg_start + kMaxSeg;      // [kMaxSeg + 1] where each starts in L
  __shared__ uint32_t s_cut, s_hazards, s_backward, s_maxcnt, s_kept;
  __shared__ int s_dlo, s_dhi;  // destinations this round's kept records go to: only that window of the histograms is live
  __shared__ int s_status;
  __shared__ unsigned long long s_pool_off, s_rounds, s_hazard_rounds, s_entries;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0)
  {
    s_hazards = 0;
    s_backward = 0;
    s_maxcnt = ctl->max_cnt;
    s_pool_off = ctl->pool_off;
    s_rounds = s_hazard_rounds = s_entries = 0;
    s_status = kResDone;
    s_dlo = nb;
    s_dhi = -1;
  }
  for (int k = tid; k < (kResidentWarps + 1) * nb; k += kResidentThreads)
    hist[k] = 0;  // hist and tot: all zero between rounds (a round clears the window it used)
  __syncthreads();
  int i = i0;
  for (; i < nb; ++i)
  {
    const uint32_t cnt = __ldcg(q.cnt + i);
    if (cnt == 0)
      continue;
    // the bucket's segments and where each lands in L
    for (uint32_t k = tid; k < cnt; k += kResidentThreads)
    {
      sg_start[k] = __ldcg(q.start + static_cast<size_t>(k) * nb + i);
      sg_out[k + 1] = __ldcg(q.len + static_cast<size_t>(k) * nb + i);
    }
    __syncthreads();
    if (tid == 0)
    {
      uint32_t out = 0;
      sg_out[0] = 0;
      for (uint32_t k = 0; k < cnt; ++k)
      {
        const uint32_t l = sg_out[k + 1];
        out = (out > 0xffffffffu - l) ? 0xffffffffu : out + l;  // saturating: a huge bucket is "too big", not a wrap-around
        sg_out[k + 1] = out;
      }
      const int S0 = i == 0 ? 26 : 6;
      const unsigned long long need = static_cast<unsigned long long>(out) * S0;
      s_status = need > kResidentRecords ? kResTooBig
                 : s_pool_off + need > pool_cap ? kResNeedPool
                 : s_maxcnt + 8 > static_cast<uint32_t>(kMaxSeg) ? kResSegFull
                                                                 : kResDone;
    }
    __syncthreads();
    if (s_status != kResDone)
      break;  // nothing of bucket i was touched
    const uint32_t total = sg_out[cnt];
    const int S = i == 0 ? 26 : 6;
    for (uint32_t j = tid; j < total; j += kResidentThreads)
    {
      uint32_t lo = 0, hi = cnt;  // the segment that holds entry j: the last k with sg_out[k] <= j
      while (hi - lo > 1)
      {
        const uint32_t mid = (lo + hi) >> 1;
        if (sg_out[mid] <= j)
          lo = mid;
        else
          hi = mid;
      }
      L[j] = __ldcg(pool + sg_start[lo] + (j - sg_out[lo]));
    }
    if (tid == 0)
    {
      q.cnt[i] = 0;  // what is pushed into the bucket while it drains never runs (:398, :441)
      s_entries += total;
    }
    __syncthreads();
    uint32_t a = 0;
    while (a < total)
    {
      const uint32_t b = total;
      const uint32_t n_rec = (b - a) * static_cast<uint32_t>(S);
      if (tid == 0)
      {
        s_cut = b;
        if (s_maxcnt + 1 > static_cast<uint32_t>(kMaxSeg))
          s_status = kResFatal;  // eight rounds of one bucket on a nearly full table: see the check above
      }
      for (uint32_t p = a + tid; p < b; p += kResidentThreads)
      {
        const uint32_t v = L[p];
        atomicMin(&st.first[v], p);
        atomicMax(&st.lastp1[v], p + 1);
      }
      __syncthreads();
      if (s_status == kResFatal)
        break;
      if (S == 26)
        for (uint32_t idx = tid; idx < n_rec; idx += kResidentThreads)
          offerBody<26>(idx, L, a, g, max_d2, st, rec_key, rec_val, rec_cp, &s_cut, &s_hazards);
      else
        for (uint32_t idx = tid; idx < n_rec; idx += kResidentThreads)
          offerBody<6>(idx, L, a, g, max_d2, st, rec_key, rec_val, rec_cp, &s_cut, &s_hazards);
      __syncthreads();
      const uint32_t cut = s_cut;
      for (uint32_t idx = tid; idx < n_rec; idx += kResidentThreads)
      {
        const uint16_t k = rec_key[idx];
        if (k != kNoKey && a + idx / S < cut)
          atomicMin(&st.dist[rec_val[idx]], static_cast<int32_t>(k));
      }
      __syncthreads();
      for (uint32_t idx = tid; idx < n_rec; idx += kResidentThreads)
      {
        const uint16_t k = rec_key[idx];
        if (k == kNoKey)
          continue;
        if (a + idx / S >= cut)
        {
          rec_key[idx] = kNoKey;
          continue;
        }
        const uint32_t t = rec_val[idx];
        if (__ldcg(st.dist + t) == static_cast<int32_t>(k))
        {
          const uint64_t r = rec_cp[idx];
          st.cp[t] = r & kPointMask;
          st.dir[t] = static_cast<uint8_t>(r >> 48);
        }
        if (static_cast<int>(k) <= i)
          atomicAdd(&s_backward, 1u);
        if (static_cast<int>(k) == i)
          rec_key[idx] = kNoKey;
      }
      for (uint32_t p = a + tid; p < b; p += kResidentThreads)
      {
        const uint32_t v = L[p];
        st.first[v] = kNone;
        st.lastp1[v] = 0;
      }
      __syncthreads();
      // stable partition of the kept records by destination: warp w owns records [w * chunk, (w + 1) * chunk)
      const uint32_t chunk = ((n_rec + kResidentWarps - 1) / kResidentWarps + 31u) & ~31u;
      const uint32_t r0 = warp * chunk, r1 = min(n_rec, r0 + chunk);
      {
        int lo = nb, hi = -1;
        for (uint32_t j = r0 + lane; j < r1; j += 32)
        {
          const uint16_t k = rec_key[j];
          if (k != kNoKey)
          {
            atomicAdd(&hist[warp * nb + k], 1u);
            lo = min(lo, static_cast<int>(k));
            hi = max(hi, static_cast<int>(k));
          }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0 && hi >= 0)
        {
          atomicMin(&s_dlo, lo);
          atomicMax(&s_dhi, hi);
        }
      }
      __syncthreads();
      const int dlo = s_dlo, dhi = s_dhi, nd = dhi - dlo + 1;  // nd <= 0: nothing was kept
      for (int d = dlo + tid; d <= dhi; d += kResidentThreads)
      {
        uint32_t run = 0;
        for (int w = 0; w < kResidentWarps; ++w)
        {
          const uint32_t t = hist[w * nb + d];
          hist[w * nb + d] = run;
          run += t;
        }
        tot[d] = run;
      }
      __syncthreads();
      if (warp == 0)
      {  // exclusive scan of tot over the window: every lane sums a contiguous stretch, the stretches are scanned across the warp
        const int per = (max(nd, 0) + 31) / 32, d0 = dlo + lane * per, d1 = min(dhi + 1, d0 + per);
        uint32_t sum = 0;
        for (int d = d0; d < d1; ++d)
          sum += tot[d];
        uint32_t incl = sum;
        for (int o = 1; o < 32; o <<= 1)
        {
          const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o)
            incl += up;
        }
        uint32_t run = incl - sum;
        for (int d = d0; d < d1; ++d)
        {
          base[d] = run;
          run += tot[d];
        }
        if (lane == 31)
          s_kept = incl;
      }
      __syncthreads();
      const unsigned long long pool_off = s_pool_off;
      for (uint32_t j0 = r0; j0 < r1; j0 += 32)
      {
        const uint32_t j = j0 + lane;
        const uint16_t k = j < r1 ? rec_key[j] : kNoKey;
        const bool kept = k != kNoKey;
        const unsigned mask = __match_any_sync(0xffffffffu, kept ? static_cast<unsigned>(k) : 0x10000u + lane);
        if (kept)
        {
          const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
          pool[pool_off + base[k] + hist[warp * nb + k] + rank] = rec_val[j];
        }
        __syncwarp();
        if (kept && lane == __ffs(mask) - 1)
          hist[warp * nb + k] += __popc(mask);
        __syncwarp();
      }
      __syncthreads();
      for (int d = dlo + tid; d <= dhi; d += kResidentThreads)
      {
        const uint32_t n = tot[d];
        if (n)
        {  // one thread per destination: no two threads touch the same bucket's list
          const uint32_t c = q.cnt[d];
          q.start[static_cast<size_t>(c) * nb + d] = static_cast<uint32_t>(pool_off) + base[d];
          q.len[static_cast<size_t>(c) * nb + d] = n;
          q.cnt[d] = c + 1;
          atomicMax(&s_maxcnt, c + 1);
        }
        tot[d] = 0;
      }
      for (int k = tid; k < kResidentWarps * max(nd, 0); k += kResidentThreads)
        hist[(k / nd) * nb + dlo + k % nd] = 0;  // the window is clean again for the next round
      __syncthreads();
      if (tid == 0)
      {
        s_dlo = nb;
        s_dhi = -1;
        s_pool_off = pool_off + s_kept;
        ++s_rounds;
        if (cut < b)
          ++s_hazard_rounds;
      }
      a = cut;
      __syncthreads();
    }
    if (s_status == kResFatal)
      break;
  }
  if (tid == 0)
  {
    ctl->status = s_status;
    ctl->bucket = i;
    ctl->pool_off = s_pool_off;
    ctl->rounds = s_rounds;
    ctl->hazard_rounds = s_hazard_rounds;
    ctl->entries = s_entries;
    ctl->hazards = s_hazards;
    ctl->backward = s_backward;
    ctl->max_cnt = s_maxcnt;
  }
}

// ---- point lists -> voxel lists -----------------------------------------------------------------------------------------
// worldToGrid (voxel_grid.hpp:395-407) in double, as scatterKernel.  mode 0: addPointsToField (:177-196) keeps valid points
// whose voxel is not an obstacle yet; mode 1: removePointsFromField (:198-219) keeps every valid point.
__global__ void pointsToVoxelsKernel(const double* __restrict__ xyz, size_t n, GridDev gd, const int32_t* __restrict__ dist, int mode,
                                     uint16_t* __restrict__ rec_key, uint32_t* __restrict__ rec_val)
{
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const int x = static_cast<int>(floor((xyz[3 * i] - gd.om[0]) * gd.oo));
  const int y = static_cast<int>(floor((xyz[3 * i + 1] - gd.om[1]) * gd.oo));
  const int z = static_cast<int>(floor((xyz[3 * i + 2] - gd.om[2]) * gd.oo));
  uint16_t k = kNoKey;
  uint32_t v = 0;
  if (x >= 0 && y >= 0 && z >= 0 && x < gd.nx && y < gd.ny && z < gd.nz)
  {
    v = (static_cast<uint32_t>(x) * gd.ny + y) * gd.nz + z;
    if (mode == 1 || dist[v] > 0)
      k = 0;
  }
  rec_key[i] = k;
  rec_val[i] = v;
}

// readFromStream (:730-758): the occupied voxels of a bit field in x, y, z order, i.e. by ascending voxel reference
__global__ void bitsToVoxelsKernel(const uint32_t* __restrict__ occ, Dims g, int wz, uint32_t v0, uint32_t n, uint16_t* __restrict__ rec_key,
                                   uint32_t* __restrict__ rec_val)
{
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n)
    return;
  const uint32_t v = v0 + j;
  const uint32_t row = v / g.nz, z = v - row * g.nz;
  const bool set = (occ[static_cast<size_t>(row) * wz + (z >> 5)] >> (z & 31)) & 1u;
  rec_key[j] = set ? static_cast<uint16_t>(0) : kNoKey;
  rec_val[j] = v;
}

// addNewObstacleVoxels :221-243 / removeObstacleVoxels :303-330: the listed voxels' own records
// (uninit: the negative half of a NEW obstacle voxel -- distance max, closest point UNINITIALIZED, direction untouched, :233-238)
__global__ void markVoxelsKernel(const uint32_t* __restrict__ vox, uint32_t n, Dims g, State s, int dist_value, int uninit = 0)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t v = vox[i];
  s.dist[v] = dist_value;
  if (uninit)
  {
    s.cp[v] = 0;
    return;
  }
  int x, y, z;
  coordsOf(g, v, x, y, z);
  s.cp[v] = packPoint(x, y, z);
  s.dir[v] = kInitialDirection;
}

// removeObstacleVoxels' flood (:332-384), one warp, the reference's order.  stack: [sp] entries; out: bucket-0 entries.
struct FloodCtl
{
  unsigned long long sp, out_n;
};
// What the walk does with a voxel it examines depends on nothing the walk itself changes, except "was it reset already":
//   closest point gone (its voxel is no obstacle any more; an uninitialised closest point counts as the voxel itself)
//     and distance != max   -> kReset: reset it and walk on from it, the first time it is examined
//     and distance == max   -> kInert
//   closest point still an obstacle -> kBoundary: queued in bucket 0, every time it is examined
// so every voxel is classified up front, in parallel, and the one warp that must walk in the reference's order reads ONE byte
// per neighbour instead of chasing closest point -> its voxel -> its distance.  What the walk would have written is applied
// afterwards, in parallel again (applyFloodKernel).
enum : uint8_t { kInert = 0, kReset = 1, kBoundary = 2, kWasReset = 3, kSeed = 4 };
__global__ void classifyKernel(State s, Dims g, size_t n, int max_d2, uint8_t* __restrict__ type)
{
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < n; v += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    int cx, cy, cz;
    unpackPoint(s.cp[v], cx, cy, cz);
    const uint32_t cpv = validCell(g, cx, cy, cz) ? refOf(g, cx, cy, cz) : static_cast<uint32_t>(v);
    type[v] = s.dist[cpv] != 0 ? (s.dist[v] != max_d2 ? kReset : kInert) : kBoundary;
  }
}
// the walk's stack holds COORDINATES (packPoint): the one warp that walks would otherwise spend a third of its instructions on
// the two integer divisions that turn a voxel reference into x, y, z
__global__ void seedTypesKernel(const uint32_t* __restrict__ seeds, uint32_t count, Dims g, uint8_t* __restrict__ type,
                                uint64_t* __restrict__ stack)
{
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count)
    return;
  const uint32_t v = seeds[i];
  int x, y, z;
  coordsOf(g, v, x, y, z);
  stack[i] = packPoint(x, y, z);
  type[v] = kSeed;  // the removed / added voxels themselves: inert when examined, but they ARE visited
}
// removeObstacleVoxels' walk (:332-384) and the negative half's after an addition (:255-298): one warp, the reference's order
__global__ void floodKernel(Dims g, uint8_t* __restrict__ type, uint64_t* stack, uint32_t* out, unsigned long long out_cap, FloodCtl* ctl)
{
  const int lane = threadIdx.x;
  unsigned long long sp = ctl->sp, out_n = ctl->out_n;
  const int ex = lane / 9 - 1, ey = (lane / 3) % 3 - 1, ez = lane % 3 - 1;  // direction_number_to_direction_[lane]
  while (sp > 0 && out_n + 27 <= out_cap)
  {
    int lx, ly, lz;
    unpackPoint(stack[sp - 1], lx, ly, lz);
    --sp;
    bool push_stack = false, push_bucket = false;
    uint32_t nv = 0;
    uint64_t nc = 0;
    if (lane < 27)
    {
      const int nx = lx + ex, ny = ly + ey, nz = lz + ez;
      if (validCell(g, nx, ny, nz))
      {
        nv = refOf(g, nx, ny, nz);
        nc = packPoint(nx, ny, nz);
        const uint8_t t = type[nv];
        if (t == kReset)
        {
          type[nv] = kWasReset;
          push_stack = true;
        }
        push_bucket = t == kBoundary;
      }
    }
    const unsigned ms = __ballot_sync(0xffffffffu, push_stack), mb = __ballot_sync(0xffffffffu, push_bucket);
    const unsigned lt = (1u << lane) - 1u;
    if (push_stack)
      stack[sp + __popc(ms & lt)] = nc;
    if (push_bucket)
      out[out_n + __popc(mb & lt)] = nv;
    sp += __popc(ms);
    out_n += __popc(mb);
    __syncwarp();
  }
  if (lane == 0)
  {
    ctl->sp = sp;
    ctl->out_n = out_n;
  }
}
// what the walk wrote on its way: a reset voxel is "max, itself" (positive half: update direction reset too; negative half:
// reset to UNINITIALIZED, then set to itself when the voxel examined itself after being popped, :268-272); an uninitialised
// closest point of ANY voxel examined -- a neighbour of a visited one -- becomes the voxel itself (:350-354)
__global__ void applyFloodKernel(State s, Dims g, size_t n, int max_d2, const uint8_t* __restrict__ type, int neg_style)
{
  for (size_t v = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; v < n; v += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    int x, y, z;
    coordsOf(g, static_cast<uint32_t>(v), x, y, z);
    const uint8_t t = type[v];
    if (t == kWasReset)
    {
      s.dist[v] = max_d2;
      s.cp[v] = packPoint(x, y, z);
      if (!neg_style)
        s.dir[v] = kInitialDirection;
      continue;
    }
    int cx, cy, cz;
    unpackPoint(s.cp[v], cx, cy, cz);
    if (validCell(g, cx, cy, cz))
      continue;
    bool examined = false;
    for (int k = 0; k < 27 && !examined; ++k)
    {
      const int ux = x + k / 9 - 1, uy = y + (k / 3) % 3 - 1, uz = z + k % 3 - 1;
      if (validCell(g, ux, uy, uz))
      {
        const uint8_t tu = type[refOf(g, ux, uy, uz)];
        examined = tu == kWasReset || tu == kSeed;
      }
    }
    if (examined)
      s.cp[v] = packPoint(x, y, z);
  }
}
__global__ void boundaryDirectionKernel(const uint32_t* __restrict__ out, unsigned long long n, State s)
{
  const unsigned long long i = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x;
  if (i < n)
    s.dir[out[i]] = kInitialDirection;  // :374 / :292
}

// ---- updatePointsInField (:130-175): VoxelSets ordered by z, y, x (propagation_distance_field.hpp:60-75) -----------------
__global__ void pointsToSetKeysKernel(const double* __restrict__ xyz, size_t n, GridDev gd, uint32_t* __restrict__ keys, uint32_t* bits)
{
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const int x = static_cast<int>(floor((xyz[3 * i] - gd.om[0]) * gd.oo));
  const int y = static_cast<int>(floor((xyz[3 * i + 1] - gd.om[1]) * gd.oo));
  const int z = static_cast<int>(floor((xyz[3 * i + 2] - gd.om[2]) * gd.oo));
  uint32_t k = kNone;
  if (x >= 0 && y >= 0 && z >= 0 && x < gd.nx && y < gd.ny && z < gd.nz)
  {
    k = (static_cast<uint32_t>(z) * gd.ny + y) * gd.nx + x;
    atomicOr(&bits[k >> 5], 1u << (k & 31));
  }
  keys[i] = k;
}
// sorted set keys -> records: kept iff first of its run, not in the other set, and (filter_obstacles) not an obstacle already
__global__ void setDifferenceKernel(const uint32_t* __restrict__ keys, size_t n, GridDev gd, const uint32_t* __restrict__ other_bits,
                                    const int32_t* __restrict__ dist, int filter_obstacles, uint16_t* __restrict__ rec_key,
                                    uint32_t* __restrict__ rec_val)
{
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n)
    return;
  const uint32_t k = keys[i];
  uint16_t out = kNoKey;
  uint32_t v = 0;
  if (k != kNone && (i == 0 || keys[i - 1] != k) && !((other_bits[k >> 5] >> (k & 31)) & 1u))
  {
    const uint32_t x = k % gd.nx, r = k / gd.nx, y = r % gd.ny, z = r / gd.ny;
    v = (x * gd.ny + y) * gd.nz + z;
    if (!filter_obstacles || dist[v] != 0)
      out = 0;
  }
  rec_key[i] = out;
  rec_val[i] = v;
}

template <typename T>
struct Buf
{
  T* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t n, bool keep = false, size_t keep_n = 0, cudaStream_t s = nullptr)
  {
    if (n <= cap)
      return cudaSuccess;
    // geometric growth either way: the rounds of a first build grow bucket by bucket, and every cudaFree is a device-wide
    // synchronisation
    const size_t want = std::max<size_t>(std::max<size_t>(n, 4096), keep ? cap * 2 : cap + cap / 2);
    T* q = nullptr;
    static const bool trace = getenv("B200SV_PROP_TRACE") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(&q, want * sizeof(T));
    if (trace)
      fprintf(stderr, "[b200sv propagation] buffer of %zu-byte elements grows %zu -> %zu (asked %zu): cudaMalloc %.2f ms\n", sizeof(T), cap,
              want, n, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    if (e != cudaSuccess)
      return e;
    if (p)
    {
      if (keep && keep_n)
      {
        e = cudaMemcpyAsync(q, p, keep_n * sizeof(T), cudaMemcpyDeviceToDevice, s);
        if (e == cudaSuccess)
          e = cudaStreamSynchronize(s);
        if (e != cudaSuccess)
        {
          cudaFree(q);
          return e;
        }
      }
      cudaFree(p);
    }
    p = q;
    cap = want;
    return cudaSuccess;
  }
  void release()
  {
    if (p)
      cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct HostSeg
{
  size_t start;
  uint32_t len;
};
}  // namespace

struct PropField
{
  GridDev gd;
  Dims g;
  size_t n_vox = 0;
  State st{};
  std::vector<std::vector<HostSeg>> buckets;  // bucket_queue_: segments of the pool, in push order
  Buf<uint32_t> pool;                         // queue entries (voxel references)
  size_t pool_off = 0;
  // signed fields (propagate_negative_): the negative half's records and queue.  The functions below always work on the
  // members above; swapHalves() makes the negative half the working one and back.
  bool is_signed = false;
  struct Half
  {
    State st{};
    std::vector<std::vector<HostSeg>> buckets;
    Buf<uint32_t> pool;
    size_t pool_off = 0;
  } neg;
  Buf<uint32_t> L;  // the bucket being drained, contiguous
  Buf<uint16_t> key_in, key_out;
  Buf<uint32_t> val_in, set_keys, set_keys_sorted, stack;
  Buf<uint64_t> rec_cp;
  Buf<uint8_t> sort_tmp;
  Buf<uint32_t> set_bits;
  Buf<uint8_t> flood_type;  // per voxel: what the removal walk does with it (classifyKernel)
  Buf<uint64_t> stack_xyz;  // the walk's stack, as coordinates
  Buf<Seg> d_segs;
  RoundCtl* d_ctl = nullptr;
  FloodCtl* d_flood = nullptr;
  uint32_t *d_starts = nullptr, *d_ends = nullptr;
  // pinned mirror: ctl | starts | ends
  uint8_t* h_pin = nullptr;
  Seg* h_segs = nullptr;
  size_t h_segs_cap = 0;
  PropStats stats{};
  // the resident kernel's side (small buckets): the queue's device-side table, its pinned mirror, the control block
  bool resident_ok = false;
  size_t resident_smem = 0;
  DevQueue dq{};
  uint32_t* h_q = nullptr;  // pinned: start [kMaxSeg][nb] | len [kMaxSeg][nb] | cnt [nb]
  ResidentCtl* d_res = nullptr;
  unsigned long long res_backward = 0, res_hazards = 0;
};

namespace
{
#define PCK(x)                    \
  do                              \
  {                               \
    cudaError_t e_ = (x);         \
    if (e_ != cudaSuccess)        \
      return e_;                  \
  } while (0)

inline unsigned blocksFor(size_t n, int threads) { return static_cast<unsigned>((n + threads - 1) / threads); }

void swapHalves(PropField* f)
{
  std::swap(f->st, f->neg.st);
  f->buckets.swap(f->neg.buckets);
  std::swap(f->pool, f->neg.pool);
  std::swap(f->pool_off, f->neg.pool_off);
}

// stable sort of n records by their 16-bit key; values land at `out_vals`
cudaError_t sortRecords(PropField* f, size_t n, uint32_t* out_vals, cudaStream_t s)
{
  size_t tmp = 0;
  PCK(cub::DeviceRadixSort::SortPairs(nullptr, tmp, f->key_in.p, f->key_out.p, f->val_in.p, out_vals, n, 0, 16, s));
  PCK(f->sort_tmp.ensure(tmp));
  return cub::DeviceRadixSort::SortPairs(f->sort_tmp.p, tmp, f->key_in.p, f->key_out.p, f->val_in.p, out_vals, n, 0, 16, s);
}

cudaError_t ensureRecords(PropField* f, size_t n, bool with_cp)
{
  PCK(f->key_in.ensure(n));
  PCK(f->key_out.ensure(n));
  PCK(f->val_in.ensure(n));
  if (with_cp)
    PCK(f->rec_cp.ensure(n));
  return cudaSuccess;
}

// sorted records -> the count of kept ones (key 0); they sit at out_vals[0 .. count)
cudaError_t keptCount(PropField* f, size_t n, cudaStream_t s, uint32_t* count)
{
  const int nb = f->gd.max_d2 + 1;
  roundInitKernel<<<blocksFor(nb, 256), 256, 0, s>>>(f->d_ctl, 0, f->d_starts, f->d_ends, nb);
  runsKernel<<<blocksFor(n, 256), 256, 0, s>>>(f->key_out.p, n, f->d_starts, f->d_ends);
  PCK(cudaGetLastError());
  PCK(cudaMemcpyAsync(f->h_pin, f->d_ends, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  PCK(cudaStreamSynchronize(s));
  *count = *reinterpret_cast<uint32_t*>(f->h_pin);
  return cudaSuccess;
}

// one bucket, host-driven: rounds of kernels, one synchronisation each
cudaError_t drainBucketHost(PropField* f, int i, cudaStream_t s)
{
  const int nb = f->gd.max_d2 + 1;
  constexpr size_t kMaxRecords = size_t(1) << 25;
  static const bool trace = getenv("B200SV_PROP_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  double marks[6] = { 0, 0, 0, 0, 0, 0 };
  auto mark = [&](int k) {
    if (trace)
      marks[k] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  };
  struct Dump
  {
    const double* m;
    const int i;
    const bool on;
    ~Dump()
    {
      if (on && m[5] > 2.0)
        fprintf(stderr, "[b200sv propagation] bucket %d: buffers %.2f | gather %.2f | round buffers %.2f | kernels queued %.2f | sort queued %.2f | synced %.2f ms\n",
                i, m[0], m[1], m[2], m[3], m[4], m[5]);
    }
  } dump{ marks, i, trace };
  {
    std::vector<HostSeg>& segs = f->buckets[i];
    if (segs.empty())
      return cudaSuccess;
    size_t total = 0;
    for (const HostSeg& sg : segs)
      total += sg.len;
    if (total == 0 || total >= 0xfffffff0ull)
    {
      segs.clear();
      return total ? cudaErrorInvalidValue : cudaSuccess;
    }
    // the bucket, contiguous; its list is empty from here on: what is pushed into it while it drains never runs (:398, :441)
    if (segs.size() > f->h_segs_cap)
    {
      if (f->h_segs)
        cudaFreeHost(f->h_segs);
      f->h_segs = nullptr;
      f->h_segs_cap = 0;
      PCK(cudaMallocHost(&f->h_segs, 2 * segs.size() * sizeof(Seg)));
      f->h_segs_cap = 2 * segs.size();
    }
    PCK(f->d_segs.ensure(segs.size()));
    PCK(f->L.ensure(total));
    mark(0);
    {
      uint32_t out = 0, longest = 0;
      for (size_t k = 0; k < segs.size(); ++k)
      {
        f->h_segs[k] = Seg{ segs[k].start, segs[k].len, out };
        out += segs[k].len;
        longest = std::max(longest, segs[k].len);
      }
      PCK(cudaMemcpyAsync(f->d_segs.p, f->h_segs, segs.size() * sizeof(Seg), cudaMemcpyHostToDevice, s));
      dim3 grid(std::max(1u, std::min(blocksFor(longest, 256), 4096u)), static_cast<unsigned>(std::min<size_t>(segs.size(), 1024)));
      gatherKernel<<<grid, 256, 0, s>>>(f->pool.p, f->d_segs.p, static_cast<int>(segs.size()), f->L.p);
      PCK(cudaGetLastError());
      // (h_segs is rewritten for the next bucket only after this bucket's rounds, each of which ends in a synchronisation)
    }
    segs.clear();
    mark(1);
    f->stats.entries += total;
    const int S = i == 0 ? 26 : 6;
    uint32_t a = 0;
    const uint32_t end = static_cast<uint32_t>(total);
    while (a < end)
    {
      const uint32_t b = static_cast<uint32_t>(std::min<size_t>(end, a + kMaxRecords / S));
      const size_t n_rec = static_cast<size_t>(b - a) * S;
      PCK(ensureRecords(f, n_rec, true));
      PCK(f->pool.ensure(f->pool_off + n_rec, true, f->pool_off, s));
      mark(2);
      stampKernel<<<blocksFor(b - a, 256), 256, 0, s>>>(f->L.p, a, b, f->st, f->d_ctl, f->d_starts, f->d_ends, nb);
      if (S == 26)
      {
        offerKernel<26><<<blocksFor(n_rec, 256), 256, 0, s>>>(f->L.p, a, b, i, f->g, f->gd.max_d2, f->st, f->key_in.p, f->val_in.p,
                                                             f->rec_cp.p, f->d_ctl);
        applyMinKernel<26><<<blocksFor(n_rec, 256), 256, 0, s>>>(a, n_rec, f->key_in.p, f->val_in.p, f->st, f->d_ctl);
        applyStateKernel<26><<<blocksFor(n_rec, 256), 256, 0, s>>>(a, n_rec, i, f->key_in.p, f->val_in.p, f->rec_cp.p, f->st, f->d_ctl);
      }
      else
      {
        offerKernel<6><<<blocksFor(n_rec, 256), 256, 0, s>>>(f->L.p, a, b, i, f->g, f->gd.max_d2, f->st, f->key_in.p, f->val_in.p,
                                                            f->rec_cp.p, f->d_ctl);
        applyMinKernel<6><<<blocksFor(n_rec, 256), 256, 0, s>>>(a, n_rec, f->key_in.p, f->val_in.p, f->st, f->d_ctl);
        applyStateKernel<6><<<blocksFor(n_rec, 256), 256, 0, s>>>(a, n_rec, i, f->key_in.p, f->val_in.p, f->rec_cp.p, f->st, f->d_ctl);
      }
      PCK(cudaGetLastError());
      mark(3);
      PCK(sortRecords(f, n_rec, f->pool.p + f->pool_off, s));
      mark(4);
      runsKernel<<<blocksFor(n_rec, 256), 256, 0, s>>>(f->key_out.p, n_rec, f->d_starts, f->d_ends, f->L.p, a, b, f->st);
      PCK(cudaGetLastError());
      uint8_t* h = f->h_pin;  // ctl | starts | ends are one device allocation: one copy
      PCK(cudaMemcpyAsync(h, f->d_ctl, sizeof(RoundCtl) + 2 * nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
      PCK(cudaStreamSynchronize(s));
      mark(5);
      const RoundCtl* ctl = reinterpret_cast<const RoundCtl*>(h);
      const uint32_t* starts = reinterpret_cast<const uint32_t*>(h + sizeof(RoundCtl));
      const uint32_t* ends = starts + nb;
      uint32_t pushed = 0;
      for (int j = 0; j < nb; ++j)
        if (ends[j] != 0)
        {
          f->buckets[j].push_back(HostSeg{ f->pool_off + starts[j], ends[j] - starts[j] });
          pushed = std::max(pushed, ends[j]);
        }
      f->pool_off += pushed;
      ++f->stats.rounds;
      if (ctl->cut <= a || ctl->cut > b)
        return cudaErrorAssert;  // cannot happen: a hazard's position lies behind the entry that raised it
      if (ctl->cut < b)
        ++f->stats.hazard_rounds;
      a = ctl->cut;
    }
  }
  return cudaSuccess;
}

// ---- the resident kernel's host side --------------------------------------------------------------------------------------
// host lists -> device table (only the columns in use); false: the lists do not fit the table's form right now
bool uploadQueue(PropField* f, cudaStream_t s, uint32_t* max_cnt, cudaError_t* err)
{
  const size_t nb = f->buckets.size();
  size_t maxc = 0;
  for (const auto& b : f->buckets)
    maxc = std::max(maxc, b.size());
  if (maxc + 8 > static_cast<size_t>(kMaxSeg))
    return false;
  uint32_t* h_start = f->h_q;
  uint32_t* h_len = f->h_q + static_cast<size_t>(kMaxSeg) * nb;
  uint32_t* h_cnt = h_len + static_cast<size_t>(kMaxSeg) * nb;
  for (size_t i = 0; i < nb; ++i)
  {
    const auto& b = f->buckets[i];
    h_cnt[i] = static_cast<uint32_t>(b.size());
    for (size_t k = 0; k < b.size(); ++k)
    {
      if (b[k].start + b[k].len > 0xffffffffull)
        return false;
      h_start[k * nb + i] = static_cast<uint32_t>(b[k].start);
      h_len[k * nb + i] = b[k].len;
    }
  }
  *err = cudaMemcpyAsync(f->dq.cnt, h_cnt, nb * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
  if (*err == cudaSuccess && maxc)
    *err = cudaMemcpyAsync(f->dq.start, h_start, maxc * nb * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
  if (*err == cudaSuccess && maxc)
    *err = cudaMemcpyAsync(f->dq.len, h_len, maxc * nb * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
  *max_cnt = static_cast<uint32_t>(maxc);
  return true;
}

cudaError_t downloadQueue(PropField* f, uint32_t max_cnt, cudaStream_t s)
{
  const size_t nb = f->buckets.size();
  uint32_t* h_start = f->h_q;
  uint32_t* h_len = f->h_q + static_cast<size_t>(kMaxSeg) * nb;
  uint32_t* h_cnt = h_len + static_cast<size_t>(kMaxSeg) * nb;
  const size_t cols = std::min<size_t>(max_cnt, kMaxSeg);
  PCK(cudaMemcpyAsync(h_cnt, f->dq.cnt, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  if (cols)
  {
    PCK(cudaMemcpyAsync(h_start, f->dq.start, cols * nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
    PCK(cudaMemcpyAsync(h_len, f->dq.len, cols * nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
  }
  PCK(cudaStreamSynchronize(s));
  for (size_t i = 0; i < nb; ++i)
  {
    auto& b = f->buckets[i];
    b.clear();
    for (uint32_t k = 0; k < h_cnt[i] && k < cols; ++k)
      b.push_back(HostSeg{ h_start[k * nb + i], h_len[k * nb + i] });
  }
  return cudaSuccess;
}

// Buckets from i0 on, on the device, until one is refused or none is left.  *next: the bucket the host has to drain itself
// (n_buckets: none).  *entered false: the queue does not fit the device table right now, nothing was done.
cudaError_t runResident(PropField* f, int i0, cudaStream_t s, int* next, bool* entered)
{
  const int nb = f->gd.max_d2 + 1;
  *entered = false;
  *next = i0;
  uint32_t max_cnt = 0;
  cudaError_t e = cudaSuccess;
  PCK(f->L.ensure(kResidentRecords / 6 + 64));
  PCK(ensureRecords(f, kResidentRecords, true));
  PCK(f->pool.ensure(f->pool_off + 4 * static_cast<size_t>(kResidentRecords), true, f->pool_off, s));
  if (f->pool.cap > 0xffffffffull || !uploadQueue(f, s, &max_cnt, &e))
    return e;
  PCK(e);
  *entered = true;
  ResidentCtl rc{};
  int at = i0;
  for (;;)
  {
    rc = ResidentCtl{};
    rc.pool_off = f->pool_off;
    rc.max_cnt = max_cnt;
    std::memcpy(f->h_pin, &rc, sizeof(rc));
    PCK(cudaMemcpyAsync(f->d_res, f->h_pin, sizeof(rc), cudaMemcpyHostToDevice, s));
    residentKernel<<<1, kResidentThreads, f->resident_smem, s>>>(f->st, f->g, f->gd.max_d2, f->pool.p, f->pool.cap, f->dq, at, f->L.p,
                                                                 f->key_in.p, f->val_in.p, f->rec_cp.p, f->d_res);
    PCK(cudaGetLastError());
    PCK(cudaMemcpyAsync(f->h_pin, f->d_res, sizeof(rc), cudaMemcpyDeviceToHost, s));
    PCK(cudaStreamSynchronize(s));
    std::memcpy(&rc, f->h_pin, sizeof(rc));
    f->pool_off = rc.pool_off;
    f->stats.rounds += rc.rounds;
    f->stats.hazard_rounds += rc.hazard_rounds;
    f->stats.entries += rc.entries;
    f->res_backward += rc.backward;
    f->res_hazards += rc.hazards;
    max_cnt = rc.max_cnt;
    if (rc.status != kResNeedPool)
      break;
    at = rc.bucket;  // grow the pool (the table's offsets stay valid), go on with the bucket it stopped in front of
    PCK(f->pool.ensure(std::max(2 * f->pool.cap, f->pool_off + 8 * static_cast<size_t>(kResidentRecords)), true, f->pool_off, s));
    if (f->pool.cap > 0xffffffffull)
      break;  // beyond 32-bit offsets: the host path takes it from here
  }
  PCK(downloadQueue(f, max_cnt, s));
  if (rc.status == kResFatal)
    return cudaErrorInvalidValue;
  *next = rc.status == kResDone ? nb : rc.bucket;
  return cudaSuccess;
}

cudaError_t propagate(PropField* f, cudaStream_t s)
{
  const int nb = f->gd.max_d2 + 1;
  static const bool trace = getenv("B200SV_PROP_TRACE") != nullptr;
  for (int i = 0; i < nb; ++i)
  {
    const auto t_bucket = std::chrono::steady_clock::now();
    struct Report
    {
      const std::chrono::steady_clock::time_point t0;
      const int i;
      const bool on;
      ~Report()
      {
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (on && ms > 2.0)
          fprintf(stderr, "[b200sv propagation] bucket %d: %.2f ms\n", i, ms);
      }
    } report{ t_bucket, i, trace };
    if (f->buckets[i].empty())
      continue;
    if (f->resident_ok)
    {
      size_t total = 0;
      for (const HostSeg& sg : f->buckets[i])
        total += sg.len;
      if (total * (i == 0 ? 26 : 6) <= kResidentRecords)
      {
        int next = i;
        bool entered = false;
        PCK(runResident(f, i, s, &next, &entered));
        if (entered)
        {
          if (next >= nb)
            break;
          if (next > i)
          {
            i = next - 1;  // the loop's ++i lands on the bucket the device refused; its size is looked at again there
            continue;
          }
          // refused the very bucket it was started on (a full pool beyond 32 bits, a full table): the host drains it
        }
      }
    }
    PCK(drainBucketHost(f, i, s));
  }
  {  // counters, and the pool: entries pushed "backwards" stay queued for the next call (rare), everything else is spent
    PCK(cudaMemcpyAsync(f->h_pin, f->d_ctl, sizeof(RoundCtl), cudaMemcpyDeviceToHost, s));
    PCK(cudaStreamSynchronize(s));
    f->stats.backward_offers = reinterpret_cast<const RoundCtl*>(f->h_pin)->backward + f->res_backward;
    f->stats.hazard_offers = reinterpret_cast<const RoundCtl*>(f->h_pin)->hazards + f->res_hazards;
    size_t left = 0;
    for (const auto& b : f->buckets)
      for (const HostSeg& sg : b)
        left += sg.len;
    f->stats.queued_for_next_call = left;
    if (left == 0)
      f->pool_off = 0;
    else
    {
      PCK(f->L.ensure(left));
      size_t out = 0;
      for (auto& b : f->buckets)
        for (HostSeg& sg : b)
        {
          PCK(cudaMemcpyAsync(f->L.p + out, f->pool.p + sg.start, sg.len * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
          sg.start = out;
          out += sg.len;
        }
      PCK(cudaMemcpyAsync(f->pool.p, f->L.p, left * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      PCK(cudaStreamSynchronize(s));
      f->pool_off = left;
    }
  }
  return cudaSuccess;
}

// the flood from the `count` voxels at the bottom of f->stack over the WORKING half; the bucket-0 entries it queues go to the
// working pool
cudaError_t floodFromStack(PropField* f, uint32_t count, int neg_style, cudaStream_t s)
{
  FloodCtl fc{ count, 0 };
  const size_t seg_start = f->pool_off;
  if (count)
  {
    PCK(f->flood_type.ensure(f->n_vox));
    PCK(f->stack_xyz.ensure(static_cast<size_t>(count) + f->n_vox + 32));  // every voxel is pushed at most once more
    classifyKernel<<<148 * 8, 256, 0, s>>>(f->st, f->g, f->n_vox, f->gd.max_d2, f->flood_type.p);
    seedTypesKernel<<<blocksFor(count, 256), 256, 0, s>>>(f->stack.p, count, f->g, f->flood_type.p, f->stack_xyz.p);
    PCK(cudaGetLastError());
  }
  while (fc.sp > 0)
  {
    PCK(f->pool.ensure(f->pool_off + std::max<size_t>(size_t(1) << 20, 64 * fc.sp), true, f->pool_off, s));
    const unsigned long long room = f->pool.cap - f->pool_off;
    fc.out_n = 0;
    PCK(cudaMemcpyAsync(f->d_flood, &fc, sizeof(fc), cudaMemcpyHostToDevice, s));
    floodKernel<<<1, 32, 0, s>>>(f->g, f->flood_type.p, f->stack_xyz.p, f->pool.p + f->pool_off, room, f->d_flood);
    PCK(cudaGetLastError());
    PCK(cudaMemcpyAsync(f->h_pin, f->d_flood, sizeof(fc), cudaMemcpyDeviceToHost, s));
    PCK(cudaStreamSynchronize(s));
    std::memcpy(&fc, f->h_pin, sizeof(fc));
    if (fc.out_n)
      boundaryDirectionKernel<<<blocksFor(fc.out_n, 256), 256, 0, s>>>(f->pool.p + f->pool_off, fc.out_n, f->st);
    f->pool_off += fc.out_n;
    f->stats.flood_pops_batches++;
  }
  if (count)
  {
    applyFloodKernel<<<148 * 8, 256, 0, s>>>(f->st, f->g, f->n_vox, f->gd.max_d2, f->flood_type.p, neg_style);
    PCK(cudaGetLastError());
  }
  // one segment, or several of at most 2^31 - 1 entries
  for (size_t at = seg_start; at < f->pool_off;)
  {
    const uint32_t len = static_cast<uint32_t>(std::min<size_t>(f->pool_off - at, 0x7fffffffu));
    f->buckets[0].push_back(HostSeg{ at, len });
    at += len;
  }
  return cudaSuccess;
}

// addNewObstacleVoxels (:221-301) for `count` voxels already at pool[pool_off ..)
cudaError_t addVoxelsAtPool(PropField* f, uint32_t count, cudaStream_t s)
{
  if (f->is_signed)
    PCK(f->stack.ensure(static_cast<size_t>(count) + 32));
  if (count)
  {
    if (f->is_signed)
    {  // their negative halves: distance max, closest free voxel unknown; they seed the negative flood below (:233-238)
      PCK(cudaMemcpyAsync(f->stack.p, f->pool.p + f->pool_off, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      markVoxelsKernel<<<blocksFor(count, 256), 256, 0, s>>>(f->stack.p, count, f->g, f->neg.st, f->gd.max_d2, 1);
    }
    markVoxelsKernel<<<blocksFor(count, 256), 256, 0, s>>>(f->pool.p + f->pool_off, count, f->g, f->st, 0);
    PCK(cudaGetLastError());
    f->buckets[0].push_back(HostSeg{ f->pool_off, count });
    f->pool_off += count;
  }
  PCK(propagate(f, s));
  if (!f->is_signed)
    return cudaSuccess;
  swapHalves(f);  // :255-300: the voxels whose closest free voxel is gone, then propagateNegative
  cudaError_t e = floodFromStack(f, count, 1, s);
  if (e == cudaSuccess)
    e = propagate(f, s);
  swapHalves(f);
  return e;
}

// removeObstacleVoxels (:303-388) for `count` voxels at pool[pool_off ..) (not consumed as queue entries)
cudaError_t removeVoxelsAtPool(PropField* f, uint32_t count, cudaStream_t s)
{
  PCK(f->stack.ensure(static_cast<size_t>(count) + 32));
  if (count)
  {
    PCK(cudaMemcpyAsync(f->stack.p, f->pool.p + f->pool_off, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
    markVoxelsKernel<<<blocksFor(count, 256), 256, 0, s>>>(f->stack.p, count, f->g, f->st, f->gd.max_d2);
    PCK(cudaGetLastError());
    if (f->is_signed)
    {  // a removed voxel is free again: negative distance 0, queued in the negative bucket 0 (:321-328)
      PCK(f->neg.pool.ensure(f->neg.pool_off + count, true, f->neg.pool_off, s));
      PCK(cudaMemcpyAsync(f->neg.pool.p + f->neg.pool_off, f->stack.p, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      markVoxelsKernel<<<blocksFor(count, 256), 256, 0, s>>>(f->stack.p, count, f->g, f->neg.st, 0);
      PCK(cudaGetLastError());
      f->neg.buckets[0].push_back(HostSeg{ f->neg.pool_off, count });
      f->neg.pool_off += count;
    }
  }
  PCK(floodFromStack(f, count, 0, s));
  PCK(propagate(f, s));
  if (!f->is_signed)
    return cudaSuccess;
  swapHalves(f);
  const cudaError_t e = propagate(f, s);
  swapHalves(f);
  return e;
}

cudaError_t pointsToPool(PropField* f, const double* d_xyz, size_t n, int mode, cudaStream_t s, uint32_t* count)
{
  *count = 0;
  if (n == 0)
    return cudaSuccess;
  PCK(ensureRecords(f, n, false));
  PCK(f->pool.ensure(f->pool_off + n, true, f->pool_off, s));
  pointsToVoxelsKernel<<<blocksFor(n, 256), 256, 0, s>>>(d_xyz, n, f->gd, f->st.dist, mode, f->key_in.p, f->val_in.p);
  PCK(cudaGetLastError());
  PCK(sortRecords(f, n, f->pool.p + f->pool_off, s));
  return keptCount(f, n, s, count);
}
}  // namespace

cudaError_t propCreate(const GridDev& gd, bool is_signed, PropField** out)
{
  *out = nullptr;
  if (gd.nx > 65534 || gd.ny > 65534 || gd.nz > 65534 || gd.max_d2 >= 0xffff)
    return cudaErrorInvalidConfiguration;
  PropField* f = new PropField();
  f->gd = gd;
  f->g = Dims{ gd.nx, gd.ny, gd.nz };
  f->n_vox = static_cast<size_t>(gd.nx) * gd.ny * gd.nz;
  f->buckets.resize(gd.max_d2 + 1);
  const NeighbourTables t = makeTables();
  cudaError_t e = cudaMemcpyToSymbol(c_nb, &t, sizeof(t));
  const int nb = gd.max_d2 + 1;
  if (e == cudaSuccess) e = cudaMalloc(&f->st.dist, f->n_vox * sizeof(int32_t));
  if (e == cudaSuccess) e = cudaMalloc(&f->st.cp, f->n_vox * sizeof(uint64_t));
  if (e == cudaSuccess) e = cudaMalloc(&f->st.dir, f->n_vox);
  if (e == cudaSuccess) e = cudaMalloc(&f->st.first, f->n_vox * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMalloc(&f->st.lastp1, f->n_vox * sizeof(uint32_t));
  f->is_signed = is_signed;
  if (is_signed)
  {
    f->neg.buckets.resize(gd.max_d2 + 1);
    if (e == cudaSuccess) e = cudaMalloc(&f->neg.st.dist, f->n_vox * sizeof(int32_t));
    if (e == cudaSuccess) e = cudaMalloc(&f->neg.st.cp, f->n_vox * sizeof(uint64_t));
    if (e == cudaSuccess) e = cudaMalloc(&f->neg.st.dir, f->n_vox);
    f->neg.st.first = f->st.first;  // the stamps belong to the round in flight: one set serves both halves
    f->neg.st.lastp1 = f->st.lastp1;
  }
  if (e == cudaSuccess) e = cudaMalloc(&f->d_ctl, sizeof(RoundCtl) + 2 * nb * sizeof(uint32_t));  // ctl | starts | ends
  if (e == cudaSuccess) e = cudaMemset(f->d_ctl, 0, sizeof(RoundCtl) + 2 * nb * sizeof(uint32_t));
  if (e == cudaSuccess)
  {
    f->d_starts = reinterpret_cast<uint32_t*>(f->d_ctl + 1);
    f->d_ends = f->d_starts + nb;
  }
  if (e == cudaSuccess) e = cudaMalloc(&f->d_flood, sizeof(FloodCtl));
  if (e == cudaSuccess) e = cudaMallocHost(&f->h_pin, sizeof(RoundCtl) + 2 * nb * sizeof(uint32_t) + 64);
  // the resident kernel (small buckets): grids of at most kResidentMaxBuckets buckets, unless B200SV_PROP_RESIDENT=0
  static const bool resident_off = getenv("B200SV_PROP_RESIDENT") && atoi(getenv("B200SV_PROP_RESIDENT")) == 0;
  if (e == cudaSuccess && !resident_off && nb <= kResidentMaxBuckets)
  {
    f->resident_smem = (static_cast<size_t>(kResidentWarps + 2) * nb + 2 * kMaxSeg + 1) * sizeof(uint32_t);
    const size_t qn = static_cast<size_t>(kMaxSeg) * nb;
    if (cudaFuncSetAttribute(residentKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(f->resident_smem)) ==
        cudaSuccess)
    {
      e = cudaMalloc(&f->dq.start, qn * sizeof(uint32_t));
      if (e == cudaSuccess) e = cudaMalloc(&f->dq.len, qn * sizeof(uint32_t));
      if (e == cudaSuccess) e = cudaMalloc(&f->dq.cnt, nb * sizeof(uint32_t));
      if (e == cudaSuccess) e = cudaMalloc(&f->d_res, sizeof(ResidentCtl));
      if (e == cudaSuccess) e = cudaMallocHost(&f->h_q, (2 * qn + nb) * sizeof(uint32_t));
      f->resident_ok = e == cudaSuccess;
    }
    else
      cudaGetLastError();  // not enough shared memory on this device: the host path does everything
  }
  if (e != cudaSuccess)
  {
    propDestroy(f);
    return e;
  }
  *out = f;
  return cudaSuccess;
}

void propDestroy(PropField* f)
{
  if (!f)
    return;
  cudaFree(f->st.dist);
  cudaFree(f->st.cp);
  cudaFree(f->st.dir);
  cudaFree(f->st.first);
  cudaFree(f->st.lastp1);
  cudaFree(f->neg.st.dist);
  cudaFree(f->neg.st.cp);
  cudaFree(f->neg.st.dir);
  f->neg.pool.release();
  cudaFree(f->d_ctl);
  cudaFree(f->d_flood);
  cudaFree(f->dq.start);
  cudaFree(f->dq.len);
  cudaFree(f->dq.cnt);
  cudaFree(f->d_res);
  if (f->h_q)
    cudaFreeHost(f->h_q);
  if (f->h_pin)
    cudaFreeHost(f->h_pin);
  if (f->h_segs)
    cudaFreeHost(f->h_segs);
  f->pool.release();
  f->L.release();
  f->key_in.release();
  f->key_out.release();
  f->val_in.release();
  f->set_keys.release();
  f->set_keys_sorted.release();
  f->stack.release();
  f->rec_cp.release();
  f->sort_tmp.release();
  f->set_bits.release();
  f->flood_type.release();
  f->stack_xyz.release();
  f->d_segs.release();
  delete f;
}

cudaError_t propReset(PropField* f, cudaStream_t s)
{  // reset() re-initialises the voxels; the bucket queues are not its business (entries left there stay queued)
  resetKernel<<<148 * 8, 256, 0, s>>>(f->st, f->n_vox, f->gd.max_d2, f->g, 0);
  if (f->is_signed)
    resetKernel<<<148 * 8, 256, 0, s>>>(f->neg.st, f->n_vox, f->gd.max_d2, f->g, 1);
  return cudaGetLastError();
}

cudaError_t propAddPoints(PropField* f, const double* d_xyz, size_t n, cudaStream_t s)
{
  uint32_t count = 0;
  PCK(pointsToPool(f, d_xyz, n, 0, s, &count));
  return addVoxelsAtPool(f, count, s);
}

cudaError_t propRemovePoints(PropField* f, const double* d_xyz, size_t n, cudaStream_t s)
{
  uint32_t count = 0;
  PCK(pointsToPool(f, d_xyz, n, 1, s, &count));
  return removeVoxelsAtPool(f, count, s);
}

cudaError_t propUpdatePoints(PropField* f, const double* d_old, size_t n_old, const double* d_new, size_t n_new, cudaStream_t s)
{
  const size_t n_words = (f->n_vox + 31) / 32;
  PCK(f->set_bits.ensure(2 * n_words));
  PCK(cudaMemsetAsync(f->set_bits.p, 0, 2 * n_words * sizeof(uint32_t), s));
  uint32_t* bits[2] = { f->set_bits.p, f->set_bits.p + n_words };
  const double* pts[2] = { d_old, d_new };
  const size_t cnt[2] = { n_old, n_new };
  const size_t n_max = std::max<size_t>(std::max(n_old, n_new), 1);
  PCK(f->set_keys.ensure(2 * n_max));
  PCK(f->set_keys_sorted.ensure(2 * n_max));
  for (int w = 0; w < 2; ++w)
    if (cnt[w])
    {
      pointsToSetKeysKernel<<<blocksFor(cnt[w], 256), 256, 0, s>>>(pts[w], cnt[w], f->gd, f->set_keys.p + w * n_max, bits[w]);
      PCK(cudaGetLastError());
      size_t tmp = 0;
      PCK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, f->set_keys.p + w * n_max, f->set_keys_sorted.p + w * n_max, cnt[w], 0, 32, s));
      PCK(f->sort_tmp.ensure(tmp));
      PCK(cub::DeviceRadixSort::SortKeys(f->sort_tmp.p, tmp, f->set_keys.p + w * n_max, f->set_keys_sorted.p + w * n_max, cnt[w], 0, 32, s));
    }
  // both differences are taken (and new \ old filtered by "is no obstacle yet") BEFORE anything is removed (:156-172)
  uint32_t n_rem = 0, n_add = 0;
  PCK(ensureRecords(f, n_max, false));
  PCK(f->pool.ensure(f->pool_off + n_old + n_new + 64, true, f->pool_off, s));
  if (n_old)
  {
    setDifferenceKernel<<<blocksFor(n_old, 256), 256, 0, s>>>(f->set_keys_sorted.p, n_old, f->gd, bits[1], f->st.dist, 0, f->key_in.p,
                                                             f->val_in.p);
    PCK(cudaGetLastError());
    PCK(sortRecords(f, n_old, f->pool.p + f->pool_off, s));
    PCK(keptCount(f, n_old, s, &n_rem));
  }
  uint32_t* add_list = nullptr;
  if (n_new)
  {
    setDifferenceKernel<<<blocksFor(n_new, 256), 256, 0, s>>>(f->set_keys_sorted.p + n_max, n_new, f->gd, bits[0], f->st.dist, 1,
                                                             f->key_in.p, f->val_in.p);
    PCK(cudaGetLastError());
    // parked in the set-key buffer while the removal uses the pool
    add_list = f->set_keys.p;
    PCK(sortRecords(f, n_new, add_list, s));
    PCK(keptCount(f, n_new, s, &n_add));
  }
  PCK(removeVoxelsAtPool(f, n_rem, s));
  PCK(f->pool.ensure(f->pool_off + n_add + 64, true, f->pool_off, s));
  if (n_add)
    PCK(cudaMemcpyAsync(f->pool.p + f->pool_off, add_list, n_add * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
  return addVoxelsAtPool(f, n_add, s);
}

cudaError_t propAddOccupancy(PropField* f, const uint32_t* d_occ, cudaStream_t s)
{
  constexpr size_t kChunk = size_t(1) << 25;
  uint32_t total = 0;
  for (size_t v0 = 0; v0 < f->n_vox; v0 += kChunk)
  {
    const size_t n = std::min(kChunk, f->n_vox - v0);
    uint32_t count = 0;
    PCK(ensureRecords(f, n, false));
    PCK(f->pool.ensure(f->pool_off + total + n, true, f->pool_off + total, s));
    bitsToVoxelsKernel<<<blocksFor(n, 256), 256, 0, s>>>(d_occ, f->g, f->gd.wz, static_cast<uint32_t>(v0), static_cast<uint32_t>(n),
                                                         f->key_in.p, f->val_in.p);
    PCK(cudaGetLastError());
    PCK(sortRecords(f, n, f->pool.p + f->pool_off + total, s));
    PCK(keptCount(f, n, s, &count));
    total += count;
  }
  return addVoxelsAtPool(f, total, s);  // addNewObstacleVoxels: no "already an obstacle" filter on this path
}

const int32_t* propDistances(const PropField* f) { return f->st.dist; }
const int32_t* propNegativeDistances(const PropField* f) { return f->is_signed ? f->neg.st.dist : nullptr; }
const PropStats& propStats(const PropField* f) { return f->stats; }
}  // namespace b200sv
