// Batched edge ("motion") validation: the consumer-side row of SURVEY.md section 8f.
//
// MoveIt has no MotionValidator of its own; OMPL's DiscreteMotionValidator is configured over ModelBasedStateSpace
// (moveit_planners/ompl/ompl_interface/src/model_based_planning_context.cpp:294-317 sets
// longest_valid_segment_fraction).  For an edge (s1, s2) it checks s2, then the interpolated states at j / nd,
// j = 1 .. nd - 1, with nd = StateSpace::validSegmentCount(s1, s2) = ceil(distance(s1, s2) / longest valid segment):
//   distance     JointModelGroup::distance (robot_model/src/joint_model_group.cpp:462-472): sum over the active joints
//                of factor * |a - b| (continuous revolute: wrapped to [0, pi], revolute_joint_model.cpp:173-182)
//   interpolate  JointModelGroup::interpolate (:474-486): per active variable from + (to - from) * t, continuous
//                revolute joints along the shorter arc (revolute_joint_model.cpp:138-171)
// OMPL itself is not vendored in the reference tree (moveit_planners/ompl/package.xml:25, no version pin): the
// discretisation is restated from its published algorithm and pinned only by our own tests (parity unpinned).
//
// Three small kernels around the state check: count the states of every edge, expand the edges into one state batch
// (FP64, no FMA contraction: the same roundings as the host expression), and after the check reduce every edge's bit
// range to (valid, index of the first invalid state).
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr double kPi = 3.14159265358979323846;

// states to check for an edge: j = 1 .. nd - 1 interpolated, then s2 itself  =>  max(nd, 1)
__global__ void motionCountKernel(MotionArgs a)
{
  for (size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; e < a.n_edges;
       e += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const double* f = a.q_from + e * a.n_vars;
    const double* t = a.q_to + e * a.n_vars;
    double d = 0.0;
    for (int k = 0; k < a.n_vars; ++k)
    {
      const double w = a.factor ? a.factor[k] : 1.0;
      if (w == 0.0)
        continue;  // mimic / passive variables do not contribute (not in active_joint_model_vector_)
      double dk = fabs(f[k] - t[k]);
      if (a.continuous && a.continuous[k])
      {
        dk = fmod(dk, 2.0 * kPi);
        dk = dk > kPi ? 2.0 * kPi - dk : dk;
      }
      d = __dadd_rn(d, __dmul_rn(w, dk));
    }
    const double nd = ceil(d / a.max_segment);
    unsigned c = nd < 1.0 ? 1u : (nd > 4.0e9 ? 0xffffffffu : static_cast<unsigned>(nd));
    if (a.max_states_per_edge && c > a.max_states_per_edge)
      c = a.max_states_per_edge;  // guarded by the caller's cap (documented): an edge this long is cut into fewer steps
    a.count[e] = c;
  }
}

// single-CTA exclusive scan of count[0..n) into offset[0..n]; n is a batch of edges, not of states.  Every WARP owns a
// contiguous stretch of the counts and walks it 32 at a time, so every load and store is one coalesced 128 / 256-byte request
// (a first version gave every THREAD a contiguous run: 32 lines per request, 98 us for 100 000 edges under ncu): pass 1 sums
// the stretches, the 32 sums are scanned once, pass 2 scans each stretch with shuffles and a running carry.
__global__ void __launch_bounds__(1024) motionScanKernel(const unsigned* __restrict__ count, unsigned long long* __restrict__ offset,
                                                         size_t n)
{
  __shared__ unsigned long long warp_sums[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warps = blockDim.x >> 5;
  const size_t per = ((n + warps - 1) / warps + 31) & ~static_cast<size_t>(31);  // a stretch is whole groups of 32
  const size_t i0 = min(n, warp * per), i1 = min(n, i0 + per);
  unsigned long long mine = 0ull;
  for (size_t i = i0 + lane; i < i1; i += 32)
    mine += count[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    mine += __shfl_xor_sync(0xffffffffu, mine, d);
  if (lane == 0)
    warp_sums[warp] = mine;
  __syncthreads();
  if (warp == 0)
  {
    unsigned long long s = lane < warps ? warp_sums[lane] : 0ull;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, s, d);
      if (lane >= d)
        s += y;
    }
    warp_sums[lane] = s;  // inclusive
  }
  __syncthreads();
  unsigned long long carry = warp ? warp_sums[warp - 1] : 0ull;
  for (size_t base = i0; base < i1; base += 32)
  {
    const size_t i = base + lane;
    const unsigned long long c = i < i1 ? count[i] : 0ull;
    unsigned long long x = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long y = __shfl_up_sync(0xffffffffu, x, d);
      if (lane >= d)
        x += y;
    }
    if (i < i1)
      offset[i] = carry + x - c;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (threadIdx.x == 0)
    offset[n] = warp_sums[warps - 1];
}

// one WARP per edge, one lane per generated state (no search for the state's edge; the lanes of a warp write one contiguous
// run of the batch): JointModelGroup::interpolate
__global__ void __launch_bounds__(256) motionExpandKernel(MotionArgs a, double* __restrict__ q_out, unsigned long long n_states)
{
  const int lane = threadIdx.x & 31;
  const size_t warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
  for (size_t e = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) >> 5; e < a.n_edges; e += warps)
  {
    const unsigned c = a.count[e];
    const unsigned long long s0 = a.offset[e];
    const double* f = a.q_from + e * a.n_vars;
    const double* t = a.q_to + e * a.n_vars;
    for (unsigned j = 1u + lane; j <= c; j += 32u)  // 1 .. c
    {
      const unsigned long long s = s0 + (j - 1u);
      if (s >= n_states)
        break;
      double* o = q_out + s * a.n_vars;
      if (j == c)
      {
        for (int k = 0; k < a.n_vars; ++k)
          o[k] = t[k];  // s2 itself
        continue;
      }
      const double tt = static_cast<double>(j) / static_cast<double>(c);
      for (int k = 0; k < a.n_vars; ++k)
      {
        double diff = __dsub_rn(t[k], f[k]);
        double v;
        if (a.continuous && a.continuous[k] && fabs(diff) > kPi)
        {  // revolute_joint_model.cpp:147-167
          diff = diff > 0.0 ? __dsub_rn(2.0 * kPi, diff) : __dsub_rn(-2.0 * kPi, diff);
          v = __dsub_rn(f[k], __dmul_rn(diff, tt));
          if (v > kPi)
            v = __dsub_rn(v, 2.0 * kPi);
          else if (v < -kPi)
            v = __dadd_rn(v, 2.0 * kPi);
        }
        else
          v = __dadd_rn(f[k], __dmul_rn(diff, tt));
        o[k] = v;
      }
    }
  }
}

// one thread per edge: first 0 bit of its range of the state bitmap
__global__ void motionReduceKernel(MotionArgs a, const uint32_t* __restrict__ state_bits, uint32_t* __restrict__ edge_bits,
                                   int32_t* __restrict__ first_invalid)
{
  const size_t e = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  const bool in = e < a.n_edges;
  int bad = -1;
  if (in)
  {
    const unsigned long long s0 = a.offset[e], s1 = a.offset[e + 1];
    for (unsigned long long w = s0 >> 5; w <= (s1 - 1) >> 5 && bad < 0; ++w)
    {
      uint32_t inv = ~state_bits[w];
      if (w == (s0 >> 5))
        inv &= 0xffffffffu << (s0 & 31);
      if (w == ((s1 - 1) >> 5) && ((s1 & 31) != 0))
        inv &= 0xffffffffu >> (32 - (s1 & 31));
      if (inv)
        bad = static_cast<int>((w << 5) + (__ffs(inv) - 1) - s0) + 1;  // j of the first invalid state, 1 .. count
    }
    if (first_invalid)
      first_invalid[e] = bad;
  }
  const unsigned word = __ballot_sync(0xffffffffu, in && bad < 0);
  if ((threadIdx.x & 31) == 0 && (e >> 5) < (a.n_edges + 31) / 32)
    edge_bits[e >> 5] = word;
}
}  // namespace

cudaError_t launchMotionCount(const MotionArgs& a, cudaStream_t s)
{
  if (a.n_edges == 0)
    return cudaSuccess;
  const int grid = static_cast<int>((a.n_edges + 255) / 256 < 4096 ? (a.n_edges + 255) / 256 : 4096);
  motionCountKernel<<<grid, 256, 0, s>>>(a);
  motionScanKernel<<<1, 1024, 0, s>>>(a.count, a.offset, a.n_edges);
  return cudaGetLastError();
}

cudaError_t launchMotionExpand(const MotionArgs& a, double* d_q, unsigned long long n_states, cudaStream_t s)
{
  if (n_states == 0)
    return cudaSuccess;
  const unsigned long long blocks = (static_cast<unsigned long long>(a.n_edges) + 7) / 8;  // 8 warps = 8 edges per CTA
  motionExpandKernel<<<static_cast<int>(blocks < 148ull * 64 ? (blocks ? blocks : 1) : 148ull * 64), 256, 0, s>>>(a, d_q, n_states);
  return cudaGetLastError();
}

cudaError_t launchMotionReduce(const MotionArgs& a, const uint32_t* state_bits, uint32_t* edge_bits, int32_t* first_invalid,
                               cudaStream_t s)
{
  if (a.n_edges == 0)
    return cudaSuccess;
  motionReduceKernel<<<static_cast<int>((a.n_edges + 255) / 256), 256, 0, s>>>(a, state_bits, edge_bits, first_invalid);
  return cudaGetLastError();
}
}  // namespace b200sv
