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

#include <cstdint>

#include "kernels.hpp"

namespace b200sv
{
namespace
{
constexpr uint16_t kInf16 = 0xFFFF;

__global__ void scatterKernel(const double* __restrict__ xyz, size_t n, GridDev g, uint32_t* __restrict__ occ, int set)
{
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    // VoxelGrid::getCellFromLocation (voxel_grid.hpp:522), in double exactly as the reference
    const int x = static_cast<int>(floor((xyz[3 * i + 0] - g.om[0]) * g.oo));
    const int y = static_cast<int>(floor((xyz[3 * i + 1] - g.om[1]) * g.oo));
    const int z = static_cast<int>(floor((xyz[3 * i + 2] - g.om[2]) * g.oo));
    if (x < 0 || y < 0 || z < 0 || x >= g.nx || y >= g.ny || z >= g.nz)
      continue;  // isCellValid (voxel_grid.hpp:405-408)
    uint32_t* w = occ + (static_cast<size_t>(x) * g.ny + y) * g.wz + (z >> 5);
    if (set)
      atomicOr(w, 1u << (z & 31));
    else
      atomicAnd(w, ~(1u << (z & 31)));
  }
}

// distance (in cells) from z to the nearest set bit of `row` within +-R; R + 1 if there is none
__device__ __forceinline__ int zDistance(const uint32_t* row, int wz, int z, int R)
{
  const int w = z >> 5, b = z & 31;
  int best = R + 1;
  const uint32_t cur = row[w] >> b;
  if (cur)
    best = __ffs(cur) - 1;
  else
  {
    int off = 32 - b;
    for (int ww = w + 1; ww < wz && off <= R; ++ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        best = off + __ffs(v) - 1;
        break;
      }
    }
  }
  const uint32_t low = b ? (row[w] & ((1u << b) - 1u)) : 0u;
  if (low)
    best = min(best, b - (31 - __clz(low)));
  else
  {
    int off = b + 1;
    for (int ww = w - 1; ww >= 0 && off <= R; --ww, off += 32)
    {
      const uint32_t v = row[ww];
      if (v)
      {
        best = min(best, off + __clz(v));
        break;
      }
    }
  }
  return min(best, R + 1);
}

// One CTA per (x plane, y chunk).  Shared memory: occupancy words of rows [y0-R, y1+R), then their z distances.
__global__ void __launch_bounds__(512) edtZYKernel(GridDev g, const uint32_t* __restrict__ occ, uint16_t* __restrict__ out,
                                                  int chunk, int n_chunks)
{
  extern __shared__ __align__(16) uint8_t smem[];
  const int x = blockIdx.x / n_chunks, ci = blockIdx.x % n_chunks;
  const int y0 = ci * chunk, y1 = min(g.ny, y0 + chunk);
  const int ya = y0 - g.R, rows = (y1 - y0) + 2 * g.R;  // haloed row range [ya, ya + rows)
  uint32_t* bits = reinterpret_cast<uint32_t*>(smem);
  uint8_t* gz = smem + static_cast<size_t>(rows) * g.wz * 4;
  for (int i = threadIdx.x; i < rows * g.wz; i += blockDim.x)
  {
    const int y = ya + i / g.wz;
    bits[i] = (y >= 0 && y < g.ny) ? occ[(static_cast<size_t>(x) * g.ny + y) * g.wz + (i % g.wz)] : 0u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < rows * g.nz; i += blockDim.x)
  {
    const int r = i / g.nz, z = i - r * g.nz;
    gz[i] = static_cast<uint8_t>(zDistance(bits + r * g.wz, g.wz, z, g.R));
  }
  __syncthreads();
  const int R = g.R, max_d2 = g.max_d2;
  for (int i = threadIdx.x; i < (y1 - y0) * g.nz; i += blockDim.x)
  {
    const int yy = i / g.nz, z = i - yy * g.nz;
    const int r = yy + R;  // row index inside the haloed block
    int c = gz[r * g.nz + z];
    int best = c * c;
    for (int k = 1; k <= R && k * k < best; ++k)
    {
      const int a = gz[(r - k) * g.nz + z], b = gz[(r + k) * g.nz + z];
      const int m = min(a, b);
      best = min(best, m * m + k * k);
    }
    out[(static_cast<size_t>(x) * g.ny + (y0 + yy)) * g.nz + z] = best <= max_d2 ? static_cast<uint16_t>(best) : kInf16;
  }
}

template <typename FT>
__global__ void __launch_bounds__(256) edtXKernel(GridDev g, const uint16_t* __restrict__ in, uint16_t* __restrict__ d2,
                                                 FT* __restrict__ compact)
{
  const size_t plane = static_cast<size_t>(g.ny) * g.nz;
  const size_t total = plane * g.nx;
  const int R = g.R, max_d2 = g.max_d2;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
  {
    const int x = static_cast<int>(i / plane);
    int best = in[i];  // kInf16 (65535) acts as +inf: it can only lose against anything within range
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
    // interleaved compact field: element 2 * i + which (the caller passes compact already offset by `which`)
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
    compact[2 * i] = static_cast<FT>(sizeof(FT) == 1 ? min(v, 255) : v);
    if (v == 0)
    {
      const int z = static_cast<int>(i % g.nz);
      const size_t row = i / g.nz;
      atomicOr(occ + row * g.wz + (z >> 5), 1u << (z & 31));
    }
  }
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
  scatterKernel<<<gridFor(n, 256, 148 * 16), 256, 0, s>>>(d_xyz, n, g, occ, set ? 1 : 0);
  return cudaGetLastError();
}

cudaError_t launchEdt(const GridDev& g, const uint32_t* occ, uint16_t* tmp, uint16_t* d2, void* compact,
                      int compact_bytes, int sm_count, cudaStream_t s)
{
  // y chunking so a haloed slab (occupancy words + z distances) fits the 227 KB of shared memory per CTA
  const size_t per_row = static_cast<size_t>(g.wz) * 4 + g.nz;
  const size_t budget = 200 * 1024;
  long max_rows = static_cast<long>(budget / per_row);
  int chunk = g.ny;
  if (max_rows < g.ny + 2L * g.R)
  {
    chunk = static_cast<int>(max_rows - 2L * g.R);
    if (chunk < 1)
      return cudaErrorInvalidConfiguration;  // a single z row does not fit: grid too deep for this kernel
  }
  const int n_chunks = (g.ny + chunk - 1) / chunk;
  const size_t smem = static_cast<size_t>(chunk + 2 * g.R) * per_row;
  cudaError_t e = cudaFuncSetAttribute(edtZYKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess)
    return e;
  edtZYKernel<<<g.nx * n_chunks, 512, smem, s>>>(g, occ, tmp, chunk, n_chunks);
  e = cudaGetLastError();
  if (e != cudaSuccess)
    return e;
  const size_t total = static_cast<size_t>(g.nx) * g.ny * g.nz;
  const int grid = gridFor(total, 256, sm_count * 32);
  if (compact_bytes == 1)
    edtXKernel<uint8_t><<<grid, 256, 0, s>>>(g, tmp, d2, static_cast<uint8_t*>(compact));
  else
    edtXKernel<uint16_t><<<grid, 256, 0, s>>>(g, tmp, d2, static_cast<uint16_t*>(compact));
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
