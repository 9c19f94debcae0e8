// Experiment: occupancy scatter of 5 M random points into a 400 x 400 x 200 grid -- atomicOr on a bitmap (what the
// library does) vs plain byte stores into a byte-per-voxel array.  nvcc -arch=sm_100a -O3 -o scatter_bytes scatter_bytes.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>
__global__ void atomicBits(const int3* __restrict__ p, size_t n, unsigned* occ)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int3 v = p[i];
    atomicOr(occ + ((size_t)v.x * 400 + v.y) * 7 + (v.z >> 5), 1u << (v.z & 31));
  }
}
__global__ void plainBytes(const int3* __restrict__ p, size_t n, unsigned char* occ)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int3 v = p[i];
    occ[((size_t)v.x * 400 + v.y) * 200 + v.z] = 1;
  }
}
__global__ void plainBytesPacked(const int3* __restrict__ p, size_t n, unsigned char* occ)
{  // one byte per 8 voxels cannot be written without an atomic; one byte per voxel but z-major rows of 256 B
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int3 v = p[i];
    occ[((size_t)v.x * 400 + v.y) * 256 + v.z] = 1;
  }
}
int main()
{
  const size_t n = 5000000;
  std::vector<int3> h(n);
  srand(5);
  for (auto& v : h)
    v = make_int3(rand() % 400, rand() % 400, rand() % 200);
  int3* d;
  unsigned* bits;
  unsigned char* bytes;
  cudaMalloc(&d, n * sizeof(int3));
  cudaMalloc(&bits, 400 * 400 * 7 * 4);
  cudaMalloc(&bytes, (size_t)400 * 400 * 256);
  cudaMemcpy(d, h.data(), n * sizeof(int3), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int k = 0; k < 3; ++k)
  {
    float ms[3];
    cudaMemset(bits, 0, 400 * 400 * 7 * 4);
    cudaMemset(bytes, 0, (size_t)400 * 400 * 256);
    cudaEventRecord(e0);
    atomicBits<<<148 * 16, 256>>>(d, n, bits);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[0], e0, e1);
    cudaEventRecord(e0);
    plainBytes<<<148 * 16, 256>>>(d, n, bytes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[1], e0, e1);
    cudaMemset(bytes, 0, (size_t)400 * 400 * 256);
    cudaEventRecord(e0);
    plainBytesPacked<<<148 * 16, 256>>>(d, n, bytes);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms[2], e0, e1);
    printf("atomicOr bits %.1f us   byte stores %.1f us   byte stores (256-byte rows) %.1f us\n", ms[0] * 1e3, ms[1] * 1e3, ms[2] * 1e3);
  }
  return 0;
}
