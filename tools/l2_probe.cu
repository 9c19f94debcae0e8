// Prints the L2 persistence limits of the device (tools/l2_probe.cu; nvcc -o gpurun_out/l2_probe tools/l2_probe.cu).
#include <cstdio>
#include <cuda_runtime.h>
int main()
{
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  printf("l2CacheSize %d\npersistingL2CacheMaxSize %d\naccessPolicyMaxWindowSize %d\n", p.l2CacheSize, p.persistingL2CacheMaxSize,
         p.accessPolicyMaxWindowSize);
  size_t lim = 0;
  cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
  printf("default cudaLimitPersistingL2CacheSize %zu\n", lim);
  cudaError_t e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)p.persistingL2CacheMaxSize);
  cudaDeviceGetLimit(&lim, cudaLimitPersistingL2CacheSize);
  printf("after set to max: %s, limit %zu\n", cudaGetErrorString(e), lim);
  return 0;
}
