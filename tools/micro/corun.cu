// Microbenchmark helper (not product code): co-running load generators for tools/corun.py
#include <cuda_runtime.h>
extern "C" {
__global__ void spin_kernel(long long cycles, int* sink) {
  const long long t0 = clock64();
  int x = 0;
  while (clock64() - t0 < cycles) x += 1;
  if (x == -1) *sink = x;
}
__global__ void zc_read_kernel(const float4* __restrict__ src, size_t n, float* sink, int reps) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
      float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 12345.678f) *sink = acc;
}
int set_carveout(int pct) {
  cudaFuncSetAttribute(spin_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  cudaFuncSetAttribute(zc_read_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  return (int)cudaGetLastError();
}
int launch_spin(int grid, int block, long long cycles, void* stream) {
  spin_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(cycles, nullptr);
  return (int)cudaGetLastError();
}
int launch_zc(const void* src, size_t nfloat4, int grid, int block, int reps, void* stream) {
  zc_read_kernel<<<grid, block, 0, (cudaStream_t)stream>>>((const float4*)src, nfloat4, nullptr, reps);
  return (int)cudaGetLastError();
}
}
