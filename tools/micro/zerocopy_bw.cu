// Microbenchmark (not product code): how fast can SMs read page-locked host memory in place over
// PCIe, by access width, against cudaMemcpyAsync of the same bytes.  Drives the design of
// field_gather_kernel (csrc/voting.cu).   nvcc -arch=sm_100a -O3 -o zerocopy_bw zerocopy_bw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <typename T>
__global__ void read_kernel(const T* __restrict__ src, T* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) dst[i] = src[i];
}
// 4 independent loads in flight per thread
template <typename T>
__global__ void read_kernel_u4(const T* __restrict__ src, T* __restrict__ dst, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n; i += 4 * stride) {
    T a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    dst[i] = a; dst[i + stride] = b; dst[i + 2 * stride] = c; dst[i + 3 * stride] = d;
  }
  for (; i < n; i += stride) dst[i] = src[i];
}
// bulk async copy global(host) -> shared -> global(device), 16 KB per CTA stage
__global__ void bulk_kernel(const char* __restrict__ src, char* __restrict__ dst, size_t bytes, int chunk) {
  extern __shared__ __align__(128) char smem[];
  __shared__ uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  const uint32_t sm_a = (uint32_t)__cvta_generic_to_shared(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint32_t phase = 0;
  for (size_t off = (size_t)blockIdx.x * chunk; off + chunk <= bytes; off += (size_t)gridDim.x * chunk) {
    if (threadIdx.x == 0) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(chunk));
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(sm_a), "l"(src + off), "r"(chunk), "r"(bar_a) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                   : "=r"(ok) : "r"(bar_a), "r"(phase) : "memory");
    }
    phase ^= 1;
    for (int i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16)
      *reinterpret_cast<uint4*>(dst + off + i) = *reinterpret_cast<const uint4*>(smem + i);
    __syncthreads();
  }
}

int main() {
  const size_t bytes = 256ull << 20;
  char *h, *d;
  CK(cudaHostAlloc(&h, bytes, cudaHostAllocDefault));
  CK(cudaMalloc(&d, bytes));
  for (size_t i = 0; i < bytes; i += 4096) h[i] = (char)i;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  float ms;
  auto report = [&](const char* name) { printf("%-34s %7.2f GB/s  (%.3f ms)\n", name, bytes / ms / 1e6, ms); };
  for (int rep = 0; rep < 2; ++rep) {
    CK(cudaEventRecord(a)); CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice)); CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
  }
  report("cudaMemcpyAsync H2D");
  const int grids[] = {148, 148 * 4, 148 * 16};
  for (int g : grids) {
    char name[96];
#define RUN(KERN, T, label)                                                                      \
    for (int rep = 0; rep < 2; ++rep) {                                                          \
      CK(cudaEventRecord(a));                                                                    \
      KERN<T><<<g, 256>>>((const T*)h, (T*)d, bytes / sizeof(T));                                \
      CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b)); \
    }                                                                                            \
    snprintf(name, sizeof name, "%s grid %d", label, g); report(name);
    RUN(read_kernel, float, "zero-copy ld.32")
    RUN(read_kernel, float4, "zero-copy ld.128")
    RUN(read_kernel_u4, float, "zero-copy ld.32 x4 in flight")
    RUN(read_kernel_u4, float4, "zero-copy ld.128 x4 in flight")
  }
  for (int chunk : {4096, 16384, 32768}) {
    for (int g : {148, 148 * 4}) {
      for (int rep = 0; rep < 2; ++rep) {
        CK(cudaEventRecord(a));
        bulk_kernel<<<g, 256, chunk>>>(h, d, bytes, chunk);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b)); CK(cudaEventElapsedTime(&ms, a, b));
      }
      char name[96]; snprintf(name, sizeof name, "cp.async.bulk %d B grid %d", chunk, g); report(name);
    }
  }
  CK(cudaGetLastError());
  return 0;
}
