// Dependent-chain latencies (SM cycles per operation, one warp alone on an SM) of what the pose solve is made of:
// FP64 add/mul/fma, conversions, reciprocal / square-root sequences, shuffles and shared-memory round trips.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

#define N 512
template <int OP>
__global__ void chain(double* out, long long* cyc, double seed, double k) {
  __shared__ double sm[64];
  double x = seed + threadIdx.x * 1e-9;
  float xf = (float)seed;
  sm[threadIdx.x] = x;
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) x = fma(x, k, 1e-3);
      if (OP == 1) x = x * k;
      if (OP == 2) x = x + k;
      if (OP == 3) x = (double)(float)x;                       // F2F down + up
      if (OP == 4) x = 1.0 / x + 0.5;                          // IEEE division
      if (OP == 5) x = sqrt(x) + 1.0;                          // IEEE square root
      if (OP == 6) x = __shfl_xor_sync(0xffffffffu, x, 1);     // double shuffle
      if (OP == 7) { sm[threadIdx.x ^ 1] = x; __syncwarp(); x = sm[threadIdx.x] * k; __syncwarp(); }
      if (OP == 8) x = fmax(x, k) * k;                         // DSETP/select + DMUL
      if (OP == 9) xf = __fmaf_rn(xf, (float)k, 1e-3f);        // FP32 reference
      if (OP == 10) xf = rsqrtf(xf) + 1.0f;                    // MUFU + FADD
      if (OP == 11) x = (double)__frcp_rn((float)x) + 0.5;     // F2F, MUFU.RCP, F2F, DADD
      if (OP == 12) x = rsqrt(x) + 1.0;                        // double rsqrt (library)
      if (OP == 13) x = __hiloint2double(__double2hiint(x) + 1, __double2loint(x));  // integer view
      if (OP == 14) x = __drcp_rn(x) + 0.5;
    }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[OP] = t1 - t0;
  out[threadIdx.x] = x + xf;
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 32 * 8);
  const char* names[] = {"DFMA", "DMUL", "DADD", "F2F f64->f32->f64", "1/x + c (IEEE div)", "sqrt(x) + c (IEEE)",
                         "double __shfl_xor", "STS + syncwarp + LDS + DMUL + syncwarp", "fmax + DMUL", "FFMA",
                         "rsqrtf + FADD", "(double)__frcp_rn((float)x) + c", "rsqrt(double) + c", "hi-word integer edit",
                         "__drcp_rn + c"};
#define RUN(OP) chain<OP><<<1, 32>>>(out, cyc, 1.37, 0.999); chain<OP><<<1, 32>>>(out, cyc, 1.37, 0.999);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) RUN(9) RUN(10) RUN(11) RUN(12) RUN(13) RUN(14)
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 15; ++i) printf("%-45s %7.1f cycles per dependent op\n", names[i], (double)h[i] / N);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
