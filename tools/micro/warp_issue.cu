// What ONE warp alone on an SM can issue: cycles per instruction for independent streams of the instruction
// classes the pose solve is made of (the solve is a single-warp serial program, so this -- not the SM's peak --
// is its speed limit).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o warp_issue warp_issue.cu
#include <cstdio>
#include <cuda_runtime.h>

#define REP 256
template <int OP>
__global__ void k(double* out, long long* cyc, double seed, int iseed) {
  __shared__ double sm[8 * 33];
  double x[8];
  int n[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { x[j] = seed + j + threadIdx.x * 1e-9; n[j] = iseed + j * 7 + threadIdx.x; sm[j * 33 + threadIdx.x] = x[j]; }
  __syncwarp();
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < REP; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (OP == 0) x[j] = fma(x[j], 0.999, 1e-3);                          // 8 independent DFMA chains
      if (OP == 1) n[j] = n[j] * 3 + i;                                    // 8 independent IMAD chains
      if (OP == 2) x[j] += sm[j * 33 + ((threadIdx.x + i) & 31)];          // LDS + DADD
      if (OP == 3) sm[j * 33 + ((threadIdx.x + i) & 31)] = x[j];           // STS
      if (OP == 4) { x[j] = fma(x[j], 0.999, 1e-3); n[j] = n[j] * 3 + i; }  // DFMA + IMAD interleaved
      if (OP == 5) x[j] = x[j] * 0.999;                                    // DMUL
      if (OP == 6) n[j] = (n[j] ^ i) + (n[j] >> 3);                        // LOP3 / SHF / IADD3
    }
  }
  long long t1 = clock64();
  double s = 0; int m = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { s += x[j]; m += n[j]; }
  if (threadIdx.x == 0) cyc[OP] = t1 - t0;
  out[threadIdx.x] = s + m + sm[threadIdx.x];
}

int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 64 * 8); cudaMalloc(&cyc, 32 * 8);
  const char* names[] = {"DFMA x8 independent", "IMAD x8 independent", "LDS+DADD x8", "STS x8", "DFMA+IMAD x8", "DMUL x8",
                         "LOP3/SHF/IADD3 x8 (3 instr each)"};
  const int per[] = {8, 8, 16, 8, 16, 8, 24};
#define RUN(OP) k<OP><<<1, 32>>>(out, cyc, 1.37, 3); k<OP><<<1, 32>>>(out, cyc, 1.37, 3);
  RUN(0) RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6)
  cudaDeviceSynchronize();
  long long h[32];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  for (int i = 0; i < 7; ++i)
    printf("%-36s %6.2f cycles per instruction (one warp, %d instr per loop trip)\n", names[i], (double)h[i] / REP / per[i], per[i]);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
