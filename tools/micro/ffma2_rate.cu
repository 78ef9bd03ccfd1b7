// Issue-rate microbenchmark for the instruction mix of vote_count_kernel on sm_100a:
// how many warp-instructions per clock per SM do FFMA, FFMA2 (fma.rn.f32x2), FADD, FSETP/LEA-class ALU
// ops and their mixes sustain?  (Found: see DESIGN.md section 5.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_rate ffma2_rate.cu && ./ffma2_rate
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
#define FMA1(acc, a, b) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(acc) : "f"(a), "f"(b))
#define FMA2(acc, a, b) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b))
#define ADD1(acc, a) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(acc) : "f"(a))
#define LEAHI(acc, a) asm volatile("{ .reg .u32 t; shr.u32 t, %1, 31; add.u32 %0, %0, t; }" : "+r"(acc) : "r"(a))
#define SETP(p, a, b) asm volatile("{ .reg .pred q; setp.gt.f32 q, %1, %2; @q add.u32 %0, %0, 1; }" : "+r"(p) : "f"(a), "f"(b))

template <int MODE>
__global__ void __launch_bounds__(128, 8) k(float* out, int iters, float a, float b) {
  float s[8]; u64 d[8]; unsigned n[8];
  for (int i = 0; i < 8; ++i) { s[i] = threadIdx.x * 1e-3f + i; d[i] = (u64)threadIdx.x * 0x3f8000013f800000ull + i; n[i] = i; }
  u64 ab; asm("mov.b64 %0, {%1, %2};" : "=l"(ab) : "f"(a), "f"(b));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { FMA1(s[i], a, b); }                                        // 8 FFMA
      if (MODE == 1) { FMA2(d[i], ab, ab); }                                      // 8 FFMA2
      if (MODE == 2) { FMA2(d[i], ab, ab); FMA1(s[i], a, b); }                    // 8 FFMA2 + 8 FFMA
      if (MODE == 3) { FMA2(d[i], ab, ab); ADD1(s[i], a); }                       // 8 FFMA2 + 8 FADD
      if (MODE == 4) { FMA2(d[i], ab, ab); LEAHI(n[i], n[(i + 1) & 7]); }         // 8 FFMA2 + 8 (SHF+IADD)
      if (MODE == 5) { ADD1(s[i], a); }                                           // 8 FADD
      if (MODE == 6) { LEAHI(n[i], n[(i + 1) & 7]); }                             // 8 x (SHF + IADD)
      if (MODE == 7) { FMA1(s[i], a, b); LEAHI(n[i], n[(i + 1) & 7]); }           // 8 FFMA + 8 x (SHF+IADD)
      if (MODE == 8) { FMA1(s[i], a, b); FMA1(s[i], b, a); }                      // 16 FFMA (same work as mode 1)
    }
    if (MODE == 9) {   // the trip of vote_count: 20 FFMA2, 8 FADD, 8 FSETP(+pred add), 8 LEA.HI-like
#pragma unroll
      for (int r = 0; r < 20; ++r) FMA2(d[r & 7], ab, ab);
#pragma unroll
      for (int r = 0; r < 8; ++r) { ADD1(s[r], a); SETP(n[r], s[r], b); LEAHI(n[r], n[(r + 1) & 7]); }
    }
    if (MODE == 10) {  // same arithmetic with scalar FFMA: 40 FFMA, 8 FADD, 8 FSETP, 8 LEA.HI-like
#pragma unroll
      for (int r = 0; r < 40; ++r) FMA1(s[r & 7], a, b);
#pragma unroll
      for (int r = 0; r < 8; ++r) { ADD1(s[r], a); SETP(n[r], s[r], b); LEAHI(n[r], n[(r + 1) & 7]); }
    }
    if (MODE == 11) {  // half of the FMA work packed, half scalar: 10 FFMA2 + 20 FFMA, 8 FADD, 8 FSETP, 8 LEA.HI-like
#pragma unroll
      for (int r = 0; r < 10; ++r) { FMA2(d[r & 7], ab, ab); FMA1(s[r & 7], a, b); FMA1(s[(r + 3) & 7], b, a); }
#pragma unroll
      for (int r = 0; r < 8; ++r) { ADD1(s[r], a); SETP(n[r], s[r], b); LEAHI(n[r], n[(r + 1) & 7]); }
    }
    if (MODE == 12) {  // mode 9's instructions, interleaved so that every few issue slots feed both pipes
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        FMA2(d[r], ab, ab); ADD1(s[r], a); FMA2(d[(r + 4) & 7], ab, ab); SETP(n[r], s[r], b);
        if (r < 4) FMA2(d[(r + 2) & 7], ab, ab);
        LEAHI(n[r], n[(r + 1) & 7]);
      }
    }
    if (MODE == 13) {  // mode 9 with the ALU share of the real kernel: 8 FSETP (no predicated add) + 8 LEA.HI
#pragma unroll
      for (int r = 0; r < 20; ++r) FMA2(d[r & 7], ab, ab);
#pragma unroll
      for (int r = 0; r < 8; ++r) { ADD1(s[r], a); LEAHI(n[r], __float_as_uint(s[r])); LEAHI(n[(r + 3) & 7], n[(r + 1) & 7]); }
    }
    if (MODE == 14) {  // mode 13 interleaved
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        FMA2(d[r], ab, ab); ADD1(s[r], a); FMA2(d[(r + 4) & 7], ab, ab); LEAHI(n[r], __float_as_uint(s[r]));
        if (r < 4) FMA2(d[(r + 2) & 7], ab, ab);
        LEAHI(n[(r + 3) & 7], n[(r + 1) & 7]);
      }
    }
    if (MODE == 15 || MODE == 16 || MODE == 17) {
      // 16: the fast path of vote_count today (20 FFMA2, 8 FADD, 8 FSETP chained, 8 LEA.HI)
      // 15: band from the record (16 FFMA2), undecided test as a 3-input |min| tree (4 FMNMX3 + 1 FSETP)
      // 17: 20 FFMA2 with the |min| tree
      const int nf = MODE == 15 ? 16 : 20;
#pragma unroll
      for (int r = 0; r < nf; ++r) FMA2(d[r & 7], ab, ab);
#pragma unroll
      for (int r = 0; r < 8; ++r) { ADD1(s[r], a); LEAHI(n[r], __float_as_uint(s[r])); }
      if (MODE == 16) {
        unsigned q = 0;
#pragma unroll
        for (int r = 0; r < 8; ++r) q |= (fabsf(s[r]) > b);
        n[0] += q;
      } else {
        float mn = fminf(fminf(fabsf(s[0]), fabsf(s[1])), fabsf(s[2]));
        mn = fminf(fminf(fabsf(s[3]), fabsf(s[4])), mn);
        mn = fminf(fminf(fabsf(s[5]), fabsf(s[6])), mn);
        mn = fminf(fabsf(s[7]), mn);
        n[0] += (mn > b);
      }
    }
    if (MODE >= 18 && MODE <= 21) {
      // 18: 20 FFMA2 + 8 FADD, nothing else      19: + 8 LEA.HI
      // 20: the subtraction packed (4 x add.f32x2 on an |.|-masked operand: 8 LOP3) + 8 LEA.HI + |min| tree
      // 21: 24 FFMA2 only
#pragma unroll
      for (int r = 0; r < 20; ++r) FMA2(d[r & 7], ab, ab);
      if (MODE == 18 || MODE == 19) {
#pragma unroll
        for (int r = 0; r < 8; ++r) { ADD1(s[r], a); if (MODE == 19) LEAHI(n[r], __float_as_uint(s[r])); }
      }
      if (MODE == 20) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          u64 t = d[r + 4] & 0x7fffffff7fffffffull, m;
          asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(m) : "l"(d[r]), "l"(t));
          float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(m));
          LEAHI(n[2 * r], __float_as_uint(lo)); LEAHI(n[2 * r + 1], __float_as_uint(hi));
          s[2 * r] = lo; s[2 * r + 1] = hi;
        }
        float mn = fminf(fminf(fabsf(s[0]), fabsf(s[1])), fabsf(s[2]));
        mn = fminf(fminf(fabsf(s[3]), fabsf(s[4])), mn);
        mn = fminf(fminf(fabsf(s[5]), fabsf(s[6])), mn);
        mn = fminf(fabsf(s[7]), mn);
        n[0] += (mn > b);
      }
      if (MODE == 21) {
#pragma unroll
        for (int r = 0; r < 4; ++r) FMA2(d[r], ab, ab);
      }
    }
  }
  float acc = 0.f;
  for (int i = 0; i < 8; ++i) acc += s[i] + (float)(d[i] & 0xffff) + (float)n[i];
  if (acc == 123.456f) out[0] = acc;
}

template <int MODE>
static void run(const char* name, double warp_instr_per_iter, double lane_fma_per_iter) {
  float* out; cudaMalloc(&out, 4);
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<sms * 8, 128>>>(out, 100, 1.0001f, 0.9999f);
  cudaEventRecord(e0);
  k<MODE><<<sms * 8, 128>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
  const double clocks = ms * 1e-3 * khz * 1e3;                       // at the nominal max clock
  const double warps_per_sm = 8 * 4;
  const double wi = warp_instr_per_iter * iters * warps_per_sm;      // warp-instructions per SM
  printf("%-52s %8.3f ms  %6.2f warp-instr/clk/SM  %7.1f lane-FMA-class ops/clk/SM  (%s)\n", name, ms, wi / clocks,
         lane_fma_per_iter * iters * warps_per_sm * 32 / clocks, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}

int main() {
  run<0>("8 FFMA", 8, 8);
  run<8>("16 FFMA", 16, 16);
  run<1>("8 FFMA2", 8, 16);
  run<2>("8 FFMA2 + 8 FFMA", 16, 24);
  run<3>("8 FFMA2 + 8 FADD", 16, 24);
  run<5>("8 FADD", 8, 8);
  run<6>("8 x (SHF + IADD)", 16, 0);
  run<4>("8 FFMA2 + 8 x (SHF + IADD)", 24, 16);
  run<7>("8 FFMA + 8 x (SHF + IADD)", 24, 8);
  run<9>("trip: 20 FFMA2 + 8 FADD + 8 FSETP/@IADD + 8 SHF/IADD", 20 + 8 + 16 + 16, 48);
  run<10>("trip: 40 FFMA + 8 FADD + 8 FSETP/@IADD + 8 SHF/IADD", 40 + 8 + 16 + 16, 48);
  run<11>("trip: 10 FFMA2 + 20 FFMA + 8 FADD + ... ", 30 + 8 + 16 + 16, 48);
  run<12>("trip (mode 9) interleaved", 20 + 8 + 16 + 16, 48);
  run<13>("trip: 20 FFMA2 + 8 FADD + 16 LEA.HI, clustered", 20 + 8 + 16, 48);
  run<14>("trip: 20 FFMA2 + 8 FADD + 16 LEA.HI, interleaved", 20 + 8 + 16, 48);
  run<16>("trip: 20 FFMA2 + 8 FADD + 8 FSETP + 8 LEA.HI (today)", 20 + 8 + 8 + 8 + 1, 48);
  run<17>("trip: 20 FFMA2 + 8 FADD + 4 FMNMX3 + 8 LEA.HI", 20 + 8 + 4 + 8 + 2, 48);
  run<15>("trip: 16 FFMA2 + 8 FADD + 4 FMNMX3 + 8 LEA.HI", 16 + 8 + 4 + 8 + 2, 40);
  run<18>("20 FFMA2 + 8 FADD", 28, 48);
  run<19>("20 FFMA2 + 8 FADD + 8 LEA.HI", 36, 48);
  run<20>("20 FFMA2 + 4 FADD2 + 8 LOP3 + 8 LEA.HI + 4 FMNMX3", 20 + 4 + 8 + 8 + 6, 48);
  run<21>("24 FFMA2", 24, 48);
  return 0;
}
