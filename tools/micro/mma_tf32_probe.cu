// Probe for the split-TF32 tensor-core form of the vote test on sm_100a:
//  (1) issue rate of mma.sync.m16n8k8 / m16n8k4 (tf32, f32 accumulate) alone and inside the
//      instruction mix of the planned inner loop (cycles per trip of 8 pair tests per lane and scheduler,
//      the unit of profiles/r1_ffma2_issue_rate_microbench.txt);
//  (2) numerics: a' = A1 hx + A2 hy + A3 and p = B1 hx + B2 hy + B3 evaluated as ONE K = 8 MMA each with a
//      hi/lo split of both operands (hi*hi + lo*hi + hi*lo), against the exact value in double: the
//      worst error in units of u * sum|terms| (u = 2^-24) is what the undecided band has to cover.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_tf32_probe mma_tf32_probe.cu && ./mma_tf32_probe
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ float tf32_hi(float x) {
  unsigned r; asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x)); return __uint_as_float(r);
}
__device__ __forceinline__ void mma_k8(float (&d)[4], const float (&a)[4], const float (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%11,%12,%13};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])), "r"(__float_as_uint(a[3])),
                 "r"(__float_as_uint(b[0])), "r"(__float_as_uint(b[1])), "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f));
}
__device__ __forceinline__ void mma_k4(float (&d)[4], const float (&a)[2], float b) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%8,%9,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(b)),
                 "f"(0.f), "f"(0.f), "f"(0.f), "f"(0.f));
}

// ---------------------------------------------------------------------------------------------- rates
// MODE 0: 6 x k8            1: 6 x k4           2: 4 x k8 + 2 x k4 (the MMAs of one trip)
//      3: mode 2 + epilogue with the band from the third MMA   (8 x: FADD, FSETP |m| > w, LEA.HI)
//      4: 4 x k8 + epilogue with a per-pair band FMA           (8 x: FADD, FFMA, FSETP, LEA.HI)
//      5: epilogue of mode 3 alone (no MMA)
template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(float* out, int iters, float seed) {
  // distinct operands per MMA (ptxas merges identical HMMAs): two A tiles, six B fragments
  float a[4], aa[4], a2[2], bb[6][2];
  for (int i = 0; i < 4; ++i) { a[i] = tf32_hi(seed + threadIdx.x * 1e-3f + i); aa[i] = tf32_hi(seed * 3.f + threadIdx.x * 2e-3f + i); }
  for (int i = 0; i < 6; ++i) { bb[i][0] = tf32_hi(seed * 0.5f + i); bb[i][1] = tf32_hi(seed * 0.25f + 2 * i); }
  a2[0] = a[0]; a2[1] = aa[1];
  float (&b)[2] = bb[0];
  unsigned neg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned amb = 0;
  float keep = 0.f;
  for (int it = 0; it < iters; ++it) {
    float d[6][4];
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) mma_k8(d[i], (i & 1) ? aa : a, bb[i]);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 6; ++i) mma_k4(d[i], a2, bb[i][i & 1]);
    } else if (MODE == 2 || MODE == 3) {
      mma_k8(d[0], a, bb[0]); mma_k8(d[1], aa, bb[0]); mma_k4(d[2], a2, bb[2][0]);
      mma_k8(d[3], a, bb[1]); mma_k8(d[4], aa, bb[1]); mma_k4(d[5], a2, bb[3][1]);
    } else if (MODE == 4) {
      mma_k8(d[0], a, bb[0]); mma_k8(d[1], aa, bb[0]); mma_k8(d[3], a, bb[1]); mma_k8(d[4], aa, bb[1]);
#pragma unroll
      for (int j = 0; j < 4; ++j) { d[2][j] = 0.f; d[5][j] = 0.f; }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) d[i][j] = a[j] + (float)it;
    }
    if (MODE >= 3) {
      // tile (d[0], d[1]) with band d[2]; tile (d[3], d[4]) with band d[5]
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const float (&x)[4] = d[3 * t], (&y)[4] = d[3 * t + 1], (&w)[4] = d[3 * t + 2];
        float m0 = x[0] - fabsf(x[2]), m1 = x[1] - fabsf(x[3]), m2 = y[0] - fabsf(y[2]), m3 = y[1] - fabsf(y[3]);
        float w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        if (MODE == 4) {
          w0 = fmaf(seed, x[0], b[0]); w1 = fmaf(seed, x[1], b[0]); w2 = fmaf(seed, y[0], b[1]); w3 = fmaf(seed, y[1], b[1]);
        }
        amb |= !(fabsf(m0) > w0) | !(fabsf(m1) > w1) | !(fabsf(m2) > w2) | !(fabsf(m3) > w3);
        neg[4 * t + 0] += __float_as_uint(m0) >> 31; neg[4 * t + 1] += __float_as_uint(m1) >> 31;
        neg[4 * t + 2] += __float_as_uint(m2) >> 31; neg[4 * t + 3] += __float_as_uint(m3) >> 31;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) keep += d[i][0] + d[i][3];
    }
    a[0] = __uint_as_float(__float_as_uint(a[0]) ^ (it & 1) << 13);   // keep the loop body from being hoisted
    aa[1] = __uint_as_float(__float_as_uint(aa[1]) ^ (it & 1) << 14); a2[0] = a[0];
  }
  unsigned s = amb;
  for (int i = 0; i < 8; ++i) s += neg[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = keep + (float)s;
}

template <int MODE>
static void run_rate(const char* what, int warps_per_sched) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int threads = 256, ctas_per_sm = warps_per_sched * 4 / (threads / 32);
  float* out; cudaMalloc(&out, (size_t)sms * ctas_per_sm * threads * 4);
  const int iters = 20000;
  rate_kernel<MODE><<<sms * ctas_per_sm, threads>>>(out, 100, 1.25f);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    rate_kernel<MODE><<<sms * ctas_per_sm, threads>>>(out, iters, 1.25f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // cycles per loop iteration and scheduler at the nominal clock (the boost clock the run actually held is
  // in the nvidia-smi record next to this file): trips = iters * warps_per_sched per scheduler
  const double cyc = best * 1e-3 * khz * 1e3 / ((double)iters * warps_per_sched);
  printf("%-64s warps/sched %d : %7.2f cycles per trip and scheduler (at %d MHz nominal), %.3f ms\n", what,
         warps_per_sched, cyc, khz / 1000, best);
  cudaFree(out);
}

// ---------------------------------------------------------------------------------------------- numerics
// One warp per 8 records x 8 hypotheses: rows 0-7 = a' of the records, rows 8-15 = p; K = 8 =
// (A1h,A2h,A3h,A1h | A1l,A2l,A3l,A2h) x (hxh,hyh,1,hxl | hxh,hyh,1,hyl).
__global__ void numerics_kernel(const float* rec /* [n][6]: A1 A2 A3 B1 B2 B3 */, const float* hyp /* [n][2] */,
                                float* out /* [n/8 tiles][8 rec][8 hyp][2] */, int ntiles) {
  const int lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
  for (int tile = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); tile < ntiles; tile += gridDim.x * (blockDim.x >> 5)) {
    const float* r = rec + ((size_t)tile * 8 + g) * 6;
    float A[3] = {r[0], r[1], r[2]}, B[3] = {r[3], r[4], r[5]};
    float Ah[3], Al[3], Bh[3], Bl[3];
    for (int i = 0; i < 3; ++i) { Ah[i] = tf32_hi(A[i]); Al[i] = tf32_hi(A[i] - Ah[i]); Bh[i] = tf32_hi(B[i]); Bl[i] = tf32_hi(B[i] - Bh[i]); }
    float a[4];
    a[0] = j < 3 ? Ah[j] : Ah[0]; a[1] = j < 3 ? Bh[j] : Bh[0];
    a[2] = j < 3 ? Al[j] : Ah[1]; a[3] = j < 3 ? Bl[j] : Bh[1];
    const float* h = hyp + ((size_t)tile * 8 + g) * 2;     // B fragment: n = g, k = j / j + 4
    const float hx = h[0], hy = h[1], hxh = tf32_hi(hx), hyh = tf32_hi(hy), hxl = tf32_hi(hx - hxh), hyl = tf32_hi(hy - hyh);
    float b[2];
    b[0] = j == 0 ? hxh : j == 1 ? hyh : j == 2 ? 1.f : hxl;
    b[1] = j == 0 ? hxh : j == 1 ? hyh : j == 2 ? 1.f : hyl;
    float d[4];
    mma_k8(d, a, b);
    float* o = out + (size_t)tile * 128;
    o[(g * 8 + 2 * j) * 2 + 0] = d[0]; o[(g * 8 + 2 * j + 1) * 2 + 0] = d[1];
    o[(g * 8 + 2 * j) * 2 + 1] = d[2]; o[(g * 8 + 2 * j + 1) * 2 + 1] = d[3];
  }
}

static double urand() { return (double)rand() / RAND_MAX; }

static void run_numerics(const char* what, double hrange, double crange, int far_every) {
  const int ntiles = 1 << 16, n = ntiles * 8;
  std::vector<float> rec((size_t)n * 6), hyp((size_t)n * 2);
  const double k = tan(acos(0.999));
  for (int i = 0; i < n; ++i) {
    const double ang = 2 * M_PI * urand(), nx = cos(ang), ny = sin(ang);
    const double cx = floor(crange * urand()), cy = floor(crange * urand());
    const float fnx = (float)nx, fny = (float)ny, kf = (float)k;
    rec[i * 6 + 0] = kf * fnx; rec[i * 6 + 1] = kf * fny;
    rec[i * 6 + 2] = -(kf * fmaf(fnx, (float)cx, fny * (float)cy));
    rec[i * 6 + 3] = -fny; rec[i * 6 + 4] = fnx; rec[i * 6 + 5] = fmaf(fny, (float)cx, -(fnx * (float)cy));
    double hx = (urand() - 0.25) * hrange, hy = (urand() - 0.25) * hrange;
    if (far_every && i % far_every == 0) { hx *= 1e3 * urand(); hy *= 1e3 * urand(); }
    hyp[i * 2] = (float)hx; hyp[i * 2 + 1] = (float)hy;
  }
  float *drec, *dhyp, *dout;
  cudaMalloc(&drec, rec.size() * 4); cudaMalloc(&dhyp, hyp.size() * 4); cudaMalloc(&dout, (size_t)ntiles * 128 * 4);
  cudaMemcpy(drec, rec.data(), rec.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dhyp, hyp.data(), hyp.size() * 4, cudaMemcpyHostToDevice);
  numerics_kernel<<<296, 256>>>(drec, dhyp, dout, ntiles);
  std::vector<float> out((size_t)ntiles * 128);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  const double u = ldexp(1.0, -24);
  double worst_a = 0, worst_p = 0, sum_a = 0, sum_p = 0;
  for (int t = 0; t < ntiles; ++t)
    for (int g = 0; g < 8; ++g)
      for (int hcol = 0; hcol < 8; ++hcol) {
        const float* r = &rec[((size_t)t * 8 + g) * 6];
        const float* h = &hyp[((size_t)t * 8 + hcol) * 2];
        const double ea = (double)r[0] * h[0] + (double)r[1] * h[1] + (double)r[2];
        const double ep = (double)r[3] * h[0] + (double)r[4] * h[1] + (double)r[5];
        const double sa = fabs((double)r[0] * h[0]) + fabs((double)r[1] * h[1]) + fabs((double)r[2]);
        const double sp = fabs((double)r[3] * h[0]) + fabs((double)r[4] * h[1]) + fabs((double)r[5]);
        const double da = fabs(out[(size_t)t * 128 + (g * 8 + hcol) * 2] - ea) / (u * sa);
        const double dp = fabs(out[(size_t)t * 128 + (g * 8 + hcol) * 2 + 1] - ep) / (u * sp);
        if (da > worst_a) worst_a = da;
        if (dp > worst_p) worst_p = dp;
        sum_a += da; sum_p += dp;
      }
  const double cnt = (double)ntiles * 64;
  printf("numerics %-40s: error / (u * sum|terms|)  a': worst %.3f mean %.3f   p: worst %.3f mean %.3f   (%g pairs)\n", what,
         worst_a, sum_a / cnt, worst_p, sum_p / cnt, cnt);
  cudaFree(drec); cudaFree(dhyp); cudaFree(dout);
}

int main() {
  for (int w : {2, 4}) {
    run_rate<0>("6 x mma.m16n8k8.tf32", w);
    run_rate<1>("6 x mma.m16n8k4.tf32", w);
    run_rate<2>("4 x k8 + 2 x k4 (MMAs of one trip of 8 pair tests per lane)", w);
    run_rate<3>("... + 8 x (FADD, FSETP, LEA.HI), band from the k4 MMA", w);
    run_rate<4>("4 x k8 + 8 x (FADD, FFMA band, FSETP, LEA.HI)", w);
    run_rate<5>("8 x (FADD, FSETP, LEA.HI) alone", w);
  }
  srand(12345);
  run_numerics("h in [-64, 192), c in [0, 256)", 256, 256, 0);
  run_numerics("h in [-192, 576), c in [0, 768)", 768, 768, 0);
  run_numerics("c in [0, 256), every 7th h up to 1e3 x farther", 256, 256, 7);
  run_numerics("h in [-1024, 3072), c in [0, 4096)", 4096, 4096, 0);
  return 0;
}
