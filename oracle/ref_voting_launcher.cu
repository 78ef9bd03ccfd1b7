// TEST INFRASTRUCTURE ONLY (oracle/): "device oracle" for the voting kernels.
//
// oracle/Makefile extracts the four __global__ kernels of the reference's
// lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu (lines 10-49, 87-126,
// 169-229, 267-310) VERBATIM into oracle/_ref/ref_kernels.cuh (git-ignored build
// output; the ATen launchers around them need THC / old ATen and cannot be built
// against torch 2.11) and compiles them for sm_100a behind the plain C launchers
// below.  The launch geometry follows getGPULayout (cuda_common.h:35-55).
// Only tests/ call this library; it is the reference's arithmetic as nvcc
// contracts it, used to pin our CUDA kernels and the C restatement bit-exactly.
#include <cuda_runtime.h>
#include <stdint.h>

#include "_ref/ref_kernels.cuh"

static int inf_two_exp(int v) { int i = 1; while (v > i) i <<= 1; return i; }

// cuda_common.h:35-55 restated (ours; the original is host code inside a header
// that also defines non-inline functions).
static void gpu_layout(int d0, int d1, int d2, dim3* b, dim3* t) {
  int t2 = 64; if (d2 < t2) t2 = inf_two_exp(d2);
  int b2 = d2 / t2 + (d2 % t2 > 0);
  int t1 = 1024 / t2; if (d1 < t1) t1 = inf_two_exp(d1);
  int b1 = d1 / t1 + (d1 % t1 > 0);
  int t0 = 1024 / (t1 * t2); if (d0 < t0) t0 = inf_two_exp(d0);
  int b0 = d0 / t0 + (d0 % t0 > 0);
  *b = dim3(b0, b1, b2); *t = dim3(t0, t1, t2);
}

extern "C" {

// hypo_pts must be zeroed by the caller (the reference allocates it with at::zeros).
int ref_generate_hypothesis(float* direct, float* coords, int* idxs, float* hypo_pts,
                            int tn, int vn, int hn) {
  dim3 b, t; gpu_layout(hn * vn, 1, 1, &b, &t);
  generate_hypothesis_kernel<<<b, t>>>(direct, coords, idxs, hypo_pts, tn, vn, hn);
  return (int)cudaGetLastError();
}

int ref_voting_for_hypothesis(float* direct, float* coords, float* hypo_pts,
                              unsigned char* inliers, int tn, int vn, int hn, float thresh) {
  dim3 b, t; gpu_layout(hn, vn * tn, 1, &b, &t);
  voting_for_hypothesis_kernel<<<b, t>>>(direct, coords, hypo_pts, inliers, tn, vn, hn, thresh);
  return (int)cudaGetLastError();
}

int ref_generate_hypothesis_vanishing_point(float* direct, float* coords, int* idxs,
                                            float* hypo_pts, int tn, int vn, int hn) {
  dim3 b, t; gpu_layout(hn * vn, 1, 1, &b, &t);
  generate_hypothesis_vanishing_point_kernel<<<b, t>>>(direct, coords, idxs, hypo_pts, tn, vn, hn);
  return (int)cudaGetLastError();
}

int ref_voting_for_hypothesis_vanishing_point(float* direct, float* coords, float* hypo_pts,
                                              unsigned char* inliers, int tn, int vn, int hn,
                                              float thresh) {
  dim3 b, t; gpu_layout(hn, vn * tn, 1, &b, &t);
  voting_for_hypothesis_vanishing_point_kernel<<<b, t>>>(direct, coords, hypo_pts, inliers, tn,
                                                         vn, hn, thresh);
  return (int)cudaGetLastError();
}

}  // extern "C"
