// LINEMOD-style pose metrics, batched over poses on the device.
//
// Replaces (paths under /root/reference):
//   evaluation.py:340-346 projection_2d        mean 2-D distance of the projected model points
//   evaluation.py:348-354 projection_2d_sym    ... to the NEAREST projected point of the other pose
//   evaluation.py:356-383 add_metric           ADD: mean 3-D distance of the transformed model points
//   evaluation.py:385-397 add_metric_sym       ADD-S: mean distance to the nearest transformed point
//   evaluation.py:399-411 cm_degree_5_metric   translation error (cm) and rotation error (degrees)
//   (twins: lib/utils/evaluation_utils.py:75-141), with
//   lib/utils/base_utils.py:293-297 project_K and the nearest-point search
//   lib/utils/extend_utils/src/nearest_neighborhood.cu:48-121 (float32 brute force, first minimum; the
//   reference copies both point sets to the device per call and back, extend_utils.py:40-61).
//
// One CTA per pose.  Distances and means in FP64 like numpy; the nearest-point SEARCH in float32 with the
// reference kernel's expression as nvcc contracts it (fma(dz,dz, fma(dx,dx, rn(dy*dy))), strict '<' from
// FLT_MAX), so the chosen indices equal the reference's (bitwise-checked on the GPU against the reference file
// compiled unmodified, tests/test_metrics_gpu.py).  The ADD-S search is a brute-force n_model^2 minimum per
// pose: reference points are transformed on the fly (FP64 -> float32) into 1024-point shared-memory tiles and
// every thread scans them for its own query point -- nothing but the model is read from HBM.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace epb {

constexpr int MET_THREADS = 256;
constexpr int MET_TILE = 1024;

struct Rt {
  double r[9], t[3];
};
__device__ __forceinline__ Rt load_rt(const double* p) {   // [3,4] row-major [R|t]
  Rt o;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    o.r[i * 3] = p[i * 4]; o.r[i * 3 + 1] = p[i * 4 + 1]; o.r[i * 3 + 2] = p[i * 4 + 2]; o.t[i] = p[i * 4 + 3];
  }
  return o;
}
// np.dot(model, R.T) + t  (evaluation.py:364)
__device__ __forceinline__ void xform(const Rt& a, const double* __restrict__ m, double (&o)[3]) {
  const double x = m[0], y = m[1], z = m[2];
#pragma unroll
  for (int i = 0; i < 3; ++i) o[i] = (x * a.r[i * 3] + y * a.r[i * 3 + 1] + z * a.r[i * 3 + 2]) + a.t[i];
}
// (pts @ K.T)[:, :2] / (pts @ K.T)[:, 2:]  (base_utils.py:295-296)
__device__ __forceinline__ void project(const double* __restrict__ K, const double (&p)[3], double (&uv)[2]) {
  const double a = p[0] * K[0] + p[1] * K[1] + p[2] * K[2];
  const double b = p[0] * K[3] + p[1] * K[4] + p[2] * K[5];
  const double c = p[0] * K[6] + p[1] * K[7] + p[2] * K[8];
  uv[0] = a / c; uv[1] = b / c;
}

__device__ __forceinline__ double block_sum(double v, double* s_red /* [8] */) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  double s = 0;
#pragma unroll
  for (int w = 0; w < MET_THREADS / 32; ++w) s += s_red[w];
  return s;
}

// sum over the query points (model under `que`) of the FP64 distance to the nearest reference point (model
// under `ref`), nearest in the reference kernel's float32 arithmetic.  DIM 3: transformed points; 2: projections.
template <int DIM>
__device__ double nearest_sum(const Rt& ref, const Rt& que, const double* __restrict__ model, int n,
                              const double* __restrict__ K, float* s_pts /* [MET_TILE * DIM] */, double* s_red) {
  double acc = 0;
  for (int q0 = 0; q0 < n; q0 += MET_THREADS) {
    const int qi = q0 + threadIdx.x;
    double qd[3] = {0, 0, 0};
    float qf[3] = {0.f, 0.f, 0.f};
    if (qi < n) {
      double p3[3];
      xform(que, model + (size_t)qi * 3, p3);
      if (DIM == 2) { double uv[2]; project(K, p3, uv); qd[0] = uv[0]; qd[1] = uv[1]; }
      else { qd[0] = p3[0]; qd[1] = p3[1]; qd[2] = p3[2]; }
#pragma unroll
      for (int d = 0; d < DIM; ++d) qf[d] = (float)qd[d];      // np.ascontiguousarray(..., np.float32)
    }
    float best = FLT_MAX;
    int bi = 0;
    for (int r0 = 0; r0 < n; r0 += MET_TILE) {
      __syncthreads();
      for (int k = threadIdx.x; k < MET_TILE && r0 + k < n; k += MET_THREADS) {
        double p3[3];
        xform(ref, model + (size_t)(r0 + k) * 3, p3);
        if (DIM == 2) { double uv[2]; project(K, p3, uv); s_pts[k * 2] = (float)uv[0]; s_pts[k * 2 + 1] = (float)uv[1]; }
        else { s_pts[k * 3] = (float)p3[0]; s_pts[k * 3 + 1] = (float)p3[1]; s_pts[k * 3 + 2] = (float)p3[2]; }
      }
      __syncthreads();
      const int cnt = min(MET_TILE, n - r0);
      for (int k = 0; k < cnt; ++k) {
        const float dx = __fsub_rn(s_pts[k * DIM], qf[0]);
        const float dy = __fsub_rn(s_pts[k * DIM + 1], qf[1]);
        float d = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
        if (DIM == 3) { const float dz = __fsub_rn(s_pts[k * DIM + 2], qf[2]); d = __fmaf_rn(dz, dz, d); }
        if (d < best) { best = d; bi = r0 + k; }
      }
    }
    if (qi < n) {
      // np.linalg.norm(pts1[idxs] - pts2, 2, 1) on the float64 points (evaluation.py:170)
      double p3[3], rd[3] = {0, 0, 0};
      xform(ref, model + (size_t)bi * 3, p3);
      if (DIM == 2) { double uv[2]; project(K, p3, uv); rd[0] = uv[0]; rd[1] = uv[1]; }
      else { rd[0] = p3[0]; rd[1] = p3[1]; rd[2] = p3[2]; }
      double ss = 0;
#pragma unroll
      for (int d = 0; d < DIM; ++d) ss += (rd[d] - qd[d]) * (rd[d] - qd[d]);
      acc += sqrt(ss);
    }
  }
  return block_sum(acc, s_red);
}

__global__ void __launch_bounds__(MET_THREADS)
pose_metrics_kernel(const double* __restrict__ pred, const double* __restrict__ gt, const double* __restrict__ model, int n,
                    const double* __restrict__ K, int sym, double* __restrict__ proj, double* __restrict__ add,
                    double* __restrict__ cm, double* __restrict__ deg) {
  __shared__ float s_pts[MET_TILE * 3];
  __shared__ double s_red[MET_THREADS / 32];
  const int b = blockIdx.x;
  const Rt P = load_rt(pred + (size_t)b * 12), G = load_rt(gt + (size_t)b * 12);
  if (threadIdx.x == 0 && (cm || deg)) {
    // evaluation.py:403-407
    const double dx = P.t[0] - G.t[0], dy = P.t[1] - G.t[1], dz = P.t[2] - G.t[2];
    if (cm) cm[b] = sqrt(dx * dx + dy * dy + dz * dz) * 100.0;
    double trace = 0;   // trace(R_pred R_gt^T) = sum of the element-wise products
#pragma unroll
    for (int i = 0; i < 9; ++i) trace += P.r[i] * G.r[i];
    trace = trace <= 3.0 ? trace : 3.0;
    if (deg) deg[b] = acos((trace - 1.0) / 2.0) * (180.0 / 3.14159265358979323846);   // NaN below -1, like np.arccos
  }
  if (n <= 0) return;
  if (!sym) {
    double s_add = 0, s_proj = 0;
    for (int i = threadIdx.x; i < n; i += MET_THREADS) {
      double a[3], g3[3];
      xform(P, model + (size_t)i * 3, a);
      xform(G, model + (size_t)i * 3, g3);
      if (add) s_add += sqrt((a[0] - g3[0]) * (a[0] - g3[0]) + (a[1] - g3[1]) * (a[1] - g3[1]) + (a[2] - g3[2]) * (a[2] - g3[2]));
      if (proj) {
        double ua[2], ug[2];
        project(K, a, ua); project(K, g3, ug);
        s_proj += sqrt((ua[0] - ug[0]) * (ua[0] - ug[0]) + (ua[1] - ug[1]) * (ua[1] - ug[1]));
      }
    }
    if (add) { const double t = block_sum(s_add, s_red); if (threadIdx.x == 0) add[b] = t / n; }
    if (proj) { const double t = block_sum(s_proj, s_red); if (threadIdx.x == 0) proj[b] = t / n; }
  } else {
    // find_nearest_point_distance(pts_pred, pts_targets): for every TARGET point the nearest PREDICTED point
    if (add) { const double t = nearest_sum<3>(P, G, model, n, K, s_pts, s_red); if (threadIdx.x == 0) add[b] = t / n; }
    if (proj) { const double t = nearest_sum<2>(P, G, model, n, K, s_pts, s_red); if (threadIdx.x == 0) proj[b] = t / n; }
  }
}

// find_nearest_point_idx itself (extend_utils.py:40-61): ref [pn1,dim], que [pn2,dim] float32 -> idxs [pn2]
template <int DIM>
__global__ void __launch_bounds__(MET_THREADS)
nearest_idx_kernel(const float* __restrict__ ref, const float* __restrict__ que, int32_t* __restrict__ idxs, int pn1, int pn2) {
  __shared__ float s_pts[MET_TILE * 3];
  const int qi = blockIdx.x * MET_THREADS + threadIdx.x;
  float qf[3] = {0.f, 0.f, 0.f};
  if (qi < pn2)
#pragma unroll
    for (int d = 0; d < DIM; ++d) qf[d] = que[(size_t)qi * DIM + d];
  float best = FLT_MAX;
  int bi = 0;
  for (int r0 = 0; r0 < pn1; r0 += MET_TILE) {
    __syncthreads();
    const int cnt = min(MET_TILE, pn1 - r0);
    for (int k = threadIdx.x; k < cnt * DIM; k += MET_THREADS) s_pts[k] = ref[(size_t)r0 * DIM + k];
    __syncthreads();
    for (int k = 0; k < cnt; ++k) {
      const float dx = __fsub_rn(s_pts[k * DIM], qf[0]);
      const float dy = __fsub_rn(s_pts[k * DIM + 1], qf[1]);
      float d = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
      if (DIM == 3) { const float dz = __fsub_rn(s_pts[k * DIM + 2], qf[2]); d = __fmaf_rn(dz, dz, d); }
      if (d < best) { best = d; bi = r0 + k; }
    }
  }
  if (qi < pn2) idxs[qi] = bi;
}

}  // namespace epb

using namespace epb;

extern "C" int epb_pose_metrics(const double* pred_rt34, const double* gt_rt34, int N, const double* model, int n_model,
                                const double* K, int symmetric, double* proj2d, double* add, double* cm, double* deg,
                                void* stream) {
  if (!pred_rt34 || !gt_rt34 || N <= 0 || n_model < 0) return EPB_ERR_INVALID;
  if ((proj2d || add) && n_model > 0 && !model) return EPB_ERR_INVALID;
  if (proj2d && !K) return EPB_ERR_INVALID;
  pose_metrics_kernel<<<N, MET_THREADS, 0, (cudaStream_t)stream>>>(pred_rt34, gt_rt34, model, n_model, K, symmetric, proj2d,
                                                                    add, cm, deg);
  return check_launch();
}

extern "C" int epb_nearest_point_idx(const float* ref_pts, const float* que_pts, int32_t* idxs, int pn1, int pn2, int dim,
                                     void* stream) {
  if (!ref_pts || !que_pts || !idxs || pn1 <= 0 || pn2 <= 0 || (dim != 2 && dim != 3)) return EPB_ERR_INVALID;
  const unsigned grid = (unsigned)((pn2 + MET_THREADS - 1) / MET_THREADS);
  if (dim == 3) nearest_idx_kernel<3><<<grid, MET_THREADS, 0, (cudaStream_t)stream>>>(ref_pts, que_pts, idxs, pn1, pn2);
  else nearest_idx_kernel<2><<<grid, MET_THREADS, 0, (cudaStream_t)stream>>>(ref_pts, que_pts, idxs, pn1, pn2);
  return check_launch();
}
