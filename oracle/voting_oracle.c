/* TEST INFRASTRUCTURE ONLY (oracle/).  Never imported by the product path: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.
 *
 * Plain-C restatement of the reference's RANSAC-voting CUDA kernels
 * (lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu).  The arithmetic
 * reproduces, with explicit fmaf(), the FMA contraction nvcc/ptxas applied to the
 * reference source (read from the shipped sm_86 cubin with `cuobjdump -sass`,
 * SURVEY.md 8a/8c); build with -ffp-contract=off so gcc adds none of its own.
 * Pinned on the GPU box against oracle/_ref/libref_voting.so (the reference
 * kernels compiled verbatim for sm_100a): tests/test_voting_gpu.py.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ransac_voting_kernel.cu:11-49.  hypo_pts [hn,vn,2] must be pre-zeroed
 * (the reference launcher allocates it with at::zeros, :75); degenerate pairs
 * are left untouched.
 *   reference source:  y=(nx1*(nx0*cx0+ny0*cy0)-nx0*(nx1*cx1+ny1*cy1))/(nx1*ny0-nx0*ny1)
 *   as compiled (nvcc 12.9, sm_100a = oracle/_ref):
 *                      s = fma(nx,cx,rn(ny*cy)); y: fma(nx1,s0,-rn(nx0*s1)); x: fma(-ny0,s1,rn(ny1*s0));
 *                      det = rn(nx1*ny0) - rn(nx0*ny1)   (un-fused), div.rn */
void orc_generate_hypothesis(const float* direct, const float* coords, const int* idxs,
                             float* hypo_pts, int tn, int vn, int hn) {
  (void)tn;
  for (int hi = 0; hi < hn; ++hi) {
    for (int vi = 0; vi < vn; ++vi) {
      const int t0 = idxs[hi * vn * 2 + vi * 2];
      const int t1 = idxs[hi * vn * 2 + vi * 2 + 1];
      const float nx0 = direct[t0 * vn * 2 + vi * 2 + 1];
      const float ny0 = -direct[t0 * vn * 2 + vi * 2];
      const float cx0 = coords[t0 * 2], cy0 = coords[t0 * 2 + 1];
      const float nx1 = direct[t1 * vn * 2 + vi * 2 + 1];
      const float ny1 = -direct[t1 * vn * 2 + vi * 2];
      const float cx1 = coords[t1 * 2], cy1 = coords[t1 * 2 + 1];
      const float a = nx1 * ny0;
      const float b = nx0 * ny1;
      const float det1 = a - b; /* nx1*ny0-nx0*ny1 */
      const float det2 = b - a; /* ny1*nx0-ny0*nx1 */
      if (fabs((double)det1) < 1e-6) continue; /* :42, float promoted to double */
      if (fabs((double)det2) < 1e-6) continue; /* :43 */
      const float s0 = fmaf(nx0, cx0, ny0 * cy0);
      const float s1 = fmaf(nx1, cx1, ny1 * cy1);
      const float y = fmaf(nx1, s0, -(nx0 * s1)) / det1;
      /* sm_100a / CUDA 12.9 contraction (oracle/_ref); the shipped sm_86 cubin has
       * fmaf(ny1, s0, -(ny0 * s1)) here -- same source, 1 ulp apart on some pairs. */
      const float x = fmaf(-ny0, s1, ny1 * s0) / det2;
      hypo_pts[hi * vn * 2 + vi * 2] = x;
      hypo_pts[hi * vn * 2 + vi * 2 + 1] = y;
    }
  }
}

/* The inlier predicate of ransac_voting_kernel.cu:100-125, as compiled (PTX pins
 * the fma placement: norm^2 = fma(x,x,rn(y*y)), dot = fma(dx,nx,rn(dy*ny))). */
static inline int orc_is_inlier(float cx, float cy, float hx, float hy, float nx, float ny,
                                float thresh) {
  const float dx = hx - cx;
  const float dy = hy - cy;
  const float norm1 = sqrtf(fmaf(nx, nx, ny * ny));
  const float norm2 = sqrtf(fmaf(dx, dx, dy * dy));
  if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) return 0;
  const float angle_dist = fmaf(dx, nx, dy * ny) / (norm1 * norm2);
  return angle_dist > thresh; /* NaN -> not an inlier */
}

/* ransac_voting_kernel.cu:88-126: writes 1 into a caller-zeroed [hn,vn,tn] buffer. */
void orc_voting_for_hypothesis(const float* direct, const float* coords, const float* hypo_pts,
                               uint8_t* inliers, int tn, int vn, int hn, float thresh) {
  for (int hi = 0; hi < hn; ++hi)
    for (int vi = 0; vi < vn; ++vi) {
      const float hx = hypo_pts[hi * vn * 2 + vi * 2];
      const float hy = hypo_pts[hi * vn * 2 + vi * 2 + 1];
      uint8_t* out = inliers + ((size_t)hi * vn + vi) * tn;
      for (int ti = 0; ti < tn; ++ti) {
        if (orc_is_inlier(coords[ti * 2], coords[ti * 2 + 1], hx, hy,
                          direct[ti * vn * 2 + vi * 2], direct[ti * vn * 2 + vi * 2 + 1], thresh))
          out[ti] = 1;
      }
    }
}

/* voting_for_hypothesis followed by torch.sum(inlier, 2)
 * (ransac_voting_gpu.py:557-561) without materialising the byte tensor. */
void orc_vote_counts(const float* direct, const float* coords, const float* hypo_pts,
                     int32_t* counts /* [hn,vn] */, int tn, int vn, int hn, float thresh) {
  for (int hi = 0; hi < hn; ++hi)
    for (int vi = 0; vi < vn; ++vi) {
      const float hx = hypo_pts[hi * vn * 2 + vi * 2];
      const float hy = hypo_pts[hi * vn * 2 + vi * 2 + 1];
      int32_t c = 0;
      for (int ti = 0; ti < tn; ++ti)
        c += orc_is_inlier(coords[ti * 2], coords[ti * 2 + 1], hx, hy,
                           direct[ti * vn * 2 + vi * 2], direct[ti * vn * 2 + vi * 2 + 1], thresh);
      counts[hi * vn + vi] = c;
    }
}

/* ransac_voting_kernel.cu:170-229 (vanishing-point hypothesis, homogeneous), with the FMA
 * contraction nvcc 12.9 / ptxas apply to the reference source for sm_100a (SASS of
 * oracle/_ref/libref_voting.so): lz = fma(dx, cy, -rn(dy cx)), z = fma(dx0, dy1, -rn(dy0 dx1)),
 * x = fma(dx1, lz0, -rn(dx0 lz1)), y = fma(dy1, lz0, -rn(dy0 lz1)), x - z c = fma(-c, z, x);
 * the two sign products are tested through one min (NaN is ignored by both forms). */
void orc_generate_hypothesis_vanishing_point(const float* direct, const float* coords,
                                             const int* idxs, float* hypo_pts, int tn, int vn,
                                             int hn) {
  (void)tn;
  for (int hi = 0; hi < hn; ++hi)
    for (int vi = 0; vi < vn; ++vi) {
      const int id0 = idxs[hi * vn * 2 + vi * 2], id1 = idxs[hi * vn * 2 + vi * 2 + 1];
      const float dx0 = direct[id0 * vn * 2 + vi * 2], dy0 = direct[id0 * vn * 2 + vi * 2 + 1];
      const float cx0 = coords[id0 * 2], cy0 = coords[id0 * 2 + 1];
      const float dx1 = direct[id1 * vn * 2 + vi * 2], dy1 = direct[id1 * vn * 2 + vi * 2 + 1];
      const float cx1 = coords[id1 * 2], cy1 = coords[id1 * 2 + 1];
      const float lz0 = fmaf(dx0, cy0, -(dy0 * cx0));
      const float lz1 = fmaf(dx1, cy1, -(dy1 * cx1));
      float z = fmaf(dx0, dy1, -(dy0 * dx1));
      float x = fmaf(dx1, lz0, -(dx0 * lz1));
      float y = fmaf(dy1, lz0, -(dy0 * lz1));
      const float val_x0 = dx0 * fmaf(-cx0, z, x), val_x1 = dx1 * fmaf(-cx1, z, x);
      const float val_y0 = dy0 * fmaf(-cy0, z, y), val_y1 = dy1 * fmaf(-cy1, z, y);
      if (val_x0 < 0 && val_x1 < 0 && val_y0 < 0 && val_y1 < 0) { z = -z; x = -x; y = -y; }
      if (val_x0 * val_x1 < 0 || val_y0 * val_y1 < 0) { x = 0.f; y = 0.f; z = 0.f; }
      hypo_pts[hi * vn * 3 + vi * 3] = x;
      hypo_pts[hi * vn * 3 + vi * 3 + 1] = y;
      hypo_pts[hi * vn * 3 + vi * 3 + 2] = z;
    }
}

/* ransac_voting_kernel.cu:268-310 as compiled: diff = fma(-c, hz, h); the numerator of the cosine is the
 * un-fused sum of the two products that the sign test reuses. */
void orc_voting_for_hypothesis_vanishing_point(const float* direct, const float* coords,
                                               const float* hypo_pts, uint8_t* inliers, int tn,
                                               int vn, int hn, float thresh) {
  for (int hi = 0; hi < hn; ++hi)
    for (int vi = 0; vi < vn; ++vi) {
      const float hx = hypo_pts[hi * vn * 3 + vi * 3], hy = hypo_pts[hi * vn * 3 + vi * 3 + 1];
      const float hz = hypo_pts[hi * vn * 3 + vi * 3 + 2];
      uint8_t* out = inliers + ((size_t)hi * vn + vi) * tn;
      for (int ti = 0; ti < tn; ++ti) {
        const float cx = coords[ti * 2], cy = coords[ti * 2 + 1];
        const float ddx = direct[ti * vn * 2 + vi * 2], ddy = direct[ti * vn * 2 + vi * 2 + 1];
        const float fx = fmaf(-cx, hz, hx), fy = fmaf(-cy, hz, hy);
        const float norm1 = sqrtf(fmaf(ddx, ddx, ddy * ddy));
        const float norm2 = sqrtf(fmaf(fx, fx, fy * fy));
        if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) continue;
        const float vx = fx * ddx, vy = fy * ddy;
        const float ad = (vx + vy) / (norm1 * norm2);
        if (vx < 0 || vy < 0) continue;
        if (fabsf(ad) > thresh) out[ti] = 1;
      }
    }
}
