/* TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the reference's nearest-point search
 * /root/reference/lib/utils/extend_utils/src/nearest_neighborhood.cu:48-121
 * (findNearestPoint3DIdxKernel / findNearestPoint2DIdxKernel; Python entry
 * lib/utils/extend_utils/extend_utils.py:40-61 find_nearest_point_idx), which the LINEMOD-style
 * symmetric metrics use (evaluation.py:162-170, :348-351, :385-397).
 *
 * float32 throughout; squared distance in the order nvcc contracts the source expression
 *   (x1-x2)*(x1-x2) + (y1-y2)*(y1-y2) + (z1-z2)*(z1-z2)   ->   fma(dz,dz, fma(dx,dx, rn(dy*dy)))
 * (a*b + c*d becomes fma(a,b, rn(c*d)): SASS of oracle/_ref/libref_nearest.so, same rule as the voting kernels)
 * strict '<' update from FLT_MAX (first minimum wins, index 0 when nothing is closer than FLT_MAX).
 * Pinned on the GPU box against the reference file itself compiled unmodified
 * (oracle/_ref/libref_nearest.so; tests/test_metrics_gpu.py). */
#include <float.h>
#include <math.h>
#include <stdint.h>

void orc_nearest_idx(const float* ref_pts, const float* que_pts, int32_t* idxs, int pn1, int pn2, int dim,
                     int exclude_self) {
  for (int p2 = 0; p2 < pn2; ++p2) {
    float best = FLT_MAX;
    int bi = 0;
    for (int p1 = 0; p1 < pn1; ++p1) {
      if (exclude_self && p1 == p2) continue;
      const float dx = ref_pts[p1 * dim] - que_pts[p2 * dim];
      const float dy = ref_pts[p1 * dim + 1] - que_pts[p2 * dim + 1];
      float d = fmaf(dx, dx, dy * dy);
      if (dim == 3) {
        const float dz = ref_pts[p1 * dim + 2] - que_pts[p2 * dim + 2];
        d = fmaf(dz, dz, d);
      }
      if (d < best) { best = d; bi = p1; }
    }
    idxs[p2] = bi;
  }
}
