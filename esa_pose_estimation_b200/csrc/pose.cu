// Per-image pose solve for sm_100a: EPnP-RANSAC initialisation + Levenberg-Marquardt refinement,
// one lane per 2-D/3-D correspondence (n <= 32), one warp per EPnP solve, 2-8 solver warps per image (RANSAC
// candidates side by side; a two-CTA cluster per image while SMs are idle), FP64 throughout.
//
// Replaces (paths under /root/reference):
//   pnp.py:46-90           pnp(): cv2.solvePnPRansac(flags=EPNP, reprojectionError=5) + Rodrigues
//   val.py:197-207         Rodrigues round trips, cpnp.cpnp_m (Ceres LM; binary/source absent)
//   lib/utils/extend_utils/src/uncertainty_pnp.cpp:7-92   the cost + ceres::Solve the LM follows
//   lib/utils/extend_utils/include/ceres/tiny_solver.h:150-293  the LM recipe
//   val.py:172-193,203-228 keypoint selection, un-crop, quaternion packaging
//   demo.py:295-310        ESA score
// The EPnP statement (control points with OpenCV's SVD sign convention, RANSAC driven by
// cv::RNG(-1) samples of 5, float32-rounded inputs) is the one validated on the CPU in
// oracle/epnp_port.py; this file transliterates it.
//
// ~360 B of traffic per pose and ~1e5 dependent FP64 operations: nothing to tile.  A solve is the program of ONE
// warp, and one warp alone issues an instruction every 3-4 cycles and a shared-memory access every ~7
// (tools/micro/warp_issue.cu, fp64_latency.cu): the design counts instructions per solve, keeps MUFU / F2F /
// divisions out of the serial chain and gives every solver warp an SM sub-partition of its own.
#include <float.h>
#include <math.h>

#include "common.cuh"

namespace epb {

constexpr int POSE_WARPS = 4;

struct __align__(16) WarpScratch {
  double A[12][14];  // M^T M, destroyed by the eigen-solver (rows padded: 16-byte aligned pairs, fewer bank conflicts)
  double V[12][14];  // eigenvectors (columns)
  double S[80];      // reduced sums / scratch
  double L[6][10];
  double vs[4][12];  // the four null-space vectors, ascending eigenvalue
  int order[12];
  // the same bytes as one flat array: transposed per-point tables of the sums below (EPnP) and of LM
  __device__ double* flat() { return reinterpret_cast<double*>(this); }
};
// Sums over the points of a frame are formed "transposed": every point lane parks its terms in a row-per-term
// table in shared memory (row stride 33 doubles: conflict-free both ways) and lane e then adds up term e over the
// points in index order, a short loop.  Against a butterfly per sum this is ~10x fewer instructions (a double
// shuffle is 2 SHFL + register moves, and 40 sums x 5 stages of them were 45 KB of straight-line code that the
// instruction cache had to stream for every solve), and the order of the additions is the sequential one of the
// CPU implementations.
constexpr int TSTRIDE = 33;
constexpr int LM_ROWS = 14, LM_SUMS = 28;
static_assert(sizeof(WarpScratch) >= (LM_ROWS * TSTRIDE + LM_SUMS) * sizeof(double), "LM scratch");
// sum of term(p) over the points p < np, in index order like the CPU implementations (cv::mulTransposed, Eigen's
// J^T J): 5-point EPnP samples are ill-conditioned enough for the order of these additions to decide whether a
// borderline point is inside RANSAC's 5 px -- one of 48 consensus masks differed from OpenCV's with two
// interleaved partial sums.  Lanes outside the point set park zeros, so no mask is needed.
template <class F>
__device__ __forceinline__ double sum_points(int np, F term) {
  double s = 0.0;
#pragma unroll 2
  for (int p = 0; p < np; ++p) s += term(p);
  return s;
}
struct Cam { double fu, fv, uc, vc; };

// Phase clocks (tuning build only): thread 0 of CTA 0 accumulates the SM cycles between successive marks, read and
// cleared through epb_debug_pose_phase_clocks (tools/pose_phases.py).  Compiles to nothing in the product library.
#ifdef EPB_TUNING
__device__ long long g_phase_clk[32];
__device__ long long g_phase_last;
#define POSE_PHASE(i)                                                   \
  do {                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                          \
      const long long t_ = clock64();                                   \
      g_phase_clk[i] += t_ - g_phase_last;                              \
      g_phase_last = t_;                                                \
    }                                                                   \
  } while (0)
#define POSE_PHASE_START()                                              \
  do { if (blockIdx.x == 0 && threadIdx.x == 0) g_phase_last = clock64(); } while (0)
#else
#define POSE_PHASE(i) do {} while (0)
#define POSE_PHASE_START() do {} while (0)
#endif


// ------------------------------------------------------------------------------ small linear algebra
// OpenCV cv::SVD (JacobiSVDImpl_) on a 3x3: rows of `at` are the columns of A.  Sign convention
// matters for the EPnP control points (oracle/epnp_port.py svd_onesided_cv).
__device__ __noinline__ void svd3_cv(const double a[9], double w[3], double ut[9], double vt[9]) {
  double at[3][3], v[3][3], ww[3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) { at[i][j] = a[j * 3 + i]; v[i][j] = (i == j) ? 1.0 : 0.0; }
#pragma unroll
  for (int i = 0; i < 3; ++i) ww[i] = at[i][0] * at[i][0] + at[i][1] * at[i][1] + at[i][2] * at[i][2];
  const double eps2 = (DBL_EPSILON * 10) * (DBL_EPSILON * 10);
  for (int it = 0; it < 30; ++it) {
    bool changed = false;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = i + 1; j < 3; ++j) {
        const double aa = ww[i], bb = ww[j];
        double p = at[i][0] * at[j][0] + at[i][1] * at[j][1] + at[i][2] * at[j][2];
        // OpenCV's test |p| <= eps sqrt(aa bb), squared (no square root on the serial path)
        if (!(p * p <= eps2 * (aa * bb))) {
          // OpenCV: p *= 2, beta = aa - bb, gamma = hypot(p, beta) and, for beta >= 0,
          //   c = sqrt((gamma + beta) / (2 gamma)), s = p / (2 gamma c)       (beta < 0: the roles of c and s swap)
          // computed here from two reciprocal square roots (1 / gamma and 1 / big) instead of a hypot, a square
          // root and two divisions: the same numbers within rounding at ~200 instead of ~500 cycles per rotation.
          p *= 2.0;
          const double beta = aa - bb;
          const double g2 = fma(p, p, beta * beta);
          const double rg = rsqrt(g2);
          const double big2 = fma(0.5 * fabs(beta), rg, 0.5);
          const double rb = rsqrt(big2);
          const double big = big2 * rb, small = (0.5 * p * rg) * rb;
          const double c = beta < 0 ? small : big, s = beta < 0 ? big : small;
          double na = 0, nb = 0;
#pragma unroll
          for (int k = 0; k < 3; ++k) {
            const double t0 = c * at[i][k] + s * at[j][k], t1 = -s * at[i][k] + c * at[j][k];
            at[i][k] = t0; at[j][k] = t1; na += t0 * t0; nb += t1 * t1;
            const double u0 = c * v[i][k] + s * v[j][k], u1 = -s * v[i][k] + c * v[j][k];
            v[i][k] = u0; v[j][k] = u1;
          }
          ww[i] = na; ww[j] = nb;
          changed = true;
        }
      }
    if (!changed) break;
  }
  double winv[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double ss = at[i][0] * at[i][0] + at[i][1] * at[i][1] + at[i][2] * at[i][2];
    const double r = ss > 0 ? rsqrt(ss) : 0.0;      // 1 / w
    winv[i] = r; ww[i] = ss * r;
  }
  auto swap_rows = [&](int i, int j) {
    double t = ww[i]; ww[i] = ww[j]; ww[j] = t;
    t = winv[i]; winv[i] = winv[j]; winv[j] = t;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      t = at[i][k]; at[i][k] = at[j][k]; at[j][k] = t;
      t = v[i][k]; v[i][k] = v[j][k]; v[j][k] = t;
    }
  };
  // selection sort, descending, first maximum (OpenCV: if (W[j] < W[k]) j = k)
  {
    const bool j1 = ww[0] < ww[1];
    const double m01 = j1 ? ww[1] : ww[0];
    const bool j2 = m01 < ww[2];
    if (j2) swap_rows(0, 2); else if (j1) swap_rows(0, 1);
    if (ww[1] < ww[2]) swap_rows(1, 2);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    w[i] = ww[i];
    const double inv = winv[i];
#pragma unroll
    for (int k = 0; k < 3; ++k) { ut[i * 3 + k] = at[i][k] * inv; vt[i * 3 + k] = v[i][k]; }
  }
}

// Parallel-order Jacobi eigen-decomposition of the symmetric 12x12 in ws.A (destroyed), the four eigenvectors of
// the smallest eigenvalues to ws.vs (ascending).  Both matrices stay in shared memory and the warp works BY 2x2
// BLOCK: a round rotates the six index pairs (2k, 2k+1) at once, A <- J^T A J, V <- V J with J = diag of six plane
// rotations, and under that the 2x2 block (pair I rows, pair J columns) of A only mixes with itself,
// B' = J_I^T B J_J.
//   1. lanes 0..5 compute the (c, s) of their pair from three elements of A and publish them;
//   2. lanes 0..20 own the 21 blocks I <= J (two 16-byte loads, two small products, stores of the block and of its
//      mirror image), and all lanes rotate the 72 (row, pair) element pairs of V, at most three each.
// So that every pair of indices meets, the contents then move by the fixed tournament permutation pi (position 0
// stays, the others go round: all 66 pairs in 11 rounds, back in place after a sweep).  The move costs nothing: the
// results are STORED at their permuted positions into a second copy of A and V (the two copies swap roles every
// round), so every lane reads and writes the same static offsets in every round -- no schedule look-ups, no index
// arithmetic.  The solve is the program of ONE warp, and one warp alone issues an instruction every 3-4 cycles and
// a shared-memory access every ~7 (tools/micro/warp_issue.cu): what counts is instructions per round, ~120 here
// against ~500 for the register-resident version of round 1 (one column per lane, rows exchanged by shuffle).
// a2 = 12 x 14 doubles of scratch for the second copy of A; the second copy of V overlays ws.S[12..], ws.L and
// ws.vs, which are all dead while the solver runs.
__device__ __noinline__ void jacobi_eigh12(WarpScratch& ws, int lane, double* a2) {
  constexpr int LD = 14;
  // pi[x]: where the content of position x goes after a round
  auto pi = [](int x) -> int { return (int)((0x9b7a58361420ull >> (4 * x)) & 15ull); };   // {0,2,4,1,6,3,8,5,10,7,11,9}
  double* A0 = &ws.A[0][0];
  double* V0 = &ws.V[0][0];
  double* V1 = ws.S + 12;
  static_assert(80 - 12 + 60 + 48 >= 12 * LD, "second copy of V");
  for (int e = lane; e < 144; e += 32) ws.V[e / 12][e % 12] = (e / 12 == e % 12) ? 1.0 : 0.0;
  int bI = 0, bJ = 0;                       // this lane's block of A (lanes 0..20)
  {
    int rem = lane;
    while (bI < 5 && rem >= 6 - bI) { rem -= 6 - bI; ++bI; }
    bJ = min(bI + rem, 5);
  }
  // static element offsets: loads of the block rows (2 bI, 2 bI + 1) x cols (2 bJ, 2 bJ + 1), stores at pi(...)
  const int oL0 = (2 * bI) * LD + 2 * bJ, oL1 = (2 * bI + 1) * LD + 2 * bJ;
  const int r0 = pi(2 * bI), r1 = pi(2 * bI + 1), c0 = pi(2 * bJ), c1 = pi(2 * bJ + 1);
  const int oS00 = r0 * LD + c0, oS01 = r0 * LD + c1, oS10 = r1 * LD + c0, oS11 = r1 * LD + c1;
  const int oM00 = c0 * LD + r0, oM01 = c1 * LD + r0, oM10 = c0 * LD + r1, oM11 = c1 * LD + r1;   // mirror image
  const bool hasA = lane < 21, mirror = hasA && bI != bJ;
  // this lane's (row, pair) items of V: item = lane + 32 m < 72, row = item / 6, pair = item % 6
  int vJ[3], oVL[3], oVp[3], oVq[3];
#pragma unroll
  for (int m = 0; m < 3; ++m) {
    const int item = min(lane + 32 * m, 71), vr = item / 6;
    vJ[m] = item - 6 * vr;
    oVL[m] = vr * LD + 2 * vJ[m];
    oVp[m] = vr * LD + pi(2 * vJ[m]); oVq[m] = vr * LD + pi(2 * vJ[m] + 1);
  }
  const bool v2 = lane + 64 < 72;
  const int kp = min(lane, 5);              // pivot pair of lanes 0..5
  const int oP0 = (2 * kp) * LD + 2 * kp, oP1 = (2 * kp + 1) * LD + 2 * kp + 1;
  double2* cs = reinterpret_cast<double2*>(ws.S);      // [6]: (c, s) of the six pairs of the round
  double* sA = A0; double* dA = a2; double* sV = V0; double* dV = V1;
  __syncwarp();
  double off_prev = INFINITY;
  for (int sweep = 0; sweep < 30; ++sweep) {
    // Convergence is quadratic and never there before the fourth sweep on a full matrix, so the first tests are
    // skipped (an input that is already diagonal just rotates by the identity for four sweeps).
    if (sweep >= 4) {
      double off = 0.0, tr = 0.0;
      for (int e = lane; e < 144; e += 32) {
        const int r = e / 12, c = e - 12 * r;
        const double v = sA[r * LD + c];
        if (r == c) tr += fabs(v); else off += v * v;
      }
      off = 0.5 * warp_sum(off);
      tr = warp_sum(tr);
      // rounding floor of an exact similarity: (12 eps tr)^2 ~ 2e-30 tr^2.  The sweeps converge quadratically
      // (off / tr^2: ... 1e-10, 1e-19, 1e-33 on this problem family), so the first sweep that lands below 1e-28 is
      // the last useful one -- the register version ran one more to SEE the stagnation.  That test stays as the
      // fallback (and `!(off > 0)` ends exact and non-finite cases).
      if (!(off > 1e-280) || off < 1e-28 * tr * tr) break;
      if (off < 1e-24 * tr * tr && off > 0.5 * off_prev) break;
      off_prev = off;
    }
    POSE_PHASE(20);
#ifdef EPB_TUNING
    if (blockIdx.x == 0 && threadIdx.x == 0) g_phase_clk[23] += 1;
#endif
#pragma unroll 1
    for (int step = 0; step < 11; ++step) {
      // everything the lane will touch is loaded up front; only the (c, s) round trip through shared memory sits
      // between the rotation parameters and the updates
      const double2 piv = *reinterpret_cast<const double2*>(sA + oP0);   // a_pp, a_pq
      const double aqq = sA[oP1];
      const double2 row0 = *reinterpret_cast<const double2*>(sA + oL0);
      const double2 row1 = *reinterpret_cast<const double2*>(sA + oL1);
      double2 vv[3];
#pragma unroll
      for (int m = 0; m < 3; ++m) vv[m] = *reinterpret_cast<const double2*>(sV + oVL[m]);
      if (lane < 6) {
        // Rotation that annihilates a_pq: t = tan(theta) is the small root of  b t^2 + 2 a t - b = 0  with
        // a = (aqq - app)/2, b = a_pq, i.e. t = sgn(a) b / (|a| + h), h = hypot(a, b), and from it
        //   c^2 = (h + |a|) / (2 h),   s = sgn(a) b / (2 h c).
        // Two reciprocal square roots (1/h and 1/c) and a handful of multiplications: ~200 cycles of dependent
        // latency, against ~800 for the float32-seeded Newton form of round 1 (MUFU and F2F conversions cost 20-45
        // cycles each on this part, tools/micro/fp64_latency.cu).  c^2 + s^2 = 1 within rounding and A is
        // transformed with those very (c, s) (an exact orthogonal similarity, no "a_pq := 0" shortcut), so any
        // residual a_pq is removed by the next sweep.
        const double app = piv.x, apq = piv.y;
        const double alpha = 0.5 * (aqq - app);
        const double h2 = fma(alpha, alpha, apq * apq);
        double c = 1.0, s = 0.0;
        if (apq != 0.0 && h2 > 1e-290 && h2 < 1e290) {
          const double rh = rsqrt(h2);                               // 1 / h
          const double c2 = fma(0.5 * fabs(alpha), rh, 0.5);        // in [0.5, 1]
          const double rc = rsqrt(c2);                               // 1 / c
          const double hb = 0.5 * apq * rh;
          c = c2 * rc;
          s = alpha < 0.0 ? -(hb * rc) : hb * rc;
        }
        cs[lane] = make_double2(c, s);
      }
      __syncwarp();
      POSE_PHASE(21);
      {
        const double2 csI = cs[bI], csJ = cs[bJ];
        const double cI = csI.x, sI = csI.y, cJ = csJ.x, sJ = csJ.y;
        const double b00 = row0.x, b01 = row0.y, b10 = row1.x, b11 = row1.y;
        // columns: [col_p, col_q] <- [c col_p - s col_q, s col_p + c col_q]; rows likewise with the row pair's (c, s)
        const double t00 = fma(cJ, b00, -sJ * b01), t01 = fma(sJ, b00, cJ * b01);
        const double t10 = fma(cJ, b10, -sJ * b11), t11 = fma(sJ, b10, cJ * b11);
        const double n00 = fma(cI, t00, -sI * t10), n10 = fma(sI, t00, cI * t10);
        const double n01 = fma(cI, t01, -sI * t11), n11 = fma(sI, t01, cI * t11);
        double np[3], nq[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          const double2 cv = cs[vJ[m]];
          np[m] = fma(cv.x, vv[m].x, -cv.y * vv[m].y);
          nq[m] = fma(cv.y, vv[m].x, cv.x * vv[m].y);
        }
        if (hasA) { dA[oS00] = n00; dA[oS01] = n01; dA[oS10] = n10; dA[oS11] = n11; }
        if (mirror) { dA[oM00] = n00; dA[oM01] = n01; dA[oM10] = n10; dA[oM11] = n11; }
        dV[oVp[0]] = np[0]; dV[oVq[0]] = nq[0];
        dV[oVp[1]] = np[1]; dV[oVq[1]] = nq[1];
        if (v2) { dV[oVp[2]] = np[2]; dV[oVq[2]] = nq[2]; }
      }
      __syncwarp();
      { double* t = sA; sA = dA; dA = t; t = sV; sV = dV; dV = t; }
      POSE_PHASE(22);
    }
  }
  // ascending order of the eigenvalues (stable); value at position x <-> column x of V (a whole number of sweeps
  // leaves every content at its original position)
  if (lane == 0) {
    for (int i = 0; i < 12; ++i) ws.order[i] = i;
    for (int i = 1; i < 12; ++i) {
      const int oi = ws.order[i];
      const double wi = sA[oi * LD + oi];
      int j = i - 1;
      while (j >= 0 && sA[ws.order[j] * LD + ws.order[j]] > wi) { ws.order[j + 1] = ws.order[j]; --j; }
      ws.order[j + 1] = oi;
    }
  }
  __syncwarp();
  // (the current copy of V may be the one that overlays ws.vs: read everything, then write)
  const int e1 = min(lane + 32, 47);
  const double x0 = sV[(lane % 12) * LD + ws.order[lane / 12]];
  const double x1 = sV[(e1 % 12) * LD + ws.order[e1 / 12]];
  __syncwarp();
  ws.vs[lane / 12][lane % 12] = x0;
  if (lane + 32 < 48) ws.vs[e1 / 12][e1 % 12] = x1;
  __syncwarp();
}

// Householder least squares, 6 x ncols (ncols <= 5; the columns beyond ncols must be zero and get x = 0).  One
// copy with the width as a run-time (per-lane) value serves the three beta initialisations (4, 3 and 5 columns,
// side by side in one pass instead of three divergent ones) and the Gauss-Newton steps (epnp_core calls it from a
// single site): the unrolled body is ~1 k instructions and used to exist four times.
// Same reflections as OpenCV's (v = a_k - alpha e_k, alpha = -sgn(a_kk) |a_k|), arranged for latency: this is a
// serial chain that runs 8 times per EPnP solve, and an IEEE square root or division costs 80-100 cycles on this
// part against 9 for an FMA (tools/micro/fp64_latency.cu).  With rs = rsqrt(|a_k|^2):  |alpha| = |a_k|^2 rs,
// v^T v = 2 |alpha| (|alpha| + |a_kk|), so 2 / v^T v = rs / (|alpha| + |a_kk|) -- one rsqrt and one reciprocal per
// column -- and the back-substitution divides by R_kk = alpha, i.e. multiplies by -+rs, already known.  The dot
// products run as two interleaved partial sums.
__device__ __forceinline__ void lstsq6(double (&a)[6][5], double (&b)[6], double (&x)[5], int ncols) {
  double rdiag[5];
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    rdiag[k] = 0.0;
    if (k >= ncols) continue;
    double n0 = 0, n1 = 0;
#pragma unroll
    for (int i = k; i < 6; ++i) { if ((i - k) & 1) n1 = fma(a[i][k], a[i][k], n1); else n0 = fma(a[i][k], a[i][k], n0); }
    const double nrm = n0 + n1;
    if (nrm != 0.0) {
      const double rs = rsqrt(nrm), aabs = nrm * rs;
      const double akk = a[k][k];
      const double vkk = akk > 0 ? akk + aabs : akk - aabs;        // a_kk - alpha
      const double inv = rs / (aabs + fabs(akk));                   // 2 / v^T v
      rdiag[k] = akk > 0 ? -rs : rs;                                // 1 / alpha = 1 / R_kk
#pragma unroll
      for (int j = k + 1; j < 6; ++j) {                             // j == 5: the right-hand side
        double d0 = vkk * (j < 5 ? a[k][j < 5 ? j : 0] : b[k]), d1 = 0;
#pragma unroll
        for (int i = k + 1; i < 6; ++i) {
          const double cij = j < 5 ? a[i][j < 5 ? j : 0] : b[i];
          if ((i - k) & 1) d1 = fma(a[i][k], cij, d1); else d0 = fma(a[i][k], cij, d0);
        }
        const double d = (d0 + d1) * inv;
        if (j < 5) {
          a[k][j < 5 ? j : 0] -= d * vkk;
#pragma unroll
          for (int i = k + 1; i < 6; ++i) a[i][j < 5 ? j : 0] -= d * a[i][k];
        } else {
          b[k] -= d * vkk;
#pragma unroll
          for (int i = k + 1; i < 6; ++i) b[i] -= d * a[i][k];
        }
      }
    } else {
      rdiag[k] = 1.0 / a[k][k];
    }
  }
#pragma unroll
  for (int i = 4; i >= 0; --i) {
    double s = b[i];
#pragma unroll
    for (int j = i + 1; j < 5; ++j) s -= a[i][j] * x[j];
    x[i] = i < ncols ? s * rdiag[i] : 0.0;
  }
}

// ------------------------------------------------------------------------------ EPnP
struct PoseRT { double R[9]; double t[3]; double err; };

// The three candidates at once.  Every point lane forms its camera-frame point under each candidate's betas and
// parks the nine coordinates next to its centred model point in the transposed table; lanes 0..8 then add up the
// nine centroid coordinates and lanes 0..26 the 3 x 9 entries of the three cross-covariances, over the points of
// the set in index order.  The 3x3 SVD / rotation / translation of candidate c is computed by the lanes with
// lane % 3 == c (three different problems side by side, lane-redundant scalar work), and every point lane then
// scores all three poses.  The pose with the smallest mean reprojection error wins, ties keep the lower index.
// Table (ws.flat()): rows 0..8 = pc[c][k], rows 9..11 = pw - pw0, then 9 centroid sums and 27 covariance sums --
// this overlays A, V, S and L, which are all spent by now, and stays clear of ws.vs.
__device__ void compute_r_and_t3(WarpScratch& ws, const double be[4], const double al[4],
                                 const double pw[3], double u, double v, unsigned amask, int n, int first_lane,
                                 const double pw0[3], const Cam& cam, int lane, PoseRT& best) {
  const int mine = lane % 3;
  const bool active = (amask >> lane) & 1u;
  double* T = ws.flat();
  double* P0 = T + 12 * TSTRIDE;        // [3][3] centroids
  double* CV = P0 + 9;                  // [3][9] cross-covariances
  static_assert(12 * TSTRIDE + 36 <= 12 * 14 * 2 + 80 + 60, "the table must end before ws.vs");
  double pcs[3][3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    double bc[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) bc[k] = __shfl_sync(FULL, be[k], c);   // lane c carries candidate c
    double pc[3] = {0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int k = 0; k < 3; ++k)
        pc[k] += al[j] * (bc[0] * ws.vs[0][3 * j + k] + bc[1] * ws.vs[1][3 * j + k] +
                          bc[2] * ws.vs[2][3 * j + k] + bc[3] * ws.vs[3][3 * j + k]);
    const double z0 = __shfl_sync(FULL, pc[2], first_lane);
#pragma unroll
    for (int k = 0; k < 3; ++k) pcs[c][k] = z0 < 0.0 ? -pc[k] : pc[k];
  }
  __syncwarp();                          // every lane is done reading ws.L / ws.S
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int k = 0; k < 3; ++k) T[(3 * c + k) * TSTRIDE + lane] = pcs[c][k];
#pragma unroll
  for (int k = 0; k < 3; ++k) T[(9 + k) * TSTRIDE + lane] = active ? pw[k] - pw0[k] : 0.0;
  __syncwarp();
  const int np = 32 - __clz(amask);
  if (lane < 9) {
    const double* row = T + lane * TSTRIDE;
    P0[lane] = sum_points(np, [&](int p) { return row[p]; }) / n;
  }
  __syncwarp();
  if (lane < 27) {
    const int c = lane / 9, j = (lane % 9) / 3, k = lane % 3;
    const double* pj = T + (3 * c + j) * TSTRIDE;
    const double* wk = T + (9 + k) * TSTRIDE;
    const double p0 = P0[3 * c + j];
    CV[lane] = sum_points(np, [&](int p) { return (pj[p] - p0) * wk[p]; });
  }
  __syncwarp();
  POSE_PHASE(8);
  double abt[9], pc0m[3];
#pragma unroll
  for (int i = 0; i < 9; ++i) abt[i] = CV[9 * mine + i];
#pragma unroll
  for (int i = 0; i < 3; ++i) pc0m[i] = P0[3 * mine + i];
  double w[3], ut[9], vt[9], R[9], t[3];
  svd3_cv(abt, w, ut, vt);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)
      R[3 * i + j] = ut[0 * 3 + i] * vt[0 * 3 + j] + ut[1 * 3 + i] * vt[1 * 3 + j] + ut[2 * 3 + i] * vt[2 * 3 + j];
  const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] -
                     R[2] * R[4] * R[6] - R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
  if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
#pragma unroll
  for (int i = 0; i < 3; ++i) t[i] = pc0m[i] - (R[3 * i] * pw0[0] + R[3 * i + 1] * pw0[1] + R[3 * i + 2] * pw0[2]);
  POSE_PHASE(9);
  best.err = INFINITY;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    PoseRT cur;
#pragma unroll
    for (int i = 0; i < 9; ++i) cur.R[i] = __shfl_sync(FULL, R[i], c);
#pragma unroll
    for (int i = 0; i < 3; ++i) cur.t[i] = __shfl_sync(FULL, t[i], c);
    const double xc = cur.R[0] * pw[0] + cur.R[1] * pw[1] + cur.R[2] * pw[2] + cur.t[0];
    const double yc = cur.R[3] * pw[0] + cur.R[4] * pw[1] + cur.R[5] * pw[2] + cur.t[1];
    const double inv_z = 1.0 / (cur.R[6] * pw[0] + cur.R[7] * pw[1] + cur.R[8] * pw[2] + cur.t[2]);
    const double du = u - (cam.uc + cam.fu * xc * inv_z), dv = v - (cam.vc + cam.fv * yc * inv_z);
    cur.err = warp_sum(active ? sqrt(du * du + dv * dv) : 0.0) / n;
    if (c == 0 || cur.err < best.err) best = cur;
  }
}

// EPnP over the lanes flagged `active` (n = popcount >= 4).  All lanes return the same pose.
// (one out-of-line copy: the RANSAC candidates, the speculative all-point solve and the final solve over the
// consensus set share it, so the W = 2 / 4 / 8 instantiations of the kernels execute the SAME machine code and
// return bit-identical poses whatever the batch size; it also takes two inlined copies out of the instruction cache)
__device__ __noinline__ void epnp_core(WarpScratch& ws, double* a2, int lane, bool active, int n, int first_lane,
                          const double pw_in[3], double u, double v, const Cam& cam, PoseRT& best) {
  double pw[3] = {active ? pw_in[0] : 0.0, active ? pw_in[1] : 0.0, active ? pw_in[2] : 0.0};
  const unsigned amask = __ballot_sync(FULL, active);
  double* T = ws.flat();                 // transposed per-point table, rows of TSTRIDE (overlays A and V)
  // control points: centroid + PCA axes (OpenCV SVD sign convention)
#pragma unroll
  for (int k = 0; k < 3; ++k) T[k * TSTRIDE + lane] = pw[k];
  __syncwarp();
  const int np = 32 - __clz(amask);       // lanes outside the set hold zeros in every table below
  if (lane < 3) {
    const double* row = T + lane * TSTRIDE;
    ws.S[lane] = sum_points(np, [&](int p) { return row[p]; }) / n;
  }
  __syncwarp();
  double c0[3], d[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) { c0[k] = ws.S[k]; d[k] = active ? pw[k] - c0[k] : 0.0; }
#pragma unroll
  for (int k = 0; k < 3; ++k) T[(3 + k) * TSTRIDE + lane] = d[k];
  __syncwarp();
  if (lane < 6) {                        // xx xy xz yy yz zz
    const int j = lane < 3 ? 0 : (lane < 5 ? 1 : 2), k = lane < 3 ? lane : (lane < 5 ? lane - 2 : 2);
    const double* dj = T + (3 + j) * TSTRIDE;
    const double* dk = T + (3 + k) * TSTRIDE;
    ws.S[4 + lane] = sum_points(np, [&](int p) { return dj[p] * dk[p]; });
  }
  __syncwarp();
  double cov[9];
  cov[0] = ws.S[4]; cov[1] = ws.S[5]; cov[2] = ws.S[6]; cov[4] = ws.S[7]; cov[5] = ws.S[8]; cov[8] = ws.S[9];
  cov[3] = cov[1]; cov[6] = cov[2]; cov[7] = cov[5];
  double w[3], uct[9], vt_unused[9];
  POSE_PHASE(1);
  svd3_cv(cov, w, uct, vt_unused);
  double kk[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) kk[i] = sqrt(w[i] / n);
  // barycentric coordinates: CC = [k_i * uct_i] has orthogonal columns, CC^-1 rows = uct_i / k_i
  double al[4];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double inv = kk[i] > 1e-12 * kk[0] ? 1.0 / kk[i] : 0.0;
    al[i + 1] = (uct[3 * i] * d[0] + uct[3 * i + 1] * d[1] + uct[3 * i + 2] * d[2]) * inv;
  }
  al[0] = 1.0 - al[1] - al[2] - al[3];
  if (!active) { al[0] = al[1] = al[2] = al[3] = 0.0; }
  POSE_PHASE(2);
  // M^T M from four families of pair sums: sum al_j al_k {1, du, dv, du^2 + dv^2}, 10 pairs j <= k each
  {
    const double du = cam.uc - u, dv = cam.vc - v;
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) T[j * TSTRIDE + lane] = al[j];
    T[4 * TSTRIDE + lane] = du; T[5 * TSTRIDE + lane] = dv; T[6 * TSTRIDE + lane] = du * du + dv * dv;
    __syncwarp();
    for (int e = lane; e < 40; e += 32) {
      const int fam = e / 10;
      int j = 0, rem = e - 10 * fam;
      while (rem >= 4 - j) { rem -= 4 - j; ++j; }
      const double* aj = T + j * TSTRIDE;
      const double* ak = T + (j + rem) * TSTRIDE;
      const double* f = T + (3 + fam) * TSTRIDE;
      ws.S[40 + e] = fam ? sum_points(np, [&](int p) { return aj[p] * ak[p] * f[p]; })
                         : sum_points(np, [&](int p) { return aj[p] * ak[p]; });
    }
    __syncwarp();
  }
  for (int e = lane; e < 144; e += 32) {
    const int r = e / 12, c = e % 12;
    const int j = r / 3, cr = r % 3, k = c / 3, cc = c % 3;
    const int lo = j < k ? j : k, hi = j < k ? k : j;
    const int pidx = 40 + lo * 4 - lo * (lo - 1) / 2 + (hi - lo);  // index of pair (lo,hi), lo<=hi
    const double s0 = ws.S[pidx], s1 = ws.S[10 + pidx], s2 = ws.S[20 + pidx], s3 = ws.S[30 + pidx];
    double val;
    if (cr == 0 && cc == 0) val = s0 * cam.fu * cam.fu;
    else if (cr == 1 && cc == 1) val = s0 * cam.fv * cam.fv;
    else if (cr == 2 && cc == 2) val = s3;
    else if ((cr == 0 && cc == 2) || (cr == 2 && cc == 0)) val = s1 * cam.fu;
    else if ((cr == 1 && cc == 2) || (cr == 2 && cc == 1)) val = s2 * cam.fv;
    else val = 0.0;
    ws.A[r][c] = val;
  }
  __syncwarp();
  POSE_PHASE(3);
  jacobi_eigh12(ws, lane, a2);
  POSE_PHASE(4);
  // L_6x10 and rho
  {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    for (int e = lane; e < 60; e += 32) {
      const int i = e / 10, c = e % 10;
      // column c <-> (m,n): B11 B12 B22 B13 B23 B33 B14 B24 B34 B44
      const int cm[10] = {0, 0, 1, 0, 1, 2, 0, 1, 2, 3}, cn[10] = {0, 1, 1, 2, 2, 2, 3, 3, 3, 3};
      const int m = cm[c], nn = cn[c];
      double dot = 0;
      for (int k = 0; k < 3; ++k) {
        const double dm = ws.vs[m][3 * pa[i] + k] - ws.vs[m][3 * pb[i] + k];
        const double dn = ws.vs[nn][3 * pa[i] + k] - ws.vs[nn][3 * pb[i] + k];
        dot += dm * dn;
      }
      ws.L[i][c] = (m == nn) ? dot : 2.0 * dot;
    }
  }
  __syncwarp();
  // rho: squared distances between control points = k_i^2 (+ k_j^2), since the axes are orthonormal;
  // computed from the control points themselves like the reference implementation
  double cw[4][3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    cw[0][k] = c0[k];
#pragma unroll
    for (int i = 0; i < 3; ++i) cw[i + 1][k] = c0[k] + kk[i] * uct[3 * i + k];
  }
  double rho[6];
  {
    auto d2 = [&](int a, int b) {
      const double x = cw[a][0] - cw[b][0], y = cw[a][1] - cw[b][1], z = cw[a][2] - cw[b][2];
      return x * x + y * y + z * z;
    };
    rho[0] = d2(0, 1); rho[1] = d2(0, 2); rho[2] = d2(0, 3); rho[3] = d2(1, 2); rho[4] = d2(1, 3); rho[5] = d2(2, 3);
  }
  double pw0[3] = {c0[0], c0[1], c0[2]};

  best.err = INFINITY;
  // The three beta initialisations and their Gauss-Newton refinements are independent and lane-redundant
  // scalar work: lane l carries candidate l % 3, so the 3 x 5 Gauss-Newton solves run side by side
  // (only the three different initial solves diverge); the betas are then broadcast per candidate.
  POSE_PHASE(5);
  double be[4] = {0, 0, 0, 0};
  {
    const int cand = lane % 3;
    // One loop, six passes through ONE inlined copy of the least-squares solver (everything stays in registers;
    // out of line its arrays lived in local memory and the loads sat on the serial path):
    //   pass 0: the initial solve -- cand 0 -> [B11 B12 B13 B14] (columns 0 1 3 6 of L), cand 1 -> [B11 B12 B22]
    //           (0 1 2), cand 2 -> [B11 B12 B22 B13 B23] (0 1 2 3 4), side by side with a per-lane width;
    //   passes 1..5: the Gauss-Newton steps on the control-point distance constraints.
    const int c2 = cand == 0 ? 3 : 2, c3 = cand == 0 ? 6 : 3;
#pragma unroll 1
    for (int pass = 0; pass < 6; ++pass) {
      double a[6][5], b[6], x[5];
      int ncols = 4;
      if (pass == 0) {
        ncols = cand == 0 ? 4 : (cand == 1 ? 3 : 5);
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          a[i][0] = ws.L[i][0]; a[i][1] = ws.L[i][1]; a[i][2] = ws.L[i][c2];
          a[i][3] = cand == 1 ? 0.0 : ws.L[i][c3];
          a[i][4] = cand == 2 ? ws.L[i][4] : 0.0;
          b[i] = rho[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double* r = ws.L[i];
          a[i][0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3];
          a[i][1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3];
          a[i][2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3];
          a[i][3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3];
          a[i][4] = 0.0;
          b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] +
                           r[3] * be[0] * be[2] + r[4] * be[1] * be[2] + r[5] * be[2] * be[2] +
                           r[6] * be[0] * be[3] + r[7] * be[1] * be[3] + r[8] * be[2] * be[3] +
                           r[9] * be[3] * be[3]);
        }
      }
      POSE_PHASE(24);
      lstsq6(a, b, x, ncols);
      POSE_PHASE(25);
      if (pass == 0) {
        if (cand == 0) {
          if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
          else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
        } else {
          if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
          else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
          if (x[1] < 0) be[0] = -be[0];
          if (cand == 2) be[2] = x[3] / be[0];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) be[k] += x[k];
      }
    }
  }
  POSE_PHASE(7);
  compute_r_and_t3(ws, be, al, pw, u, v, amask, n, first_lane, pw0, cam, lane, best);
  POSE_PHASE(10);
}

#ifdef EPB_TUNING
__constant__ int c_ransac_round0_warps = 4;   // EPB_RANSAC_R0: warps of round 0 (measurements only)
#endif

// cv::RNG (multiply-with-carry), seed (uint64)-1 as RANSACPointSetRegistrator::run uses
struct CvRng {
  unsigned long long state;
  __device__ unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ull + (state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

__device__ int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.0), 1.0);
  ep = fmin(fmax(ep, 0.0), 1.0);
  double num = fmax(1.0 - p, DBL_MIN);
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num); denom = log(denom);
  return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// cv2.solvePnPRansac(flags=EPNP) restated (oracle/epnp_port.py solve_pnp_ransac_epnp).
// pw/u/v are this lane's correspondence (already rounded to float32 by the caller, as
// OpenCV converts its inputs to CV_32F).  Returns status; all lanes of warp 0 (of CTA 0 of the frame) hold the result.
//
// W warps per CTA and C CTAs (a thread-block cluster) per image, W * C in {2, 4, 8} candidate slots.  OpenCV's loop
// is sequential -- iteration k draws its 5-point sample from cv::RNG, solves, scores, and a better consensus
// shortens `niters` -- but the SAMPLES do not depend on the data: every warp runs the same generator and knows the
// sample of every iteration.  So the iterations are evaluated W * C at a time, one per warp, and the sequential
// bookkeeping (strictly-better consensus, niters update, stop) is then REPLAYED in iteration order over the results
// by every warp identically.  A result beyond the shortened niters is ignored, exactly as if it had never been
// computed: the consensus set equals OpenCV's for every input, while one outlier in 11 points costs one round
// instead of ~5 serial solves and the no-consensus case ceil(100 / (W C)) rounds instead of 100.
// Round 0 keeps the round-1 trick: OpenCV ends with one EPnP over the consensus set; on clean frames that set is
// "all points" and is known only after the first sample has been solved and scored, so warp 1 solves EPnP over
// ALL points speculatively during round 0 and warp 0 takes its result when the consensus turns out to be everything.
//
// Why clusters (round 2): one warp alone is bound by what a single warp can issue (3-4 cycles per instruction,
// ~7 per shared-memory access: tools/micro/warp_issue.cu), and two solver warps on one SM sub-partition share its
// FP64 and shared-memory pipes and each run at about half speed.  A round of 8 candidates on one SM therefore took
// twice as long as a round of 4.  While the batch leaves SMs idle (2 B <= SMs) a frame gets a cluster of two CTAs
// of four warps -- one warp per sub-partition on two SMs -- and the eight candidates of a round run at full speed;
// the per-slot results are written into both CTAs' shared memory (DSMEM) and one cluster barrier per round
// publishes them (the slots are double-buffered by round parity).
struct RansacShared {
  PoseRT spec;          // EPnP over all points (warp 1 of CTA 0, round 0)
  int cnt[2][8];        // [round parity][slot]: consensus size of the candidate, -1 = not evaluated / non-finite model
  unsigned mask[2][8];
};

__device__ __forceinline__ unsigned cluster_cta_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
// store to the same shared-memory variable in CTA `rank` of the cluster
__device__ __forceinline__ void st_cluster_u32(void* local, unsigned rank, unsigned v) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(ra), "r"(v) : "memory");
}
template <int C>
__device__ __forceinline__ void frame_barrier() {
  if (C == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int W, int C>
__device__ int pnp_ransac_epnp(WarpScratch& ws, double* a2, int lane, int n, const double pw[3], double u, double v,
                               const Cam& cam, double reproj_err, int max_iters, double confidence,
                               PoseRT& out, unsigned& inlier_mask, int warp, RansacShared* sh) {
  constexpr int WT = W * C;                          // candidate slots per round
  static_assert(WT >= 2 && WT <= 8 && (C == 1 || C == 2), "slots of RansacShared");
  const int rank = C == 1 ? 0 : (int)cluster_cta_rank();
  const int gw = rank * W + warp;                    // warp index within the frame
  const int model_points = 5;
  inlier_mask = 0;
  if (n < model_points) return EPB_POSE_TOO_FEW;   // (uniform over the frame's CTAs: no barrier has been executed)
  // a non-finite correspondence (decode's NaN-is-max policy hands a NaN keypoint on) fails the frame as a whole
  // instead of silently becoming an outlier: NaN pose, EPB_POSE_FAILED (uniform over the frame as well)
  if (__any_sync(FULL, lane < n && !(isfinite(pw[0]) && isfinite(pw[1]) && isfinite(pw[2]) && isfinite(u) && isfinite(v))))
    return EPB_POSE_FAILED;
  const unsigned all = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
  if (n == model_points) {       // the minimal sample is the whole set: OpenCV solves it once
    if (gw == 1) {
      PoseRT sp;
      epnp_core(ws, a2, lane, lane < n, n, 0, pw, u, v, cam, sp);
      if (lane == 0) sh->spec = sp;
    }
    frame_barrier<C>();
    out = sh->spec;              // (meaningful in CTA 0, which is where warp 0 lives)
    inlier_mask = all;
    return EPB_POSE_OK;
  }
  CvRng rng{0xFFFFFFFFFFFFFFFFull};
  int niters = max_iters;
  unsigned best_mask = 0;
  int best_count = 0;
  const float thr2 = (float)(reproj_err * reproj_err);
  int it0 = 0;                                       // iterations replayed so far
  for (int round = 0; it0 < niters; ++round) {
    // Round 0 is what a clean frame pays: at most four warps per CTA work in it (one per scheduler: the critical
    // warps 0 and 1 do not share issue slots), warp 1 on the speculative solve; the others join from round 1 on.
#ifdef EPB_TUNING
    const int W0 = min(W, max(2, c_ransac_round0_warps));
#else
    constexpr int W0 = W < 4 ? W : 4;
#endif
    const int g0 = rank * W0 + warp;                 // index among the warps of round 0
    const int slots = round == 0 ? W0 * C - 1 : WT;
    const int my_slot = round == 0 ? (g0 == 0 ? 0 : g0 - 1) : gw;
    const bool speculative = round == 0 && gw == 1;
    const bool idle = round == 0 && warp >= W0;
    const int par = round & 1;
    // advance the generator through the round; keep the sample of this warp's iteration
    int idx0 = 0;
    unsigned my_m = 0;
    for (int sl = 0; sl < slots; ++sl) {
      unsigned m = 0;
      int first = 0;
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        int k = rng.uniform(0, n);
        while (m & (1u << k)) k = rng.uniform(0, n);
        if (i == 0) first = k;
        m |= 1u << k;
      }
      if (sl == my_slot && !speculative) { my_m = m; idx0 = first; }
    }
    if (speculative) {
      PoseRT sp;
      epnp_core(ws, a2, lane, lane < n, n, 0, pw, u, v, cam, sp);
      if (lane == 0) sh->spec = sp;
    } else {
      int cnt = -1;
      unsigned gm = 0;
      if (!idle && it0 + my_slot < niters) {
        PoseRT cur;
        epnp_core(ws, a2, lane, (my_m >> lane) & 1u, 5, idx0, pw, u, v, cam, cur);
        bool finite = true;
#pragma unroll
        for (int i = 0; i < 9; ++i) finite = finite && isfinite(cur.R[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i) finite = finite && isfinite(cur.t[i]);
        if (finite) {
          const double xc = cur.R[0] * pw[0] + cur.R[1] * pw[1] + cur.R[2] * pw[2] + cur.t[0];
          const double yc = cur.R[3] * pw[0] + cur.R[4] * pw[1] + cur.R[5] * pw[2] + cur.t[1];
          const double zc = cur.R[6] * pw[0] + cur.R[7] * pw[1] + cur.R[8] * pw[2] + cur.t[2];
          const double du = u - (cam.uc + cam.fu * xc / zc), dv = v - (cam.vc + cam.fv * yc / zc);
          const bool good = lane < n && ((float)(du * du + dv * dv) <= thr2);
          gm = __ballot_sync(FULL, good);
          cnt = __popc(gm);
        }
      }
      if (lane == 0 && !idle) {
        if (C == 1) {
          sh->cnt[par][my_slot] = cnt; sh->mask[par][my_slot] = gm;
        } else {
#pragma unroll
          for (int r = 0; r < C; ++r) {
            st_cluster_u32(&sh->cnt[par][my_slot], r, (unsigned)cnt);
            st_cluster_u32(&sh->mask[par][my_slot], r, gm);
          }
        }
      }
    }
    frame_barrier<C>();
    // replay OpenCV's sequential bookkeeping over the round (identical in every warp); the next round writes the
    // other parity, and the one after that is separated from these reads by the next barrier
    for (int sl = 0; sl < slots && it0 + sl < niters; ++sl) {
      const int cnt = sh->cnt[par][sl];
      if (cnt > max(best_count, model_points - 1)) {
        best_mask = sh->mask[par][sl]; best_count = cnt;
        niters = ransac_update_num_iters(confidence, (double)(n - cnt) / n, model_points, niters);
      }
    }
    it0 += slots;
  }
  if (max_iters <= 0) {   // (no round ran: the speculative result does not exist either)
    return EPB_POSE_FAILED;
  }
  if (best_mask == 0) return EPB_POSE_FAILED;
  if (gw != 0) return EPB_POSE_OK;                   // only warp 0 carries the pose on
  if (best_mask == all) out = sh->spec;              // == epnp_core over all points, first_lane 0
  else epnp_core(ws, a2, lane, (best_mask >> lane) & 1u, best_count, __ffs(best_mask) - 1, pw, u, v, cam, out);
  inlier_mask = best_mask;
  return EPB_POSE_OK;
}

// ------------------------------------------------------------------------------ rotations
// cv2.Rodrigues, vector -> matrix
__device__ void rodrigues_to_matrix(const double r[3], double R[9]) {
  const double theta = sqrt(r[0] * r[0] + r[1] * r[1] + r[2] * r[2]);
  if (theta < DBL_EPSILON) {
    R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = 0; R[7] = 0; R[8] = 1;
    return;
  }
  const double c = cos(theta), s = sin(theta), c1 = 1.0 - c, it = 1.0 / theta;
  const double x = r[0] * it, y = r[1] * it, z = r[2] * it;
  R[0] = c + c1 * x * x;     R[1] = c1 * x * y - s * z; R[2] = c1 * x * z + s * y;
  R[3] = c1 * x * y + s * z; R[4] = c + c1 * y * y;     R[5] = c1 * y * z - s * x;
  R[6] = c1 * x * z - s * y; R[7] = c1 * y * z + s * x; R[8] = c + c1 * z * z;
}
// cv2.Rodrigues, matrix -> vector (R assumed orthonormal; OpenCV first re-orthogonalises by SVD)
__device__ void matrix_to_rodrigues(const double R[9], double r[3]) {
  double rx = R[7] - R[5], ry = R[2] - R[6], rz = R[3] - R[1];
  const double s = sqrt((rx * rx + ry * ry + rz * rz) * 0.25);
  double c = (R[0] + R[4] + R[8] - 1.0) * 0.5;
  c = c > 1.0 ? 1.0 : (c < -1.0 ? -1.0 : c);
  double theta = acos(c);
  if (s < 1e-5) {
    if (c > 0) { r[0] = r[1] = r[2] = 0.0; return; }
    double t = (R[0] + 1) * 0.5; rx = sqrt(fmax(t, 0.0));
    t = (R[4] + 1) * 0.5; ry = sqrt(fmax(t, 0.0)) * (R[1] < 0 ? -1.0 : 1.0);
    t = (R[8] + 1) * 0.5; rz = sqrt(fmax(t, 0.0)) * (R[2] < 0 ? -1.0 : 1.0);
    if (fabs(rx) < fabs(ry) && fabs(rx) < fabs(rz) && ((R[5] > 0) != (ry * rz > 0))) rz = -rz;
    theta /= sqrt(rx * rx + ry * ry + rz * rz);
    r[0] = rx * theta; r[1] = ry * theta; r[2] = rz * theta;
  } else {
    const double vth = theta / (2.0 * s);
    r[0] = rx * vth; r[1] = ry * vth; r[2] = rz * vth;
  }
}
// scipy Rotation.from_matrix(R).as_quat() -> (x,y,z,w); returned as (w,x,y,z)
__device__ void matrix_to_quat_wxyz(const double R[9], double q[4]) {
  const double tr = R[0] + R[4] + R[8];
  double x, y, z, w;
  int choice = 0; double best = R[0];
  if (R[4] > best) { best = R[4]; choice = 1; }
  if (R[8] > best) { best = R[8]; choice = 2; }
  if (tr > best) { choice = 3; }
  if (choice == 0) { x = 1 - tr + 2 * R[0]; y = R[3] + R[1]; z = R[6] + R[2]; w = R[7] - R[5]; }
  else if (choice == 1) { y = 1 - tr + 2 * R[4]; z = R[7] + R[5]; x = R[1] + R[3]; w = R[2] - R[6]; }
  else if (choice == 2) { z = 1 - tr + 2 * R[8]; x = R[2] + R[6]; y = R[5] + R[7]; w = R[3] - R[1]; }
  else { x = R[7] - R[5]; y = R[2] - R[6]; z = R[3] - R[1]; w = 1 + tr; }
  const double inv = 1.0 / sqrt(x * x + y * y + z * z + w * w);
  q[0] = w * inv; q[1] = x * inv; q[2] = y * inv; q[3] = z * inv;
}

// ------------------------------------------------------------------------------ LM (TinySolver recipe)
struct J3 { double a, v0, v1, v2; };
__device__ __forceinline__ J3 jc(double a) { return {a, 0, 0, 0}; }
__device__ __forceinline__ J3 operator+(J3 x, J3 y) { return {x.a + y.a, x.v0 + y.v0, x.v1 + y.v1, x.v2 + y.v2}; }
__device__ __forceinline__ J3 operator-(J3 x, J3 y) { return {x.a - y.a, x.v0 - y.v0, x.v1 - y.v1, x.v2 - y.v2}; }
__device__ __forceinline__ J3 operator*(J3 x, J3 y) {
  return {x.a * y.a, x.a * y.v0 + x.v0 * y.a, x.a * y.v1 + x.v1 * y.a, x.a * y.v2 + x.v2 * y.a};
}
__device__ __forceinline__ J3 operator*(J3 x, double s) { return {x.a * s, x.v0 * s, x.v1 * s, x.v2 * s}; }

// rotation.h:563-620 with forward-mode duals on the angle-axis parameters
__device__ void rotate_point_jet(const double aa[3], const double pt[3], J3 out[3]) {
  const J3 w0 = {aa[0], 1, 0, 0}, w1 = {aa[1], 0, 1, 0}, w2 = {aa[2], 0, 0, 1};
  const J3 theta2 = w0 * w0 + w1 * w1 + w2 * w2;
  if (theta2.a > DBL_EPSILON) {
    const double ith = rsqrt(theta2.a), th = theta2.a * ith;   // one rsqrt instead of a square root and two divisions
    const double dth = 0.5 * ith;
    const J3 theta = {th, theta2.v0 * dth, theta2.v1 * dth, theta2.v2 * dth};
    double s, c;
    sincos(th, &s, &c);
    const J3 ct = {c, -s * theta.v0, -s * theta.v1, -s * theta.v2};
    const J3 st = {s, c * theta.v0, c * theta.v1, c * theta.v2};
    const double dith = -ith * ith;
    const J3 ti = {ith, dith * theta.v0, dith * theta.v1, dith * theta.v2};
    const J3 n0 = w0 * ti, n1 = w1 * ti, n2 = w2 * ti;
    const J3 x0 = n1 * pt[2] - n2 * pt[1], x1 = n2 * pt[0] - n0 * pt[2], x2 = n0 * pt[1] - n1 * pt[0];
    const J3 tmp = (n0 * pt[0] + n1 * pt[1] + n2 * pt[2]) * (jc(1.0) - ct);
    out[0] = ct * pt[0] + x0 * st + n0 * tmp;
    out[1] = ct * pt[1] + x1 * st + n1 * tmp;
    out[2] = ct * pt[2] + x2 * st + n2 * tmp;
  } else {
    out[0] = jc(pt[0]) + (w1 * pt[2] - w2 * pt[1]);
    out[1] = jc(pt[1]) + (w2 * pt[0] - w0 * pt[2]);
    out[2] = jc(pt[2]) + (w0 * pt[1] - w1 * pt[0]);
  }
}

// uncertainty_pnp.cpp:16-34: this lane's two residuals and (optionally) their 2x6 jacobian
__device__ void residual_lane(const double x[6], const double pt[3], double u, double v, double wxx,
                              double wxy, double wyy, const Cam& cam, bool active, double r[2],
                              double (*J)[6]) {
  if (!active) {
    r[0] = r[1] = 0.0;
    if (J) for (int k = 0; k < 6; ++k) { J[0][k] = 0.0; J[1][k] = 0.0; }
    return;
  }
  J3 p[3];
  rotate_point_jet(x, pt, p);
  const double X = p[0].a + x[3], Y = p[1].a + x[4], Z = p[2].a + x[5];
  const double iz = 1.0 / Z;
  const double dx = cam.fu * X * iz + cam.uc - u, dy = cam.fv * Y * iz + cam.vc - v;
  r[0] = wxx * dx + wxy * dy;
  r[1] = wxy * dx + wyy * dy;
  if (J) {
    double jx[6], jy[6];
    const double fxz = cam.fu * iz, fyz = cam.fv * iz, xz = X * iz, yz = Y * iz;
    jx[0] = fxz * (p[0].v0 - xz * p[2].v0); jx[1] = fxz * (p[0].v1 - xz * p[2].v1); jx[2] = fxz * (p[0].v2 - xz * p[2].v2);
    jy[0] = fyz * (p[1].v0 - yz * p[2].v0); jy[1] = fyz * (p[1].v1 - yz * p[2].v1); jy[2] = fyz * (p[1].v2 - yz * p[2].v2);
    jx[3] = fxz; jx[4] = 0.0; jx[5] = -fxz * xz;
    jy[3] = 0.0; jy[4] = fyz; jy[5] = -fyz * yz;
#pragma unroll
    for (int k = 0; k < 6; ++k) { J[0][k] = wxx * jx[k] + wxy * jy[k]; J[1][k] = wxy * jx[k] + wyy * jy[k]; }
  }
}

__device__ bool ldlt_solve6(const double (&A)[6][6], const double (&b)[6], double (&x)[6]) {
  double L[6][6], D[6], rD[6];
  bool ok = true;
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    if (d == 0.0 || d != d) ok = false;
    const double rd = 1.0 / d;          // one reciprocal per pivot (a division is ~20 instructions and 80 cycles)
    rD[j] = rd;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * D[k];
      L[i][j] = s * rd;
    }
  }
  double y[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) { double s = b[i];
#pragma unroll
    for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s; }
#pragma unroll
  for (int i = 0; i < 6; ++i) y[i] *= rD[i];
#pragma unroll
  for (int i = 5; i >= 0; --i) { double s = y[i];
#pragma unroll
    for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s; }
  return ok;
}

struct LmState { double scale[6]; double jtj[6][6]; double g[6]; double cost, gmax; };

// One evaluation of the cost at x: residuals and Jacobian of every point, and the 28 sums over the points that
// everything else is made of.  `red` = (LM_ROWS * TSTRIDE + LM_SUMS) doubles of this warp's shared memory.
// Rows 0..5 / 7..12 of the table: d r0 / d x_k and d r1 / d x_k of the lane's point, rows 6 / 13: -r0, -r1; the pair
// sums (a <= b < 7) of  row_a . row_b + row_{7+a} . row_{7+b}  are J^T J (21), the gradient J^T (-r) (6) and
// |r|^2 (1), unscaled; e[a * 7 - a (a - 1) / 2 + (b - a)] = sum (a, b).
__device__ __forceinline__ void lm_eval(const double x[6], const double pt[3], double u, double v, double wxx,
                                        double wxy, double wyy, const Cam& cam, bool active,
                                        double (&e)[LM_SUMS], double* red, int lane) {
  {
    double r[2], J[2][6];
    residual_lane(x, pt, u, v, wxx, wxy, wyy, cam, active, r, J);
#pragma unroll
    for (int k = 0; k < 6; ++k) { red[k * TSTRIDE + lane] = J[0][k]; red[(7 + k) * TSTRIDE + lane] = J[1][k]; }
    red[6 * TSTRIDE + lane] = -r[0];
    red[13 * TSTRIDE + lane] = -r[1];
  }
  POSE_PHASE(26);
  const int np = 32 - __clz(__ballot_sync(FULL, active));
  __syncwarp();
  if (lane < LM_SUMS) {
    int a = 0, rem = lane;
    while (rem >= 7 - a) { rem -= 7 - a; ++a; }
    const double* xa = red + a * TSTRIDE;
    const double* xb = red + (a + rem) * TSTRIDE;
    red[LM_ROWS * TSTRIDE + lane] =
        sum_points(np, [&](int p) { return fma(xa[p], xb[p], xa[7 * TSTRIDE + p] * xb[7 * TSTRIDE + p]); });
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < LM_SUMS; ++i) e[i] = red[LM_ROWS * TSTRIDE + i];
  __syncwarp();                      // the table is rewritten by the next evaluation
  POSE_PHASE(27);
}

// tiny_solver.h:166-195 (Update) from the sums of lm_eval.  The Jacobi column scaling is applied to the sums,
// (J^T J)_ab s_a s_b, instead of to J first: same numbers up to rounding, and the unscaled diagonal that defines
// the scale on the first call comes out of the same pass.
__device__ __forceinline__ void lm_accept(const double (&e)[LM_SUMS], bool first, LmState& S) {
  auto at = [&](int a, int b) -> double { return e[a * 7 - a * (a - 1) / 2 + (b - a)]; };
  if (first) {
#pragma unroll
    for (int k = 0; k < 6; ++k) S.scale[k] = 1.0 / (1.0 + sqrt(at(k, k)));
  }
  S.gmax = 0.0;
#pragma unroll
  for (int a = 0; a < 6; ++a) {
#pragma unroll
    for (int b = a; b < 6; ++b) {
      const double sab = at(a, b) * (S.scale[a] * S.scale[b]);
      S.jtj[a][b] = sab; S.jtj[b][a] = sab;
    }
    const double gg = at(a, 6) * S.scale[a];
    S.g[a] = gg; S.gmax = fmax(S.gmax, fabs(gg));
  }
  S.cost = 0.5 * at(6, 6);
}

// tiny_solver.h:197-293 (Solve).  Returns TinySolver's status (0 gradient, 1 step, 2 cost, 3 max-iter).
// TinySolver evaluates the residuals at the trial point and, when the step is accepted, evaluates residuals AND
// Jacobian at the same point again (Update).  Here the trial point gets one full evaluation whose sums serve both
// the acceptance test (|r|^2 is one of them) and the next iteration, and that evaluation exists once in the
// machine code (it is most of the solver).  The sequence of tests and the iteration count are TinySolver's.
__device__ int lm_solve(double x[6], const double pt[3], double u, double v, double wxx, double wxy,
                        double wyy, const Cam& cam, bool active, int* iterations, double* final_cost,
                        double* red, int lane) {
  LmState S;
  int status = 3, it = 0;
  double uu = 1.0 / 1e4, vv = 2.0, model = 1.0;
  double xe[6];
#pragma unroll
  for (int i = 0; i < 6; ++i) xe[i] = x[i];
  bool first = true;
#pragma unroll 1
  for (;;) {
    double e[LM_SUMS];
    lm_eval(xe, pt, u, v, wxx, wxy, wyy, cam, active, e, red, lane);
    const double rho = first ? 1.0 : (2 * S.cost - e[LM_SUMS - 1]) / model;
    POSE_PHASE(15);
    if (rho > 0) {
#pragma unroll
      for (int i = 0; i < 6; ++i) x[i] = xe[i];
      lm_accept(e, first, S);
      POSE_PHASE(13);
      if (S.gmax < 1e-10) { status = 0; break; }
      if (S.cost < DBL_EPSILON) { status = 2; break; }
      if (!first) {
        const double tmp = 2 * rho - 1;
        uu = uu * fmax(1 / 3., 1 - tmp * tmp * tmp);
        vv = 2;
      }
      first = false;
    } else {
      uu *= vv; vv *= 2;
    }
    // next trial point from the accepted state; a failed factorisation counts as a rejected step
    bool have_trial = false;
#pragma unroll 1
    while (!have_trial && ++it < 50) {
      double A[6][6], step[6], dx[6];
#pragma unroll
      for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) A[i][j] = S.jtj[i][j];
        A[i][i] += uu * fmin(fmax(S.jtj[i][i], 1e-6), 1e32);     // (sqrt(u d))^2 in TinySolver
      }
      const bool ok = ldlt_solve6(A, S.g, step);
      double nx = 0, ndx = 0;
#pragma unroll
      for (int i = 0; i < 6; ++i) { dx[i] = S.scale[i] * step[i]; nx += x[i] * x[i]; ndx += dx[i] * dx[i]; }
      if (ok && sqrt(ndx) < 1e-8 * (sqrt(nx) + 1e-8)) { status = 1; break; }
      if (ok) {
#pragma unroll
        for (int i = 0; i < 6; ++i) xe[i] = x[i] + dx[i];
        model = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          double s = 2 * S.g[a];
#pragma unroll
          for (int b = 0; b < 6; ++b) s -= S.jtj[a][b] * step[b];
          model += step[a] * s;
        }
        have_trial = true;
      } else {
        uu *= vv; vv *= 2;
      }
    }
    POSE_PHASE(14);
    if (!have_trial) break;   // a converged step (status 1) or the iteration limit (status 3)
  }
  if (iterations) *iterations = it;
  if (final_cost) *final_cost = S.cost;
  return status;
}

// ------------------------------------------------------------------------------ kernels
__device__ __forceinline__ Cam load_cam(const double* K, int batched, int img) {
  const double* k = K + (batched ? (size_t)img * 9 : 0);
  return {k[0], k[4], k[2], k[5]};
}
__device__ __forceinline__ double round_f32(double x) { return (double)(float)x; }

template <int W, int C>
__global__ void __cluster_dims__(C, 1, 1) __launch_bounds__(32 * W)
pnp_kernel(const double* __restrict__ p3d, int p3d_batched, const double* __restrict__ p2d,
           const double* __restrict__ K, int K_batched, const int32_t* __restrict__ npts, int B,
           int n_max, double reproj_err, int max_iters, double confidence, double* __restrict__ rt34,
           unsigned long long* __restrict__ inlier_mask, int32_t* __restrict__ status) {
  __shared__ WarpScratch scratch[W];
  __shared__ RansacShared s_ransac;
  __shared__ __align__(16) double s_a2[W][12 * 14];              // the eigen-solver's second copy of M^T M
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp 0 carries the result; all evaluate candidates
  const int img = blockIdx.x / C;                                // C CTAs (one cluster) per image
  const bool lead = warp == 0 && (C == 1 || blockIdx.x % C == 0);
  if (img >= B) return;
  WarpScratch& ws = scratch[warp];
  const int n = npts ? min(npts[img], n_max) : n_max;
  double pw[3] = {0, 0, 0}, u = 0, v = 0;
  if (lane < n) {
    const double* q3 = p3d + ((p3d_batched ? (size_t)img * n_max : 0) + lane) * 3;
    const double* q2 = p2d + ((size_t)img * n_max + lane) * 2;
    // OpenCV's solvePnPRansac converts its inputs to CV_32F
    pw[0] = round_f32(q3[0]); pw[1] = round_f32(q3[1]); pw[2] = round_f32(q3[2]);
    u = round_f32(q2[0]); v = round_f32(q2[1]);
  }
  const Cam cam = load_cam(K, K_batched, img);
  PoseRT out;
  unsigned mask = 0;
  const int st = pnp_ransac_epnp<W, C>(ws, s_a2[warp], lane, n, pw, u, v, cam, reproj_err, max_iters, confidence, out, mask, warp,
                                       &s_ransac);
  if (lead && lane == 0) {
    double* o = rt34 + (size_t)img * 12;
    if (st == EPB_POSE_OK) {
#pragma unroll
      for (int i = 0; i < 3; ++i) { o[4 * i] = out.R[3 * i]; o[4 * i + 1] = out.R[3 * i + 1]; o[4 * i + 2] = out.R[3 * i + 2]; o[4 * i + 3] = out.t[i]; }
    } else {
      for (int i = 0; i < 12; ++i) o[i] = NAN;
    }
    if (inlier_mask) inlier_mask[img] = mask;
    if (status) status[img] = st;
  }
}

// (three CTAs per SM: 168 registers, no spills worth mentioning -- 110 M poses/s on the LM sweep against 99 M with the
// 232 registers the compiler takes unasked, and 86 M when forced down to 128)
__global__ void __launch_bounds__(POSE_WARPS * 32, 3)
lm_kernel(const double* __restrict__ p2d, const double* __restrict__ p3d, int p3d_batched,
          const double* __restrict__ w2d, const double* __restrict__ K, int K_batched,
          const double* __restrict__ init_rt, const int32_t* __restrict__ npts, int B, int n_max,
          double* __restrict__ result_rt, int32_t* __restrict__ iters, double* __restrict__ final_cost) {
  __shared__ double s_red[POSE_WARPS][LM_ROWS * TSTRIDE + LM_SUMS];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int img = blockIdx.x * (blockDim.x >> 5) + warp;
  if (img >= B) return;
  const int n = npts ? min(npts[img], n_max) : n_max;
  const bool active = lane < n;
  double pt[3] = {0, 0, 0}, u = 0, v = 0, wxx = 0, wxy = 0, wyy = 0;
  if (active) {
    const double* q3 = p3d + ((p3d_batched ? (size_t)img * n_max : 0) + lane) * 3;
    const double* q2 = p2d + ((size_t)img * n_max + lane) * 2;
    const double* qw = w2d + ((size_t)img * n_max + lane) * 3;
    pt[0] = q3[0]; pt[1] = q3[1]; pt[2] = q3[2]; u = q2[0]; v = q2[1];
    wxx = qw[0]; wxy = qw[1]; wyy = qw[2];
  }
  const Cam cam = load_cam(K, K_batched, img);
  double x[6];
#pragma unroll
  for (int k = 0; k < 6; ++k) x[k] = init_rt[(size_t)img * 6 + k];
  int it = 0; double fc = 0;
  lm_solve(x, pt, u, v, wxx, wxy, wyy, cam, active, &it, &fc, s_red[warp], lane);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < 6; ++k) result_rt[(size_t)img * 6 + k] = x[k];
    if (iters) iters[img] = it;
    if (final_cost) final_cost[img] = fc;
  }
}

// ------------------------------------------------------------------------------ P3P (4 points)
// lib/utils/extend_utils/extend_utils.py:85-95: cv2.solvePnP(flags=SOLVEPNP_P3P) on the four best-weighted
// correspondences -- the LM initialiser of uncertainty_pnp, and its whole answer when pn == 4.  OpenCV's P3P
// source is third-party and absent from the reference tree; its contract is: the pose that maps points 0..2
// onto their image rays exactly and reprojects point 3 best.  Restated from the published problem (Grunert's
// quartic in the depth ratio v = s3/s1, Haralick et al. 1994, for the candidates; every candidate is then
// polished by Newton's method on the three inter-point distance equations, which is what makes the result
// agree with cv2 to 1e-10 px on the three points -- the quartic alone loses 1e-2 px with a 3000 px focal length).
struct Cplx { double re, im; };
__device__ __forceinline__ Cplx cmul(Cplx a, Cplx b) { return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
__device__ __forceinline__ Cplx csub(Cplx a, Cplx b) { return {a.re - b.re, a.im - b.im}; }
__device__ __forceinline__ Cplx cdiv(Cplx a, Cplx b) {
  const double d = b.re * b.re + b.im * b.im;
  return {(a.re * b.re + a.im * b.im) / d, (a.im * b.re - a.re * b.im) / d};
}
// all four roots of the monic quartic z^4 + c3 z^3 + c2 z^2 + c1 z + c0 (Durand-Kerner)
__device__ void quartic_roots(double c3, double c2, double c1, double c0, Cplx z[4]) {
  const double rad = 1.0 + fmax(fmax(fabs(c3), fabs(c2)), fmax(fabs(c1), fabs(c0)));   // Cauchy bound
  Cplx w = {1.0, 0.0};
  const Cplx seed = {0.4, 0.9};
  for (int i = 0; i < 4; ++i) { z[i] = {w.re * rad * 0.5, w.im * rad * 0.5}; w = cmul(w, seed); }
  for (int it = 0; it < 200; ++it) {
    double move = 0.0, size = 0.0;
    for (int i = 0; i < 4; ++i) {
      Cplx pz = {1.0, 0.0};
      pz = cmul(pz, z[i]); pz.re += c3;
      pz = cmul(pz, z[i]); pz.re += c2;
      pz = cmul(pz, z[i]); pz.re += c1;
      pz = cmul(pz, z[i]); pz.re += c0;
      Cplx den = {1.0, 0.0};
      for (int j = 0; j < 4; ++j) if (j != i) den = cmul(den, csub(z[i], z[j]));
      const Cplx d = cdiv(pz, den);
      if (isfinite(d.re) && isfinite(d.im)) { z[i] = csub(z[i], d); move = fmax(move, fabs(d.re) + fabs(d.im)); }
      size = fmax(size, fabs(z[i].re) + fabs(z[i].im));
    }
    if (move <= 1e-15 * fmax(size, 1.0)) break;
  }
}
__device__ __forceinline__ void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
// orthonormal frame of three points: e1 along X1 - X0, e3 normal to the triangle; columns of F (row-major 3x3)
__device__ bool tri_frame(const double X[3][3], double F[9]) {
  double e1[3], d2[3], e3[3], e2[3];
  for (int k = 0; k < 3; ++k) { e1[k] = X[1][k] - X[0][k]; d2[k] = X[2][k] - X[0][k]; }
  const double n1 = sqrt(e1[0] * e1[0] + e1[1] * e1[1] + e1[2] * e1[2]);
  cross3(e1, d2, e3);
  const double n3 = sqrt(e3[0] * e3[0] + e3[1] * e3[1] + e3[2] * e3[2]);
  if (!(n1 > 0.0) || !(n3 > 0.0)) return false;
  for (int k = 0; k < 3; ++k) { e1[k] /= n1; e3[k] /= n3; }
  cross3(e3, e1, e2);
  for (int k = 0; k < 3; ++k) { F[3 * k] = e1[k]; F[3 * k + 1] = e2[k]; F[3 * k + 2] = e3[k]; }
  return true;
}
// P [4][3] world, uv [4][2] pixels.  Returns false when no candidate has positive depths.
__device__ bool p3p_solve(const double P[4][3], const double uv[4][2], const Cam& cam, double R[9], double t[3]) {
  double f[3][3];
  for (int i = 0; i < 3; ++i) {
    const double x = (uv[i][0] - cam.uc) / cam.fu, y = (uv[i][1] - cam.vc) / cam.fv;
    const double n = sqrt(x * x + y * y + 1.0);
    f[i][0] = x / n; f[i][1] = y / n; f[i][2] = 1.0 / n;
  }
  auto d2 = [&](int i, int j) {
    const double a = P[i][0] - P[j][0], b = P[i][1] - P[j][1], c = P[i][2] - P[j][2];
    return a * a + b * b + c * c;
  };
  auto dotf = [&](int i, int j) { return f[i][0] * f[j][0] + f[i][1] * f[j][1] + f[i][2] * f[j][2]; };
  const double a2 = d2(1, 2), b2 = d2(0, 2), c2 = d2(0, 1);
  const double ca = dotf(1, 2), cb = dotf(0, 2), cg = dotf(0, 1);
  if (!(b2 > 0.0)) return false;
  const double q = (a2 - c2) / b2, r = (a2 + c2) / b2;
  const double A4 = (q - 1) * (q - 1) - 4 * c2 / b2 * ca * ca;
  const double A3 = 4 * (q * (1 - q) * cb - (1 - r) * ca * cg + 2 * c2 / b2 * ca * ca * cb);
  const double A2 = 2 * (q * q - 1 + 2 * q * q * cb * cb + 2 * (b2 - c2) / b2 * ca * ca - 4 * r * ca * cb * cg +
                         2 * (b2 - a2) / b2 * cg * cg);
  const double A1 = 4 * (-q * (1 + q) * cb + 2 * a2 / b2 * cg * cg * cb - (1 - r) * ca * cg);
  const double A0 = (1 + q) * (1 + q) - 4 * a2 / b2 * cg * cg;
  Cplx z[4];
  quartic_roots(A3 / A4, A2 / A4, A1 / A4, A0 / A4, z);
  double FP[9];
  if (!tri_frame(P, FP)) return false;
  double best = INFINITY;
  bool found = false;
  for (int k = 0; k < 4; ++k) {
    const double v = z[k].re;
    if (!isfinite(v) || !(fabs(z[k].im) <= 1e-4 * fmax(1.0, fabs(v)))) continue;
    const double den = 2 * (cg - v * ca), s1sq = b2 / (1 + v * v - 2 * v * cb);
    if (!(s1sq > 0.0) || den == 0.0) continue;
    const double u = ((q - 1) * v * v - 2 * q * cb * v + 1 + q) / den;
    double s[3] = {sqrt(s1sq), 0, 0};
    s[1] = u * s[0]; s[2] = v * s[0];
    // Newton on  s2^2 + s3^2 - 2 s2 s3 cos(alpha) = a^2,  s1^2 + s3^2 - 2 s1 s3 cos(beta) = b^2,
    //            s1^2 + s2^2 - 2 s1 s2 cos(gamma) = c^2
    bool ok = true;
    for (int it = 0; it < 8 && ok; ++it) {
      const double g0 = s[1] * s[1] + s[2] * s[2] - 2 * s[1] * s[2] * ca - a2;
      const double g1 = s[0] * s[0] + s[2] * s[2] - 2 * s[0] * s[2] * cb - b2;
      const double g2 = s[0] * s[0] + s[1] * s[1] - 2 * s[0] * s[1] * cg - c2;
      const double J[3][3] = {{0, 2 * s[1] - 2 * s[2] * ca, 2 * s[2] - 2 * s[1] * ca},
                              {2 * s[0] - 2 * s[2] * cb, 0, 2 * s[2] - 2 * s[0] * cb},
                              {2 * s[0] - 2 * s[1] * cg, 2 * s[1] - 2 * s[0] * cg, 0}};
      const double det = J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
      if (!(fabs(det) > 0.0) || !isfinite(det)) { ok = false; break; }
      const double dx0 = (g0 * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (g1 * J[2][2] - J[1][2] * g2) +
                          J[0][2] * (g1 * J[2][1] - J[1][1] * g2)) / det;
      const double dx1 = (J[0][0] * (g1 * J[2][2] - J[1][2] * g2) - g0 * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
                          J[0][2] * (J[1][0] * g2 - g1 * J[2][0])) / det;
      const double dx2 = (J[0][0] * (J[1][1] * g2 - g1 * J[2][1]) - J[0][1] * (J[1][0] * g2 - g1 * J[2][0]) +
                          g0 * (J[1][0] * J[2][1] - J[1][1] * J[2][0])) / det;
      s[0] -= dx0; s[1] -= dx1; s[2] -= dx2;
    }
    if (!ok || !(s[0] > 0.0) || !(s[1] > 0.0) || !(s[2] > 0.0)) continue;
    double Q[3][3], FQ[9], Rc[9], tc[3];
    for (int i = 0; i < 3; ++i) for (int d = 0; d < 3; ++d) Q[i][d] = s[i] * f[i][d];
    if (!tri_frame(Q, FQ)) continue;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) Rc[3 * i + j] = FQ[3 * i] * FP[3 * j] + FQ[3 * i + 1] * FP[3 * j + 1] + FQ[3 * i + 2] * FP[3 * j + 2];
    for (int i = 0; i < 3; ++i) tc[i] = Q[0][i] - (Rc[3 * i] * P[0][0] + Rc[3 * i + 1] * P[0][1] + Rc[3 * i + 2] * P[0][2]);
    const double xc = Rc[0] * P[3][0] + Rc[1] * P[3][1] + Rc[2] * P[3][2] + tc[0];
    const double yc = Rc[3] * P[3][0] + Rc[4] * P[3][1] + Rc[5] * P[3][2] + tc[1];
    const double zc = Rc[6] * P[3][0] + Rc[7] * P[3][1] + Rc[8] * P[3][2] + tc[2];
    const double du = cam.uc + cam.fu * xc / zc - uv[3][0], dv = cam.vc + cam.fv * yc / zc - uv[3][1];
    const double e = du * du + dv * dv;
    if (e < best) {
      best = e; found = true;
      for (int i = 0; i < 9; ++i) R[i] = Rc[i];
      for (int i = 0; i < 3; ++i) t[i] = tc[i];
    }
  }
  return found;
}

// One thread per problem.  w2d == nullptr: n must be 4 and the points are used in the given order.  Otherwise
// the four correspondences with the largest wxx + wxy are taken in ascending order of that key
// (extend_utils.py:83: np.argsort(weights_2d[:,0] + weights_2d[:,1])[-4:]; ties: lower index first).
__global__ void p3p_kernel(const double* __restrict__ p3d, int p3d_batched, const double* __restrict__ p2d,
                           const double* __restrict__ w2d, const double* __restrict__ K, int K_batched, int B, int n,
                           double* __restrict__ rt34, int32_t* __restrict__ status) {
  const int img = blockIdx.x * blockDim.x + threadIdx.x;
  if (img >= B) return;
  int sel[4] = {0, 1, 2, 3};
  if (w2d) {
    // selection sort of the four largest keys, then ascending order (stable argsort semantics)
    unsigned taken = 0;
    for (int r = 3; r >= 0; --r) {
      int bi = -1; double bk = 0;
      for (int i = 0; i < n; ++i) {
        if ((taken >> i) & 1u) continue;
        const double key = w2d[((size_t)img * n + i) * 3] + w2d[((size_t)img * n + i) * 3 + 1];
        if (bi < 0 || key > bk || (key == bk && i > bi)) { bi = i; bk = key; }
      }
      sel[r] = bi; taken |= 1u << bi;
    }
  }
  double P[4][3], uv[4][2];
  for (int k = 0; k < 4; ++k) {
    const double* q3 = p3d + ((p3d_batched ? (size_t)img * n : 0) + sel[k]) * 3;
    const double* q2 = p2d + ((size_t)img * n + sel[k]) * 2;
    P[k][0] = q3[0]; P[k][1] = q3[1]; P[k][2] = q3[2]; uv[k][0] = q2[0]; uv[k][1] = q2[1];
  }
  const Cam cam = load_cam(K, K_batched, img);
  double R[9], t[3];
  const bool ok = p3p_solve(P, uv, cam, R, t);
  double* o = rt34 + (size_t)img * 12;
  for (int i = 0; i < 3; ++i) {
    o[4 * i] = ok ? R[3 * i] : NAN; o[4 * i + 1] = ok ? R[3 * i + 1] : NAN; o[4 * i + 2] = ok ? R[3 * i + 2] : NAN;
    o[4 * i + 3] = ok ? t[i] : NAN;
  }
  if (status) status[img] = ok ? EPB_POSE_OK : EPB_POSE_FAILED;
}

__global__ void pose_pack_kernel(const double* __restrict__ rt6, int B, float* __restrict__ pose7,
                                 double* __restrict__ rt34) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  double r[3] = {rt6[6 * i], rt6[6 * i + 1], rt6[6 * i + 2]}, R[9], q[4];
  rodrigues_to_matrix(r, R);
  if (pose7) {
    matrix_to_quat_wxyz(R, q);
    float* o = pose7 + 7 * (size_t)i;
    o[0] = (float)q[0]; o[1] = (float)q[1]; o[2] = (float)q[2]; o[3] = (float)q[3];
    o[4] = (float)rt6[6 * i + 3]; o[5] = (float)rt6[6 * i + 4]; o[6] = (float)rt6[6 * i + 5];
  }
  if (rt34) {
    double* o = rt34 + 12 * (size_t)i;
    for (int a = 0; a < 3; ++a) { o[4 * a] = R[3 * a]; o[4 * a + 1] = R[3 * a + 1]; o[4 * a + 2] = R[3 * a + 2]; o[4 * a + 3] = rt6[6 * i + 3 + a]; }
  }
}

__global__ void rt34_to_rt6_kernel(const double* __restrict__ rt34, int B, double* __restrict__ rt6) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const double* m = rt34 + 12 * (size_t)i;
  const double R[9] = {m[0], m[1], m[2], m[4], m[5], m[6], m[8], m[9], m[10]};
  double r[3];
  matrix_to_rodrigues(R, r);
  double* o = rt6 + 6 * (size_t)i;
  o[0] = r[0]; o[1] = r[1]; o[2] = r[2]; o[3] = m[3]; o[4] = m[7]; o[5] = m[11];
}

// lib/utils/evaluation_utils.py:170-181: per keypoint W = inv(sqrtm(cov)) -> (wxx, wxy, wyy); zero
// weights when cov[0][0] < 1e-6 or any entry is NaN (the reference's guard) -- and, by definition here,
// when cov is not positive definite (scipy's sqrtm turns complex there and the reference breaks).
// 2x2 SPD closed form: sqrtm(A) = (A + s I) / t with s = sqrt(det A), t = sqrt(tr A + 2 s).
__global__ void cov_to_weights_kernel(const float* __restrict__ cov, int n, int mode, double* __restrict__ w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* c = cov + 4 * (size_t)i;
  const double a = c[0], b01 = c[1], b10 = c[2], d = c[3];
  double* o = w + 3 * (size_t)i;
  o[0] = o[1] = o[2] = 0.0;
  if (mode == EPB_WEIGHTS_INV_MAX_EIG) {
    // extend_utils.py:133-141 (uncertainty_pnp_v2): weight = 1 / largest eigenvalue, 0 when cov[0][0] < 1e-5
    if (c[0] < 1e-5f) return;
    const double bb = 0.5 * (b01 + b10);
    const double lmax = 0.5 * (a + d) + sqrt(0.25 * (a - d) * (a - d) + bb * bb);
    o[0] = o[2] = 1.0 / lmax;
    return;
  }
  if (c[0] < 1e-6f || isnan(c[0]) || isnan(c[1]) || isnan(c[2]) || isnan(c[3])) return;
  const double b = 0.5 * (b01 + b10);
  const double det = a * d - b * b;
  if (!(det > 0.0) || !(d > 0.0)) return;
  const double s = sqrt(det), t = sqrt(a + d + 2.0 * s);
  // S = (A + s I) / t;  det S = s;  inv(S) = [[S11, -S01], [-S01, S00]] / s
  const double s00 = (a + s) / t, s01 = b / t, s11 = (d + s) / t;
  o[0] = s11 / s; o[1] = -s01 / s; o[2] = s00 / s;
}

// val.py:172-228 for a batch: one warp per frame, lane k <-> keypoint k (K <= 32)
template <int W, int C>
__global__ void __cluster_dims__(C, 1, 1) __launch_bounds__(32 * W)
pose_pipeline_kernel(const float* __restrict__ preds, const float* __restrict__ maxvals,
                     const double* __restrict__ bbox_xy, const double* __restrict__ rate,
                     const double* __restrict__ p3d_model, const double* __restrict__ Kmat, int B, int K,
                     int min_k, double sel_thresh, int weighted, float* __restrict__ pose7,
                     double* __restrict__ rt6_out, double* __restrict__ epnp_rt34,
                     int32_t* __restrict__ status) {
  __shared__ WarpScratch scratch[W];
  __shared__ RansacShared s_ransac;
  // x3d,y3d,z3d,u,v,maxval in rank order (one copy per warp: no barrier needed); once the lanes hold their point
  // the 192 doubles serve as the eigen-solver's second copy of M^T M
  __shared__ __align__(16) double s_pts[W][32][6];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;   // warp 0: RANSAC bookkeeping + LM; all: candidates
  const int img = blockIdx.x / C;                                // C CTAs (one cluster) per image
  const bool lead = warp == 0 && (C == 1 || blockIdx.x % C == 0);
  if (img >= B) return;
  POSE_PHASE_START();
  WarpScratch& ws = scratch[warp];
  // --- keypoint selection (val.py:172-177): large_k = max(#(maxval > 0.8), 24), top large_k by maxval
  const bool have = lane < K;
  const float mv = have ? maxvals[(size_t)img * K + lane] : -INFINITY;
  int large_k = __popc(__ballot_sync(FULL, have && (double)mv > sel_thresh));
  large_k = min(max(large_k, min_k), K);
  // heapq.nlargest order: by maxval, ties keep the lower index.  NaN compares false with everything, which would
  // give every NaN lane rank 0: it ranks as -inf instead (a NaN keypoint that is still selected makes every
  // candidate model non-finite, i.e. a clean EPB_POSE_FAILED with a NaN pose)
  const float mk = (mv == mv) ? mv : -INFINITY;
  int rank = 0;
  for (int j = 0; j < K; ++j) {
    const float mj = __shfl_sync(FULL, mk, j);
    rank += (mj > mk) || (mj == mk && j < lane);
  }
  // --- un-crop (val.py:180): float32 pred * (1/rate) + (x, y), in float64
  if (have) {
    const double inv_rate = 1.0 / rate[img];
    double* d = s_pts[warp][rank];
    d[0] = p3d_model[3 * lane]; d[1] = p3d_model[3 * lane + 1]; d[2] = p3d_model[3 * lane + 2];
    d[3] = (double)preds[((size_t)img * K + lane) * 2] * inv_rate + bbox_xy[2 * img];
    d[4] = (double)preds[((size_t)img * K + lane) * 2 + 1] * inv_rate + bbox_xy[2 * img + 1];
    d[5] = (double)mv;
  }
  __syncwarp();
  const int n = large_k;
  const bool active = lane < n;
  double pt[3] = {0, 0, 0}, u = 0, v = 0, w = 0;
  if (active) {
    const double* d = s_pts[warp][lane];
    pt[0] = d[0]; pt[1] = d[1]; pt[2] = d[2]; u = d[3]; v = d[4]; w = d[5];
  }
  const Cam cam = {Kmat[0], Kmat[4], Kmat[2], Kmat[5]};
  // --- pnp(): EPnP-RANSAC on float32-rounded correspondences
  const double pwf[3] = {round_f32(pt[0]), round_f32(pt[1]), round_f32(pt[2])};
  PoseRT init;
  unsigned mask = 0;
  POSE_PHASE(0);
  const int st = pnp_ransac_epnp<W, C>(ws, &s_pts[warp][0][0], lane, n, pwf, round_f32(u), round_f32(v), cam, 5.0, 100, 0.99, init, mask,
                                       warp, &s_ransac);
  if (!lead) return;
  POSE_PHASE(11);
  double x[6];
  if (st == EPB_POSE_OK) {
    matrix_to_rodrigues(init.R, x);
    x[3] = init.t[0]; x[4] = init.t[1]; x[5] = init.t[2];
    if (epnp_rt34 && lane == 0) {
      double* o = epnp_rt34 + 12 * (size_t)img;
      for (int i = 0; i < 3; ++i) { o[4 * i] = init.R[3 * i]; o[4 * i + 1] = init.R[3 * i + 1]; o[4 * i + 2] = init.R[3 * i + 2]; o[4 * i + 3] = init.t[i]; }
    }
    // --- cpnp_m: LM with maxval weights on the unrounded float64 correspondences
    // cpnp_m's weighting is not recoverable (binary and source absent): maxval itself (1, the documented
    // assumption), sqrt(maxval) (2) or unit weights (0 = cpnp)
    const double ww = weighted == 1 ? w : (weighted == 2 ? sqrt(fmax(w, 0.0)) : 1.0);
    POSE_PHASE(12);
    lm_solve(x, pt, u, v, ww, 0.0, ww, cam, active, nullptr, nullptr, ws.flat(), lane);
    POSE_PHASE(16);
  } else {
    for (int k = 0; k < 6; ++k) x[k] = NAN;
    if (epnp_rt34 && lane == 0) for (int i = 0; i < 12; ++i) epnp_rt34[12 * (size_t)img + i] = NAN;
  }
  if (lane == 0) {
    if (rt6_out) for (int k = 0; k < 6; ++k) rt6_out[6 * (size_t)img + k] = x[k];
    if (pose7) {
      double R[9], q[4];
      rodrigues_to_matrix(x, R);
      matrix_to_quat_wxyz(R, q);
      float* o = pose7 + 7 * (size_t)img;
      o[0] = (float)q[0]; o[1] = (float)q[1]; o[2] = (float)q[2]; o[3] = (float)q[3];
      o[4] = (float)x[3]; o[5] = (float)x[4]; o[6] = (float)x[5];
    }
    if (status) status[img] = st;
  }
  POSE_PHASE(17);
}

// demo.py:295-310
__global__ void esa_score_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int B,
                                 double* __restrict__ score_t, double* __restrict__ score_r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float* p = pred + 7 * (size_t)i;
  const float* g = gt + 7 * (size_t)i;
  double dt = 0, nt = 0, dq = 0;
  for (int k = 0; k < 3; ++k) {
    const double d = (double)p[4 + k] - (double)g[4 + k];
    dt += d * d; nt += (double)g[4 + k] * (double)g[4 + k];
  }
  // pred_qua is float32 and target float (np.matmul in the dtype numpy promotes to)
  for (int k = 0; k < 4; ++k) dq += (double)p[k] * (double)g[k];
  dq = fabs(dq);
  if (score_t) score_t[i] = sqrt(dt) / sqrt(nt);
  if (score_r) score_r[i] = dq > 1.0 ? 0.0 : 2.0 * acos(dq);  // Re(arccos(x + 0j)) = 0 for x > 1
}

}  // namespace epb

using namespace epb;

// one warp per image; small batches get one warp per CTA so that they spread over more SMs
static inline void pose_launch_shape(int B, int* grid, int* block) {
  const int warps = B <= 148 * 8 ? 1 : POSE_WARPS;
  *block = warps * 32;
  *grid = (B + warps - 1) / warps;
}

// RANSAC candidates evaluated side by side per image: the kernel is one long serial FP64 chain per warp (255
// registers), so extra warps are free while SMs are idle and cost throughput once the batch fills the machine
static inline int ransac_warps(int B) {
  const int sms = device_sm_count();
  return B <= sms ? 8 : (B <= 2 * sms ? 4 : 2);
}
// ... and while two SMs per image are to be had, the eight warps are a cluster of two CTAs of four (one warp per
// SM sub-partition: see pnp_ransac_epnp)
static inline bool ransac_cluster(int B) { return 2 * B <= device_sm_count(); }

static void pose_kernel_attributes() {
#ifdef EPB_TUNING
  { const int r0 = tuning_int("EPB_RANSAC_R0", 4); cudaMemcpyToSymbol(c_ransac_round0_warps, &r0, sizeof(int)); }
#endif
  prefer_max_shared(pnp_kernel<2, 1>); prefer_max_shared(pnp_kernel<4, 1>); prefer_max_shared(pnp_kernel<8, 1>);
  prefer_max_shared(pnp_kernel<4, 2>); prefer_max_shared(pose_pipeline_kernel<4, 2>);
  prefer_max_shared(lm_kernel); prefer_max_shared(pose_pack_kernel);
  prefer_max_shared(rt34_to_rt6_kernel); prefer_max_shared(cov_to_weights_kernel);
  prefer_max_shared(pose_pipeline_kernel<2, 1>); prefer_max_shared(pose_pipeline_kernel<4, 1>);
  prefer_max_shared(pose_pipeline_kernel<8, 1>);
  prefer_max_shared(esa_score_kernel);
  cudaGetLastError();
}

extern "C" int epb_pnp_epnp_ransac(const double* p3d, int p3d_batched, const double* p2d, const double* K,
                                   int K_batched, const int32_t* npts, int B, int n_max, double reproj_err,
                                   int max_iters, double confidence, double* rt34,
                                   unsigned long long* inlier_mask, int32_t* status, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!p3d || !p2d || !K || !rt34 || B < 0 || n_max <= 0 || n_max > 32 || max_iters < 0) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  // W warps per image evaluate RANSAC candidates side by side (pnp_ransac_epnp): 8 while every image still gets
  // an SM of its own, fewer when the batch has to share them
  const int W = ransac_warps(B);
#define EPB_PNP_LAUNCH(WW, CC) pnp_kernel<WW, CC><<<B * CC, 32 * WW, 0, (cudaStream_t)stream>>>(             \
      p3d, p3d_batched, p2d, K, K_batched, npts, B, n_max, reproj_err, max_iters, confidence, rt34, inlier_mask, status)
  if (ransac_cluster(B)) EPB_PNP_LAUNCH(4, 2);
  else if (W == 8) EPB_PNP_LAUNCH(8, 1); else if (W == 4) EPB_PNP_LAUNCH(4, 1); else EPB_PNP_LAUNCH(2, 1);
#undef EPB_PNP_LAUNCH
  return check_launch();
}

extern "C" int epb_lm_refine(const double* p2d, const double* p3d, int p3d_batched, const double* w2d,
                             const double* K, int K_batched, const double* init_rt, const int32_t* npts,
                             int B, int n_max, double* result_rt, int32_t* iters, double* final_cost,
                             void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!p2d || !p3d || !w2d || !K || !init_rt || !result_rt || B < 0 || n_max <= 0 || n_max > 32)
    return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  int grid, block; pose_launch_shape(B, &grid, &block);
  lm_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(
      p2d, p3d, p3d_batched, w2d, K, K_batched, init_rt, npts, B, n_max, result_rt, iters, final_cost);
  return check_launch();
}

extern "C" int epb_pose_pack(const double* rt6, int B, float* pose7, double* rt34, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!rt6 || B < 0 || (!pose7 && !rt34)) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  pose_pack_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rt6, B, pose7, rt34);
  return check_launch();
}

extern "C" int epb_rt34_to_rt6(const double* rt34, int B, double* rt6, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!rt34 || !rt6 || B < 0) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  rt34_to_rt6_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(rt34, B, rt6);
  return check_launch();
}

extern "C" int epb_p3p(const double* p3d, int p3d_batched, const double* p2d, const double* w2d, const double* K,
                       int K_batched, int B, int n, double* rt34, int32_t* status, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!p3d || !p2d || !K || !rt34 || B < 0 || n < 4 || n > 32 || (!w2d && n != 4)) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  p3p_kernel<<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(p3d, p3d_batched, p2d, w2d, K, K_batched, B, n, rt34, status);
  return check_launch();
}

extern "C" int epb_cov_to_weights(const float* cov, int n, int mode, double* w2d, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!cov || !w2d || n < 0 || mode < EPB_WEIGHTS_INV_SQRTM || mode > EPB_WEIGHTS_INV_MAX_EIG) return EPB_ERR_INVALID;
  if (n == 0) return EPB_OK;
  cov_to_weights_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(cov, n, mode, w2d);
  return check_launch();
}

extern "C" int epb_pose_pipeline(const float* preds, const float* maxvals, const double* bbox_xy,
                                 const double* rate, const double* p3d_model, const double* Kmat, int B,
                                 int K, int min_k, double sel_thresh, int weighted, float* pose7, double* rt6,
                                 double* epnp_rt34, int32_t* status, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!preds || !maxvals || !bbox_xy || !rate || !p3d_model || !Kmat || B < 0 || K <= 0 || K > 32)
    return EPB_ERR_INVALID;
  if (!pose7 && !rt6) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  ProfScope ps(PROF_POSE, (cudaStream_t)stream);
  const int W = ransac_warps(B);
#define EPB_POSE_LAUNCH(WW, CC) pose_pipeline_kernel<WW, CC><<<B * CC, 32 * WW, 0, (cudaStream_t)stream>>>(       \
      preds, maxvals, bbox_xy, rate, p3d_model, Kmat, B, K, min_k, sel_thresh, weighted, pose7, rt6, epnp_rt34, status)
  if (ransac_cluster(B)) EPB_POSE_LAUNCH(4, 2);
  else if (W == 8) EPB_POSE_LAUNCH(8, 1); else if (W == 4) EPB_POSE_LAUNCH(4, 1); else EPB_POSE_LAUNCH(2, 1);
#undef EPB_POSE_LAUNCH
  return check_launch();
}

extern "C" int epb_esa_score(const float* pose7_pred, const float* pose7_gt, int B, double* score_t,
                             double* score_r, void* stream) {
  EPB_INIT_ONCE_PER_DEVICE(pose_kernel_attributes);
  if (!pose7_pred || !pose7_gt || B < 0) return EPB_ERR_INVALID;
  if (B == 0) return EPB_OK;
  esa_score_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(pose7_pred, pose7_gt, B, score_t, score_r);
  return check_launch();
}

#ifdef EPB_TUNING
// tuning build only: read and clear the phase clocks of CTA 0 (see POSE_PHASE)
extern "C" int epb_debug_pose_phase_clocks(long long* out32) {
  cudaDeviceSynchronize();
  if (cudaMemcpyFromSymbol(out32, epb::g_phase_clk, 32 * sizeof(long long)) != cudaSuccess) return 1;
  long long zero[32] = {0};
  return cudaMemcpyToSymbol(epb::g_phase_clk, zero, sizeof(zero)) != cudaSuccess;
}
#endif
