/* TEST INFRASTRUCTURE ONLY (oracle/).  Never imported by the product path.
 *
 * Plain-C restatement of the reference's pose refinement:
 *   cost      lib/utils/extend_utils/src/uncertainty_pnp.cpp:16-34
 *             (ReprojectionErrorArray: r = W (pi(R(w) X + t) - x), W = [[wxx,wxy],[wxy,wyy]])
 *   rotation  lib/utils/extend_utils/include/ceres/rotation.h:563-620 (AngleAxisRotatePoint,
 *             first-order branch when theta^2 <= DBL_EPSILON)
 *   solver    lib/utils/extend_utils/include/ceres/tiny_solver.h:150-293 (the only complete
 *             LM loop shipped in the reference tree): Jacobi column scaling
 *             1/(1+||J_j||) fixed at iteration 0, u = 1/1e4, diagonal clamp [1e-6,1e32],
 *             accept iff rho > 0, u *= max(1/3, 1-(2rho-1)^3), reject -> u *= v, v *= 2,
 *             stops: max|g| < 1e-10, ||dx|| < 1e-8 (||x|| + 1e-8), cost < DBL_EPSILON, 50 its.
 *   entry     uncertainty_pnp(pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn), :61-92
 *             (K read as K[0],K[4],K[2],K[5], :80)
 * Derivatives are forward-mode duals (what Ceres' Jets compute).  The 6x6 solve is an
 * unpivoted LDL^T (Eigen::LDLT pivots; identical for SPD systems up to rounding).
 * PARITY PINNING: checked against oracle/_ref/libuncertainty_pnp_ref.so (the
 * reference file compiled over the vendored TinySolver, oracle/ceres_shim) in
 * tests/test_oracle_pose.py.  The cpnp binary/source is absent from the reference,
 * so parity with cpnp itself is UNPINNED.
 */
#include <float.h>
#include <math.h>
#include <string.h>

#define NP 6

typedef struct { double a; double v[NP]; } jet;

static jet j_const(double a) { jet r; r.a = a; memset(r.v, 0, sizeof r.v); return r; }
static jet j_var(double a, int k) { jet r = j_const(a); r.v[k] = 1.0; return r; }
static jet j_add(jet x, jet y) { jet r; r.a = x.a + y.a; for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] + y.v[i]; return r; }
static jet j_sub(jet x, jet y) { jet r; r.a = x.a - y.a; for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] - y.v[i]; return r; }
static jet j_mul(jet x, jet y) { jet r; r.a = x.a * y.a; for (int i = 0; i < NP; ++i) r.v[i] = x.a * y.v[i] + x.v[i] * y.a; return r; }
static jet j_div(jet x, jet y) {
  jet r; const double inv = 1.0 / y.a; r.a = x.a * inv; const double q = x.a * inv;
  for (int i = 0; i < NP; ++i) r.v[i] = (x.v[i] - q * y.v[i]) * inv; return r;
}
static jet j_scale(jet x, double s) { jet r; r.a = x.a * s; for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] * s; return r; }
static jet j_sqrt(jet x) { jet r; r.a = sqrt(x.a); const double d = 1.0 / (2.0 * r.a); for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] * d; return r; }
static jet j_cos(jet x) { jet r; r.a = cos(x.a); const double d = -sin(x.a); for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] * d; return r; }
static jet j_sin(jet x) { jet r; r.a = sin(x.a); const double d = cos(x.a); for (int i = 0; i < NP; ++i) r.v[i] = x.v[i] * d; return r; }

/* rotation.h:563-620 */
static void angle_axis_rotate_point(const jet aa[3], const double pt[3], jet out[3]) {
  const jet theta2 = j_add(j_add(j_mul(aa[0], aa[0]), j_mul(aa[1], aa[1])), j_mul(aa[2], aa[2]));
  if (theta2.a > DBL_EPSILON) {
    const jet theta = j_sqrt(theta2);
    const jet costheta = j_cos(theta), sintheta = j_sin(theta);
    const jet theta_inverse = j_div(j_const(1.0), theta);
    const jet w[3] = { j_mul(aa[0], theta_inverse), j_mul(aa[1], theta_inverse), j_mul(aa[2], theta_inverse) };
    const jet wxp[3] = {
      j_sub(j_scale(w[1], pt[2]), j_scale(w[2], pt[1])),
      j_sub(j_scale(w[2], pt[0]), j_scale(w[0], pt[2])),
      j_sub(j_scale(w[0], pt[1]), j_scale(w[1], pt[0])) };
    const jet wdotp = j_add(j_add(j_scale(w[0], pt[0]), j_scale(w[1], pt[1])), j_scale(w[2], pt[2]));
    const jet tmp = j_mul(wdotp, j_sub(j_const(1.0), costheta));
    for (int k = 0; k < 3; ++k)
      out[k] = j_add(j_add(j_scale(costheta, pt[k]), j_mul(wxp[k], sintheta)), j_mul(w[k], tmp));
  } else {
    const jet wxp[3] = {
      j_sub(j_scale(aa[1], pt[2]), j_scale(aa[2], pt[1])),
      j_sub(j_scale(aa[2], pt[0]), j_scale(aa[0], pt[2])),
      j_sub(j_scale(aa[0], pt[1]), j_scale(aa[1], pt[0])) };
    for (int k = 0; k < 3; ++k) out[k] = j_add(j_const(pt[k]), wxp[k]);
  }
}

typedef struct {
  const double *p2, *p3, *w, *K; int pn;
} problem;

/* uncertainty_pnp.cpp:16-34; res [2pn], jac row-major [2pn x 6] or NULL */
static void evaluate(const problem* P, const double* x, double* res, double* jac) {
  const double fx = P->K[0], fy = P->K[4], px = P->K[2], py = P->K[5];
  jet pose[6];
  for (int k = 0; k < 6; ++k) pose[k] = j_var(x[k], k);
  for (int i = 0; i < P->pn; ++i) {
    jet tp[3];
    angle_axis_rotate_point(pose, P->p3 + 3 * i, tp);
    tp[0] = j_add(tp[0], pose[3]); tp[1] = j_add(tp[1], pose[4]); tp[2] = j_add(tp[2], pose[5]);
    const jet proj_x = j_add(j_div(j_scale(tp[0], fx), tp[2]), j_const(px));
    const jet proj_y = j_add(j_div(j_scale(tp[1], fy), tp[2]), j_const(py));
    const jet dx = j_sub(proj_x, j_const(P->p2[2 * i])), dy = j_sub(proj_y, j_const(P->p2[2 * i + 1]));
    const double wxx = P->w[3 * i], wxy = P->w[3 * i + 1], wyy = P->w[3 * i + 2];
    const jet r0 = j_add(j_scale(dx, wxx), j_scale(dy, wxy));
    const jet r1 = j_add(j_scale(dx, wxy), j_scale(dy, wyy));
    res[2 * i] = r0.a; res[2 * i + 1] = r1.a;
    if (jac) for (int k = 0; k < 6; ++k) { jac[(2 * i) * 6 + k] = r0.v[k]; jac[(2 * i + 1) * 6 + k] = r1.v[k]; }
  }
}

static int ldlt_solve6(const double A[6][6], const double b[6], double x[6]) {
  double L[6][6], D[6];
  memset(L, 0, sizeof L);
  for (int j = 0; j < 6; ++j) {
    double d = A[j][j];
    for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    if (d == 0.0 || d != d) return 0;
    L[j][j] = 1.0;
    for (int i = j + 1; i < 6; ++i) {
      double s = A[i][j];
      for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k] * D[k];
      L[i][j] = s / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i][k] * y[k]; y[i] = s; }
  for (int i = 0; i < 6; ++i) y[i] /= D[i];
  for (int i = 5; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < 6; ++k) s -= L[k][i] * x[k]; x[i] = s; }
  return 1;
}

#define MAXPN 256

typedef struct {
  double err[2 * MAXPN], jac[2 * MAXPN * 6], scale[6], jtj[6][6], g[6], cost, gmax;
} lm_state;

/* tiny_solver.h:166-195 (Update) */
static void lm_update(const problem* P, const double* x, lm_state* S, int first) {
  const int nr = 2 * P->pn;
  evaluate(P, x, S->err, S->jac);
  for (int r = 0; r < nr; ++r) S->err[r] = -S->err[r];
  if (first) {
    for (int k = 0; k < 6; ++k) {
      double n2 = 0; for (int r = 0; r < nr; ++r) n2 += S->jac[r * 6 + k] * S->jac[r * 6 + k];
      S->scale[k] = 1.0 / (1.0 + sqrt(n2));
    }
  }
  for (int r = 0; r < nr; ++r) for (int k = 0; k < 6; ++k) S->jac[r * 6 + k] *= S->scale[k];
  S->gmax = 0; S->cost = 0;
  for (int a = 0; a < 6; ++a) {
    for (int b = 0; b < 6; ++b) {
      double s = 0; for (int r = 0; r < nr; ++r) s += S->jac[r * 6 + a] * S->jac[r * 6 + b];
      S->jtj[a][b] = s;
    }
    double s = 0; for (int r = 0; r < nr; ++r) s += S->jac[r * 6 + a] * S->err[r];
    S->g[a] = s; if (fabs(s) > S->gmax) S->gmax = fabs(s);
  }
  for (int r = 0; r < nr; ++r) S->cost += S->err[r] * S->err[r];
  S->cost *= 0.5;
}

/* tiny_solver.h:197-293 (Solve).  Returns the TinySolver status
 * (0 gradient, 1 step, 2 cost, 3 max-iterations), -1 on bad input. */
int orc_lm_refine(const double* pts2d, const double* pts3d, const double* wgt2d, const double* K,
                  const double* init_rt, double* result_rt, int pn, int* iterations,
                  double* final_cost) {
  if (pn <= 0 || pn > MAXPN) return -1;
  problem P = { pts2d, pts3d, wgt2d, K, pn };
  static __thread lm_state S;
  double x[6]; memcpy(x, init_rt, sizeof x);
  int status = 3, it = 0;
  lm_update(&P, x, &S, 1);
  if (S.gmax < 1e-10) { status = 0; goto done; }
  if (S.cost < DBL_EPSILON) { status = 2; goto done; }
  {
    double u = 1.0 / 1e4, v = 2.0;
    for (it = 1; it < 50; ++it) {
      double A[6][6], step[6], dx[6], xn[6], fnew[2 * MAXPN];
      memcpy(A, S.jtj, sizeof A);
      for (int i = 0; i < 6; ++i) {
        const double d = fmin(fmax(S.jtj[i][i], 1e-6), 1e32);
        const double lm = sqrt(u * d);
        A[i][i] += lm * lm;
      }
      const int ok = ldlt_solve6(A, S.g, step);
      double nx = 0, ndx = 0;
      for (int i = 0; i < 6; ++i) { dx[i] = S.scale[i] * step[i]; nx += x[i] * x[i]; ndx += dx[i] * dx[i]; }
      if (ok && sqrt(ndx) < 1e-8 * (sqrt(nx) + 1e-8)) { status = 1; break; }
      double rho = -1.0;
      if (ok) {
        for (int i = 0; i < 6; ++i) xn[i] = x[i] + dx[i];
        evaluate(&P, xn, fnew, 0);
        double f2 = 0; for (int r = 0; r < 2 * pn; ++r) f2 += fnew[r] * fnew[r];
        const double cost_change = 2 * S.cost - f2;
        double model = 0;
        for (int a = 0; a < 6; ++a) {
          double s = 2 * S.g[a]; for (int b = 0; b < 6; ++b) s -= S.jtj[a][b] * step[b];
          model += step[a] * s;
        }
        rho = cost_change / model;
      }
      if (rho > 0) {
        memcpy(x, xn, sizeof x);
        lm_update(&P, x, &S, 0);
        if (S.gmax < 1e-10) { status = 0; break; }
        if (S.cost < DBL_EPSILON) { status = 2; break; }
        const double tmp = 2 * rho - 1;
        u = u * fmax(1 / 3., 1 - tmp * tmp * tmp);
        v = 2;
        continue;
      }
      u *= v; v *= 2;
    }
  }
done:
  memcpy(result_rt, x, sizeof x);
  if (iterations) *iterations = it;
  if (final_cost) *final_cost = S.cost;
  return status;
}

/* Same argument order as the reference's C entry (uncertainty_pnp.cpp:61-69). */
void orc_uncertainty_pnp(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_rt,
                         double* result_rt, int pn) {
  orc_lm_refine(pts2d, pts3d, wgt2d, K, init_rt, result_rt, pn, 0, 0);
}
