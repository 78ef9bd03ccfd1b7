"""TEST INFRASTRUCTURE ONLY: numpy statement of the published EPnP algorithm.

The reference's ``pnp.pnp`` (pnp.py:68-84) calls ``cv2.solvePnPRansac(flags=SOLVEPNP_EPNP)``;
EPnP itself lives in OpenCV (third-party; not under /root/reference; version unpinned by the
reference, 4.13.0 in this image).  This file restates the published algorithm
(Lepetit, Moreno-Noguer, Fua, "EPnP: An Accurate O(n) Solution to the PnP Problem", IJCV 2009,
as implemented in OpenCV's calib3d epnp): control points from PCA, barycentric alphas,
M^T M null-space, three beta initialisations + 5 Gauss-Newton steps, Horn/Arun alignment,
best-of-three by reprojection error.  It uses the same numerical building blocks as the
CUDA kernel (cyclic Jacobi eigen-solver, Householder least squares, one-sided Jacobi SVD) so
that csrc/pose.cu is a transliteration of this file.  tests/test_oracle_pose.py pins it
against cv2.solvePnP(EPNP) and against vectors produced by the reference's own pnp.py.
"""
import numpy as np


def jacobi_eigh(a, sweeps=30):
    """Cyclic Jacobi for a symmetric matrix.  -> (w ascending, v columns)."""
    a = np.array(a, np.float64)
    n = a.shape[0]
    v = np.eye(n)
    for _ in range(sweeps):
        off = 0.0
        for p in range(n - 1):
            for q in range(p + 1, n):
                off += a[p, q] * a[p, q]
        if off < 1e-300:
            break
        for p in range(n - 1):
            for q in range(p + 1, n):
                apq = a[p, q]
                if apq == 0.0:
                    continue
                theta = (a[q, q] - a[p, p]) / (2.0 * apq)
                t = (1.0 if theta >= 0 else -1.0) / (abs(theta) + np.sqrt(theta * theta + 1.0))
                c = 1.0 / np.sqrt(t * t + 1.0)
                s = t * c
                app, aqq = a[p, p], a[q, q]
                a[p, p] = app - t * apq
                a[q, q] = aqq + t * apq
                a[p, q] = a[q, p] = 0.0
                for k in range(n):
                    if k != p and k != q:
                        akp, akq = a[k, p], a[k, q]
                        a[k, p] = a[p, k] = c * akp - s * akq
                        a[k, q] = a[q, k] = s * akp + c * akq
                    vkp, vkq = v[k, p], v[k, q]
                    v[k, p] = c * vkp - s * vkq
                    v[k, q] = s * vkp + c * vkq
    w = np.diag(a).copy()
    order = np.argsort(w, kind="stable")
    return w[order], v[:, order]


def householder_lstsq(a, b):
    """Least squares min ||a x - b|| by Householder QR (a: m x n, m >= n)."""
    a = np.array(a, np.float64)
    b = np.array(b, np.float64)
    m, n = a.shape
    for k in range(n):
        x = a[k:, k]
        alpha = np.linalg.norm(x)
        if alpha == 0.0:
            continue
        if x[0] > 0:
            alpha = -alpha
        vk = x.copy()
        vk[0] -= alpha
        vn2 = vk @ vk
        if vn2 == 0.0:
            continue
        for j in range(k, n):
            a[k:, j] -= 2.0 * vk * (vk @ a[k:, j]) / vn2
        b[k:] -= 2.0 * vk * (vk @ b[k:]) / vn2
    x = np.zeros(n)
    for i in range(n - 1, -1, -1):
        x[i] = (b[i] - a[i, i + 1:n] @ x[i + 1:]) / a[i, i]
    return x


def svd_onesided_cv(a, sweeps=30):
    """One-sided (Hestenes) Jacobi SVD with the rotation formulas, pair order, V = I start and
    descending selection sort of OpenCV's cv::SVD (JacobiSVDImpl_).  Rows of ``at`` are the
    columns of ``a``.  -> (w descending, ut rows = left singular vectors, vt rows = right).
    The SIGN convention matters for EPnP: the control points are c0 + sqrt(w/n) * ut[i], and a
    flipped axis gives a (slightly) different answer on noisy data, so the PCA of
    choose_control_points must reproduce OpenCV's signs (checked against cv2.SVDecomp)."""
    at = np.array(a, np.float64).T.copy()
    n = at.shape[0]
    vt = np.eye(n)
    eps = np.finfo(np.float64).eps * 10
    w = np.array([at[i] @ at[i] for i in range(n)])
    for _ in range(sweeps):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                aa, p, bb = w[i], at[i] @ at[j], w[j]
                if abs(p) <= eps * np.sqrt(aa * bb):
                    continue
                p *= 2.0
                beta = aa - bb
                gamma = np.hypot(p, beta)
                if beta < 0:
                    delta = (gamma - beta) * 0.5
                    s = np.sqrt(delta / gamma)
                    c = p / (gamma * s * 2.0)
                else:
                    c = np.sqrt((gamma + beta) / (gamma * 2.0))
                    s = p / (gamma * c * 2.0)
                t0 = c * at[i] + s * at[j]
                t1 = -s * at[i] + c * at[j]
                at[i], at[j] = t0, t1
                w[i], w[j] = t0 @ t0, t1 @ t1
                changed = True
                t0 = c * vt[i] + s * vt[j]
                t1 = -s * vt[i] + c * vt[j]
                vt[i], vt[j] = t0, t1
        if not changed:
            break
    w = np.array([np.sqrt(at[i] @ at[i]) for i in range(n)])
    for i in range(n - 1):
        j = i + int(np.argmax(w[i:]))
        if j != i:
            w[[i, j]] = w[[j, i]]
            at[[i, j]] = at[[j, i]]
            vt[[i, j]] = vt[[j, i]]
    ut = np.zeros_like(at)
    for i in range(n):
        if w[i] > 0:
            ut[i] = at[i] / w[i]
    return w, ut, vt


def _dist2(p, q):
    d = p - q
    return d @ d


def epnp(pws, us, fu, fv, uc, vc):
    """EPnP on n >= 4 correspondences.  pws [n,3], us [n,2] (pixels).  -> (R [3,3], t [3], err)."""
    pws = np.asarray(pws, np.float64)
    us = np.asarray(us, np.float64)
    n = pws.shape[0]
    # --- control points
    cws = np.zeros((4, 3))
    cws[0] = pws.mean(0)
    pw0 = pws - cws[0]
    w, uct, _ = svd_onesided_cv(pw0.T @ pw0)    # OpenCV sign convention (see docstring)
    for i in range(1, 4):
        cws[i] = cws[0] + np.sqrt(w[i - 1] / n) * uct[i - 1]
    # --- barycentric coordinates
    cc = (cws[1:] - cws[0]).T                  # cc[i][j-1] = cws[j][i]-cws[0][i]
    cc_inv = np.linalg.inv(cc)
    alphas = np.zeros((n, 4))
    alphas[:, 1:] = (pws - cws[0]) @ cc_inv.T
    alphas[:, 0] = 1.0 - alphas[:, 1:].sum(1)
    # --- M^T M
    m = np.zeros((2 * n, 12))
    for j in range(4):
        m[0::2, 3 * j] = alphas[:, j] * fu
        m[0::2, 3 * j + 2] = alphas[:, j] * (uc - us[:, 0])
        m[1::2, 3 * j + 1] = alphas[:, j] * fv
        m[1::2, 3 * j + 2] = alphas[:, j] * (vc - us[:, 1])
    _, ev = jacobi_eigh(m.T @ m)               # ascending: ev[:,0] smallest
    vs = [ev[:, i] for i in range(4)]          # v[0] = null vector, ... (OpenCV: ut rows 11..8)
    # --- L_6x10 and rho
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        for j, (a, b) in enumerate(pairs):
            dv[i, j] = vs[i][3 * a:3 * a + 3] - vs[i][3 * b:3 * b + 3]
    l = np.zeros((6, 10))
    for i in range(6):
        d0, d1, d2, d3 = dv[0, i], dv[1, i], dv[2, i], dv[3, i]
        l[i] = [d0 @ d0, 2 * d0 @ d1, d1 @ d1, 2 * d0 @ d2, 2 * d1 @ d2, d2 @ d2,
                2 * d0 @ d3, 2 * d1 @ d3, 2 * d2 @ d3, d3 @ d3]
    rho = np.array([_dist2(cws[a], cws[b]) for a, b in pairs])

    def betas_approx_1():
        b4 = householder_lstsq(l[:, [0, 1, 3, 6]], rho)
        be = np.zeros(4)
        if b4[0] < 0:
            be[0] = np.sqrt(-b4[0]); be[1:] = -b4[1:] / be[0]
        else:
            be[0] = np.sqrt(b4[0]); be[1:] = b4[1:] / be[0]
        return be

    def betas_approx_2():
        b3 = householder_lstsq(l[:, [0, 1, 2]], rho)
        be = np.zeros(4)
        if b3[0] < 0:
            be[0] = np.sqrt(-b3[0]); be[1] = np.sqrt(-b3[2]) if b3[2] < 0 else 0.0
        else:
            be[0] = np.sqrt(b3[0]); be[1] = np.sqrt(b3[2]) if b3[2] > 0 else 0.0
        if b3[1] < 0:
            be[0] = -be[0]
        return be

    def betas_approx_3():
        b5 = householder_lstsq(l[:, [0, 1, 2, 3, 4]], rho)
        be = np.zeros(4)
        if b5[0] < 0:
            be[0] = np.sqrt(-b5[0]); be[1] = np.sqrt(-b5[2]) if b5[2] < 0 else 0.0
        else:
            be[0] = np.sqrt(b5[0]); be[1] = np.sqrt(b5[2]) if b5[2] > 0 else 0.0
        if b5[1] < 0:
            be[0] = -be[0]
        be[2] = b5[3] / be[0]
        return be

    def gauss_newton(be):
        be = be.copy()
        for _ in range(5):
            a = np.zeros((6, 4)); b = np.zeros(6)
            for i in range(6):
                r = l[i]
                a[i, 0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3]
                a[i, 1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3]
                a[i, 2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3]
                a[i, 3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3]
                b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1]
                                 + r[3] * be[0] * be[2] + r[4] * be[1] * be[2] + r[5] * be[2] * be[2]
                                 + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] + r[8] * be[2] * be[3]
                                 + r[9] * be[3] * be[3])
            be += householder_lstsq(a, b)
        return be

    def compute_r_and_t(be):
        ccs = np.zeros((4, 3))
        for i in range(4):
            ccs += be[i] * vs[i].reshape(4, 3)
        pcs = alphas @ ccs
        if pcs[0, 2] < 0.0:
            ccs, pcs = -ccs, -pcs
        pc0, pw0c = pcs.mean(0), pws.mean(0)
        abt = (pcs - pc0).T @ (pws - pw0c)
        _, ut, vt = svd_onesided_cv(abt)
        r = ut.T @ vt                              # R = U V^T
        if np.linalg.det(r) < 0:
            r[2] = -r[2]
        t = pc0 - r @ pw0c
        xc = pws @ r.T + t
        ue = uc + fu * xc[:, 0] / xc[:, 2]
        ve = vc + fv * xc[:, 1] / xc[:, 2]
        err = np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2).sum() / n
        return r, t, err

    best = None
    for init in (betas_approx_1, betas_approx_2, betas_approx_3):
        cand = compute_r_and_t(gauss_newton(init()))
        if best is None or cand[2] < best[2]:
            best = cand
    return best


def reprojection_errors(pws, us, r, t, fu, fv, uc, vc):
    xc = np.asarray(pws) @ r.T + t
    ue = uc + fu * xc[:, 0] / xc[:, 2]
    ve = vc + fv * xc[:, 1] / xc[:, 2]
    return np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2)


class CvRng:
    """OpenCV's cv::RNG (multiply-with-carry), used with seed (uint64)-1 by
    RANSACPointSetRegistrator::run; uniform(a,b) = next() % (b-a) + a."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * 4164903690 + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def ransac_update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num, denom = np.log(num), np.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))


def solve_pnp_ransac_epnp(p3d, p2d, fu, fv, uc, vc, reproj_err=5.0, iters=100, confidence=0.99):
    """Restates cv2.solvePnPRansac(flags=SOLVEPNP_EPNP) as called by pnp.py:68-73: inputs rounded to
    float32, minimal samples of 5 drawn with cv::RNG(-1), EPnP per sample, inliers by squared
    reprojection error <= 25, adaptive iteration count, final EPnP on the consensus set.
    The 5-point EPnP has a 2-dimensional numerical null space whose basis is decided by rounding
    noise inside OpenCV's SVD, so per-sample models can differ from OpenCV's within the noise;
    the consensus set (and hence the result) agrees whenever outliers are gross.
    -> (R, t, inlier mask) or None."""
    p3 = np.asarray(p3d, np.float32).astype(np.float64)
    p2 = np.asarray(p2d, np.float32).astype(np.float64)
    n = p3.shape[0]
    model_points = 5
    if n < model_points:
        return None
    if n == model_points:
        r, t, _ = epnp(p3, p2, fu, fv, uc, vc)
        return r, t, np.ones(n, bool)
    rng = CvRng()
    niters = iters
    best_mask, best_count = None, 0
    thr2 = reproj_err * reproj_err
    it = 0
    while it < niters:
        idx = []
        for i in range(model_points):
            k = rng.uniform(0, n)
            while k in idx:
                k = rng.uniform(0, n)
            idx.append(k)
        r, t, _ = epnp(p3[idx], p2[idx], fu, fv, uc, vc)
        if np.all(np.isfinite(r)) and np.all(np.isfinite(t)):
            err = reprojection_errors(p3, p2, r, t, fu, fv, uc, vc) ** 2
            mask = err.astype(np.float32) <= np.float32(thr2)
            good = int(mask.sum())
            if good > max(best_count, model_points - 1):
                best_mask, best_count = mask, good
                niters = ransac_update_num_iters(confidence, (n - good) / n, model_points, niters)
        it += 1
    if best_mask is None:
        return None
    r, t, _ = epnp(p3[best_mask], p2[best_mask], fu, fv, uc, vc)
    return r, t, best_mask
