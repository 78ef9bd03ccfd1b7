"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's per-image pose solve.

  pnp()            /root/reference/pnp.py:46-90  (cv2.solvePnPRansac, EPNP, 5 px) -- OpenCV is a
                   third-party dependency the reference does not pin (3.x era; 4.13.0 here).
  lm_refine()      lib/utils/extend_utils/src/uncertainty_pnp.cpp:61-92 via oracle/lm_oracle.c
                   (C restatement) or oracle/_ref (reference file over vendored TinySolver).
  cpnp / cpnp_m    val.py:200-202 call sites; the cpnp binary and source are ABSENT from the
                   reference.  ASSUMPTION: weights wxx = wyy = maxval, wxy = 0
                   for cpnp_m and 1,0,1 for cpnp.  Parity with cpnp itself is UNPINNED.
  uncertainty_pnp  lib/utils/extend_utils/extend_utils.py:64-115 (P3P on the 4 best weights).
  frame_pose()     val.py:168-228 glue: select keypoints, un-crop, pnp, Rodrigues, LM, quaternion.
  esa_score()      demo.py:295-310.
"""
import ctypes

import cv2
import numpy as np

from . import _lib, decode

# ESA/SPEED intrinsics, lib/utils/base_utils.py:250-252 / utils.py:28-39
ESA_K = np.array([[0.0176 / 5.86e-6, 0, 960.0], [0, 0.0176 / 5.86e-6, 600.0], [0, 0, 1.0]])


def pnp(points_3d, points_2d, camera_matrix, method=cv2.SOLVEPNP_ITERATIVE):
    """pnp.py:46-90.  The solver is always RANSAC-EPnP; `method` only adds a batch axis."""
    dist_coeffs = np.zeros(shape=[8, 1], dtype="float64")
    assert points_3d.shape[0] == points_2d.shape[0]
    if method == cv2.SOLVEPNP_EPNP:
        points_3d = np.expand_dims(points_3d, 0)
        points_2d = np.expand_dims(points_2d, 0)
    points_2d = np.ascontiguousarray(points_2d.astype(np.float64))
    points_3d = np.ascontiguousarray(points_3d.astype(np.float64))
    camera_matrix = camera_matrix.astype(np.float64)
    _, r_exp, t, _ = cv2.solvePnPRansac(points_3d, points_2d, camera_matrix, dist_coeffs,
                                        reprojectionError=5.0, flags=cv2.SOLVEPNP_EPNP)
    rot, _ = cv2.Rodrigues(r_exp)
    return np.concatenate([rot, t], axis=-1)


def lm_refine(points_2d, points_3d, weights_2d, camera_matrix, init_rt, use_ref=False):
    """uncertainty_pnp C entry.  weights_2d [pn,3] = (wxx, wxy, wyy).  -> rt[6] float64."""
    p2 = np.ascontiguousarray(points_2d, np.float64)
    p3 = np.ascontiguousarray(points_3d, np.float64)
    w = np.ascontiguousarray(weights_2d, np.float64)
    k = np.ascontiguousarray(np.asarray(camera_matrix, np.float64).reshape(-1)[:9])
    init = np.ascontiguousarray(init_rt, np.float64).reshape(6)
    out = np.empty(6, np.float64)
    pn = ctypes.c_int(p2.shape[0])
    if use_ref:
        _lib.ref_pnp_lib().uncertainty_pnp(_lib.dptr(p2), _lib.dptr(p3), _lib.dptr(w), _lib.dptr(k),
                                           _lib.dptr(init), _lib.dptr(out), pn)
    else:
        _lib.oracle_lib().orc_uncertainty_pnp(_lib.dptr(p2), _lib.dptr(p3), _lib.dptr(w),
                                              _lib.dptr(k), _lib.dptr(init), _lib.dptr(out), pn)
    return out


def cpnp(p3d, p2d, K, camera, use_ref=False):
    """cpnp.cpnp(p3d, p2d, K, camera[6]) -> camera[6] (val.py:200); unit weights (assumption)."""
    n = np.asarray(p3d).shape[0]
    w = np.stack([np.ones(n), np.zeros(n), np.ones(n)], 1)
    return lm_refine(p2d, p3d, w, np.asarray(K, np.float64).reshape(-1)[-9:], camera, use_ref)


def cpnp_m(p3d, p2d, maxvals, K, camera, use_ref=False):
    """cpnp.cpnp_m(p3d, p2d, maxvals, K, camera[6]) -> camera[6] (val.py:202).
    K may be the batched [1,3,3] tensor the reference passes; read as a flat buffer."""
    mv = np.asarray(maxvals, np.float64).reshape(-1)
    w = np.stack([mv, np.zeros_like(mv), mv], 1)
    return lm_refine(p2d, p3d, w, np.asarray(K, np.float64).reshape(-1)[-9:], camera, use_ref)


def uncertainty_pnp(points_2d, weights_2d, points_3d, camera_matrix, use_ref=False):
    """extend_utils.py:64-115."""
    pn = points_2d.shape[0]
    assert points_3d.shape[0] == pn and pn >= 4
    dist = np.zeros([8, 1], np.float64)
    points_3d = points_3d.astype(np.float64)
    points_2d = points_2d.astype(np.float64)
    weights_2d = weights_2d.astype(np.float64)
    camera_matrix = camera_matrix.astype(np.float64)
    idxs = np.argsort(weights_2d[:, 0] + weights_2d[:, 1])[-4:]
    _, r_exp, t = cv2.solvePnP(np.expand_dims(points_3d[idxs, :], 0),
                               np.expand_dims(points_2d[idxs, :], 0),
                               camera_matrix, dist, None, None, False, flags=cv2.SOLVEPNP_P3P)
    if pn == 4:
        rot, _ = cv2.Rodrigues(r_exp)
        return np.concatenate([rot, t], axis=-1)
    init_rt = np.concatenate([r_exp, t], 0).reshape(6)
    rt = lm_refine(points_2d, points_3d, weights_2d, camera_matrix, init_rt, use_ref)
    rot, _ = cv2.Rodrigues(rt[:3])
    return np.concatenate([rot, rt[3:, None]], axis=-1)


def quat_wxyz_from_matrix(rot):
    """val.py:221-224: scipy Rotation.from_matrix(R).as_quat() (x,y,z,w) -> (w,x,y,z) f32."""
    from scipy.spatial.transform import Rotation
    q = Rotation.from_matrix(rot).as_quat()
    return np.asarray([q[3], q[0], q[1], q[2]], dtype=np.float32)


def frame_pose(preds, maxvals, bbox_xy, rate, points_3d, camera_k, use_ref=False,
               min_k=24, thresh=0.8, weighted=True):
    """val.py:172-228 for one frame.  preds [K,2] f32 crop px, maxvals [K], bbox_xy (x,y),
    rate scalar.  -> dict(q wxyz f32, t f64, rt6, pose34, idxs)."""
    maxvals = [float(m) for m in maxvals]
    idxs = decode.select_keypoints(maxvals, min_k, thresh)
    ori = preds * (1 / np.asarray(rate)) + [bbox_xy[0], bbox_xy[1]]         # :180 (float64)
    p3d = np.asarray(points_3d)[idxs]
    p2d = ori[idxs]
    mav = np.asarray(maxvals)[idxs]
    pose_pred = pnp(p3d, p2d, np.asarray(camera_k), cv2.SOLVEPNP_EPNP)      # :194
    r_exp, _ = cv2.Rodrigues(pose_pred[:, :3])                              # :197
    camera = np.concatenate([r_exp.reshape(3), pose_pred[:, 3].reshape(3)])
    camera = cpnp_m(p3d, p2d, mav, camera_k, camera, use_ref) if weighted else \
        cpnp(p3d, p2d, camera_k, camera, use_ref)                           # :200-202
    rot, _ = cv2.Rodrigues(camera[:3])                                      # :207
    pose34 = np.concatenate((rot, camera[3:].reshape(3, 1)), axis=1)
    return dict(q=quat_wxyz_from_matrix(rot), t=camera[3:].copy(), rt6=camera, pose34=pose34,
                idxs=idxs, epnp34=pose_pred)


def cov_to_weights(covar):
    """lib/utils/evaluation_utils.py:170-181: per keypoint inv(sqrtm(cov)) -> [pn,3] (wxx,wxy,wyy);
    zeros when cov[0,0] < 1e-6 or NaN."""
    import scipy.linalg
    covar = np.asarray(covar)
    cov_invs = []
    for vi in range(covar.shape[0]):
        if covar[vi, 0, 0] < 1e-6 or np.sum(np.isnan(covar)[vi]) > 0:
            cov_invs.append(np.zeros([2, 2]).astype(np.float32))
            continue
        cov_invs.append(np.linalg.inv(scipy.linalg.sqrtm(covar[vi])))
    weights = np.asarray(cov_invs).reshape([-1, 4])
    return weights[:, (0, 1, 3)]


def isotropic_weights(covars):
    """extend_utils.py:133-141 (uncertainty_pnp_v2): 1 / max eigenvalue, 0 when cov[0,0] < 1e-5 -> [pn,3]."""
    covars = np.asarray(covars)
    w = []
    for pi in range(covars.shape[0]):
        w.append(0.0 if covars[pi, 0, 0] < 1e-5 else 1.0 / np.max(np.linalg.eigvals(covars[pi])))
    w = np.asarray(w, np.float64)[:, None]
    return np.concatenate([w, np.zeros_like(w), w], 1)


def esa_score(q_pred, t_pred, q_gt, t_gt):
    """demo.py:295-310: per-frame ||t^-t||/||t|| + 2 Re(arccos(|q^.q| + 0j))."""
    q_pred, q_gt = np.asarray(q_pred, np.float64), np.asarray(q_gt, np.float64)
    t_pred, t_gt = np.asarray(t_pred, np.float64), np.asarray(t_gt, np.float64)
    score_t = np.linalg.norm(t_pred - t_gt) / np.linalg.norm(t_gt)
    score_r = 2 * np.real(np.arccos(np.abs(np.dot(q_pred, q_gt)) + 0j))
    return score_t, score_r
