"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's LINEMOD-style pose metrics
(/root/reference/evaluation.py:340-411, twins in lib/utils/evaluation_utils.py:75-141).

  project_K             lib/utils/base_utils.py:293-297  (pts @ R^T + t) @ K^T, divide by z
  nearest_idx           lib/utils/extend_utils/extend_utils.py:40-61 -> nearest_neighborhood.cu:48-121
                        (float32 brute force, first minimum) via oracle/metrics_oracle.c
  projection_2d[_sym]   evaluation.py:340-354   mean 2-D reprojection distance (sym: to the NEAREST projected point)
  add_metric[_sym]      evaluation.py:356-397   mean 3-D distance of the transformed model points (ADD / ADD-S)
  cm_degree_5_metric    evaluation.py:399-411   translation error in cm, rotation error in degrees
Each returns the value the reference appends to its recorders (mean distance and the pass flag).
Pinned by tests/golden/metrics_ref.npz = the reference's own Evaluator methods run by
tests/golden/make_golden_metrics.py."""
import ctypes

import numpy as np

from . import _lib


def project_K(pts_3d, RT, K):
    pts_2d = np.matmul(pts_3d, RT[:, :3].T) + RT[:, 3:].T
    pts_2d = np.matmul(pts_2d, K.T)
    return pts_2d[:, :2] / pts_2d[:, 2:]


def nearest_idx(ref_pts, que_pts):
    """for every point of que_pts the index of the nearest point of ref_pts (float32 search, first minimum)."""
    assert ref_pts.shape[1] == que_pts.shape[1] and 1 < que_pts.shape[1] <= 3
    r = np.ascontiguousarray(ref_pts, np.float32)
    q = np.ascontiguousarray(que_pts, np.float32)
    idx = np.zeros(q.shape[0], np.int32)
    _lib.oracle_lib().orc_nearest_idx(_lib.fptr(r), _lib.fptr(q), _lib.iptr(idx), ctypes.c_int(r.shape[0]),
                                      ctypes.c_int(q.shape[0]), ctypes.c_int(r.shape[1]), ctypes.c_int(0))
    return idx


def nearest_distance(pts1, pts2):
    """evaluation.py:162-170: distances from every point of pts2 to its nearest point of pts1 (float64 norm)."""
    idxs = nearest_idx(pts1, pts2)
    return np.linalg.norm(pts1[idxs] - pts2, 2, 1)


def projection_2d(pose_pred, pose_targets, model, K, threshold=5, sym=False):
    a, b = project_K(model, pose_pred, K), project_K(model, pose_targets, K)
    d = np.mean(nearest_distance(a, b)) if sym else np.mean(np.linalg.norm(a - b, axis=-1))
    return d, bool(d < threshold)


def add_metric(pose_pred, pose_targets, model, diameter, percentage=0.1, sym=False):
    a = np.dot(model, pose_pred[:, :3].T) + pose_pred[:, 3]
    b = np.dot(model, pose_targets[:, :3].T) + pose_targets[:, 3]
    d = np.mean(nearest_distance(a, b)) if sym else np.mean(np.linalg.norm(a - b, axis=-1))
    return d, bool(d < diameter * percentage)


def cm_degree_5_metric(pose_pred, pose_targets):
    cm = np.linalg.norm(pose_pred[:, 3] - pose_targets[:, 3]) * 100
    trace = np.trace(np.dot(pose_pred[:, :3], pose_targets[:, :3].T))
    trace = trace if trace <= 3 else 3
    with np.errstate(invalid="ignore"):
        deg = np.rad2deg(np.arccos((trace - 1.) / 2.))
    return cm, deg, bool(cm < 5 and deg < 5)
