"""PnP initialisation and LM refinement -- drop-ins for the reference's ``pnp.py``, the ``cpnp``
extension and ``lib/utils/extend_utils`` ``uncertainty_pnp`` on the CUDA path.

  pnp(points_3d, points_2d, camera_matrix, method)      /root/reference/pnp.py:46-90
  cpnp(p3d, p2d, K, camera) / cpnp_m(p3d, p2d, maxvals, K, camera)   val.py:200-202
  uncertainty_pnp(points_2d, weights_2d, points_3d, camera_matrix)   extend_utils.py:64-115
and the batched, device-resident forms ``pnp_batch`` / ``lm_refine_batch`` / ``pose_pipeline``
that the B200 path is built for.  All arithmetic is csrc/pose.cu (one warp per image, FP64).
"""
import numpy as np
import torch

from . import _lib

# cv2 constants (cv2 is not imported on the product path)
SOLVEPNP_ITERATIVE = 0
SOLVEPNP_EPNP = 1


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("esa_pose_estimation_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def _f64(x, device):
    if isinstance(x, torch.Tensor):
        return x.detach().to(device=device, dtype=torch.float64).contiguous()
    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float64))).to(device)


# ------------------------------------------------------------------------------------ batched
def pnp_batch(p3d, p2d, K, npts=None, reproj_err=5.0, max_iters=100, confidence=0.99,
              return_status=False):
    """p3d [B,n,3] or [n,3] (shared model), p2d [B,n,2], K [B,3,3] or [3,3]; CUDA float64.
    -> rt34 [B,3,4] float64 ([R|t]); with return_status also (inlier_mask [B] int64, status [B])."""
    dev = p2d.device if isinstance(p2d, torch.Tensor) and p2d.is_cuda else _device()
    p3d, p2d, K = _f64(p3d, dev), _f64(p2d, dev), _f64(K, dev)
    b, n = p2d.shape[0], p2d.shape[1]
    rt = torch.empty((b, 3, 4), dtype=torch.float64, device=dev)
    mask = torch.zeros((b,), dtype=torch.int64, device=dev)
    status = torch.zeros((b,), dtype=torch.int32, device=dev)
    if npts is not None:
        npts = npts.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        st = _lib.load().epb_pnp_epnp_ransac(
            _lib.ptr(p3d), int(p3d.dim() == 3), _lib.ptr(p2d), _lib.ptr(K), int(K.dim() == 3),
            _lib.ptr(npts), b, n, float(reproj_err), int(max_iters), float(confidence), _lib.ptr(rt),
            _lib.ptr(mask), _lib.ptr(status), _lib.stream_ptr())
    _lib.check(st, "epb_pnp_epnp_ransac")
    return (rt, mask, status) if return_status else rt


def lm_refine_batch(p2d, p3d, w2d, K, init_rt, npts=None, return_info=False):
    """Batched uncertainty_pnp: p2d [B,n,2], p3d [B,n,3] | [n,3], w2d [B,n,3] (wxx,wxy,wyy),
    K [B,3,3] | [3,3], init_rt [B,6] -> rt [B,6] (angle-axis, t), CUDA float64."""
    dev = p2d.device if isinstance(p2d, torch.Tensor) and p2d.is_cuda else _device()
    p2d, p3d, w2d, K, init_rt = (_f64(a, dev) for a in (p2d, p3d, w2d, K, init_rt))
    b, n = p2d.shape[0], p2d.shape[1]
    out = torch.empty((b, 6), dtype=torch.float64, device=dev)
    iters = torch.zeros((b,), dtype=torch.int32, device=dev)
    cost = torch.zeros((b,), dtype=torch.float64, device=dev)
    if npts is not None:
        npts = npts.to(device=dev, dtype=torch.int32).contiguous()
    with torch.cuda.device(dev):
        st = _lib.load().epb_lm_refine(_lib.ptr(p2d), _lib.ptr(p3d), int(p3d.dim() == 3), _lib.ptr(w2d),
                                       _lib.ptr(K), int(K.dim() == 3), _lib.ptr(init_rt), _lib.ptr(npts), b, n,
                                       _lib.ptr(out), _lib.ptr(iters), _lib.ptr(cost), _lib.stream_ptr())
    _lib.check(st, "epb_lm_refine")
    return (out, iters, cost) if return_info else out


def rt34_to_rt6(rt34):
    """[B,3,4] -> [B,6] (cv2.Rodrigues matrix -> vector, then t)."""
    rt34 = rt34.contiguous()
    out = torch.empty((rt34.shape[0], 6), dtype=torch.float64, device=rt34.device)
    with torch.cuda.device(rt34.device):
        _lib.check(_lib.load().epb_rt34_to_rt6(_lib.ptr(rt34), rt34.shape[0], _lib.ptr(out), _lib.stream_ptr()),
                   "epb_rt34_to_rt6")
    return out


def pose_pack(rt6):
    """[B,6] -> (pose7 [B,7] f32 = (qw,qx,qy,qz,tx,ty,tz), rt34 [B,3,4] f64)  (val.py:203-224)."""
    rt6 = rt6.contiguous()
    b = rt6.shape[0]
    pose7 = torch.empty((b, 7), dtype=torch.float32, device=rt6.device)
    rt34 = torch.empty((b, 3, 4), dtype=torch.float64, device=rt6.device)
    with torch.cuda.device(rt6.device):
        _lib.check(_lib.load().epb_pose_pack(_lib.ptr(rt6), b, _lib.ptr(pose7), _lib.ptr(rt34), _lib.stream_ptr()),
                   "epb_pose_pack")
    return pose7, rt34


# LM weights of the heatmap path.  cpnp_m's weighting is an ASSUMPTION (its binary and source are absent from the
# reference): True / "maxval" = wxx = wyy = maxval (default), "sqrt" = sqrt(maxval), False / "unit" = cpnp.
_WEIGHTING = {False: 0, True: 1, 0: 0, 1: 1, 2: 2, "unit": 0, "maxval": 1, "sqrt": 2}


def pose_pipeline(preds, maxvals, bbox_xy, rate, p3d_model, K, min_k=24, sel_thresh=0.8, weighted=True):
    """val.py:172-228 for a batch, one stream-ordered call.  `weighted`: True / "maxval", "sqrt", False / "unit".
    preds [B,Kp,2] f32 (crop px), maxvals [B,Kp] f32, bbox_xy [B,2], rate [B], p3d_model [Kp,3], K [3,3].
    -> dict(pose7 [B,7] f32, rt6 [B,6] f64, epnp_rt34 [B,3,4] f64, status [B] i32)."""
    dev = preds.device
    preds = preds.to(torch.float32).contiguous()
    maxvals = maxvals.to(torch.float32).contiguous()
    bbox_xy, rate, p3d_model, K = _f64(bbox_xy, dev), _f64(rate, dev), _f64(p3d_model, dev), _f64(K, dev)
    b, kp = preds.shape[0], preds.shape[1]
    pose7 = torch.empty((b, 7), dtype=torch.float32, device=dev)
    rt6 = torch.empty((b, 6), dtype=torch.float64, device=dev)
    epnp = torch.empty((b, 3, 4), dtype=torch.float64, device=dev)
    status = torch.zeros((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().epb_pose_pipeline(_lib.ptr(preds), _lib.ptr(maxvals), _lib.ptr(bbox_xy), _lib.ptr(rate),
                                           _lib.ptr(p3d_model), _lib.ptr(K), b, kp, int(min_k), float(sel_thresh),
                                           _WEIGHTING[weighted], _lib.ptr(pose7), _lib.ptr(rt6), _lib.ptr(epnp),
                                           _lib.ptr(status), _lib.stream_ptr())
    _lib.check(st, "epb_pose_pipeline")
    return dict(pose7=pose7, rt6=rt6, epnp_rt34=epnp, status=status)


def esa_score(pose7_pred, pose7_gt):
    """demo.py:295-310 -> (score_t [B], score_r [B]) float64 on the device."""
    pred = pose7_pred.to(torch.float32).contiguous()
    gt = pose7_gt.to(device=pred.device, dtype=torch.float32).contiguous()
    b = pred.shape[0]
    st_ = torch.empty((b,), dtype=torch.float64, device=pred.device)
    sr_ = torch.empty((b,), dtype=torch.float64, device=pred.device)
    with torch.cuda.device(pred.device):
        _lib.check(_lib.load().epb_esa_score(_lib.ptr(pred), _lib.ptr(gt), b, _lib.ptr(st_), _lib.ptr(sr_),
                                             _lib.stream_ptr()), "epb_esa_score")
    return st_, sr_


def cov_to_weights(cov, isotropic=False):
    """cov [...,2,2] f32 -> LM weights [...,3] f64 (wxx,wxy,wyy) on the device.
    isotropic=False: inv(sqrtm(cov)) (evaluation_utils.py:170-181);
    isotropic=True:  1 / largest eigenvalue on the diagonal (uncertainty_pnp_v2, extend_utils.py:133-141)."""
    cov = cov.to(torch.float32).contiguous()
    n = cov.numel() // 4
    w = torch.empty(cov.shape[:-2] + (3,), dtype=torch.float64, device=cov.device)
    mode = _lib.WEIGHTS_INV_MAX_EIG if isotropic else _lib.WEIGHTS_INV_SQRTM
    with torch.cuda.device(cov.device):
        _lib.check(_lib.load().epb_cov_to_weights(_lib.ptr(cov), n, mode, _lib.ptr(w), _lib.stream_ptr()),
                   "epb_cov_to_weights")
    return w


def p3p_batch(p3d, p2d, K, w2d=None, return_status=False):
    """cv2.solvePnP(flags=SOLVEPNP_P3P) on four correspondences, batched on the device (extend_utils.py:85-89).
    p3d [n,3] | [B,n,3], p2d [B,n,2], K [3,3] | [B,3,3].  w2d None: n == 4, points in the given order (0..2 are met
    exactly, 3 picks the root).  w2d [B,n,3]: the four with the largest wxx + wxy, ascending
    (np.argsort(w[:,0] + w[:,1])[-4:], extend_utils.py:83).  -> rt34 [B,3,4] f64 (NaN where P3P has no root)."""
    dev = p2d.device
    p2d, p3d, K = _f64(p2d, dev), _f64(p3d, dev), _f64(K, dev)
    b, n = p2d.shape[0], p2d.shape[1]
    w = None if w2d is None else _f64(w2d, dev)
    rt34 = torch.empty((b, 3, 4), dtype=torch.float64, device=dev)
    status = torch.zeros((b,), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().epb_p3p(_lib.ptr(p3d), int(p3d.dim() == 3), _lib.ptr(p2d), _lib.ptr(w) if w is not None else None,
                                       _lib.ptr(K), int(K.dim() == 3), b, n, _lib.ptr(rt34), _lib.ptr(status),
                                       _lib.stream_ptr()), "epb_p3p")
    return (rt34, status) if return_status else rt34


def uncertainty_pnp_v2(points_2d, covars, points_3d, camera_matrix, type='single'):
    """extend_utils.py:117-178 -> [3,4]: isotropic weights 1 / lambda_max(cov) per keypoint, then the same
    P3P-initialised weighted LM as uncertainty_pnp."""
    dev = _device()
    cov = torch.from_numpy(np.ascontiguousarray(np.asarray(covars, np.float32))).to(dev)
    w = cov_to_weights(cov, isotropic=True).cpu().numpy()
    return uncertainty_pnp(points_2d, w, points_3d, camera_matrix)


def uncertainty_pnp_batch(mean_pts2d, covar, points_3d, K, return_info=False, init="p3p"):
    """Evaluator.evaluate_uncertainty (evaluation_utils.py:165-188) batched and device-resident:
    mean_pts2d [B,n,2], covar [B,n,2,2] (estimate_voting_distribution output), points_3d [n,3] | [B,n,3],
    K [3,3] | [B,3,3] -> rt34 [B,3,4] f64.  As in the reference the LM starts from P3P on the four
    best-weighted points (extend_utils.py:83-89); `init="epnp"` starts from RANSAC-EPnP on all points instead
    (both lie in the basin of the same weighted-reprojection minimiser on well-posed input)."""
    dev = mean_pts2d.device
    w = cov_to_weights(covar)
    p2, p3, Kd = _f64(mean_pts2d, dev), _f64(points_3d, dev), _f64(K, dev)
    start = p3p_batch(p3, p2, Kd, w2d=w) if init == "p3p" else pnp_batch(p3, p2, Kd)
    if p2.shape[1] == 4 and init == "p3p" and not return_info:
        return start                          # extend_utils.py:91-95: "no other points", the P3P pose is the answer
    init_rt = rt34_to_rt6(start)
    res = lm_refine_batch(p2, p3, w, Kd, init_rt, return_info=return_info)
    rt6 = res[0] if return_info else res
    _, rt34 = pose_pack(rt6)
    return (rt34, rt6) + tuple(res[1:]) if return_info else rt34


# ------------------------------------------------------------------------------------ drop-ins
def pnp(points_3d, points_2d, camera_matrix, method=SOLVEPNP_ITERATIVE):
    """pnp.py:46-90: RANSAC-EPnP (reprojectionError 5 px) -> [3,4] float64 [R|t].
    As in the reference the solver does not depend on `method` and a failed RANSAC is not
    reported (cv2 leaves stale memory there; this returns NaN)."""
    assert points_3d.shape[0] == points_2d.shape[0], 'points 3D and points 2D must have same number of vertices'
    dev = _device()
    rt = pnp_batch(_f64(points_3d, dev)[None], _f64(points_2d, dev)[None], _f64(camera_matrix, dev))
    return rt[0].cpu().numpy()


def _flat_k(K):
    k = K.detach().cpu().numpy() if isinstance(K, torch.Tensor) else np.asarray(K)
    return np.asarray(k, np.float64).reshape(-1)[-9:].reshape(3, 3)   # accepts the batched [1,3,3] of val.py:140


def cpnp(p3d, p2d, K, camera):
    """cpnp.cpnp(p3d, p2d, K, camera[6]) -> camera[6] (val.py:200): unit-weight LM refinement."""
    n = np.asarray(p3d).shape[0]
    w = np.stack([np.ones(n), np.zeros(n), np.ones(n)], 1)
    dev = _device()
    out = lm_refine_batch(_f64(p2d, dev)[None], _f64(p3d, dev)[None], _f64(w, dev)[None], _f64(_flat_k(K), dev),
                          _f64(camera, dev).reshape(1, 6))
    return out[0].cpu().numpy()


def cpnp_m(p3d, p2d, maxvals, K, camera):
    """cpnp.cpnp_m(p3d, p2d, maxvals, K, camera[6]) -> camera[6] (val.py:202): LM refinement weighted
    by the heatmap maxima (wxx = wyy = maxval, wxy = 0 -- the cpnp source is absent from the
    reference, so this weighting is an assumption)."""
    mv = np.asarray(maxvals, np.float64).reshape(-1)
    w = np.stack([mv, np.zeros_like(mv), mv], 1)
    dev = _device()
    out = lm_refine_batch(_f64(p2d, dev)[None], _f64(p3d, dev)[None], _f64(w, dev)[None], _f64(_flat_k(K), dev),
                          _f64(camera, dev).reshape(1, 6))
    return out[0].cpu().numpy()


def uncertainty_pnp(points_2d, weights_2d, points_3d, camera_matrix, init_rt=None):
    """extend_utils.py:64-115 -> [3,4]: P3P on the four correspondences with the largest wxx + wxy
    (np.argsort(...)[-4:], :83-89) gives the initial pose -- and the answer itself when pn == 4 (:91-95) --
    then the weighted LM of uncertainty_pnp.cpp.  `init_rt` ([6] angle-axis, t) overrides the P3P start."""
    pn = points_2d.shape[0]
    assert points_3d.shape[0] == pn and pn >= 4
    dev = _device()
    p2, p3, w = _f64(points_2d, dev)[None], _f64(points_3d, dev)[None], _f64(weights_2d, dev)[None]
    K = _f64(camera_matrix, dev)
    if init_rt is None:
        rt34 = p3p_batch(p3, p2, K, w2d=w)
        if pn == 4:
            return rt34[0].cpu().numpy()
        init = rt34_to_rt6(rt34)
    else:
        init = _f64(init_rt, dev).reshape(1, 6)
    rt6 = lm_refine_batch(p2, p3, w, K, init)
    _, rt34 = pose_pack(rt6)
    return rt34[0].cpu().numpy()
