"""Drop-in for the reference's pybind11 module ``ransac_voting``
(/root/reference/lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:102-107): same four
functions, same tensor layouts, same CHECK_INPUT behaviour (RuntimeError unless CUDA and
contiguous).  Kernels: csrc/voting.cu, launched on torch's current stream."""
import torch

from . import _lib


def _check(t, name, dtype):
    _lib.require_cuda(t, name)
    _lib.require_contiguous(t, name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s" % (name, dtype))   # the reference's .data<T>() throws
    return t


def generate_hypothesis(direct, coords, idxs):
    """direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hypo_pts [hn,vn,2] f32."""
    _check(direct, "direct", torch.float32); _check(coords, "coords", torch.float32)
    _check(idxs, "idxs", torch.int32)
    tn, vn, hn = direct.size(0), direct.size(1), idxs.size(0)
    hyp = torch.empty((hn, vn, 2), dtype=torch.float32, device=direct.device)
    with torch.cuda.device(direct.device):
        st = _lib.load().epb_generate_hypothesis(_lib.ptr(direct), _lib.ptr(coords), _lib.ptr(idxs), _lib.ptr(hyp),
                                                 tn, vn, hn, _lib.stream_ptr())
    _lib.check(st, "epb_generate_hypothesis")
    return hyp


def voting_for_hypothesis(direct, coords, hypo_pts, inliers, inlier_thresh):
    """Writes 1 into the caller-zeroed inliers [hn,vn,tn] u8 where pixel t votes for hypothesis (h,v)."""
    _check(direct, "direct", torch.float32); _check(coords, "coords", torch.float32)
    _check(hypo_pts, "hypo_pts", torch.float32); _check(inliers, "inliers", torch.uint8)
    tn, vn, hn = direct.size(0), direct.size(1), hypo_pts.size(0)
    with torch.cuda.device(direct.device):
        st = _lib.load().epb_voting_for_hypothesis(_lib.ptr(direct), _lib.ptr(coords), _lib.ptr(hypo_pts),
                                                   _lib.ptr(inliers), tn, vn, hn, float(inlier_thresh),
                                                   _lib.stream_ptr())
    _lib.check(st, "epb_voting_for_hypothesis")


def generate_hypothesis_vanishing_point(direct, coords, idxs):
    """-> hypo_pts [hn,vn,3] f32 (homogeneous)."""
    _check(direct, "direct", torch.float32); _check(coords, "coords", torch.float32)
    _check(idxs, "idxs", torch.int32)
    tn, vn, hn = direct.size(0), direct.size(1), idxs.size(0)
    hyp = torch.empty((hn, vn, 3), dtype=torch.float32, device=direct.device)
    with torch.cuda.device(direct.device):
        st = _lib.load().epb_generate_hypothesis_vanishing_point(_lib.ptr(direct), _lib.ptr(coords), _lib.ptr(idxs),
                                                                 _lib.ptr(hyp), tn, vn, hn, _lib.stream_ptr())
    _lib.check(st, "epb_generate_hypothesis_vanishing_point")
    return hyp


def voting_for_hypothesis_vanishing_point(direct, coords, hypo_pts, inliers, inlier_thresh):
    _check(direct, "direct", torch.float32); _check(coords, "coords", torch.float32)
    _check(hypo_pts, "hypo_pts", torch.float32); _check(inliers, "inliers", torch.uint8)
    tn, vn, hn = direct.size(0), direct.size(1), hypo_pts.size(0)
    with torch.cuda.device(direct.device):
        st = _lib.load().epb_voting_for_hypothesis_vanishing_point(
            _lib.ptr(direct), _lib.ptr(coords), _lib.ptr(hypo_pts), _lib.ptr(inliers), tn, vn, hn,
            float(inlier_thresh), _lib.stream_ptr())
    _lib.check(st, "epb_voting_for_hypothesis_vanishing_point")
