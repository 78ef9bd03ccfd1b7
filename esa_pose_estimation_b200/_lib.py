"""ctypes binding of libesa_pose_b200.so (include/esa_pose_b200.h).

There is no CPU fallback: if the library is missing or a launch fails, calls raise.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libesa_pose_b200.so")

c_int, c_float, c_double, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_double, ctypes.c_void_p
c_ll, c_ull, c_size_t = ctypes.c_longlong, ctypes.c_ulonglong, ctypes.c_size_t

EPB_OK = 0
STATUS_NAMES = {0: "EPB_OK", 1: "EPB_ERR_INVALID", 2: "EPB_ERR_CUDA", 3: "EPB_ERR_WORKSPACE",
                4: "EPB_ERR_NO_DEVICE"}

DECODE_REFINE, DECODE_ZERO_NONPOS = 1, 2
VOTE_V3, VOTE_V4, VOTE_V5, VOTE_HYPOTHESIS, VOTE_DISTRIBUTION, VOTE_DISTRIBUTION_WITH_MEAN, VOTE_V1, VOTE_V2, VOTE_MOTION = range(9)
MASK_NONZERO, MASK_EQ1, MASK_CLASS = 0, 1, 2
RNG_IDXS, RNG_RAW32, RNG_PHILOX = 0, 1, 2
STAGE_ALL, STAGE_GATHER, STAGE_VOTE = 0, 1, 2
WEIGHTS_INV_SQRTM, WEIGHTS_INV_MAX_EIG = 0, 1
POSE_OK, POSE_FAILED, POSE_TOO_FEW = 0, 1, 2


class VotingParams(ctypes.Structure):
    _fields_ = [("mode", c_int), ("B", c_int), ("H", c_int), ("W", c_int), ("vn", c_int),
                ("hn", c_int), ("rounds", c_int), ("inlier_thresh", c_float),
                ("min_num", c_int), ("max_num", c_int), ("topk", c_int), ("mask_mode", c_int),
                ("sb", c_ll), ("sy", c_ll), ("sx", c_ll), ("sv", c_ll), ("sc", c_ll),
                ("rng_mode", c_int), ("philox_seed", c_ull), ("philox_offset", c_ull),
                ("philox_sm_count", c_int), ("philox_threads_per_sm", c_int), ("stage", c_int),
                ("classes", c_int), ("refine_iters", c_int)]


class VotingIO(ctypes.Structure):
    _fields_ = [("mask", c_void_p), ("vertex", c_void_p), ("idxs", c_void_p), ("selection", c_void_p),
                ("mean_in", c_void_p), ("pts", c_void_p), ("var_or_conf", c_void_p), ("hyp", c_void_p),
                ("counts", c_void_p), ("mean", c_void_p), ("cov", c_void_p), ("tn_out", c_void_p),
                ("status", c_void_p), ("philox_consumed", c_void_p), ("philox_state", c_void_p)]


# name -> (restype, argtypes); every symbol include/esa_pose_b200.h declares
SIGNATURES = {
    "epb_version": (c_int, []),
    "epb_last_cuda_error": (c_int, []),
    "epb_last_cuda_error_string": (ctypes.c_char_p, []),
    "epb_launch_count": (c_ull, []),
    "epb_device_info": (c_int, [ctypes.POINTER(c_int), ctypes.POINTER(c_int), ctypes.POINTER(c_size_t)]),
    "epb_profile_enable": (c_int, [c_int]),
    "epb_profile_read": (c_int, [c_int, ctypes.POINTER(c_double), ctypes.POINTER(c_int)]),
    "epb_decode_heatmaps": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "epb_refine_keypoints": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "epb_refine_keypoints_dark": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "epb_generate_hypothesis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "epb_voting_for_hypothesis": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "epb_generate_hypothesis_vanishing_point": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "epb_voting_for_hypothesis_vanishing_point": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p]),
    "epb_voting_workspace_bytes": (c_size_t, [ctypes.POINTER(VotingParams)]),
    "epb_voting_run": (c_int, [ctypes.POINTER(VotingParams), ctypes.POINTER(VotingIO), c_void_p, c_size_t, c_void_p]),
    "epb_pnp_epnp_ransac": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_double,
                                    c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "epb_lm_refine": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int,
                              c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "epb_pose_pack": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "epb_rt34_to_rt6": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "epb_pose_pipeline": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                  c_double, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "epb_cov_to_weights": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "epb_p3p": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "epb_esa_score": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "epb_pose_metrics": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p]),
    "epb_nearest_point_idx": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
}

_lib = None


def load():
    """Loads the CUDA library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "esa_pose_estimation_b200: %s is missing. Build it with "
            "`python -m esa_pose_estimation_b200.build` (needs nvcc); there is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the header and the library diverge
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status, what):
    if status != EPB_OK:
        lib = load()
        extra = ""
        if status == 2:
            extra = " (CUDA error %d: %s)" % (lib.epb_last_cuda_error(),
                                               lib.epb_last_cuda_error_string().decode())
        raise RuntimeError("%s failed: %s%s" % (what, STATUS_NAMES.get(status, status), extra))


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(t, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)   # mirrors CHECK_CUDA, ransac_voting.cpp:7
    return t


def require_contiguous(t, name):
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)      # mirrors CHECK_CONTIGUOUS, :8
    return t
