"""Heatmap decoding -- drop-in for the reference's ``inference.py`` on the CUDA path.

Same names, argument meaning and return values as /root/reference/inference.py:
``get_max_preds`` (:22), ``get_final`` (:136), ``get_final2`` (:154), ``getPrediction`` (:171); plus the batched
device-resident entry ``decode_heatmaps`` that replaces the per-frame two-stage ``torch.max``
and 150 ``.item()`` syncs of val.py:151-166.  Everything computes in
csrc/decode.cu through the C ABI; there is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("esa_pose_estimation_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", torch.cuda.current_device())


def decode_heatmaps(hms, refine=True, zero_nonpositive=False):
    """hms: CUDA float32 tensor [B,K,H,W].  -> (xy [B,K,2] f32, maxval [B,K] f32, idx [B,K] i32),
    all on the device, stream-ordered, no host synchronisation."""
    _lib.require_cuda(hms, "hms")
    if hms.dim() != 4:
        raise AssertionError("Score maps should be 4-dim")
    if hms.dtype != torch.float32:
        hms = hms.float()
    hms = hms.contiguous()
    b, k, h, w = hms.shape
    xy = torch.empty((b, k, 2), dtype=torch.float32, device=hms.device)
    mv = torch.empty((b, k), dtype=torch.float32, device=hms.device)
    idx = torch.empty((b, k), dtype=torch.int32, device=hms.device)
    flags = (_lib.DECODE_REFINE if refine else 0) | (_lib.DECODE_ZERO_NONPOS if zero_nonpositive else 0)
    with torch.cuda.device(hms.device):
        st = _lib.load().epb_decode_heatmaps(_lib.ptr(hms), b * k, h, w, flags, _lib.ptr(xy), _lib.ptr(mv),
                                             _lib.ptr(idx), _lib.stream_ptr())
    _lib.check(st, "epb_decode_heatmaps")
    return xy, mv, idx


def get_max_preds(batch_heatmaps):
    """inference.py:22-51.  numpy [B,K,H,W] -> (preds [B,K,2] f32, maxvals [B,K,1])."""
    assert isinstance(batch_heatmaps, np.ndarray), 'batch_heatmaps should be numpy.ndarray'
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    hm = torch.from_numpy(np.ascontiguousarray(batch_heatmaps, dtype=np.float32)).to(_device())
    xy, mv, _ = decode_heatmaps(hm, refine=False)
    preds = xy.cpu().numpy()
    maxvals = mv.cpu().numpy().astype(batch_heatmaps.dtype, copy=False)
    return preds, maxvals.reshape(batch_heatmaps.shape[0], batch_heatmaps.shape[1], 1)


def get_final(hm, coords):
    """inference.py:136-152.  hm numpy [1,K,H,W]; coords: list/array of K float32[2] integer peaks
    (mutated in place like the reference); returns the refined [K,2] array."""
    hm_t = torch.from_numpy(np.ascontiguousarray(hm[0], dtype=np.float32)).to(_device())
    k, h, w = hm_t.shape
    n = len(coords)
    xy = torch.from_numpy(np.ascontiguousarray(np.asarray(coords, dtype=np.float32).reshape(n, 2))).to(hm_t.device)
    st = _lib.load().epb_refine_keypoints(_lib.ptr(hm_t), min(n, k), h, w, _lib.ptr(xy), _lib.stream_ptr())
    _lib.check(st, "epb_refine_keypoints")
    out = xy.cpu().numpy()
    for p in range(n):
        coords[p] = out[p]
    return out.copy() if isinstance(coords, list) else coords.copy()


def refine_dark(hms, xy):
    """Batched get_final2 on the device: hms [B,K,H,W] CUDA f32, xy [B,K,2] f32 integer peaks -> refined copy."""
    hms = hms.contiguous()
    b, k, h, w = hms.shape
    out = xy.to(torch.float32).contiguous().clone()
    with torch.cuda.device(hms.device):
        st = _lib.load().epb_refine_keypoints_dark(_lib.ptr(hms), b * k, h, w, _lib.ptr(out), _lib.stream_ptr())
    _lib.check(st, "epb_refine_keypoints_dark")
    return out


def get_final2(hm, coords):
    """inference.py:154-170 (DARK-style decode).  hm numpy [1,K,H,W]; coords: K float32[2] integer peaks,
    mutated in place like the reference; returns the refined [K,2] array."""
    hm_t = torch.from_numpy(np.ascontiguousarray(hm[:1], dtype=np.float32)).to(_device())
    n = len(coords)
    xy = torch.from_numpy(np.ascontiguousarray(np.asarray(coords, dtype=np.float32).reshape(1, n, 2))).to(hm_t.device)
    out = refine_dark(hm_t[:, :n], xy)[0].cpu().numpy()
    for p in range(n):
        coords[p] = out[p]
    return out.copy() if isinstance(coords, list) else coords.copy()


def getPrediction(hms, inpH=128, inpW=128):
    """inference.py:171-186.  CUDA tensor [B,K,H,W] -> (preds [B,K,2] f32, maxval [B,K,1]);
    coordinates are zeroed where maxval <= 0."""
    assert hms.dim() == 4, 'Score maps should be 4-dim'
    xy, mv, _ = decode_heatmaps(hms, refine=False, zero_nonpositive=True)
    return xy, mv.view(hms.size(0), hms.size(1), 1)
