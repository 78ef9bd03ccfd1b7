"""RANSAC keypoint voting -- drop-in for the reference's
``lib/ransac_voting_gpu_layer/ransac_voting_gpu.py`` drivers, batched on the device.

Same positional signatures and defaults as the reference:
  ransac_voting_layer (:10)       ransac_voting_layer_v2 (:99)    (multi-class)
  ransac_voting_layer_v3 (:514)   ransac_voting_layer_v4 (:669)   ransac_voting_layer_v5 (:763)
  ransac_voting_hypothesis (:218) estimate_voting_distribution (:263)   ransac_motion_voting (:960)
  estimate_voting_distribution_with_mean (:333)      vertex_layer_reshape (base_utils.py:311)
Each call is ONE stream-ordered ``epb_voting_run`` over the whole batch (csrc/voting.cu): no
per-image Python loop, no ``.item()`` sync, no hn*vn*tn byte tensor.

Random numbers.  The reference draws ``idxs`` (and the ``max_num`` subsample) with torch's CUDA
generator.  By default the kernels regenerate exactly those draws on the fly (Philox4x32-10 in
torch's kernel layout, seeded from ``torch.cuda.default_generators``), so under
``torch.manual_seed(s)`` the hypothesis indices equal the reference's bit for bit, and the
generator is advanced by what the reference would have consumed.  Keyword-only extras:
  idxs=        int32 [b,rounds,hn,vn,2] explicit indices (parity tests)
  raw32=True   treat ``idxs`` as raw 32-bit draws, index = draw % tn on the device
  selection=   float32 [b,h,w] uniform draws for the max_num subsample
  vertex may also be a float32 tensor in PINNED host memory (``.pin_memory()``): the kernels then
  read the foreground pixels of the field in place over PCIe (zero-copy) instead of the caller
  copying the whole field to the device; ``mask`` may be a host tensor in that case too.
  sync_rng=    True (default): read back how much of the generator stream was consumed (one
               8-byte D2H copy) so later torch draws continue exactly like the reference;
               False: advance by the data-independent maximum, no synchronisation.
"""
import math

import torch

from . import _lib

_workspaces = {}


def vertex_layer_reshape(vertex_pred):
    """base_utils.py:311-316: [b,2vn,h,w] -> [b,h,w,vn,2] view (no copy; the kernels take strides)."""
    b, vn2, h, w = vertex_pred.shape
    return vertex_pred.permute(0, 2, 3, 1).view(b, h, w, vn2 // 2, 2)


def _workspace(device, nbytes):
    # one scratch buffer per (device, caller stream): calls issued on different streams (or from threads that
    # use different streams) never share fgpix / records / counts, and a buffer is only ever freed into the pool
    # of the stream that used it
    key = (device.type, device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        ws = torch.empty((int(nbytes * 1.25) + 4096,), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _mask_u8(mask, mask_mode):
    if mask.dtype == torch.uint8:
        return mask.contiguous()
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    if mask_mode == _lib.MASK_EQ1:
        return (mask == 1).contiguous().view(torch.uint8)
    return mask.byte().contiguous()                       # (mask[bi]).byte(), :526


def _rng_layout(numel, props):
    grid = min((numel + 255) // 256, props.multi_processor_count * (props.max_threads_per_multi_processor // 256))
    grid = max(grid, 1)
    return ((numel - 1) // (256 * grid * 4) + 1) * 4 if numel else 0


def _run(mode, mask, vertex, hn, rounds, thresh, min_num, max_num, topk=0, mean_in=None, idxs=None,
         raw32=False, selection=None, sync_rng=True, want_hyp=False, want_status=False,
         stage=_lib.STAGE_ALL, workspace=None, rng_state=None, classes=1, refine_iters=1):
    """One epb_voting_run.  `stage` / `workspace` split a run in two stream-ordered halves
    (STAGE_GATHER fills `workspace`, STAGE_VOTE consumes it; same arguments both times) so that
    pipeline.py can overlap the PCIe read of one batch chunk with the voting of the previous one.
    The torch generator is read and advanced by the gather half only; with `rng_state` (a device
    int64 [1] tensor holding the generator offset) consecutive runs chain their Philox offsets on the
    device and the caller advances the torch generator once at the end."""
    if isinstance(vertex, torch.Tensor) and not vertex.is_cuda and vertex.is_pinned():
        # Host-resident network output: the field is read in place over PCIe by field_gather_kernel
        # (zero-copy, foreground pixels only); only the mask is copied to the device.
        if vertex.dtype != torch.float32:
            raise RuntimeError("a pinned host vertex field must be float32")
        dev = mask.device if mask.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if not mask.is_cuda:
            mask = mask.to(dev, non_blocking=True)
    else:
        _lib.require_cuda(mask, "mask")
        _lib.require_cuda(vertex, "vertex")       # pageable host memory is rejected like CHECK_CUDA
        if vertex.dtype != torch.float32:
            vertex = vertex.float()
        dev = vertex.device
    if vertex.dim() != 5 or vertex.shape[-1] != 2:
        raise RuntimeError("vertex must be [b,h,w,vn,2]")
    b, h, w, vn, _ = vertex.shape
    multi = mode in (_lib.VOTE_V1, _lib.VOTE_V2)
    mask_mode = _lib.MASK_CLASS if multi else (_lib.MASK_NONZERO if (mode <= _lib.VOTE_V5 or mode == _lib.VOTE_MOTION)
                                               else _lib.MASK_EQ1)
    mask_u8 = mask.to(torch.uint8).contiguous() if multi else _mask_u8(mask, mask_mode)   # class labels (mask == k+1, :25)
    assert mask_u8.shape == (b, h, w), "mask must be [b,h,w]"
    hn_total = hn * rounds
    classes = max(int(classes), 1) if multi else 1
    bv = b * classes                                      # (image, class) pairs: one virtual image each

    p = _lib.VotingParams()
    p.mode, p.B, p.H, p.W, p.vn, p.hn, p.rounds = mode, b, h, w, vn, int(hn), int(rounds)
    p.inlier_thresh, p.min_num, p.max_num, p.topk, p.mask_mode = float(thresh), int(min_num), int(max_num), int(topk), mask_mode
    sb, sy, sx, sv, sc = vertex.stride()
    p.sb, p.sy, p.sx, p.sv, p.sc = sb, sy, sx, sv, sc
    p.classes, p.refine_iters = classes, int(refine_iters)
    props = torch.cuda.get_device_properties(dev)
    p.philox_sm_count = props.multi_processor_count
    p.philox_threads_per_sm = props.max_threads_per_multi_processor

    io = _lib.VotingIO()
    keep = [mask_u8, vertex]
    gen = None
    if idxs is not None:
        idxs = idxs.to(device=dev)
        if idxs.dtype != torch.int32:
            # raw32: keep the low 32 bits (bit pattern); indices: plain conversion
            idxs = (idxs & 0xFFFFFFFF).to(torch.int64).to(torch.int32) if raw32 else idxs.to(torch.int32)
        idxs = idxs.contiguous()
        assert idxs.numel() == bv * rounds * hn * vn * 2, "idxs must be [b(*classes),rounds,hn,vn,2]"
        p.rng_mode = _lib.RNG_RAW32 if raw32 else _lib.RNG_IDXS
        io.idxs = _lib.ptr(idxs)
        keep.append(idxs)
        if selection is not None:
            selection = selection.to(device=dev, dtype=torch.float32).contiguous()
            io.selection = _lib.ptr(selection)
            keep.append(selection)
    else:
        p.rng_mode = _lib.RNG_PHILOX
        gen = torch.cuda.default_generators[dev.index if dev.index is not None else torch.cuda.current_device()]
        p.philox_seed = gen.initial_seed() & 0xFFFFFFFFFFFFFFFF
        p.philox_offset = gen.get_offset()
        if stage == _lib.STAGE_VOTE or rng_state is not None:
            gen = None                                    # advanced by the gather half / by the caller
        if rng_state is not None:
            io.philox_state = _lib.ptr(rng_state)
            keep.append(rng_state)
        if selection is not None:
            selection = selection.to(device=dev, dtype=torch.float32).contiguous()
            io.selection = _lib.ptr(selection)
            keep.append(selection)

    out = {}
    f32 = dict(dtype=torch.float32, device=dev)
    p.stage = stage
    if stage == _lib.STAGE_GATHER:
        mode_outputs = False
    else:
        mode_outputs = True
    if mode_outputs and multi:
        out["pts"] = torch.empty((b, classes, vn, 2), **f32)
        io.pts = _lib.ptr(out["pts"])
    elif mode_outputs and (mode <= _lib.VOTE_V5 or mode == _lib.VOTE_MOTION):
        out["pts"] = torch.empty((b, vn, 2), **f32)
        io.pts = _lib.ptr(out["pts"])
        if mode != _lib.VOTE_V3:
            out["aux"] = torch.empty((b, vn), **f32)
            io.var_or_conf = _lib.ptr(out["aux"])
    if mode_outputs and (mode == _lib.VOTE_HYPOTHESIS or want_hyp):
        out["hyp"] = torch.empty((bv, hn_total, vn, 2), **f32)
        out["counts"] = torch.empty((bv, hn_total, vn), dtype=torch.int32, device=dev)
        io.hyp, io.counts = _lib.ptr(out["hyp"]), _lib.ptr(out["counts"])
    if mode_outputs and mode >= _lib.VOTE_DISTRIBUTION:
        out["mean"] = torch.empty((b, vn, 2), **f32)
        out["cov"] = torch.empty((b, vn, 2, 2), **f32)
        io.mean, io.cov = _lib.ptr(out["mean"]), _lib.ptr(out["cov"])
        if mode == _lib.VOTE_DISTRIBUTION_WITH_MEAN:
            mean_in = mean_in.to(device=dev, dtype=torch.float32).contiguous()
            io.mean_in = _lib.ptr(mean_in)
            keep.append(mean_in)
    if mode_outputs:
        out["tn"] = torch.empty((bv,), dtype=torch.int32, device=dev)
        io.tn_out = _lib.ptr(out["tn"])
    if want_status and mode_outputs:
        out["status"] = torch.zeros((bv, vn), dtype=torch.int32, device=dev)
        io.status = _lib.ptr(out["status"])
    consumed = None
    if gen is not None:
        consumed = torch.empty((1,), dtype=torch.int64, device=dev)   # always written by rng_offsets_kernel
        io.philox_consumed = _lib.ptr(consumed)
    io.mask, io.vertex = _lib.ptr(mask_u8), _lib.ptr(vertex)

    lib = _lib.load()
    nbytes = lib.epb_voting_workspace_bytes(p)
    if nbytes == 0:
        raise RuntimeError("epb_voting_workspace_bytes: invalid parameters")
    if workspace is None:
        if stage != _lib.STAGE_ALL:
            raise RuntimeError("a split run needs an explicit workspace shared by both halves")
        ws = _workspace(dev, nbytes + 256)
    else:
        ws = workspace
        if ws.numel() < nbytes + 256 or not ws.is_cuda:
            raise RuntimeError("workspace too small: need %d bytes" % (nbytes + 256))
    base = (ws.data_ptr() + 255) & ~255
    with torch.cuda.device(dev):
        st = lib.epb_voting_run(p, io, _lib.c_void_p(base), nbytes, _lib.stream_ptr())
    _lib.check(st, "epb_voting_run")
    if gen is not None:
        if sync_rng:
            gen.set_offset(p.philox_offset + int(consumed.item()))
        else:
            inc = rounds * _rng_layout(hn * vn * 2, props) + (_rng_layout(h * w, props) if h * w > max_num else 0)
            gen.set_offset(p.philox_offset + bv * inc)
    return out


def workspace_bytes(b, h, w, vn, hn, rounds=1, max_num=30000, philox=True):
    """Bytes of caller-owned workspace one run over [b,h,w,vn,2] needs (+ alignment slack).  With the torch-compatible
    Philox draws (the default) the size is bounded by max_num, with explicit idxs / selection by h * w."""
    p = _lib.VotingParams()
    p.mode, p.B, p.H, p.W, p.vn, p.hn, p.rounds = _lib.VOTE_V3, b, h, w, vn, int(hn), int(rounds)
    p.max_num = int(max_num)
    p.rng_mode = _lib.RNG_PHILOX if philox else _lib.RNG_IDXS
    n = _lib.load().epb_voting_workspace_bytes(p)
    if n == 0:
        raise RuntimeError("epb_voting_workspace_bytes: invalid parameters")
    return int(n) + 256


def ransac_voting_layer(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                        min_num=5, max_num=30000, **kw):
    """:10-97, multi-class: for every image and every class k+1 in 1..class_num-1 the winning hypothesis of
    the pixels with mask == k+1.  -> [b, class_num-1, vn, 2] (zeros where a class has < min_num pixels)."""
    return _run(_lib.VOTE_V1, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num,
                classes=class_num - 1, **kw)["pts"]


def ransac_voting_layer_v2(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, refine_iter_num=1, **kw):
    """:99-216: v1 followed by `refine_iter_num` least-squares refinements over the winner's inliers
    (torch.pinverse: zeros without inliers, minimum-norm solution for parallel normals).  -> [b,cn-1,vn,2]."""
    return _run(_lib.VOTE_V2, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num,
                classes=class_num - 1, refine_iters=refine_iter_num, **kw)["pts"]


def ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, **kw):
    """:514-598.  mask [b,h,w], vertex [b,h,w,vn,2] -> win_pts [b,vn,2].
    `confidence` / `max_iter` only bound how often the reference re-scores the SAME hypotheses
    (idxs is drawn once, :547); the result is fixed after round 1, so they are accepted and unused."""
    return _run(_lib.VOTE_V3, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num, **kw)["pts"]


def ransac_voting_layer_v4(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999, max_iter=20,
                           min_num=5, max_num=30000, **kw):
    """:669-761 -> (win_pts [b,vn,2], var [b,vn])."""
    o = _run(_lib.VOTE_V4, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num, **kw)
    return o["pts"], o["aux"]


def ransac_voting_layer_v5(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=100, **kw):
    """:763-858 -> (win_pts [b,vn,2], confidence [b,vn])."""
    o = _run(_lib.VOTE_V5, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num, **kw)
    return o["pts"], o["aux"]


def ransac_voting_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, min_num=5, max_num=30000, **kw):
    """:218-261 -> (hyp_pts [b,hn,vn,2] f32, inlier_counts [b,hn,vn] int64)."""
    o = _run(_lib.VOTE_HYPOTHESIS, mask, vertex, round_hyp_num, 1, inlier_thresh, min_num, max_num, **kw)
    return o["hyp"], o["counts"].long()


def estimate_voting_distribution(mask, vertex, round_hyp_num=256, min_hyp_num=4096, topk=128, inlier_thresh=0.99,
                                 min_num=5, max_num=30000, **kw):
    """:263-331 -> (mean [b,vn,2], cov [b,vn,2,2])."""
    rounds = int(math.ceil(min_hyp_num / round_hyp_num))
    o = _run(_lib.VOTE_DISTRIBUTION, mask, vertex, round_hyp_num, rounds, inlier_thresh, min_num, max_num,
             topk=topk, **kw)
    return o["mean"], o["cov"]


def estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=256, min_hyp_num=4096, topk=128,
                                           inlier_thresh=0.99, min_num=5, max_num=30000, output_hyp=False, **kw):
    """:333-406 -> (mean, cov [b,vn,2,2])."""
    rounds = int(math.ceil(min_hyp_num / round_hyp_num))
    o = _run(_lib.VOTE_DISTRIBUTION_WITH_MEAN, mask, vertex, round_hyp_num, rounds, inlier_thresh, min_num,
             max_num, topk=topk, mean_in=mean, **kw)
    return o["mean"], o["cov"]


def ransac_motion_voting(mask, vertex):
    """:960-981: mean over the foreground (mask.byte() != 0) of vertex + pixel coordinate -> [b,vn,2];
    zeros for an image without foreground."""
    dummy = torch.zeros((vertex.shape[0], 1, 1, vertex.shape[3], 2), dtype=torch.int32, device=mask.device if mask.is_cuda else None)
    return _run(_lib.VOTE_MOTION, mask, vertex, 1, 1, 0.999, 1, 2 ** 31 - 1, idxs=dummy)["pts"]


def voting_debug(mode, mask, vertex, round_hyp_num, rounds=1, inlier_thresh=0.999, min_num=5, max_num=30000,
                 topk=128, mean_in=None, **kw):
    """Everything one run produces (pts/aux/hyp/counts/mean/cov/tn/status) -- used by the parity tests."""
    return _run(mode, mask, vertex, round_hyp_num, rounds, inlier_thresh, min_num, max_num, topk=topk,
                mean_in=mean_in, want_hyp=True, want_status=mode <= _lib.VOTE_V5, **kw)
