"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's RANSAC voting drivers.

Follows /root/reference/lib/ransac_voting_gpu_layer/ransac_voting_gpu.py:
  ransac_voting_layer_v3 :514-598, _v4 :669-761, _v5 :763-858,
  ransac_voting_hypothesis :218-261, estimate_voting_distribution :263-331,
  estimate_voting_distribution_with_mean :333-406, b_inv :503-512.
The two kernels it calls are restated in oracle/voting_oracle.c (ctypes).

The reference draws its random numbers with torch's CUDA generator; here the
caller passes them in (``idxs_fn`` / ``selection_fn``) so that parity never depends on
an RNG re-implementation.

Pinned: the reference's ransac_voting_gpu.py itself is device-agnostic Python; imported unmodified
and run on CPU tensors over the C restatement of its kernels (its pybind binary is not shipped; the
removed ``torch.solve`` and uint8 ``masked_select`` are supplied from outside), it produces
tests/golden/voting_drivers.npz (tests/golden/make_golden_voting.py).  Every driver below reproduces
those vectors with the reference's recorded random draws replayed: hypotheses, counts, v1 winners and
v5 confidence bit-exact, keypoints within 1e-3 px (tests/test_oracle_voting.py).  The kernel
arithmetic underneath is pinned against oracle/_ref/libref_voting.so on the GPU box.
"""
import ctypes

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------- kernels
def generate_hypothesis(direct, coords, idxs):
    """ransac_voting.generate_hypothesis (ransac_voting.cpp:20-31 -> .cu:11-86)."""
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    idxs = np.ascontiguousarray(idxs, np.int32)
    tn, vn, _ = direct.shape
    hn = idxs.shape[0]
    hyp = np.zeros((hn, vn, 2), np.float32)
    _lib.oracle_lib().orc_generate_hypothesis(
        _lib.fptr(direct), _lib.fptr(coords), _lib.iptr(idxs), _lib.fptr(hyp),
        ctypes.c_int(tn), ctypes.c_int(vn), ctypes.c_int(hn))
    return hyp


def voting_for_hypothesis(direct, coords, hyp, inliers, thresh):
    """ransac_voting.voting_for_hypothesis (.cpp:41-55 -> .cu:88-167); writes 1s in place."""
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    hyp = np.ascontiguousarray(hyp, np.float32)
    tn, vn, _ = direct.shape
    hn = hyp.shape[0]
    assert inliers.dtype == np.uint8 and inliers.shape == (hn, vn, tn) and inliers.flags.c_contiguous
    _lib.oracle_lib().orc_voting_for_hypothesis(
        _lib.fptr(direct), _lib.fptr(coords), _lib.fptr(hyp), _lib.u8ptr(inliers),
        ctypes.c_int(tn), ctypes.c_int(vn), ctypes.c_int(hn), ctypes.c_float(thresh))


def generate_hypothesis_vanishing_point(direct, coords, idxs):
    """ransac_voting.generate_hypothesis_vanishing_point (.cpp:64 -> .cu:170-229) -> [hn,vn,3]."""
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    idxs = np.ascontiguousarray(idxs, np.int32)
    tn, vn, _ = direct.shape
    hn = idxs.shape[0]
    hyp = np.zeros((hn, vn, 3), np.float32)
    _lib.oracle_lib().orc_generate_hypothesis_vanishing_point(
        _lib.fptr(direct), _lib.fptr(coords), _lib.iptr(idxs), _lib.fptr(hyp),
        ctypes.c_int(tn), ctypes.c_int(vn), ctypes.c_int(hn))
    return hyp


def voting_for_hypothesis_vanishing_point(direct, coords, hyp, inliers, thresh):
    """ransac_voting.voting_for_hypothesis_vanishing_point (.cpp:85 -> .cu:268-310); writes 1s in place."""
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    hyp = np.ascontiguousarray(hyp, np.float32)
    tn, vn, _ = direct.shape
    hn = hyp.shape[0]
    assert inliers.dtype == np.uint8 and inliers.shape == (hn, vn, tn) and inliers.flags.c_contiguous
    _lib.oracle_lib().orc_voting_for_hypothesis_vanishing_point(
        _lib.fptr(direct), _lib.fptr(coords), _lib.fptr(hyp), _lib.u8ptr(inliers),
        ctypes.c_int(tn), ctypes.c_int(vn), ctypes.c_int(hn), ctypes.c_float(thresh))


def vote_counts(direct, coords, hyp, thresh):
    """voting_for_hypothesis + torch.sum(inlier, 2) (ransac_voting_gpu.py:557-561)."""
    direct = np.ascontiguousarray(direct, np.float32)
    coords = np.ascontiguousarray(coords, np.float32)
    hyp = np.ascontiguousarray(hyp, np.float32)
    tn, vn, _ = direct.shape
    hn = hyp.shape[0]
    counts = np.zeros((hn, vn), np.int32)
    _lib.oracle_lib().orc_vote_counts(
        _lib.fptr(direct), _lib.fptr(coords), _lib.fptr(hyp), _lib.iptr(counts),
        ctypes.c_int(tn), ctypes.c_int(vn), ctypes.c_int(hn), ctypes.c_float(thresh))
    return counts


# ----------------------------------------------------------------------------- helpers
def default_idxs_fn(seed):
    """CPU stand-in for ``zeros([hn,vn,2], int32).random_(0, tn)`` (:547, :297)."""
    def fn(bi, round_idx, hn, vn, tn):
        rng = np.random.default_rng([seed, bi, round_idx])
        return rng.integers(0, tn, size=(hn, vn, 2), dtype=np.int64).astype(np.int32)
    return fn


def default_selection_fn(seed):
    """CPU stand-in for ``zeros(mask.shape).uniform_(0, 1)`` (:538)."""
    def fn(bi, h, w):
        rng = np.random.default_rng([seed, bi, 12345])
        return rng.random((h, w), dtype=np.float32)
    return fn


def compact(cur_mask, vertex_bi, max_num, selection_fn, bi):
    """Foreground compaction, ransac_voting_gpu.py:527-546.
    cur_mask bool [h,w]; vertex_bi [h,w,vn,2].  -> (fg_before, coords [tn,2] (x,y), direct [tn,vn,2])."""
    cur_mask = cur_mask.copy()
    fg = int(cur_mask.sum())
    if fg > max_num:
        h, w = cur_mask.shape
        selection = np.asarray(selection_fn(bi, h, w), np.float32)
        ratio = np.float32(max_num) / np.float32(fg)       # python int / float32 tensor
        cur_mask &= selection < ratio
    ys, xs = np.nonzero(cur_mask)                          # row-major, like torch.nonzero
    coords = np.stack([xs, ys], 1).astype(np.float32)      # coords[:, [1, 0]]
    direct = np.ascontiguousarray(vertex_bi[cur_mask], np.float32)   # [tn,vn,2]
    return fg, coords, direct


def _refine(direct, coords, win_pts, thresh, dtype=np.float64):
    """ransac_voting_gpu.py:578-595.  -> (pts [vn,2], all_inlier [vn,tn] uint8, normal, b)."""
    tn, vn, _ = direct.shape
    all_inlier = np.zeros((1, vn, tn), np.uint8)
    voting_for_hypothesis(direct, coords, win_pts[None].astype(np.float32), all_inlier, thresh)
    inl = all_inlier[0].astype(dtype)                      # [vn,tn]
    normal = np.zeros((tn, vn, 2), dtype)
    normal[:, :, 0] = direct[:, :, 1]
    normal[:, :, 1] = -direct[:, :, 0]
    normal = normal.transpose(1, 0, 2) * inl[:, :, None]   # [vn,tn,2]
    b = np.sum(normal * coords[None].astype(dtype), 2)     # [vn,tn]
    ata = np.matmul(normal.transpose(0, 2, 1), normal)     # [vn,2,2]
    atb = np.sum(normal * b[:, :, None], 1)                # [vn,2]
    pts = np.full((vn, 2), np.nan, dtype)
    for v in range(vn):
        try:                                               # b_inv = solve(ATA, I), :503-512
            inv = np.linalg.solve(ata[v], np.eye(2, dtype=dtype))
            pts[v] = inv @ atb[v]
        except np.linalg.LinAlgError:
            pass                                           # torch.solve raises on singular ATA
    return pts, all_inlier[0], normal, b


def _ransac_core(cur_mask, vertex_bi, hn, thresh, min_num, max_num, idxs_fn, selection_fn, bi):
    """Shared body of v3/v4/v5 up to the winner (ransac_voting_gpu.py:523-576).  Because
    idxs is drawn once (:547) every round re-scores the same hypotheses, so the state after
    round 1 is final (strict '<' at :567); the extra rounds are not replayed."""
    vn = vertex_bi.shape[2]
    if int(cur_mask.sum()) < min_num:
        return None
    _, coords, direct = compact(cur_mask, vertex_bi, max_num, selection_fn, bi)
    tn = coords.shape[0]
    idxs = np.asarray(idxs_fn(bi, 0, hn, vn, tn), np.int32)
    hyp = generate_hypothesis(direct, coords, idxs)
    counts = vote_counts(direct, coords, hyp, thresh)      # [hn,vn]
    win_idx = np.argmax(counts, 0)                         # first max
    win_counts = counts[win_idx, np.arange(vn)]
    cur_ratio = win_counts.astype(np.float32) / np.float32(tn)
    all_ratio = np.zeros(vn, np.float32)
    all_pts = np.zeros((vn, 2), np.float32)
    larger = all_ratio < cur_ratio
    all_pts[larger] = hyp[win_idx, np.arange(vn)][larger]
    all_ratio[larger] = cur_ratio[larger]
    return coords, direct, tn, all_pts, all_ratio, hyp, counts


def ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99,
                           max_iter=20, min_num=5, max_num=30000, idxs_fn=None, selection_fn=None,
                           dtype=np.float64):
    b, h, w, vn, _ = vertex.shape
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    out = np.zeros((b, vn, 2), np.float32)
    for bi in range(b):
        cur_mask = mask[bi].astype(np.uint8) != 0          # .byte()
        core = _ransac_core(cur_mask, vertex[bi], round_hyp_num, inlier_thresh, min_num, max_num,
                            idxs_fn, selection_fn, bi)
        if core is None:
            continue
        coords, direct, tn, all_pts, _, _, _ = core
        pts, _, _, _ = _refine(direct, coords, all_pts, inlier_thresh, dtype)
        out[bi] = pts
    return out


def ransac_voting_layer_v4(mask, vertex, round_hyp_num, inlier_thresh=0.99, confidence=0.999,
                           max_iter=20, min_num=5, max_num=30000, idxs_fn=None, selection_fn=None,
                           dtype=np.float64):
    b, h, w, vn, _ = vertex.shape
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    out = np.zeros((b, vn, 2), np.float32)
    var = np.ones((b, vn), np.float32)
    for bi in range(b):
        cur_mask = mask[bi].astype(np.uint8) != 0
        core = _ransac_core(cur_mask, vertex[bi], round_hyp_num, inlier_thresh, min_num, max_num,
                            idxs_fn, selection_fn, bi)
        if core is None:
            continue
        coords, direct, tn, all_pts, _, _, _ = core
        pts, inl, normal, bvec = _refine(direct, coords, all_pts, inlier_thresh, dtype)
        residual = np.matmul(normal, pts[:, :, None])[:, :, 0] - bvec    # :752
        with np.errstate(divide="ignore", invalid="ignore"):
            var[bi] = np.sum(residual ** 2, 1) / np.sum(inl.astype(dtype), 1)
        out[bi] = pts
    return out, var


def ransac_voting_layer_v5(mask, vertex, round_hyp_num, inlier_thresh=0.999, confidence=0.99,
                           max_iter=20, min_num=5, max_num=100, idxs_fn=None, selection_fn=None,
                           dtype=np.float64):
    b, h, w, vn, _ = vertex.shape
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    out = np.zeros((b, vn, 2), np.float32)
    conf = np.zeros((b, vn), np.float32)
    for bi in range(b):
        cur_mask = mask[bi].astype(np.uint8) != 0
        core = _ransac_core(cur_mask, vertex[bi], round_hyp_num, inlier_thresh, min_num, max_num,
                            idxs_fn, selection_fn, bi)
        if core is None:
            continue
        coords, direct, tn, all_pts, _, _, _ = core
        pts, _, _, _ = _refine(direct, coords, all_pts, inlier_thresh, dtype)
        out[bi] = pts
        cnt = vote_counts(direct, coords, out[bi][None], 0.999)          # :848-850, literal 0.999
        conf[bi] = cnt[0].astype(np.float32) / np.float32(tn)
    return out, conf


def ransac_voting_hypothesis(mask, vertex, round_hyp_num, inlier_thresh=0.999, min_num=5,
                             max_num=30000, idxs_fn=None, selection_fn=None):
    """:218-261 -> (hyp [b,hn,vn,2], counts [b,hn,vn])."""
    b, h, w, vn, _ = vertex.shape
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    hyps = np.zeros((b, round_hyp_num, vn, 2), np.float32)
    counts = np.ones((b, round_hyp_num, vn), np.int64)
    for bi in range(b):
        cur_mask = mask[bi] == 1
        if int(cur_mask.sum()) < min_num:
            continue
        _, coords, direct = compact(cur_mask, vertex[bi], max_num, selection_fn, bi)
        tn = coords.shape[0]
        idxs = np.asarray(idxs_fn(bi, 0, round_hyp_num, vn, tn), np.int32)
        hyps[bi] = generate_hypothesis(direct, coords, idxs)
        counts[bi] = vote_counts(direct, coords, hyps[bi], inlier_thresh)
    return hyps, counts


def _multiclass(mask, vertex, class_num, round_hyp_num, inlier_thresh, min_num, max_num, idxs_fn, selection_fn,
                refine_iter_num):
    """Common body of ransac_voting_layer (:10-97) and _v2 (:99-216).  The random draws are indexed by
    the (image, class) pair in loop order: pair index = bi * (class_num - 1) + k."""
    b, h, w, vn, _ = vertex.shape
    out = np.zeros((b, class_num - 1, vn, 2), np.float32)
    for bi in range(b):
        for k in range(class_num - 1):
            pair = bi * (class_num - 1) + k
            cur_mask = mask[bi] == k + 1
            if int(cur_mask.sum()) < min_num:
                continue
            _, coords, direct = compact(cur_mask, vertex[bi], max_num, selection_fn, pair)
            tn = coords.shape[0]
            idxs = np.asarray(idxs_fn(pair, 0, round_hyp_num, vn, tn), np.int32)
            hyp = generate_hypothesis(direct, coords, idxs)
            counts = vote_counts(direct, coords, hyp, inlier_thresh)
            win = np.argmax(counts, 0)                                   # first maximum (:69)
            ratio = counts[win, np.arange(vn)].astype(np.float32) / np.float32(tn)
            pts = np.where((ratio > 0)[:, None], hyp[win, np.arange(vn)], 0).astype(np.float32)
            normal = np.stack([direct[:, :, 1], -direct[:, :, 0]], 2)
            for _ in range(refine_iter_num if refine_iter_num is not None else 0):
                inl = np.zeros((1, vn, tn), np.uint8)
                voting_for_hypothesis(direct, coords, pts[None], inl, inlier_thresh)
                new = np.zeros((vn, 2), np.float32)
                for vi in range(vn):
                    sel = inl[0, vi] != 0
                    if not sel.any():
                        continue                                         # zeros (:193-195)
                    a = normal[sel, vi, :].astype(np.float64)
                    rhs = np.sum(a * coords[sel].astype(np.float64), 1)
                    new[vi] = (np.linalg.pinv(a) @ rhs).astype(np.float32)    # :200
                pts = new
            out[bi, k] = pts
    return out


def ransac_voting_layer(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                        min_num=5, max_num=30000, idxs_fn=None, selection_fn=None):
    """:10-97 -> [b, class_num-1, vn, 2]."""
    return _multiclass(mask, vertex, class_num, round_hyp_num, inlier_thresh, min_num, max_num,
                       idxs_fn or default_idxs_fn(0), selection_fn or default_selection_fn(0), None)


def ransac_voting_layer_v2(mask, vertex, class_num, round_hyp_num, inlier_thresh=0.999, confidence=0.99, max_iter=20,
                           min_num=5, max_num=30000, refine_iter_num=1, idxs_fn=None, selection_fn=None):
    """:99-216 -> [b, class_num-1, vn, 2]."""
    return _multiclass(mask, vertex, class_num, round_hyp_num, inlier_thresh, min_num, max_num,
                       idxs_fn or default_idxs_fn(0), selection_fn or default_selection_fn(0), refine_iter_num)


def ransac_motion_voting(mask, vertex):
    """:960-981 -> [b,vn,2]: mean over the foreground of vertex + (x, y); zeros for an empty mask."""
    b, h, w, vn, _ = vertex.shape
    out = np.zeros((b, vn, 2), np.float32)
    for bi in range(b):
        cur = mask[bi].astype(np.uint8) != 0
        ys, xs = np.nonzero(cur)
        if ys.shape[0] < 1:
            continue
        coords = np.stack([xs, ys], 1).astype(np.float32)
        out[bi] = np.mean(vertex[bi][cur] + coords[:, None, :], 0)
    return out


def _distribution_inputs(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh, min_num, max_num,
                         idxs_fn, selection_fn, degenerate_hn):
    b, h, w, vn, _ = vertex.shape
    round_num = int(np.ceil(min_hyp_num / round_hyp_num))
    hn_total = round_num * round_hyp_num
    # A4: the reference's degenerate branch emits `degenerate_hn` rows and torch.cat then
    # fails unless that equals hn_total; we define the degenerate output at the full count.
    all_hyp = np.zeros((b, hn_total, vn, 2), np.float32)
    all_ratio = np.ones((b, hn_total, vn), np.float32)
    for bi in range(b):
        cur_mask = mask[bi] == 1
        if int(cur_mask.sum()) < min_num:
            continue
        _, coords, direct = compact(cur_mask, vertex[bi], max_num, selection_fn, bi)
        tn = coords.shape[0]                                # == foreground after subsample (:287)
        for r in range(round_num):
            idxs = np.asarray(idxs_fn(bi, r, round_hyp_num, vn, tn), np.int32)   # fresh per round (:297)
            hyp = generate_hypothesis(direct, coords, idxs)
            cnt = vote_counts(direct, coords, hyp, inlier_thresh)
            sl = slice(r * round_hyp_num, (r + 1) * round_hyp_num)
            all_hyp[bi, sl] = hyp
            all_ratio[bi, sl] = cnt.astype(np.float32) / np.float32(tn)
    return all_hyp.transpose(0, 2, 1, 3), all_ratio.transpose(0, 2, 1)   # [b,vn,hn,2], [b,vn,hn]


def topk_mask_lowest_index(ratio, topk):
    """Membership of torch.topk(..., sorted=False) is unspecified among values tied at the
    k-th place; the oracle (and the CUDA kernel) keep the lowest hypothesis indices."""
    order = np.argsort(-ratio, axis=-1, kind="stable")
    keep = np.zeros(ratio.shape, bool)
    np.put_along_axis(keep, order[..., :topk], True, axis=-1)
    return keep


def estimate_voting_distribution(mask, vertex, round_hyp_num=256, min_hyp_num=4096, topk=128,
                                 inlier_thresh=0.99, min_num=5, max_num=30000, idxs_fn=None,
                                 selection_fn=None):
    """:263-331 -> (mean [b,vn,2], cov [b,vn,2,2])."""
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    hyp, ratio = _distribution_inputs(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh,
                                      min_num, max_num, idxs_fn, selection_fn, round_hyp_num)
    keep = topk_mask_lowest_index(ratio, topk)
    wgt = np.where(keep, ratio, 0).astype(np.float64)
    hyp = hyp.astype(np.float64)
    wsum = wgt.sum(2)
    mean = (wgt[..., None] * hyp).sum(2) / wsum[..., None]
    diff = hyp - mean[:, :, None]
    cov = np.matmul(diff.transpose(0, 1, 3, 2), diff * wgt[..., None]) / wsum[..., None, None]
    return mean.astype(np.float32), cov.astype(np.float32)


def estimate_voting_distribution_with_mean(mask, vertex, mean, round_hyp_num=256, min_hyp_num=4096,
                                           topk=128, inlier_thresh=0.99, min_num=5, max_num=30000,
                                           idxs_fn=None, selection_fn=None):
    """:333-406: keep ratio >= max-0.1, +1e-3 in the denominator."""
    idxs_fn = idxs_fn or default_idxs_fn(0)
    selection_fn = selection_fn or default_selection_fn(0)
    hyp, ratio = _distribution_inputs(mask, vertex, round_hyp_num, min_hyp_num, inlier_thresh,
                                      min_num, max_num, idxs_fn, selection_fn, min_hyp_num)
    thresh = ratio.max(2) - np.float32(0.1)
    wgt = np.where(ratio < thresh[..., None], 0, ratio).astype(np.float64)
    hyp = hyp.astype(np.float64)
    diff = hyp - np.asarray(mean, np.float64)[:, :, None]
    cov = np.matmul(diff.transpose(0, 1, 3, 2), diff * wgt[..., None])
    cov /= (wgt.sum(2)[..., None, None] + 1e-3)
    return np.asarray(mean, np.float32), cov.astype(np.float32)
