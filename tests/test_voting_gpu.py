"""GPU parity for csrc/voting.cu through the C ABI:
  * our kernels and the C restatement against the reference's own kernels compiled verbatim for
    sm_100a (oracle/_ref/libref_voting.so): hypothesis points and inlier bytes bit-exact;
  * the fused batched driver against the oracle driver on identical indices: counts bit-exact,
    refined keypoints within 1e-3 px;
  * the in-kernel Philox stream against torch's own random_/uniform_ draws.
"""
import ctypes

import numpy as np
import pytest
import torch

from oracle import _lib as olib
from oracle import voting as ov
from tests.synth import make_vertex_field, vertex_hwvn2

pytestmark = pytest.mark.gpu


def _compact_case(seed, h, w, vn, frac, kind="structured"):
    mask, vertex, kpts = make_vertex_field(seed, 1, h, w, vn, frac, kind)
    vx = vertex_hwvn2(vertex)
    _, coords, direct = ov.compact(mask[0] != 0, vx[0], 10 ** 9, ov.default_selection_fn(0), 0)
    return coords, direct


def _ref_lib():
    lib = olib.ref_voting_lib(required=False)
    if lib is None:
        pytest.skip("oracle/_ref/libref_voting.so not built")
    return lib


def _dp(t):
    return ctypes.c_void_p(t.data_ptr())


@pytest.mark.parametrize("kind,seed", [("structured", 1), ("randinit", 2), ("structured", 3)])
def test_primitives_bit_exact_vs_reference_kernels(cuda_dev, kind, seed):
    from esa_pose_estimation_b200 import ransac_voting
    ref = _ref_lib()
    coords, direct = _compact_case(seed, 48, 56, 5, 0.6, kind)
    tn, vn, _ = direct.shape
    hn = 96
    rng = np.random.default_rng(seed)
    idxs = rng.integers(0, tn, (hn, vn, 2)).astype(np.int32)
    idxs[0, :, 1] = idxs[0, :, 0]                                      # same pixel twice: parallel rays
    d_t, c_t, i_t = (torch.from_numpy(a).to(cuda_dev) for a in (direct, coords, idxs))
    # reference kernels (device oracle)
    hyp_ref = torch.zeros((hn, vn, 2), device=cuda_dev)
    assert ref.ref_generate_hypothesis(_dp(d_t), _dp(c_t), _dp(i_t), _dp(hyp_ref), tn, vn, hn) == 0
    torch.cuda.synchronize()
    hyp = ransac_voting.generate_hypothesis(d_t, c_t, i_t)
    hyp_c = ov.generate_hypothesis(direct, coords, idxs)
    assert torch.equal(hyp.view(torch.int32), hyp_ref.view(torch.int32))          # ours == reference, bitwise
    np.testing.assert_array_equal(hyp_c.view(np.int32), hyp_ref.cpu().numpy().view(np.int32))  # C oracle too
    for thresh in (0.999, 0.99, 0.5):
        inl_ref = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
        assert ref.ref_voting_for_hypothesis(_dp(d_t), _dp(c_t), _dp(hyp_ref), _dp(inl_ref), tn, vn, hn,
                                             ctypes.c_float(thresh)) == 0
        torch.cuda.synchronize()
        inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
        ransac_voting.voting_for_hypothesis(d_t, c_t, hyp_ref, inl, thresh)
        assert torch.equal(inl, inl_ref)
        inl_c = np.zeros((hn, vn, tn), np.uint8)
        ov.voting_for_hypothesis(direct, coords, hyp_ref.cpu().numpy(), inl_c, thresh)
        np.testing.assert_array_equal(inl_c, inl_ref.cpu().numpy())


def test_pybind_level_checks(cuda_dev):
    from esa_pose_estimation_b200 import ransac_voting
    d = torch.zeros((8, 2, 2), device=cuda_dev)
    c = torch.zeros((8, 2), device=cuda_dev)
    i = torch.zeros((4, 2, 2), dtype=torch.int32, device=cuda_dev)
    with pytest.raises(RuntimeError):
        ransac_voting.generate_hypothesis(d.cpu(), c, i)               # CHECK_CUDA
    with pytest.raises(RuntimeError):
        ransac_voting.generate_hypothesis(d.transpose(0, 1), c, i)     # CHECK_CONTIGUOUS
    hyp = ransac_voting.generate_hypothesis(d, c, i)
    assert hyp.shape == (4, 2, 2) and (hyp == 0).all()                 # degenerate pairs stay zero


def _idxs_for(mask, vx, hn, rounds, max_num, seed, eq1=False):
    """Per-image indices drawn exactly as the oracle's idxs_fn will see them."""
    b, h, w, vn, _ = vx.shape
    fn = ov.default_idxs_fn(seed)
    sel = ov.default_selection_fn(seed)
    out = np.zeros((b, rounds, hn, vn, 2), np.int32)
    selection = np.zeros((b, h, w), np.float32)
    for bi in range(b):
        cur = (mask[bi] == 1) if eq1 else (mask[bi] != 0)
        selection[bi] = sel(bi, h, w)
        if cur.sum() < 5:
            continue
        _, coords, _ = ov.compact(cur, vx[bi], max_num, sel, bi)
        for r in range(rounds):
            out[bi, r] = fn(bi, r, hn, vn, coords.shape[0])
    return out, selection, fn, sel


@pytest.mark.parametrize("layout", ["nchw_view", "contiguous"])
@pytest.mark.parametrize("cfg", [dict(b=3, h=64, w=72, vn=4, hn=128, frac=0.3, max_num=30000),
                                 dict(b=2, h=96, w=96, vn=11, hn=512, frac=0.5, max_num=30000),
                                 dict(b=2, h=80, w=64, vn=3, hn=64, frac=1.0, max_num=1500),
                                 dict(b=2, h=33, w=47, vn=2, hn=300, frac=0.4, max_num=30000)])
def test_layer_v3_v4_v5_match_oracle(cuda_dev, cfg, layout):
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, h, w, vn, hn = cfg["b"], cfg["h"], cfg["w"], cfg["vn"], cfg["hn"]
    mask, vertex, kpts = make_vertex_field(11 + hn, b, h, w, vn, cfg["frac"])
    mask[-1, :, : w // 2] = 0                                            # ragged foreground
    vx = vertex_hwvn2(vertex)
    idxs, selection, fn, sel = _idxs_for(mask, vx, hn, 1, cfg["max_num"], 5)
    v_t = torch.from_numpy(vertex).to(cuda_dev)
    vert = rv.vertex_layer_reshape(v_t) if layout == "nchw_view" else torch.from_numpy(vx).to(cuda_dev)
    m_t = torch.from_numpy(mask).to(cuda_dev)
    kw = dict(idxs=torch.from_numpy(idxs).to(cuda_dev), selection=torch.from_numpy(selection).to(cuda_dev))
    # fused driver, everything it computes
    dbg = rv.voting_debug(_lib.VOTE_V3, m_t, vert, hn, inlier_thresh=0.999, max_num=cfg["max_num"], **kw)
    hyp_o, cnt_o = ov.ransac_voting_hypothesis((mask != 0).astype(np.uint8), vx, hn, 0.999, max_num=cfg["max_num"],
                                               idxs_fn=fn, selection_fn=sel)
    np.testing.assert_array_equal(dbg["hyp"].cpu().numpy().view(np.int32), hyp_o.view(np.int32))   # bit-exact
    np.testing.assert_array_equal(dbg["counts"].cpu().numpy(), cnt_o)                               # bit-exact
    p3_o = ov.ransac_voting_layer_v3(mask, vx, hn, max_num=cfg["max_num"], idxs_fn=fn, selection_fn=sel)
    p3 = rv.ransac_voting_layer_v3(m_t, vert, hn, max_num=cfg["max_num"], **kw).cpu().numpy()
    np.testing.assert_allclose(p3, p3_o, rtol=0, atol=1e-3)
    p4_o, var_o = ov.ransac_voting_layer_v4(mask, vx, hn, max_num=cfg["max_num"], idxs_fn=fn, selection_fn=sel)
    p4, var = rv.ransac_voting_layer_v4(m_t, vert, hn, max_num=cfg["max_num"], **kw)
    np.testing.assert_allclose(p4.cpu().numpy(), p4_o, rtol=0, atol=1e-3)
    np.testing.assert_allclose(var.cpu().numpy(), var_o, rtol=1e-3, atol=1e-6)
    mx5 = min(cfg["max_num"], 400)
    idxs5, selection5, fn5, sel5 = _idxs_for(mask, vx, hn, 1, mx5, 6)
    kw5 = dict(idxs=torch.from_numpy(idxs5).to(cuda_dev), selection=torch.from_numpy(selection5).to(cuda_dev))
    p5_o, conf_o = ov.ransac_voting_layer_v5(mask, vx, hn, max_num=mx5, idxs_fn=fn5, selection_fn=sel5)
    p5, conf = rv.ransac_voting_layer_v5(m_t, vert, hn, max_num=mx5, **kw5)
    np.testing.assert_allclose(p5.cpu().numpy(), p5_o, rtol=0, atol=1e-3)
    np.testing.assert_allclose(conf.cpu().numpy(), conf_o, rtol=0, atol=1e-6)


def test_degenerate_and_thresholds(cuda_dev):
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    mask, vertex, _ = make_vertex_field(21, 3, 40, 40, 3, 0.4, "randinit")
    mask[1] = 0
    mask[1, 0, :3] = 1                                                   # < min_num
    mask[2] = 0                                                          # empty
    vx = vertex_hwvn2(vertex)
    idxs, selection, fn, sel = _idxs_for(mask, vx, 64, 1, 30000, 9)
    vert = torch.from_numpy(vx).to(cuda_dev)
    m_t = torch.from_numpy(mask).to(cuda_dev)
    kw = dict(idxs=torch.from_numpy(idxs).to(cuda_dev))
    p3 = rv.ransac_voting_layer_v3(m_t, vert, 64, **kw)
    assert (p3[1:] == 0).all()
    _, var = rv.ransac_voting_layer_v4(m_t, vert, 64, **kw)
    assert (var[1:] == 1).all()
    _, conf = rv.ransac_voting_layer_v5(m_t, vert, 64, max_num=30000, **kw)
    assert (conf[1:] == 0).all()
    hyp, cnt = rv.ransac_voting_hypothesis(m_t, vert, 64, **kw)
    assert (hyp[1:] == 0).all() and (cnt[1:] == 1).all() and cnt.dtype == torch.int64
    # thresholds outside the fast-filter range take the exact path: still bit-exact counts
    for thr in (0.0, -0.5, 1e-4, 0.2, 0.999999):
        _, c = rv.ransac_voting_hypothesis(m_t, vert, 64, inlier_thresh=thr, **kw)
        _, c_o = ov.ransac_voting_hypothesis(mask, vx, 64, thr, idxs_fn=fn, selection_fn=sel)
        np.testing.assert_array_equal(c.cpu().numpy(), c_o)


def test_workspace_bound_drops_an_overfull_image_instead_of_overrunning(cuda_dev):
    """The workspace holds max_num + 8 sigma + 1024 voting pixels per image when the subsample draws come from Philox
    (DESIGN.md section 3).  Caller-supplied `selection` draws can defeat that bound (all zeros keep every pixel): the image
    must come back flagged (NaN keypoints, status 3) and its neighbours must be untouched -- never an out-of-bounds write."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, s, vn, hn, max_num = 3, 96, 3, 64, 100
    mask, vertex, _ = make_vertex_field(33, b, s, s, vn, 0.9)            # ~8300 foreground pixels > max_num: subsample applies
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    m_t = torch.from_numpy(mask).to(cuda_dev)
    sel = torch.rand((b, s, s), device=cuda_dev)                          # ordinary draws: ~100 pixels kept per image
    sel[1] = 0.0                                                          # image 1 keeps all ~8300 > cap (1280)
    torch.manual_seed(5)
    out = rv.voting_debug(_lib.VOTE_V3, m_t, vert, hn, max_num=max_num, selection=sel)
    st, pts, tn = out["status"].cpu().numpy(), out["pts"].cpu().numpy(), out["tn"].cpu().numpy()
    assert (st[1] == 3).all() and np.isnan(pts[1]).all() and tn[1] == 0
    assert (st[[0, 2]] == 0).all() and np.isfinite(pts[[0, 2]]).all() and (tn[[0, 2]] > 40).all() and (tn[[0, 2]] < 200).all()
    # the same call without the overfull image gives the same keypoints for the others (same draws, same generator offsets)
    torch.manual_seed(5)
    sel2 = sel.clone(); sel2[1] = 1.0                                     # image 1 keeps nothing -> fewer than min_num -> skipped
    out2 = rv.voting_debug(_lib.VOTE_V3, m_t, vert, hn, max_num=max_num, selection=sel2)
    np.testing.assert_array_equal(out2["pts"].cpu().numpy()[[0, 2]], pts[[0, 2]])


def test_workspace_is_not_overrun(cuda_dev):
    """compute-sanitizer is closed on this pool: a caller-owned workspace of exactly the advertised size, followed by a
    guard region, must come back with the guard intact for ragged batches, the subsample and every driver mode."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, s, vn, hn = 5, 72, 4, 96
    mask, vertex, _ = make_vertex_field(41, b, s, s, vn, 0.6)
    mask[1, 10:, :] = 0                                                  # ragged: few pixels
    mask[3] = 1                                                          # everything foreground
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    m_t = torch.from_numpy(mask).to(cuda_dev)
    guard = 1 << 16
    for mode, kw in ((_lib.VOTE_V3, dict(max_num=30000)), (_lib.VOTE_V5, dict(max_num=100)), (_lib.VOTE_V3, dict(max_num=900)),
                     (_lib.VOTE_DISTRIBUTION, dict(max_num=30000, rounds=3, topk=32))):
        need = rv.workspace_bytes(b, s, s, vn, hn, rounds=kw.get("rounds", 1), max_num=kw["max_num"])
        ws = torch.full((need + guard,), 0xA5, dtype=torch.uint8, device=cuda_dev)
        torch.manual_seed(3)
        out = rv._run(mode, m_t, vert, hn, kw.get("rounds", 1), 0.99, 5, kw["max_num"], topk=kw.get("topk", 0), workspace=ws)
        torch.cuda.synchronize()
        assert bool((ws[need:] == 0xA5).all()), (mode, kw)
        assert all(bool(torch.isfinite(v.float()).all()) for k, v in out.items() if k in ("pts", "mean", "cov"))


def test_distribution_matches_oracle(cuda_dev):
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    b, h, w, vn = 2, 64, 64, 5
    mask, vertex, kpts = make_vertex_field(31, b, h, w, vn, 0.5, noise_deg=1.5)
    vx = vertex_hwvn2(vertex)
    rounds, hn = 4, 64
    idxs, selection, fn, sel = _idxs_for(mask, vx, hn, rounds, 30000, 12, eq1=True)
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    m_t = torch.from_numpy(mask).to(cuda_dev)
    kw = dict(idxs=torch.from_numpy(idxs).to(cuda_dev))
    mean, cov = rv.estimate_voting_distribution(m_t, vert, round_hyp_num=hn, min_hyp_num=hn * rounds, topk=32, **kw)
    mean_o, cov_o = ov.estimate_voting_distribution(mask, vx, hn, hn * rounds, 32, idxs_fn=fn, selection_fn=sel)
    np.testing.assert_allclose(mean.cpu().numpy(), mean_o, rtol=0, atol=1e-3)
    np.testing.assert_allclose(cov.cpu().numpy(), cov_o, rtol=1e-4, atol=1e-4)
    m2, cov2 = rv.estimate_voting_distribution_with_mean(m_t, vert, mean, round_hyp_num=hn, min_hyp_num=hn * rounds, **kw)
    _, cov2_o = ov.estimate_voting_distribution_with_mean(mask, vx, mean_o, hn, hn * rounds, idxs_fn=fn, selection_fn=sel)
    np.testing.assert_allclose(cov2.cpu().numpy(), cov2_o, rtol=1e-4, atol=1e-4)


def test_philox_matches_torch_generator(cuda_dev):
    """Hypothesis indices regenerated in-kernel equal torch's random_(0, tn) under the same seed, in
    the reference's per-image call order, including the uniform_ subsample draw; and the generator
    ends where the reference's would."""
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    b, h, w, vn, hn = 3, 48, 40, 4, 96
    mask, vertex, _ = make_vertex_field(41, b, h, w, vn, 0.8, "randinit")
    mask[1] = 0                                                          # skipped image consumes nothing
    vx = vertex_hwvn2(vertex)
    max_num = 700                                                        # forces the subsample on image 0 and 2
    m_t = torch.from_numpy(mask).to(cuda_dev)
    vert = torch.from_numpy(vx).to(cuda_dev)
    # reference call order with torch itself (ransac_voting_gpu.py:527-547)
    torch.manual_seed(1234)
    idxs_ref = np.zeros((b, 1, hn, vn, 2), np.int32)
    sel_ref = np.zeros((b, h, w), np.float32)
    for bi in range(b):
        cur = m_t[bi].byte()
        fg = torch.sum(cur)
        if fg < 5:
            continue
        if fg > max_num:
            selection = torch.zeros(cur.shape, dtype=torch.float32, device=cuda_dev).uniform_(0, 1)
            sel_ref[bi] = selection.cpu().numpy()
            cur = cur * (selection < (max_num / fg.float()))
        tn = int(torch.nonzero(cur).shape[0])
        idxs_ref[bi, 0] = torch.zeros([hn, vn, 2], dtype=torch.int32, device=cuda_dev).random_(0, tn).cpu().numpy()
    end_offset = torch.cuda.default_generators[0].get_offset()
    after_ref = torch.rand(4, device=cuda_dev).cpu()
    # explicit-index run
    hyp_a, cnt_a = rv.ransac_voting_hypothesis(m_t, vert, hn, max_num=max_num,
                                               idxs=torch.from_numpy(idxs_ref).to(cuda_dev),
                                               selection=torch.from_numpy(sel_ref).to(cuda_dev))
    # Philox run under the same seed
    torch.manual_seed(1234)
    hyp_b, cnt_b = rv.ransac_voting_hypothesis(m_t, vert, hn, max_num=max_num)
    assert torch.cuda.default_generators[0].get_offset() == end_offset
    after = torch.rand(4, device=cuda_dev).cpu()
    assert torch.equal(hyp_a.view(torch.int32), hyp_b.view(torch.int32))
    assert torch.equal(cnt_a, cnt_b)
    assert torch.equal(after, after_ref)


def test_full_size_properties(cuda_dev):
    """BASELINE config 2 size (64 x 256x256, vn=11, hn=512): properties instead of a full oracle pass."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, h, w, vn, hn = 64, 256, 256, 11, 512
    mask, vertex, kpts = make_vertex_field(51, 4, h, w, vn, 0.25, noise_deg=2.0)
    mask = np.tile(mask, (16, 1, 1)); vertex = np.tile(vertex, (16, 1, 1, 1)); kpts = np.tile(kpts, (16, 1, 1))
    m_t = torch.from_numpy(mask).to(cuda_dev)
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    torch.manual_seed(7)
    dbg = rv.voting_debug(_lib.VOTE_V3, m_t, vert, hn)
    tn = dbg["tn"].cpu().numpy()
    np.testing.assert_array_equal(tn, mask.reshape(b, -1).sum(1))
    cnt = dbg["counts"].cpu().numpy()
    assert cnt.min() >= 0 and (cnt.max(axis=(1, 2)) <= tn).all()
    pts = dbg["pts"].cpu().numpy()
    assert np.abs(pts - kpts).max() < 1.5                               # recovers the planted keypoints (2 deg noise)
    # five images checked exactly against the oracle, every one of their 512 x 11 counts (92 M pair tests each),
    # using the counts' own hypotheses (drawn by the in-kernel Philox stream)
    for bi in (0, 1, 18, 35, 63):
        vxb = vertex_hwvn2(vertex[bi:bi + 1])
        _, coords, direct = ov.compact(mask[bi] != 0, vxb[0], 30000, ov.default_selection_fn(0), 0)
        hyp_b = dbg["hyp"][bi].cpu().numpy()
        np.testing.assert_array_equal(ov.vote_counts(direct, coords, hyp_b, 0.999), cnt[bi])
    # same seed -> identical result (determinism, no atomics races in the count)
    torch.manual_seed(7)
    dbg2 = rv.voting_debug(_lib.VOTE_V3, m_t, vert, hn)
    assert torch.equal(dbg2["counts"], dbg["counts"]) and torch.equal(dbg2["pts"].view(torch.int32), dbg["pts"].view(torch.int32))


@pytest.mark.parametrize("thresh", [0.999, 0.99, 0.9, 0.6])
def test_counts_exact_on_the_cone_boundary(cuda_dev, thresh):
    """Adversarial field for vote_count's division-free fast test: every hypothesis is (nearly) the
    same point P, and every pixel's direction is the direction towards P rotated by the cone
    half-angle acos(T) times (1 + eps), eps from 0 to +-1e-3 -- i.e. the cosine the reference
    computes sits within rounding error of the threshold.  Counts must still be the reference's."""
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    h, w, vn, hn = 64, 80, 3, 256
    rng = np.random.default_rng(int(thresh * 1e4))
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    mask = np.ones((1, h, w), np.uint8)
    vertex = np.zeros((1, 2 * vn, h, w), np.float32)
    theta = np.arccos(np.float64(np.float32(thresh)))
    anchors = rng.choice(h * w, 24, replace=False)                       # pixels whose rays meet in P
    for v in range(vn):
        px, py = [(31.37, 29.61), (-40.5, 17.25), (200.123, 90.77)][v]
        ang = np.arctan2(py - ys, px - xs)
        eps = rng.choice([0.0, 1e-8, -1e-8, 1e-7, -1e-7, 3e-7, -3e-7, 1e-6, -1e-6, 1e-5, -1e-5, 1e-4, -1e-4,
                          1e-3, -1e-3], size=(h, w))
        rot = ang + rng.choice([-1.0, 1.0], size=(h, w)) * theta * (1.0 + eps)
        scale = rng.choice([1.0, 0.37, 2.5, 1e-3, 40.0], size=(h, w))      # |n| != 1 must not matter
        rot.reshape(-1)[anchors] = ang.reshape(-1)[anchors]
        vertex[0, 2 * v] = (np.cos(rot) * scale).astype(np.float32)
        vertex[0, 2 * v + 1] = (np.sin(rot) * scale).astype(np.float32)
    vx = vertex_hwvn2(vertex)
    idxs = anchors[rng.integers(0, len(anchors), (1, 1, hn, vn, 2))].astype(np.int32)
    idxs[0, 0, :8] = rng.integers(0, h * w, (8, vn, 2))                   # a few arbitrary hypotheses too
    fn = lambda bi, r, hn_, vn_, tn: idxs[bi, r]
    hyp, cnt = rv.ransac_voting_hypothesis(torch.from_numpy(mask).to(cuda_dev), torch.from_numpy(vx).to(cuda_dev), hn,
                                           inlier_thresh=thresh, idxs=torch.from_numpy(idxs).to(cuda_dev))
    hyp_o, cnt_o = ov.ransac_voting_hypothesis(mask, vx, hn, thresh, idxs_fn=fn, selection_fn=ov.default_selection_fn(0))
    np.testing.assert_array_equal(hyp.cpu().numpy().view(np.int32), hyp_o.view(np.int32))
    np.testing.assert_array_equal(cnt.cpu().numpy(), cnt_o)
    assert cnt_o.max() > 100 and cnt_o.min() < cnt_o.max()               # the case is not vacuous


@pytest.mark.parametrize("h,w", [(72, 88), (33, 47)])
def test_pinned_host_field_is_read_in_place(cuda_dev, h, w):
    """A float32 field in pinned HOST memory (zero-copy gather over PCIe) gives bit-identical results
    to the device-resident field; pageable host memory is rejected like the reference's CHECK_CUDA.
    72x88 takes the planar whole-line gather, 33x47 (H*W not a multiple of 16) the generic strided one."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, vn, hn = 3, 5, 128
    mask, vertex, _ = make_vertex_field(61, b, h, w, vn, 0.4)
    mask[1, :, : w // 3] = 0
    idxs, _, _, _ = _idxs_for(mask, vertex_hwvn2(vertex), hn, 1, 30000, 3)
    i_t = torch.from_numpy(idxs).to(cuda_dev)
    v_h = torch.from_numpy(vertex).pin_memory()
    m_h = torch.from_numpy(mask).pin_memory()
    dev_out = rv.voting_debug(_lib.VOTE_V4, m_h.to(cuda_dev), rv.vertex_layer_reshape(v_h.to(cuda_dev)), hn, idxs=i_t)
    host_out = rv.voting_debug(_lib.VOTE_V4, m_h, rv.vertex_layer_reshape(v_h), hn, idxs=i_t)
    for k in ("pts", "aux", "hyp", "counts", "tn"):
        assert host_out[k].is_cuda
        assert torch.equal(host_out[k].view(torch.int32), dev_out[k].view(torch.int32)), k
    with pytest.raises(RuntimeError):
        rv.ransac_voting_layer_v3(m_h.to(cuda_dev), rv.vertex_layer_reshape(torch.from_numpy(vertex)), hn, idxs=i_t)


def test_multiclass_v1_v2_match_oracle(cuda_dev):
    """ransac_voting_layer (:10) / _v2 (:99): one voting run per (image, class) pair, mask == k+1."""
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    b, h, w, vn, hn, class_num = 2, 64, 72, 4, 128, 4
    mask, vertex, _ = make_vertex_field(81, b, h, w, vn, 0.9, noise_deg=1.0)
    # three foreground classes as vertical bands of the ellipse; class 3 of image 1 nearly empty
    lab = np.zeros_like(mask)
    lab[:, :, : w // 3] = 1; lab[:, :, w // 3: 2 * w // 3] = 2; lab[:, :, 2 * w // 3:] = 3
    mask = (mask * lab).astype(np.uint8)
    mask[1][mask[1] == 3] = 0
    mask[1, 5, w - 3:] = 3                                               # 3 pixels < min_num
    vx = vertex_hwvn2(vertex)
    classes = class_num - 1
    fn = ov.default_idxs_fn(17)
    idxs = np.zeros((b * classes, 1, hn, vn, 2), np.int32)
    for bi in range(b):
        for k in range(classes):
            tn = int((mask[bi] == k + 1).sum())
            if tn >= 5:
                idxs[bi * classes + k, 0] = fn(bi * classes + k, 0, hn, vn, tn)
    m_t = torch.from_numpy(mask).to(cuda_dev)
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    i_t = torch.from_numpy(idxs).to(cuda_dev)
    p1 = rv.ransac_voting_layer(m_t, vert, class_num, hn, idxs=i_t).cpu().numpy()
    p1_o = ov.ransac_voting_layer(mask, vx, class_num, hn, idxs_fn=fn)
    assert p1.shape == (b, classes, vn, 2)
    np.testing.assert_array_equal(p1.view(np.int32), p1_o.view(np.int32))     # winners are hypotheses: bit-exact
    assert (p1[1, 2] == 0).all()
    for iters in (1, 2):
        p2 = rv.ransac_voting_layer_v2(m_t, vert, class_num, hn, refine_iter_num=iters, idxs=i_t).cpu().numpy()
        p2_o = ov.ransac_voting_layer_v2(mask, vx, class_num, hn, refine_iter_num=iters, idxs_fn=fn)
        np.testing.assert_allclose(p2, p2_o, rtol=0, atol=1e-3)


@pytest.mark.parametrize("kind,seed", [("structured", 4), ("randinit", 5)])
def test_vanishing_point_primitives_bit_exact_vs_reference_kernels(cuda_dev, kind, seed):
    """ransac_voting_kernel.cu:170-229 / :268-310 (homogeneous hypotheses): ours and the C restatement
    against the reference's own kernels compiled for sm_100a, bitwise."""
    from esa_pose_estimation_b200 import ransac_voting
    ref = _ref_lib()
    coords, direct = _compact_case(seed, 40, 48, 4, 0.7, kind)
    tn, vn, _ = direct.shape
    hn = 80
    rng = np.random.default_rng(seed)
    idxs = rng.integers(0, tn, (hn, vn, 2)).astype(np.int32)
    idxs[0, :, 1] = idxs[0, :, 0]                                      # the same ray twice
    d_t, c_t, i_t = (torch.from_numpy(a).to(cuda_dev) for a in (direct, coords, idxs))
    hyp_ref = torch.zeros((hn, vn, 3), device=cuda_dev)
    assert ref.ref_generate_hypothesis_vanishing_point(_dp(d_t), _dp(c_t), _dp(i_t), _dp(hyp_ref), tn, vn, hn) == 0
    torch.cuda.synchronize()
    hyp = ransac_voting.generate_hypothesis_vanishing_point(d_t, c_t, i_t)
    assert torch.equal(hyp.view(torch.int32), hyp_ref.view(torch.int32))
    hyp_c = ov.generate_hypothesis_vanishing_point(direct, coords, idxs)
    np.testing.assert_array_equal(hyp_c.view(np.int32), hyp_ref.cpu().numpy().view(np.int32))
    assert (hyp_ref != 0).any()
    for thresh in (0.999, 0.99, 0.5):
        inl_ref = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
        assert ref.ref_voting_for_hypothesis_vanishing_point(_dp(d_t), _dp(c_t), _dp(hyp_ref), _dp(inl_ref), tn, vn, hn,
                                                             ctypes.c_float(thresh)) == 0
        torch.cuda.synchronize()
        inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
        ransac_voting.voting_for_hypothesis_vanishing_point(d_t, c_t, hyp_ref, inl, thresh)
        assert torch.equal(inl, inl_ref)
        inl_c = np.zeros((hn, vn, tn), np.uint8)
        ov.voting_for_hypothesis_vanishing_point(direct, coords, hyp_ref.cpu().numpy(), inl_c, thresh)
        np.testing.assert_array_equal(inl_c, inl_ref.cpu().numpy())


def test_vote_primitives_slab_past_the_grid_limit(cuda_dev):
    """hn * vn > 65535 (the grid.y limit): both pybind-level vote entries process the hypotheses in slabs and must
    equal the C restatement (the vanishing-point twin used to reject such calls)."""
    from esa_pose_estimation_b200 import ransac_voting
    coords, direct = _compact_case(9, 24, 24, 3, 0.5, "structured")
    tn, vn, _ = direct.shape
    hn = 65535 // vn + 37
    rng = np.random.default_rng(9)
    idxs = rng.integers(0, tn, (hn, vn, 2)).astype(np.int32)
    d_t, c_t, i_t = (torch.from_numpy(a).to(cuda_dev) for a in (direct, coords, idxs))
    hyp = ransac_voting.generate_hypothesis(d_t, c_t, i_t)
    inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
    ransac_voting.voting_for_hypothesis(d_t, c_t, hyp, inl, 0.99)
    inl_c = np.zeros((hn, vn, tn), np.uint8)
    ov.voting_for_hypothesis(direct, coords, hyp.cpu().numpy(), inl_c, 0.99)
    np.testing.assert_array_equal(inl.cpu().numpy(), inl_c)
    hyp3 = ransac_voting.generate_hypothesis_vanishing_point(d_t, c_t, i_t)
    inl = torch.zeros((hn, vn, tn), dtype=torch.uint8, device=cuda_dev)
    ransac_voting.voting_for_hypothesis_vanishing_point(d_t, c_t, hyp3, inl, 0.99)
    inl_c = np.zeros((hn, vn, tn), np.uint8)
    ov.voting_for_hypothesis_vanishing_point(direct, coords, hyp3.cpu().numpy(), inl_c, 0.99)
    np.testing.assert_array_equal(inl.cpu().numpy(), inl_c)
    assert inl_c[-30:].any()                                            # the last slab did real work


@pytest.mark.parametrize("hn,rounds", [(1024, 1), (384, 3), (130, 1)])
def test_counts_bit_exact_for_large_hypothesis_sets(cuda_dev, hn, rounds):
    """More than 512 hypotheses per keypoint switch vote_count to 8 hypotheses per thread and several
    chunks; odd sizes exercise the partial last pair.  Counts against the C restatement, bit-exact."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    b, h, w, vn = 2, 56, 64, 3
    mask, vertex, _ = make_vertex_field(91 + hn, b, h, w, vn, 0.5, noise_deg=1.5)
    mask = (mask != 0).astype(np.uint8)
    vx = vertex_hwvn2(vertex)
    idxs, _, fn, sel = _idxs_for(mask, vx, hn, rounds, 30000, 23, eq1=True)
    dbg = rv.voting_debug(_lib.VOTE_DISTRIBUTION, torch.from_numpy(mask).to(cuda_dev),
                          rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev)), hn, rounds=rounds,
                          inlier_thresh=0.99, topk=64, idxs=torch.from_numpy(idxs).to(cuda_dev))
    cnt = dbg["counts"].cpu().numpy()
    hyp = dbg["hyp"].cpu().numpy()
    for bi in range(b):
        _, coords, direct = ov.compact(mask[bi] == 1, vx[bi], 30000, sel, bi)
        for r in range(rounds):
            sl = slice(r * hn, (r + 1) * hn)
            hyp_o = ov.generate_hypothesis(direct, coords, idxs[bi, r])
            np.testing.assert_array_equal(hyp[bi, sl].view(np.int32), hyp_o.view(np.int32))
            np.testing.assert_array_equal(cnt[bi, sl], ov.vote_counts(direct, coords, hyp_o, 0.99))


@pytest.mark.parametrize("seed", range(8))
def test_fused_driver_fuzz_against_oracle(cuda_dev, seed):
    """Random shapes (odd sizes, tiny and full foregrounds, ragged batches, subsample on/off, both field
    layouts, random thresholds): hypotheses and counts bit-exact, refined points within 1e-3 px."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    rng = np.random.default_rng(1000 + seed)
    b = int(rng.integers(1, 5)); h = int(rng.integers(9, 90)); w = int(rng.integers(9, 90))
    vn = int(rng.integers(1, 9)); hn = int(rng.choice([1, 7, 32, 100, 257, 512, 600]))
    frac = float(rng.choice([0.02, 0.2, 0.6, 1.0])); thresh = float(rng.choice([0.999, 0.99, 0.9]))
    kind = str(rng.choice(["structured", "randinit"]))
    max_num = int(rng.choice([30000, max(6, h * w // 7)]))
    mask, vertex, _ = make_vertex_field(2000 + seed, b, h, w, vn, frac, kind)
    if b > 1:
        mask[-1, : h // 2] = 0
    vx = vertex_hwvn2(vertex)
    idxs, selection, fn, sel = _idxs_for(mask, vx, hn, 1, max_num, 40 + seed)
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev)) if seed % 2 else torch.from_numpy(vx).to(cuda_dev)
    kw = dict(idxs=torch.from_numpy(idxs).to(cuda_dev), selection=torch.from_numpy(selection).to(cuda_dev))
    dbg = rv.voting_debug(_lib.VOTE_V4, torch.from_numpy(mask).to(cuda_dev), vert, hn, inlier_thresh=thresh,
                          max_num=max_num, **kw)
    hyp_o, cnt_o = ov.ransac_voting_hypothesis((mask != 0).astype(np.uint8), vx, hn, thresh, max_num=max_num,
                                               idxs_fn=fn, selection_fn=sel)
    live = np.array([(mask[i] != 0).sum() >= 5 for i in range(b)])
    np.testing.assert_array_equal(dbg["hyp"].cpu().numpy()[live].view(np.int32), hyp_o[live].view(np.int32))
    np.testing.assert_array_equal(dbg["counts"].cpu().numpy()[live], cnt_o[live])
    p4_o, var_o = ov.ransac_voting_layer_v4(mask, vx, hn, inlier_thresh=thresh, max_num=max_num, idxs_fn=fn, selection_fn=sel)
    pts = dbg["pts"].cpu().numpy()
    ok = np.isfinite(p4_o).all(axis=-1) & np.isfinite(pts).all(axis=-1)       # singular refinements: NaN on both sides
    assert (np.isfinite(p4_o).all(axis=-1) == np.isfinite(pts).all(axis=-1)).all()
    np.testing.assert_allclose(pts[ok], p4_o[ok], rtol=1e-5, atol=1e-3)


def test_motion_voting_matches_oracle(cuda_dev):
    """ransac_motion_voting (:960-981): foreground mean of vector + coordinate, zeros for an empty mask."""
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    b, h, w, vn = 3, 40, 56, 4
    mask, vertex, _ = make_vertex_field(95, b, h, w, vn, 0.3, "randinit")
    mask[1] = 0
    mask[2, :, ::2] = 0
    vx = vertex_hwvn2(vertex)
    out = rv.ransac_motion_voting(torch.from_numpy(mask).to(cuda_dev), rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev)))
    ref = ov.ransac_motion_voting(mask, vx)
    assert (out[1] == 0).all()
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=1e-5, atol=1e-3)


def test_drivers_match_the_reference_python_drivers(cuda_dev):
    """The fused CUDA drivers against tests/golden/voting_drivers.npz: outputs of the reference's own
    ransac_voting_gpu.py (unmodified, run on CPU over the C restatement of its kernels, see
    tests/golden/make_golden_voting.py) with the reference's recorded random draws passed in.
    Hypotheses / inlier counts / v1 winners / confidence bit-exact, keypoints within 1e-3 px."""
    from esa_pose_estimation_b200 import ransac_voting_gpu as rv
    from tests.golden_voting import dense_draws, load
    g = load()
    hn = int(g["a_hn"])
    m_t = torch.from_numpy(g["a_mask"]).to(cuda_dev)
    vert = torch.from_numpy(g["a_vertex"]).to(cuda_dev)
    live = [True, True, False]                                   # image 2 has 3 foreground pixels < min_num

    def kw(name, live=live, rounds=1):
        idxs, sel = dense_draws(g, name, live, rounds)
        d = dict(idxs=torch.from_numpy(idxs).to(cuda_dev))
        if sel is not None:
            d["selection"] = torch.from_numpy(sel).to(cuda_dev)
        return d

    def close(a, ref, atol=1e-3, rtol=0.0):
        np.testing.assert_allclose(a.cpu().numpy(), ref, rtol=rtol, atol=atol)

    close(rv.ransac_voting_layer_v3(m_t, vert, hn, **kw("v3")), g["v3_pts"])
    close(rv.ransac_voting_layer_v3(m_t, vert, hn, max_num=150, **kw("v3sub")), g["v3sub_pts"])
    close(rv.ransac_voting_layer_v3(m_t, vert, hn, inlier_thresh=0.99, **kw("v3t")), g["v3t_pts"])
    p, var = rv.ransac_voting_layer_v4(m_t, vert, hn, **kw("v4"))
    close(p, g["v4_pts"])
    close(var, g["v4_var"], atol=1e-6, rtol=2e-3)
    p, conf = rv.ransac_voting_layer_v5(m_t, vert, hn, **kw("v5"))
    close(p, g["v5_pts"])
    np.testing.assert_array_equal(conf.cpu().numpy(), g["v5_conf"])
    hyp, cnt = rv.ransac_voting_hypothesis(m_t, vert, hn, **kw("hyp"))
    np.testing.assert_array_equal(hyp.cpu().numpy().view(np.uint32), g["hyp_hyp"].view(np.uint32))
    np.testing.assert_array_equal(cnt.cpu().numpy(), g["hyp_counts"])
    assert cnt.dtype == torch.int64
    close(rv.ransac_motion_voting(m_t, vert), g["motion_pts"], atol=1e-4)
    # distributions (own input, tie-free at the k-th ratio)
    dm, dv = torch.from_numpy(g["d_mask"]).to(cuda_dev), torch.from_numpy(g["d_vertex"]).to(cuda_dev)
    mean, cov = rv.estimate_voting_distribution(dm, dv, round_hyp_num=32, min_hyp_num=96, topk=16,
                                                **kw("dist", [True, True], 3))
    close(mean, g["dist_mean"])
    close(cov, g["dist_cov"], atol=1e-3, rtol=1e-3)
    mean_in = torch.from_numpy(g["d_kpts"].astype(np.float32)).to(cuda_dev)
    mean, cov = rv.estimate_voting_distribution_with_mean(dm, dv, mean_in, round_hyp_num=32, min_hyp_num=96,
                                                          **kw("distm", [True, True], 3))
    np.testing.assert_array_equal(mean.cpu().numpy(), g["distm_mean"])
    close(cov, g["distm_cov"], atol=1e-3, rtol=1e-3)
    # multi-class: (image, class) pairs in loop order; live = pairs with >= min_num pixels
    lab = g["c_mask"]
    live_c = [bool((lab[bi] == k + 1).sum() >= 5) for bi in range(3) for k in range(2)]
    l_t = torch.from_numpy(lab).to(cuda_dev)
    p1 = rv.ransac_voting_layer(l_t, vert, 3, hn, **kw("v1", live_c))
    np.testing.assert_array_equal(p1.cpu().numpy().view(np.uint32), g["v1_pts"].view(np.uint32))
    close(rv.ransac_voting_layer_v2(l_t, vert, 3, hn, refine_iter_num=2, **kw("v2", live_c)), g["v2_pts"])


def test_config4_size_distribution_matches_oracle(cuda_dev):
    """BASELINE config[3] size for one image: 768x768 field, 11 keypoints, 8 rounds x 256 = 2048 hypotheses, foreground
    147 k pixels subsampled to ~30000 (max_num), top-128 mean / covariance.  676 M pair tests through the C restatement:
    hypotheses and counts bit-exact, mean within 1e-3 px, covariance to float32 accuracy."""
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    h = w = 768
    vn, hn, rounds = 11, 256, 8
    mask, vertex, _ = make_vertex_field(91, 1, h, w, vn, 0.25, noise_deg=2.0)
    vx = vertex_hwvn2(vertex)
    idxs, selection, fn, sel = _idxs_for(mask, vx, hn, rounds, 30000, 21, eq1=True)
    m_t = torch.from_numpy(mask).to(cuda_dev)
    vert = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev))
    kw = dict(idxs=torch.from_numpy(idxs).to(cuda_dev), selection=torch.from_numpy(selection).to(cuda_dev))
    dbg = rv.voting_debug(_lib.VOTE_DISTRIBUTION, m_t, vert, hn, rounds=rounds, inlier_thresh=0.99, topk=128, **kw)
    _, coords, direct = ov.compact(mask[0] == 1, vx[0], 30000, sel, 0)
    assert int(dbg["tn"][0]) == coords.shape[0] and 25000 < coords.shape[0] < 35000
    hyp = dbg["hyp"][0].cpu().numpy()                                      # [rounds*hn, vn, 2]
    hyp_o = np.concatenate([ov.generate_hypothesis(direct, coords, idxs[0, r]) for r in range(rounds)], 0)
    np.testing.assert_array_equal(hyp.view(np.int32), hyp_o.view(np.int32))
    np.testing.assert_array_equal(dbg["counts"][0].cpu().numpy(), ov.vote_counts(direct, coords, hyp_o, 0.99))
    mean_o, cov_o = ov.estimate_voting_distribution(mask, vx, hn, hn * rounds, 128, idxs_fn=fn, selection_fn=sel)
    np.testing.assert_allclose(dbg["mean"].cpu().numpy(), mean_o, rtol=0, atol=1e-3)
    np.testing.assert_allclose(dbg["cov"].cpu().numpy(), cov_o, rtol=1e-3, atol=1e-3)
