"""Voting oracle: the C restatement against an independent pure-numpy evaluation on small
cases, plus driver-level properties (the reference driver cannot run anywhere)."""
import numpy as np

from oracle import voting as ov
from tests.synth import make_vertex_field, vertex_hwvn2


def _np_counts(direct, coords, hyp, thresh):
    """Independent float32 numpy evaluation of ransac_voting_kernel.cu:100-125 (no FMA:
    agrees with the contracted C restatement except within an ulp of the threshold)."""
    d = hyp[:, :, None, :] - coords[None, None, :, :]                  # [hn,vn,tn,2]
    n = direct.transpose(1, 0, 2)[None]                                 # [1,vn,tn,2]
    norm1 = np.sqrt((n ** 2).sum(-1, dtype=np.float32))
    norm2 = np.sqrt((d ** 2).sum(-1, dtype=np.float32))
    cos = (d * n).sum(-1, dtype=np.float32) / (norm1 * norm2)
    ok = (norm1 >= 1e-6) & (norm2 >= 1e-6)
    return (ok & (cos > np.float32(thresh)))


def _case(seed, h=24, w=28, vn=3, frac=0.5):
    mask, vertex, kpts = make_vertex_field(seed, 1, h, w, vn, frac)
    vx = vertex_hwvn2(vertex)
    fg, coords, direct = ov.compact(mask[0] != 0, vx[0], 30000, ov.default_selection_fn(0), 0)
    return mask, vx, kpts, coords, direct


def test_compaction_is_row_major_xy():
    mask, vx, _, coords, direct = _case(1)
    ys, xs = np.nonzero(mask[0])
    np.testing.assert_array_equal(coords[:, 0], xs)
    np.testing.assert_array_equal(coords[:, 1], ys)
    np.testing.assert_array_equal(direct, vx[0][ys, xs])


def test_generate_hypothesis_intersects_rays():
    _, _, kpts, coords, direct = _case(2, frac=1.0)
    tn, vn, _ = direct.shape
    rng = np.random.default_rng(0)
    idxs = rng.integers(0, tn, (64, vn, 2)).astype(np.int32)
    hyp = ov.generate_hypothesis(direct, coords, idxs)
    # float64 line intersection
    for h in range(64):
        for v in range(vn):
            t0, t1 = idxs[h, v]
            n0 = np.array([direct[t0, v, 1], -direct[t0, v, 0]], np.float64)
            n1 = np.array([direct[t1, v, 1], -direct[t1, v, 0]], np.float64)
            a = np.stack([n0, n1])
            det = np.float32(n1[0] * n0[1]) - np.float32(n0[0] * n1[1])
            if abs(det) < 1e-6:
                assert hyp[h, v, 0] == 0 and hyp[h, v, 1] == 0
                continue
            if abs(np.linalg.det(a)) < 1e-3:
                continue
            x = np.linalg.solve(a, [n0 @ coords[t0], n1 @ coords[t1]])
            np.testing.assert_allclose(hyp[h, v], x, rtol=2e-3, atol=2e-2)


def test_vote_counts_match_numpy_and_mask_sum():
    _, _, _, coords, direct = _case(3)
    tn, vn, _ = direct.shape
    rng = np.random.default_rng(1)
    idxs = rng.integers(0, tn, (32, vn, 2)).astype(np.int32)
    hyp = ov.generate_hypothesis(direct, coords, idxs)
    counts = ov.vote_counts(direct, coords, hyp, 0.99)
    inl = np.zeros((32, vn, tn), np.uint8)
    ov.voting_for_hypothesis(direct, coords, hyp, inl, 0.99)
    np.testing.assert_array_equal(counts, inl.sum(2))
    ref = _np_counts(direct, coords, hyp, 0.99)
    assert np.abs(ref.sum(2) - counts).max() <= 1          # only threshold-ulp cases may differ
    # writes only ones: pre-filled bytes survive
    inl2 = np.full((32, vn, tn), 7, np.uint8)
    ov.voting_for_hypothesis(direct, coords, hyp, inl2, 0.99)
    assert set(np.unique(inl2)) <= {1, 7}


def test_layer_v3_recovers_keypoints_and_variants_agree():
    mask, vertex, kpts = make_vertex_field(4, 2, 40, 44, 4, 0.4, noise_deg=1.0)
    vx = vertex_hwvn2(vertex)
    p3 = ov.ransac_voting_layer_v3(mask, vx, 64)
    assert np.abs(p3 - kpts).max() < 1.0
    p4, var = ov.ransac_voting_layer_v4(mask, vx, 64, inlier_thresh=0.999, confidence=0.99)
    np.testing.assert_array_equal(p3, p4)
    assert (var >= 0).all() and (var < 1.0).all()
    p5, conf = ov.ransac_voting_layer_v5(mask, vx, 64, max_num=30000)
    np.testing.assert_array_equal(p3, p5)
    assert ((conf > 0) & (conf <= 1)).all()


def test_degenerate_images():
    mask, vertex, _ = make_vertex_field(5, 2, 16, 16, 2, 0.3)
    mask[1] = 0
    mask[1, 0, :3] = 1                                   # 3 foreground pixels < min_num
    vx = vertex_hwvn2(vertex)
    p3 = ov.ransac_voting_layer_v3(mask, vx, 16)
    assert (p3[1] == 0).all()
    _, var = ov.ransac_voting_layer_v4(mask, vx, 16)
    assert (var[1] == 1).all()
    _, conf = ov.ransac_voting_layer_v5(mask, vx, 16)
    assert (conf[1] == 0).all()
    hyp, cnt = ov.ransac_voting_hypothesis(mask, vx, 16)
    assert (hyp[1] == 0).all() and (cnt[1] == 1).all()


def test_subsample_uses_selection_ratio():
    mask, vertex, _ = make_vertex_field(6, 1, 32, 32, 2, 1.0)
    vx = vertex_hwvn2(vertex)
    sel = ov.default_selection_fn(3)
    fg, coords, _ = ov.compact(mask[0] != 0, vx[0], 100, sel, 0)
    assert fg == 1024
    keep = sel(0, 32, 32) < np.float32(100) / np.float32(1024)
    assert coords.shape[0] == int(keep.sum())


def test_distribution_mean_cov():
    mask, vertex, kpts = make_vertex_field(7, 1, 40, 40, 3, 0.5, noise_deg=1.0)
    vx = vertex_hwvn2(vertex)
    mean, cov = ov.estimate_voting_distribution(mask, vx, round_hyp_num=32, min_hyp_num=128, topk=16)
    assert np.abs(mean - kpts).max() < 1.5
    assert (np.linalg.eigvalsh(cov.astype(np.float64)) > -1e-6).all()
    m2, cov2 = ov.estimate_voting_distribution_with_mean(mask, vx, mean, round_hyp_num=32, min_hyp_num=128)
    np.testing.assert_array_equal(m2, mean)
    assert np.isfinite(cov2).all()


def test_multiclass_oracle_reduces_to_single_class():
    """With one foreground class the multi-class drivers (:10, :99) follow the single-class v3 path:
    same winner, and v2's pinverse refinement equals v3's normal equations."""
    mask, vertex, kpts = make_vertex_field(5, 1, 48, 48, 3, 0.5, noise_deg=1.0)
    vx = vertex_hwvn2(vertex)
    fn = ov.default_idxs_fn(2)
    v3 = ov.ransac_voting_layer_v3(mask, vx, 96, idxs_fn=fn)
    v2 = ov.ransac_voting_layer_v2(mask, vx, 2, 96, idxs_fn=fn)
    v1 = ov.ransac_voting_layer(mask, vx, 2, 96, idxs_fn=fn)
    assert v2.shape == (1, 1, 3, 2) and v1.shape == (1, 1, 3, 2)
    np.testing.assert_allclose(v2[:, 0], v3, atol=2e-3)
    assert np.abs(v1[0, 0] - kpts[0]).max() < 3.0


# ----------------------------------------------------------------------------- drivers vs the reference's Python
def test_drivers_match_the_reference_python_drivers():
    """oracle/voting.py against tests/golden/voting_drivers.npz = outputs of the reference's own
    ransac_voting_gpu.py (imported unmodified, run on CPU over the C restatement of its kernels;
    tests/golden/make_golden_voting.py), with the reference's recorded random draws replayed.
    Hypotheses and inlier counts bit-exact; keypoints within 1e-3 px (float32 cuBLAS/torch sums in
    the reference, float64 here); variance / confidence / mean / covariance to float32 accuracy."""
    from tests.golden_voting import Replay, load
    g = load()
    mask, vx, hn = g["a_mask"], g["a_vertex"], int(g["a_hn"])

    def rp(name):
        r = Replay(g, name)
        return r, dict(idxs_fn=r.idxs_fn, selection_fn=r.selection_fn)

    r, kw = rp("v3")
    np.testing.assert_allclose(ov.ransac_voting_layer_v3(mask, vx, hn, **kw), g["v3_pts"], atol=1e-3, rtol=0)
    assert r.exhausted()
    r, kw = rp("v3sub")
    np.testing.assert_allclose(ov.ransac_voting_layer_v3(mask, vx, hn, max_num=150, **kw), g["v3sub_pts"], atol=1e-3, rtol=0)
    assert r.exhausted()
    r, kw = rp("v3t")
    np.testing.assert_allclose(ov.ransac_voting_layer_v3(mask, vx, hn, inlier_thresh=0.99, **kw), g["v3t_pts"], atol=1e-3, rtol=0)
    r, kw = rp("v4")
    pts, var = ov.ransac_voting_layer_v4(mask, vx, hn, **kw)
    np.testing.assert_allclose(pts, g["v4_pts"], atol=1e-3, rtol=0)
    np.testing.assert_allclose(var, g["v4_var"], rtol=2e-3, atol=1e-6)
    r, kw = rp("v5")
    pts, conf = ov.ransac_voting_layer_v5(mask, vx, hn, **kw)
    np.testing.assert_allclose(pts, g["v5_pts"], atol=1e-3, rtol=0)
    np.testing.assert_array_equal(conf, g["v5_conf"])
    assert r.exhausted()
    r, kw = rp("hyp")
    hyp, counts = ov.ransac_voting_hypothesis(mask, vx, hn, **kw)
    np.testing.assert_array_equal(hyp.view(np.uint32), g["hyp_hyp"].view(np.uint32))
    np.testing.assert_array_equal(counts, g["hyp_counts"])
    r, kw = rp("dist")
    mean, cov = ov.estimate_voting_distribution(g["d_mask"], g["d_vertex"], round_hyp_num=32, min_hyp_num=96, topk=16, **kw)
    np.testing.assert_allclose(mean, g["dist_mean"], atol=1e-3, rtol=0)
    np.testing.assert_allclose(cov, g["dist_cov"], rtol=1e-3, atol=1e-3)
    assert r.exhausted()
    r, kw = rp("distm")
    mean, cov = ov.estimate_voting_distribution_with_mean(g["d_mask"], g["d_vertex"], g["d_kpts"].astype(np.float32),
                                                          round_hyp_num=32, min_hyp_num=96, **kw)
    np.testing.assert_array_equal(mean, g["distm_mean"])
    np.testing.assert_allclose(cov, g["distm_cov"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(ov.ransac_motion_voting(mask, vx), g["motion_pts"], atol=1e-4, rtol=0)
    lab = g["c_mask"]
    r, kw = rp("v1")
    np.testing.assert_array_equal(ov.ransac_voting_layer(lab, vx, 3, hn, **kw), g["v1_pts"])
    assert r.exhausted()
    r, kw = rp("v2")
    np.testing.assert_allclose(ov.ransac_voting_layer_v2(lab, vx, 3, hn, refine_iter_num=2, **kw), g["v2_pts"], atol=1e-3, rtol=0)
    assert r.exhausted()
