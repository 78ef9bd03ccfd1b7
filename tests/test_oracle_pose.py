"""Pose oracle: pnp() against vectors from the reference's pnp.py; LM restatement against the
reference's uncertainty_pnp.cpp compiled over the vendored TinySolver (oracle/_ref); the EPnP /
RANSAC restatement (the blueprint of csrc/pose.cu) against cv2."""
import os

import cv2
import numpy as np
import pytest

from oracle import _lib as olib
from oracle import epnp_port, pose as opose
from tests.synth import ESA_K, make_pose_case, rodrigues


def _ang(r1, r2):
    c = (np.trace(r1 @ r2.T) - 1) / 2
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


def test_pnp_oracle_matches_reference_vectors(golden_dir):
    g = np.load(os.path.join(golden_dir, "pnp_ref.npz"))
    for i in range(int(g["n_cases"])):
        rt = opose.pnp(g["p3d_%d" % i], g["p2d_%d" % i], g["K"], cv2.SOLVEPNP_EPNP)
        np.testing.assert_allclose(rt, g["rt_%d" % i], rtol=0, atol=1e-9)


def test_epnp_port_matches_reference_vectors(golden_dir):
    """The from-scratch EPnP-RANSAC statement reproduces the reference's pnp() output."""
    g = np.load(os.path.join(golden_dir, "pnp_ref.npz"))
    K = g["K"]
    for i in range(int(g["n_cases"])):
        r, t, mask = epnp_port.solve_pnp_ransac_epnp(g["p3d_%d" % i], g["p2d_%d" % i], K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        ref = g["rt_%d" % i]
        assert _ang(r, ref[:, :3]) < 1e-4
        assert np.linalg.norm(t - ref[:, 3]) / np.linalg.norm(ref[:, 3]) < 1e-6


def test_epnp_port_vs_cv2_random():
    rng = np.random.default_rng(11)
    K = ESA_K
    for trial in range(40):
        n = int(rng.choice([6, 8, 11, 24, 30]))
        c = make_pose_case(1000 + trial, n, 0.5, 0)
        ok, rvec, tvec = cv2.solvePnP(c["p3d"][None], c["p2d"][None], K, np.zeros((8, 1)), flags=cv2.SOLVEPNP_EPNP)
        rc, _ = cv2.Rodrigues(rvec)
        r, t, _ = epnp_port.epnp(c["p3d"], c["p2d"], K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        assert _ang(rc, r) < 1e-4
        assert np.linalg.norm(t - tvec.ravel()) / np.linalg.norm(t) < 1e-8


def test_ransac_port_consensus_matches_cv2():
    K = ESA_K
    for trial in range(25):
        c = make_pose_case(2000 + trial, 11 if trial % 2 else 24, 0.7, trial % 3)
        ok, rv, tv, inl = cv2.solvePnPRansac(c["p3d"][None], c["p2d"][None], K, np.zeros((8, 1)),
                                             reprojectionError=5.0, flags=cv2.SOLVEPNP_EPNP)
        res = epnp_port.solve_pnp_ransac_epnp(c["p3d"], c["p2d"], K[0, 0], K[1, 1], K[0, 2], K[1, 2])
        assert ok and res is not None
        m = np.zeros(len(c["p3d"]), bool)
        m[inl.ravel()] = True
        np.testing.assert_array_equal(m, res[2])
        assert not m[c["outliers"]].any()
        rc, _ = cv2.Rodrigues(rv)
        assert _ang(rc, res[0]) < 1e-4


def test_lm_restatement_matches_reference_build():
    ref = olib.ref_pnp_lib(required=False)
    if ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rng = np.random.default_rng(5)
    for trial in range(60):
        n = int(rng.integers(5, 31))
        c = make_pose_case(3000 + trial, n, 0.5, 0)
        w = np.stack([rng.uniform(0.3, 1, n), rng.uniform(-0.1, 0.1, n), rng.uniform(0.3, 1, n)], 1)
        init = np.concatenate([c["rvec"] + rng.normal(0, 0.03, 3), c["t"] * (1 + rng.normal(0, 0.02, 3))])
        a = opose.lm_refine(c["p2d"], c["p3d"], w, ESA_K, init, use_ref=False)
        b = opose.lm_refine(c["p2d"], c["p3d"], w, ESA_K, init, use_ref=True)
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-9)


def test_lm_recovers_pose_and_is_fixed_point():
    c = make_pose_case(77, 11, 0.0, 0)
    w = np.stack([np.ones(11), np.zeros(11), np.ones(11)], 1)
    init = np.concatenate([c["rvec"] + 0.02, c["t"] * 1.02])
    rt = opose.lm_refine(c["p2d"], c["p3d"], w, ESA_K, init)
    assert _ang(rodrigues(rt[:3]), rodrigues(c["rvec"])) < 1e-6
    np.testing.assert_allclose(rt[3:], c["t"], rtol=1e-8)
    rt2 = opose.lm_refine(c["p2d"], c["p3d"], w, ESA_K, rt)
    np.testing.assert_allclose(rt2, rt, atol=1e-10)


def test_frame_pose_and_score():
    c = make_pose_case(91, 30, 0.3, 0)
    # crop geometry like data_load_val: preds are crop pixels, ori = preds/rate + (x, y)
    rate, bx, by = 0.25, 400.0, 300.0
    preds = ((c["p2d"] - [bx, by]) * rate).astype(np.float32)
    maxvals = np.linspace(0.95, 0.3, 30)
    out = opose.frame_pose(preds, maxvals, (bx, by), rate, c["p3d"], ESA_K)
    assert len(out["idxs"]) == 24
    assert _ang(out["pose34"][:, :3], rodrigues(c["rvec"])) < 0.5
    q_gt = opose.quat_wxyz_from_matrix(rodrigues(c["rvec"]))
    st, sr = opose.esa_score(out["q"], out["t"], q_gt, c["t"])
    assert st < 0.02 and sr < 0.02


def test_cov_to_weights_oracle_matches_closed_form():
    """evaluation_utils.py:170-181 restated with scipy's sqrtm == the 2x2 SPD closed form the kernel uses."""
    from oracle import pose as opose
    rng = np.random.default_rng(5)
    a = rng.normal(size=(40, 2, 2))
    cov = (a @ a.transpose(0, 2, 1) + 1e-3 * np.eye(2)).astype(np.float32)
    cov[3, 0, 0] = 1e-7                                   # guard: zero weights
    cov[5, 1, 1] = np.nan
    w = opose.cov_to_weights(cov)
    assert (w[3] == 0).all() and (w[5] == 0).all()
    for i in (0, 1, 2, 7, 20):
        c = cov[i].astype(np.float64)
        s = np.sqrt(np.linalg.det(c)); t = np.sqrt(np.trace(c) + 2 * s)
        winv = np.linalg.inv((c + s * np.eye(2)) / t)
        np.testing.assert_allclose(w[i], [winv[0, 0], winv[0, 1], winv[1, 1]], rtol=2e-4, atol=1e-6)


def test_uncertainty_pnp_oracle_matches_the_reference_python(golden_dir):
    """oracle/pose.py: uncertainty_pnp / isotropic_weights against tests/golden/uncertainty_pnp.npz = the
    reference's own extend_utils.uncertainty_pnp / uncertainty_pnp_v2 (imported unmodified, bound to the
    reference's uncertainty_pnp.cpp; tests/golden/make_golden_uncertainty.py)."""
    g = np.load(os.path.join(golden_dir, "uncertainty_pnp.npz"))
    K = g["K"]
    for i in range(int(g["n_cases"])):
        p2d, p3d, cov, w = g["p2d_%d" % i], g["p3d_%d" % i], g["cov_%d" % i], g["w_%d" % i]
        np.testing.assert_allclose(opose.cov_to_weights(cov), w, rtol=1e-12, atol=0)
        for use_ref in (False, True):
            if use_ref and olib.ref_pnp_lib(required=False) is None:
                continue
            rt = opose.uncertainty_pnp(p2d, w, p3d, K, use_ref=use_ref)
            assert _ang(rt[:, :3], g["rt_%d" % i][:, :3]) < 1e-5      # arccos near 1: ~1e-6 deg of rounding
            np.testing.assert_allclose(rt[:, 3], g["rt_%d" % i][:, 3], rtol=1e-8)
            rt2 = opose.uncertainty_pnp(p2d, opose.isotropic_weights(cov), p3d, K, use_ref=use_ref)
            assert _ang(rt2[:, :3], g["rt_v2_%d" % i][:, :3]) < 1e-5
            np.testing.assert_allclose(rt2[:, 3], g["rt_v2_%d" % i][:, 3], rtol=1e-8)
