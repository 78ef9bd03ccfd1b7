"""Golden vectors for the uncertainty-PnP front end, produced by the reference's own Python + C++.

Run in the build container only (needs /root/reference and a built oracle/_ref):
    python tests/golden/make_golden_uncertainty.py

/root/reference/lib/utils/extend_utils/extend_utils.py is imported UNMODIFIED.  Its cffi import is
commented out in the shipped file (:3-4), so its wrappers look up the globals ``lib`` / ``ffi`` and
raise NameError as shipped; here the two names are bound, from outside the file, to a cffi handle on
oracle/_ref/libuncertainty_pnp_ref.so = the reference's own src/uncertainty_pnp.cpp compiled unmodified
(oracle/Makefile; over its vendored header-only ceres::TinySolver, because libceres cannot link in
this image -- DESIGN.md section 2).  cv2's P3P initialiser (third-party, 4.13.0 here) runs as the reference calls it.
Output: tests/golden/uncertainty_pnp.npz -- inputs, and the [3,4] poses of ``uncertainty_pnp`` (:64)
and ``uncertainty_pnp_v2`` (:117).
"""
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import cffi  # noqa: E402
import lib.utils.extend_utils.extend_utils as ref_ext  # noqa: E402  (the reference, unmodified)

from oracle import pose as opose  # noqa: E402
from tests.synth import ESA_K, make_pose_case  # noqa: E402

ffi = cffi.FFI()
# utils_python_binding.h:23-31
ffi.cdef("void uncertainty_pnp(double* pts2d, double* pts3d, double* wgt2d, double* K, double* init_rt, double* result_rt, int pn);")
ref_ext.ffi = ffi
ref_ext.lib = ffi.dlopen(os.path.join(ROOT, "oracle", "_ref", "libuncertainty_pnp_ref.so"))


def main():
    out = {"K": ESA_K}
    cases = [(400, 11, 0.6), (401, 11, 1.0), (402, 8, 0.4), (403, 30, 0.8), (404, 11, 0.2), (405, 16, 0.7),
             (406, 4, 0.5), (407, 4, 0.3), (408, 5, 0.4)]          # pn == 4: the P3P pose itself (:91-95)
    for i, (seed, n, noise) in enumerate(cases):
        c = make_pose_case(seed, n, noise, 0)
        rng = np.random.default_rng(seed)
        # anisotropic per-keypoint covariances (px^2), like estimate_voting_distribution's output
        ang = rng.uniform(0, np.pi, n)
        s1, s2 = rng.uniform(0.3, 4.0, n), rng.uniform(0.3, 4.0, n)
        rot = np.stack([np.stack([np.cos(ang), -np.sin(ang)], 1), np.stack([np.sin(ang), np.cos(ang)], 1)], 1)
        cov = (rot @ (np.stack([s1, s2], 1)[:, :, None] * np.eye(2)) @ rot.transpose(0, 2, 1)).astype(np.float32)
        if i == 1:
            cov[3] = 0                                    # degenerate keypoint: zero weight in both variants
        weights = opose.cov_to_weights(cov)               # evaluation_utils.py:170-181 (method of a class that needs the LINEMOD db)
        rt1 = ref_ext.uncertainty_pnp(c["p2d"], weights, c["p3d"], ESA_K)
        rt2 = ref_ext.uncertainty_pnp_v2(c["p2d"], cov, c["p3d"], ESA_K)
        out.update({"p2d_%d" % i: c["p2d"], "p3d_%d" % i: c["p3d"], "cov_%d" % i: cov, "w_%d" % i: np.asarray(weights, np.float64),
                    "rt_%d" % i: rt1, "rt_v2_%d" % i: rt2, "gt_rvec_%d" % i: c["rvec"], "gt_t_%d" % i: c["t"]})
        print(i, n, np.round(rt1[:, 3], 4), np.round(rt2[:, 3], 4), np.round(c["t"], 4))
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "uncertainty_pnp.npz"), **out)


if __name__ == "__main__":
    main()
