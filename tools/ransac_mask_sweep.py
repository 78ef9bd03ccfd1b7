#!/usr/bin/env python
"""Consensus sets of the device EPnP-RANSAC against cv2.solvePnPRansac over a large random sweep (not a test: 5-point
EPnP samples are ill-conditioned, so a borderline point can legitimately fall on either side of the 5 px threshold
depending on the last bit of a sum; this script measures how often)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("", "tests"):
    sys.path.insert(0, os.path.join(ROOT, d))
import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402
from synth import make_pose_case, ESA_K  # noqa: E402
from esa_pose_estimation_b200 import _lib  # noqa: E402
if os.environ.get("EPB_LIB"):            # another build of the library (A/B against an older solver)
    _lib.LIB_PATH = os.path.join(ROOT, os.environ["EPB_LIB"])
from esa_pose_estimation_b200 import pnp as P  # noqa: E402


def main():
    n_cases = int(os.environ.get("CASES", "2000"))
    dev = torch.device("cuda:0")
    cases = [make_pose_case(20000 + i, 11 if i % 2 else 24, 0.7, i % 4) for i in range(n_cases)]
    nmax = 24
    p3 = np.zeros((n_cases, nmax, 3)); p2 = np.zeros((n_cases, nmax, 2)); npts = np.zeros(n_cases, np.int32)
    for i, c in enumerate(cases):
        n = len(c["p3d"]); npts[i] = n
        p3[i, :n], p2[i, :n] = c["p3d"], c["p2d"]
    rt, mask, status = P.pnp_batch(torch.from_numpy(p3).to(dev), torch.from_numpy(p2).to(dev), torch.from_numpy(ESA_K).to(dev),
                                   npts=torch.from_numpy(npts).to(dev), return_status=True)
    rt, mask, status = rt.cpu().numpy(), mask.cpu().numpy(), status.cpu().numpy()
    diff, fail_both, fail_one, worst_rot, worst_t = [], 0, 0, 0.0, 0.0
    for i, c in enumerate(cases):
        ok, rv, tv, inl = cv2.solvePnPRansac(c["p3d"][None], c["p2d"][None], ESA_K, np.zeros((8, 1)),
                                             reprojectionError=5.0, flags=cv2.SOLVEPNP_EPNP)
        if not ok or status[i] != 0:
            fail_both += (not ok) and status[i] != 0
            fail_one += (not ok) != (status[i] != 0)
            continue
        m = 0
        for k in inl.ravel():
            m |= 1 << int(k)
        R, _ = cv2.Rodrigues(rv)
        cosang = np.clip((np.trace(rt[i, :, :3] @ R.T) - 1) / 2, -1, 1)
        if int(mask[i]) != m:
            # the point(s) in question: reprojection error under OpenCV's final pose (inliers of either set are fine
            # points; what differs is the 5-point candidate model that scored them)
            pr = (ESA_K @ (R @ c["p3d"].T + tv)).T
            err = np.linalg.norm(pr[:, :2] / pr[:, 2:] - c["p2d"], axis=1)
            bits = [k for k in range(int(npts[i])) if ((int(mask[i]) ^ m) >> k) & 1]
            diff.append({"case": i, "n": int(npts[i]), "outliers": len(c["outliers"]), "points": bits,
                         "px_error_under_final_pose": [round(float(err[k]), 2) for k in bits],
                         "rot_deg_vs_cv2": float(np.degrees(np.arccos(cosang))),
                         "rel_t_vs_cv2": float(np.linalg.norm(rt[i, :, 3] - tv.ravel()) / np.linalg.norm(tv))})
            continue
        worst_rot = max(worst_rot, float(np.degrees(np.arccos(cosang))))
        worst_t = max(worst_t, float(np.linalg.norm(rt[i, :, 3] - tv.ravel()) / np.linalg.norm(tv)))
    print(json.dumps({"cases": n_cases, "consensus_sets_differ": len(diff), "differing": diff[:10],
                      "both_fail": int(fail_both), "one_fails": int(fail_one),
                      "worst_rotation_deg_where_sets_agree": worst_rot, "worst_rel_translation_where_sets_agree": worst_t}))


if __name__ == "__main__":
    main()
