"""Generates the committed golden vectors by running the REFERENCE's own Python code.

Run in the build container only (needs /root/reference; the GPU box has no reference tree):
    python tests/golden/make_golden.py
Imports /root/reference/{inference,pnp,submission}.py unmodified (all three import
as-is with numpy/cv2/torch/scipy) and stores inputs + the reference's outputs as small .npz files
that tests/ compare the oracle and the CUDA path against.
"""
import io
import os
import sys
import contextlib

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(HERE, "..", ".."))

import cv2  # noqa: E402
import inference as ref_inference  # noqa: E402  (reference)
import pnp as ref_pnp  # noqa: E402            (reference)
import submission as ref_submission  # noqa: E402

from tests.synth import ESA_K, make_heatmaps, make_pose_case  # noqa: E402


def decode_cases():
    for name, seed, k, hw, kind in [("gauss11", 100, 11, 48, "gauss"), ("gauss30", 101, 30, 40, "gauss"),
                                    ("randinit11", 102, 11, 48, "randinit"), ("edges11", 103, 11, 32, "edges")]:
        hm, _ = make_heatmaps(seed, 1, k, hw, hw, kind)
        preds, maxvals = ref_inference.get_max_preds(hm.copy())
        # val.py:151-168 flow: integer peaks -> get_final (mutates a list of float32[2])
        co = [preds[0, i].copy() for i in range(k)]
        final = ref_inference.get_final(hm.copy(), co)
        import torch
        gp, gm = ref_inference.getPrediction(torch.from_numpy(hm.copy()))
        # DARK-style decode, inference.py:154-170 (np.matrix and all): same integer peaks
        co2 = [preds[0, i].copy() for i in range(k)]
        final2 = ref_inference.get_final2(hm.copy(), co2)
        np.savez_compressed(os.path.join(HERE, "decode_%s.npz" % name), hm=hm, preds=preds, maxvals=maxvals,
                            final=np.asarray(final, np.float32), getpred=gp.numpy(), getpred_max=gm.numpy(),
                            final2=np.asarray(final2, np.float32))
        print("decode", name, hm.shape, float(np.abs(np.asarray(final) - preds[0]).max()))


def pnp_cases():
    out = {}
    for i, (seed, n, noise, nout) in enumerate([(200, 11, 0.5, 0), (201, 11, 1.0, 1), (202, 30, 0.7, 0),
                                                (203, 24, 0.7, 2), (204, 8, 0.3, 0), (205, 11, 0.0, 0),
                                                (206, 29, 0.8, 1), (207, 11, 0.6, 0)]):
        c = make_pose_case(seed, n, noise, nout)
        rt = ref_pnp.pnp(c["p3d"], c["p2d"], ESA_K, cv2.SOLVEPNP_EPNP)
        out["p3d_%d" % i], out["p2d_%d" % i], out["rt_%d" % i] = c["p3d"], c["p2d"], rt
        out["gt_rvec_%d" % i], out["gt_t_%d" % i] = c["rvec"], c["t"]
        print("pnp", i, n, nout, np.round(rt[:, 3], 4))
    out["K"] = ESA_K
    out["n_cases"] = np.array(8)
    np.savez_compressed(os.path.join(HERE, "pnp_ref.npz"), **out)


def submission_case():
    w = ref_submission.SubmissionWriter()
    rng = np.random.default_rng(300)
    rows = []
    for name in ["img000010.jpg", "img000002.jpg", "img000007.jpg"]:
        q = rng.normal(size=4).astype(np.float32); q /= np.linalg.norm(q)
        r = rng.normal(size=3)
        w.append_test(name, q, r)
        rows.append((name, q, r, False))
    for name in ["img000003real.jpg", "img000001real.jpg"]:
        q = rng.normal(size=4).astype(np.float32); q /= np.linalg.norm(q)
        r = rng.normal(size=3)
        w.append_real_test(name, q, r)
        rows.append((name, q, r, True))
    tmp = os.path.join(HERE, "_tmp")
    os.makedirs(tmp, exist_ok=True)
    with contextlib.redirect_stdout(io.StringIO()):
        w.export(out_dir=tmp, suffix="golden")
    text = open(os.path.join(tmp, "submission_golden.csv")).read()
    os.remove(os.path.join(tmp, "submission_golden.csv")); os.rmdir(tmp)
    np.savez_compressed(os.path.join(HERE, "submission_ref.npz"), csv=np.array(text),
                        names=np.array([r[0] for r in rows]), q=np.stack([r[1] for r in rows]),
                        r=np.stack([r[2] for r in rows]), real=np.array([r[3] for r in rows]))
    print("submission", len(text.splitlines()), "rows")


if __name__ == "__main__":
    decode_cases()
    pnp_cases()
    submission_case()
