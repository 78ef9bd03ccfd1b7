#!/usr/bin/env python
"""Where the cycles of one frame's pose solve go (tuning build: POSE_PHASE marks in csrc/pose.cu).  Runs the
single-frame pipeline a few times and prints the mean SM cycles per phase for CTA 0, thread 0."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esa_pose_estimation_b200 import _lib, build  # noqa: E402

PRODUCT = os.environ.get("PRODUCT") == "1"      # PRODUCT=1: the shipped library, no clocks (for ncu captures)
if not PRODUCT:
    _lib.LIB_PATH = build.build_tuning()
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

NAMES = ["select/un-crop", "epnp: sampling, centroid + covariance sums", "epnp: svd3(cov) + alphas", "epnp: MtM sums + fill",
         "epnp: eigen-solver entry/exit", "epnp: L, rho", "(unused)", "epnp: beta passes, rest", "crt3: pc + sums",
         "crt3: svd3 + R,t", "crt3: scoring", "ransac: score, barrier, replay", "rodrigues", "lm: accept", "lm: solve + step",
         "lm: rho", "lm: exit", "pack"]

def main():
    from synth import make_pose_case, tango_model, ESA_K
    from esa_pose_estimation_b200 import pnp as P
    dev = torch.device("cuda:0")
    lib = None if PRODUCT else ctypes.CDLL(_lib.LIB_PATH)
    n_out = int(os.environ.get("OUTLIERS", "0"))
    B = int(os.environ.get("FRAMES", "1"))
    model = tango_model(11, seed=9)
    cases = [make_pose_case(7000 + i, 11, 0.5, max(n_out, 0), model=model) for i in range(B)]
    kp = np.stack([c["p2d"] for c in cases])
    if n_out < 0:                      # OUTLIERS=-1: random keypoints, no consensus, all 100 RANSAC iterations run
        kp = np.random.default_rng(4).uniform(0, 1900, kp.shape)
    preds = torch.from_numpy(kp.astype(np.float32)).to(dev)
    maxv = torch.full((B, 11), 0.9, dtype=torch.float32, device=dev)
    z2 = torch.zeros((B, 2), dtype=torch.float64, device=dev)
    ones = torch.ones((B,), dtype=torch.float64, device=dev)
    m = torch.from_numpy(model).to(dev)
    K = torch.from_numpy(ESA_K).to(dev)
    buf = (ctypes.c_longlong * 32)()
    for _ in range(5):
        P.pose_pipeline(preds, maxv, z2, ones, m, K, min_k=11)
    if PRODUCT:
        torch.cuda.synchronize()
        return
    lib.epb_debug_pose_phase_clocks(buf)
    reps = 20
    for _ in range(reps):
        P.pose_pipeline(preds, maxv, z2, ones, m, K, min_k=11)
    lib.epb_debug_pose_phase_clocks(buf)
    cyc = [buf[i] / reps for i in range(len(NAMES))]
    tot = sum(cyc)
    extra = {"jacobi: convergence checks": buf[20] / reps, "jacobi: rotation parameters": buf[21] / reps,
             "jacobi: block updates": buf[22] / reps, "jacobi: sweeps": buf[23] / reps,
             "gn: build a, b": buf[24] / reps, "gn: lstsq6": buf[25] / reps,
             "lm eval: residual + jets": buf[26] / reps, "lm eval: table sums": buf[27] / reps}
    print(json.dumps({"frames": B, "outliers": n_out, "total_cycles": tot,
                      "phases": {n: round(c) for n, c in zip(NAMES, cyc)}, "inside_jacobi": extra}))


if __name__ == "__main__":
    main()
