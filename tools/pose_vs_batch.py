#!/usr/bin/env python
"""Pose kernel time against the batch size (launch shapes: two-CTA cluster per frame up to SMs / 2 frames, then 8, 4 and
2 solver warps in one CTA per frame)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("", "tests", "tools"):
    sys.path.insert(0, os.path.join(ROOT, d))
import numpy as np, torch
from esa_pose_estimation_b200 import _lib
if os.environ.get("OLDLIB"): _lib.LIB_PATH = os.environ["OLDLIB"]
from synth import make_pose_case, tango_model, ESA_K
from esa_pose_estimation_b200 import pipeline
import bench_configs as bc
DEV = torch.device("cuda:0")
model = tango_model(11, seed=9)
K = torch.from_numpy(ESA_K).to(DEV); m = torch.from_numpy(model).to(DEV)
base = np.stack([make_pose_case(6000 + i, 11, 0.5, 0, model=model)["p2d"] for i in range(64)])
for B in (1, 64, 74, 75, 148, 149, 296, 297, 375, 1500, 3000):
    kp = np.concatenate([base] * (B // 64 + 1))[:B]
    k_t = torch.from_numpy(kp.astype(np.float32)).to(DEV)
    ms = bc.timed(lambda: pipeline.poses_from_keypoints(k_t, m, K)["pose7"], 10)
    print("B %5d  %.3f ms  %.1f us/frame" % (B, ms, ms * 1e3 / B))
