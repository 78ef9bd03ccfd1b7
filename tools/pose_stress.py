#!/usr/bin/env python
"""Pose solve under RANSAC pressure against the TUNING build (EPB_RANSAC_R0 = warps of round 0): prints
tools/bench_configs.pose_stress() and the single-frame latency."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esa_pose_estimation_b200 import _lib, build  # noqa: E402

_lib.LIB_PATH = build.build_tuning()
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench_configs as bc  # noqa: E402

if __name__ == "__main__":
    print(json.dumps({"EPB_RANSAC_R0": os.environ.get("EPB_RANSAC_R0"), "stress": bc.pose_stress(),
                      "c1_us": bc.c1()["value"]}))
