#!/usr/bin/env python
"""A/B runs of the device-resident bench step against the TUNING build of the library (the only build in
which the EPB_* environment knobs exist, csrc/common.cuh):

    EPB_VOTE_IMPL=2 python tools/vote_ab.py --steps 50          # round-1 FP32 vote kernel
    EPB_VOTE_ITEM=1024 python tools/vote_ab.py --steps 50       # tensor-core kernel, other work-unit size

Prints bench.py's JSON line (no e2e leg, no CPU baseline); `kernel_ms_per_step` holds the per-class times."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from esa_pose_estimation_b200 import _lib, build  # noqa: E402

_lib.LIB_PATH = build.build_tuning()

import bench  # noqa: E402

if __name__ == "__main__":
    sys.argv += ["--no-cpu-baseline", "--no-configs"] + ([] if os.environ.get("EPB_AB_E2E") else ["--no-e2e"])
    bench.run_ours(bench.parse())
