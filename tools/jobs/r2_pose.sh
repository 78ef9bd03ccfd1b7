#!/bin/bash
mkdir -p gpurun_out
for r0 in 4 8; do EPB_RANSAC_R0=$r0 timeout 300 python tools/pose_stress.py 2>&1 | tail -1; done
timeout 600 python -m pytest tests/test_pose_gpu.py tests/test_voting_gpu.py -x -q -m gpu 2>&1 | tail -3
