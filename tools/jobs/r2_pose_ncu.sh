#!/bin/bash
# GPU job: ncu full capture of the single-frame pose kernel (product library).
mkdir -p gpurun_out
PRODUCT=1 FRAMES=${FRAMES:-1} python tools/pose_phases.py && \
PRODUCT=1 FRAMES=${FRAMES:-1} ncu --set full --clock-control none --import-source on -k regex:pose_pipeline -s 3 -c 1 -f -o gpurun_out/r2_pose_single python tools/pose_phases.py > gpurun_out/r2_ncu_pose_single.log 2>&1
echo "capture rc=$?"
