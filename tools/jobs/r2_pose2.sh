#!/bin/bash
# GPU job: pose / LM / pipeline parity tests, then the pose kernel under RANSAC pressure, single-frame latency and
# the LM sweep (tuning build = same sources).
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pose_gpu.py tests/test_pipeline_gpu.py tests/test_decode_gpu.py -x -q -m gpu 2>&1 | tail -5
timeout 300 python tools/pose_stress.py 2>&1 | tail -1
timeout 300 python - <<'PY' 2>&1 | tail -3
import sys; sys.path.insert(0, "tools")
import bench_configs as bc
print("c5", bc.c5())
PY
for o in 0 1; do OUTLIERS=$o timeout 300 python tools/pose_phases.py 2>&1 | tail -1; done
