#!/bin/bash
# GPU job (round 2): tensor-core vote kernel -- probe, parity tests, A/B timing.  Run from the repo root under gpurun.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_env.txt
timeout 120 ./tools/micro/mma_tf32_probe > gpurun_out/mma_tf32_probe.txt 2>&1
echo "== probe rc=$?"; cat gpurun_out/mma_tf32_probe.txt
timeout 900 python -m pytest tests/test_voting_gpu.py -x -q -m gpu > gpurun_out/r2a_voting_tests.log 2>&1
echo "== voting tests rc=$?"; tail -15 gpurun_out/r2a_voting_tests.log
for impl in 1 2; do
  EPB_VOTE_IMPL=$impl timeout 300 python tools/vote_ab.py --steps 50 > gpurun_out/r2a_ab_impl$impl.json 2> gpurun_out/r2a_ab_impl$impl.err
  echo "== impl $impl rc=$?"; python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2a_ab_impl$impl.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms/step", d["ms_per_step"], d["kernel_ms_per_step"], d["clocks"])
except Exception as e:
    print("no line:", e); print(open("gpurun_out/r2a_ab_impl$impl.err").read()[-2000:])
PY
done
for item in 512 1024 4096; do
  EPB_VOTE_ITEM=$item timeout 300 python tools/vote_ab.py --steps 50 > gpurun_out/r2a_ab_item$item.json 2> gpurun_out/r2a_ab_item$item.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2a_ab_item$item.json").read().strip().splitlines()[-1])
    print("item $item value", d["value"], "ms/step", d["ms_per_step"], d["kernel_ms_per_step"])
except Exception as e:
    print("item $item no line:", e)
PY
done
