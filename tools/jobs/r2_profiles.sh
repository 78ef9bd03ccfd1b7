#!/bin/bash
# GPU job: the ncu evidence for the final build (launch list, full captures of the top kernels) + one sustained run.
mkdir -p gpurun_out
CMD="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
$CMD > gpurun_out/r2_plain.json 2> gpurun_out/r2_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2_plain2.json 2> gpurun_out/r2_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:vote_count -s 3 -c 1 -o gpurun_out/r2_vote_count $CMD > gpurun_out/r2_ncu_vote.log 2>&1
echo "vote_count capture rc=$?"
$CMD > gpurun_out/r2_plain3.json 2> gpurun_out/r2_plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:pose_pipeline -s 3 -c 1 -o gpurun_out/r2_pose_pipeline $CMD > gpurun_out/r2_ncu_pose.log 2>&1
echo "pose capture rc=$?"
# sustained: >= 5 s of back-to-back steps with the clock record bench.py takes during the timed region
python bench.py --steps 3000 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/r2_sustained.json 2> gpurun_out/r2_sustained.err
echo "sustained rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_sustained.json").read().strip().splitlines()[-1])
print("sustained value %.0f e2e %.0f ms/step %.3f clocks %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))
PY
