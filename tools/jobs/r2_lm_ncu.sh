#!/bin/bash
# GPU job: ncu full capture of lm_kernel at 10^5 poses (config[4]).
mkdir -p gpurun_out
python tools/lm_sweep_step.py && \
ncu --set full --clock-control none --import-source on -k regex:lm_kernel -s 3 -c 1 -f -o gpurun_out/r2_lm_kernel python tools/lm_sweep_step.py > gpurun_out/r2_ncu_lm.log 2>&1
echo "capture rc=$?"
