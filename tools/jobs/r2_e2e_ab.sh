#!/bin/bash
# GPU job: e2e (host buffers) and device-resident step under different shared-memory carve-outs of vote_count.
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %s ms/step %.3f" % (d["value"], d["e2e"].get("value"), d["ms_per_step"]), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
for c in 60 75 100; do
  EPB_AB_E2E=1 EPB_VOTE_CARVEOUT=$c timeout 300 python tools/vote_ab.py --steps 100 > gpurun_out/r2_e2e_c$c.json 2> gpurun_out/r2_e2e_c$c.err
  show "carveout=$c" gpurun_out/r2_e2e_c$c.json
done
EPB_AB_E2E=1 EPB_CARVEOUT=100 EPB_VOTE_CARVEOUT=100 timeout 300 python tools/vote_ab.py --steps 100 > gpurun_out/r2_e2e_all100.json 2> gpurun_out/r2_e2e_all100.err
show "all=100" gpurun_out/r2_e2e_all100.json
timeout 600 python -m pytest tests/test_voting_gpu.py tests/test_pipeline_gpu.py -x -q -m gpu 2>&1 | tail -3
