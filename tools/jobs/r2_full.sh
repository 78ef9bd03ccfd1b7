#!/bin/bash
# GPU job: full GPU test suite + bench (default build) + A/B of the vote implementations (tuning build).
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
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1
echo "== gpu tests rc=$?"; tail -8 gpurun_out/r2_gpu_tests.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "== bench rc=$?"; show bench gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
for impl in 0 1; do
  EPB_VOTE_IMPL=$impl timeout 300 python tools/vote_ab.py --steps 50 > gpurun_out/r2_ab_impl$impl.json 2> gpurun_out/r2_ab_impl$impl.err
  show "impl=$impl" gpurun_out/r2_ab_impl$impl.json
done
