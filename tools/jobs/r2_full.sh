#!/bin/bash
# GPU job: full GPU test suite + bench (default build) with ncu evidence of the final kernels.
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.0f e2e %s ms/step %.3f" % (d["value"], d["e2e"].get("value"), d["ms_per_step"]), {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
    for k, v in (d.get("configs") or {}).items():
        print("   ", k, json.dumps(v)[:400])
    print("    reference_gpu", d.get("reference_gpu", {}).get("ms_per_image"), d.get("reference_gpu", {}).get("speedup_voting"))
    print("    cpu_baseline", d.get("cpu_baseline"))
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/r2_gpu_tests.log 2>&1
echo "== gpu tests rc=$?"; tail -8 gpurun_out/r2_gpu_tests.log
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
echo "== bench rc=$?"; show bench gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
