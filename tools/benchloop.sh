for cfg in "$@"; do
  env $cfg python bench.py --steps ${STEPS:-10} --warmup 3 --no-cpu-baseline > gpurun_out/bl.log 2>&1
  python - "$cfg" <<PY
import json,sys
try:
    d=json.loads(open("gpurun_out/bl.log").read().strip().splitlines()[-1])
    k=d["kernel_ms_per_step"]
    print(sys.argv[1], "value=%.0f step=%.3f vote=%.3f compact=%.3f refine=%.3f pose=%.3f e2e=%.0f (%.2f ms)"%(d["value"],d["ms_per_step"],k["vote_count"],k["compaction"],k["winner_refine"],k["pose"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d["clocks"], d["e2e"].get("clocks"))
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open("gpurun_out/bl.log").read()[-800:])
PY
done
