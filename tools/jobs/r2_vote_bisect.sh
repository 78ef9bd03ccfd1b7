#!/bin/bash
# GPU job: where does vote_mma_kernel spend its time?  Bisecting knobs of the tuning build + one ncu capture.
mkdir -p gpurun_out
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step %.3f" % d["ms_per_step"], {k: round(v, 3) for k, v in d["kernel_ms_per_step"].items()})
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
for dbg in 0 1 2 3 4 6 7; do
  EPB_VM_DEBUG=$dbg timeout 200 python tools/vote_ab.py --steps 20 > gpurun_out/r2b_dbg$dbg.json 2> gpurun_out/r2b_dbg$dbg.err
  show "dbg=$dbg" gpurun_out/r2b_dbg$dbg.json
done
timeout 200 python tools/vote_ab.py --steps 3 --warmup 3 > gpurun_out/r2b_plain.json 2> gpurun_out/r2b_plain.err &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vote_mma -s 3 -c 1 -o gpurun_out/r2b_vote_mma \
    python tools/vote_ab.py --steps 3 --warmup 3 > gpurun_out/r2b_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/r2b_ncu.log
