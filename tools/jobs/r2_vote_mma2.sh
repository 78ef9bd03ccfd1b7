#!/bin/bash
# GPU job: vote_mma after the code-size fix -- parity, then timing with the bisecting knobs.
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
timeout 900 python -m pytest tests/test_voting_gpu.py -x -q -m gpu > gpurun_out/r2c_voting_tests.log 2>&1
echo "== voting tests rc=$?"; tail -5 gpurun_out/r2c_voting_tests.log
for dbg in 0 4; do
  EPB_VM_DEBUG=$dbg timeout 200 python tools/vote_ab.py --steps 30 > gpurun_out/r2c_dbg$dbg.json 2> gpurun_out/r2c_dbg$dbg.err
  show "dbg=$dbg" gpurun_out/r2c_dbg$dbg.json
done
for item in 1024 4096; do
  EPB_VOTE_ITEM=$item timeout 200 python tools/vote_ab.py --steps 30 > gpurun_out/r2c_item$item.json 2> gpurun_out/r2c_item$item.err
  show "item=$item" gpurun_out/r2c_item$item.json
done
timeout 200 python tools/vote_ab.py --steps 3 --warmup 3 > gpurun_out/r2c_plain.json 2> gpurun_out/r2c_plain.err &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:vote_mma -s 3 -c 1 -o gpurun_out/r2c_vote_mma \
    python tools/vote_ab.py --steps 3 --warmup 3 > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"
