#!/bin/bash
# GPU job (gpurun --gpus 8): NCCL correctness test and bench at N = 2, 4 and 8 for the final build.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_nccl_gpu.py -x -q -m gpu > gpurun_out/r2_nccl_test.log 2>&1
echo "== nccl test rc=$?"; tail -3 gpurun_out/r2_nccl_test.log
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "== bench n=$n rc=$?"
  python - gpurun_out/r2_bench_n$n.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("value %.0f e2e %.0f ms/step %.3f e2e ms %.3f pcie %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"].get("pcie")))
    for k, v in (d.get("configs") or {}).items():
        print("   ", k, json.dumps(v)[:300])
except Exception as e:
    print("no line:", e)
PY
done
