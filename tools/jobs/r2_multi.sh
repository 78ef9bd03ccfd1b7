#!/bin/bash
# GPU job (gpurun --gpus 8): NCCL correctness test, host->device delivery at 1/2/4/8 concurrent GPUs, bench at N = 4 and 8.
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/r2_multi_env.txt
nvidia-smi topo -m >> gpurun_out/r2_multi_env.txt 2>&1
lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" >> gpurun_out/r2_multi_env.txt
timeout 600 python -m pytest tests/test_nccl_gpu.py -x -q -m gpu > gpurun_out/r2_nccl_test.log 2>&1
echo "== nccl test rc=$?"; tail -3 gpurun_out/r2_nccl_test.log
: > gpurun_out/r2_h2d_concurrent.jsonl
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      tools/h2d_concurrent.py >> gpurun_out/r2_h2d_concurrent.jsonl 2> gpurun_out/r2_h2d_$n.err
  echo "== h2d n=$n rc=$?"
done
cat gpurun_out/r2_h2d_concurrent.jsonl
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) \
      bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "== bench n=$n rc=$?"
  python - gpurun_out/r2_bench_n$n.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    print("value %.0f e2e %.0f ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    for k, v in (d.get("configs") or {}).items():
        print("   ", k, json.dumps(v)[:300])
except Exception as e:
    print("no line:", e)
PY
done
