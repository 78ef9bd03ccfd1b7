#!/bin/bash
# GPU job (gpurun --gpus 8): bench at N = 4 and 8 for the final build.
mkdir -p gpurun_out
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) \
      bench.py --gpus $n --steps 100 --warmup 3 > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "== bench n=$n rc=$?"
  python - gpurun_out/r2_bench_n$n.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[1]).read().splitlines() if l.startswith("{")][-1])
    c = d["configs"]
    print("value %.0f e2e %.0f ms/step %.3f e2e ms %.3f pcie %.1f of %.1f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["pcie"]["achieved"], d["e2e"]["pcie"]["peak"]))
    print("   config2 %.0f poses/s (median pass %.3f ms, mean %.3f)  config4 %.0f poses/s" % (c["config2_val_set_3000_frames"]["poses_per_s"], c["config2_val_set_3000_frames"]["ms_per_pass"], c["config2_val_set_3000_frames"]["ms_per_pass_mean"], c["config4_lm_sweep"]["poses_per_s"]))
except Exception as e:
    print("no line:", e)
PY
done
