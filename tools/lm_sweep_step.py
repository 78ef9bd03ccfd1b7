#!/usr/bin/env python
"""One LM sweep launch (config[4], 10^5 poses) for an ncu capture of lm_kernel:
   ncu --set full -k regex:lm_kernel -s 3 -c 1 python tools/lm_sweep_step.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch  # noqa: E402
import bench_configs as bc  # noqa: E402

step, _ = bc.make_c5_step(int(os.environ.get("POSES", "100000")))
for _ in range(6):
    step()
torch.cuda.synchronize()
