import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import bench_configs as bc
r = bc.c3()
print({k: r[k] for k in ("value", "ms_per_call", "kernel_ms")})
from esa_pose_estimation_b200 import pipeline, inference
from synth import ESA_K
hm, bbox, rate, model = bc.heatmap_batch(125, 11, 384, 2)
K = torch.from_numpy(ESA_K).to(bc.DEV)
out = pipeline.poses_from_heatmaps(hm, bbox, rate, model, K, min_k=8)
st = out["status"].cpu().numpy()
print("status counts", np.unique(st, return_counts=True))
print("keys", list(out.keys()))
