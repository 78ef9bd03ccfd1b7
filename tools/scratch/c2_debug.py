import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import bench_configs as bc
from esa_pose_estimation_b200 import pipeline
from synth import ESA_K
n = int(os.environ.get("NFRAMES", "3000"))
hm, bbox, rate, model = bc.heatmap_batch(n, 11, 384, 2)
K = torch.from_numpy(ESA_K).to(bc.DEV)
for cf in (384, 750, 10**9):
    ms = bc.timed(lambda: pipeline.poses_from_heatmaps(hm, bbox, rate, model, K, min_k=8, chunk_frames=cf)["pose7"], 5)
    print("frames", n, "chunk_frames", cf, "ms %.3f" % ms)
from esa_pose_estimation_b200 import _lib
lib = _lib.load()
lib.epb_profile_enable(1)
for cf in (384, 10**9):
    ms = bc.timed(lambda: pipeline.poses_from_heatmaps(hm, bbox, rate, model, K, min_k=8, chunk_frames=cf)["pose7"], 5)
    print("PROFILING ON frames", n, "chunk_frames", cf, "ms %.3f" % ms)
lib.epb_profile_enable(0)
