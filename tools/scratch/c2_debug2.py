import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
import bench_configs as bc
from esa_pose_estimation_b200 import pipeline, inference, pnp as P
from synth import ESA_K
n = 3000
hm, bbox, rate, model = bc.heatmap_batch(n, 11, 384, 2)
K = torch.from_numpy(ESA_K).to(bc.DEV)
xy, maxval, _ = inference.decode_heatmaps(hm, refine=True)
out = P.pose_pipeline(xy, maxval, bbox, rate, model, K, min_k=8)
st = out["status"].cpu().numpy()
print("status", np.unique(st, return_counts=True))
# time the pose kernel per block of 375 frames
for s in range(0, n, 375):
    sl = slice(s, s + 375)
    ms = bc.timed(lambda: P.pose_pipeline(xy[sl], maxval[sl], bbox[sl], rate[sl], model, K, min_k=8), 5)
    print("frames %d..%d pose %.3f ms" % (s, s + 375, ms))
# the slowest frames of the slowest block: time single frames of block with max time
mv = maxval.cpu().numpy(); xyc = xy.cpu().numpy()
print("maxval min per frame <0.8 count:", int((mv < 0.8).any(1).sum()), "min", mv.min())
np.save("gpurun_out/c2_xy.npy", xyc); np.save("gpurun_out/c2_mv.npy", mv)
np.save("gpurun_out/c2_bbox.npy", bbox.cpu().numpy()); np.save("gpurun_out/c2_rate.npy", rate.cpu().numpy())
