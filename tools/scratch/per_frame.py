import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests"); sys.path.insert(0, "/root/repo/tools")
import numpy as np, torch
from synth import make_pose_case, tango_model, ESA_K
from esa_pose_estimation_b200 import pipeline
DEV = torch.device("cuda:0")
model = tango_model(11, seed=9)
K = torch.from_numpy(ESA_K).to(DEV); m = torch.from_numpy(model).to(DEV)
for n_out in (0, 1, 2):
    kp = np.stack([make_pose_case(6000 + i, 11, 0.5, n_out, model=model)["p2d"] for i in range(64)])
    ts = []
    for i in range(64):
        k_t = torch.from_numpy(kp[i:i + 1].astype(np.float32)).to(DEV)
        for _ in range(3): pipeline.poses_from_keypoints(k_t, m, K)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): pipeline.poses_from_keypoints(k_t, m, K)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) / 10 * 1000)
    ts = np.array(ts)
    print("outliers", n_out, "us per frame: min %.0f median %.0f p90 %.0f max %.0f" % (ts.min(), np.median(ts), np.percentile(ts, 90), ts.max()), "argmax", int(ts.argmax()))
