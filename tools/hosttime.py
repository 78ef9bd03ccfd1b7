import time, sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
from esa_pose_estimation_b200 import _lib, pipeline, ransac_voting_gpu as rv
from tests.synth import make_vertex_field, tango_model
dev = torch.device("cuda", 0)
B, vn, S, hn = 64, 11, 256, 512
mask, vertex, _ = make_vertex_field(1, 4, S, S, vn, 0.25)
mask = np.tile(mask, (16, 1, 1)); vertex = np.tile(vertex, (16, 1, 1, 1))
m_h, v_h = torch.from_numpy(mask).pin_memory(), torch.from_numpy(vertex).pin_memory()
m_d, v_d = m_h.to(dev), v_h.to(dev)
model = torch.from_numpy(tango_model(vn, seed=9)).to(dev)
K = torch.from_numpy(np.array([[3000.0, 0, 128], [0, 3000.0, 128], [0, 0, 1]])).to(dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / n * 1e3, (t2 - t0) / n * 1e3
vd = rv.vertex_layer_reshape(v_d); vh = rv.vertex_layer_reshape(v_h)
print("v3 device  enqueue/total ms", t(lambda: rv.ransac_voting_layer_v3(m_d, vd, hn, sync_rng=False)))
kp = rv.ransac_voting_layer_v3(m_d, vd, hn, sync_rng=False)
print("pose       enqueue/total ms", t(lambda: pipeline.poses_from_keypoints(kp, model, K)))
for c in (1, 2, 4, 8):
    print("host chunks", c, t(lambda: pipeline.poses_from_vertex(m_h, vh, model, K, round_hyp_num=hn, sync_rng=False, chunks=c)))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(20): pipeline.poses_from_vertex(m_h, vh, model, K, round_hyp_num=hn, sync_rng=False, chunks=4)
pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
