"""Diagnostic (not product): which host call blocks for ~100 ms in the host-input pipeline?"""
import sys, os, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from esa_pose_estimation_b200 import _lib, pipeline, pnp as gp, ransac_voting_gpu as rv
from tests.synth import ESA_K

class A: batch=64; size=256; vn=11; hn=512; fg=0.25
a = A(); dev = torch.device("cuda", 0)
mask_np, vertex_np, model_np, geom_np, _ = bench.make_batch_numpy(a, 11, 8)
mask_np, vertex_np, geom_np = (bench.tile_to(x, 64) for x in (mask_np, vertex_np, geom_np))
m_h, v_h = torch.from_numpy(mask_np).pin_memory(), torch.from_numpy(vertex_np).pin_memory()
model = torch.from_numpy(model_np).to(dev); K = torch.from_numpy(ESA_K).to(dev)
bbox = torch.from_numpy(np.ascontiguousarray(geom_np[:, :2])).to(dev); rate = torch.from_numpy(np.ascontiguousarray(geom_np[:, 2])).to(dev)
vh = rv.vertex_layer_reshape(v_h)
slow = []
def wrap(obj, name, label=None):
    f = getattr(obj, name)
    def g(*args, **kw):
        t0 = time.perf_counter(); r = f(*args, **kw); dt = (time.perf_counter() - t0) * 1e3
        if dt > 3.0: slow.append((label or name, round(dt, 1)))
        return r
    setattr(obj, name, g)
lib = _lib.load()
class LibProxy:
    def __init__(self, l): self._l = l
    def __getattr__(self, n):
        f = getattr(self._l, n)
        def g(*args):
            t0 = time.perf_counter(); r = f(*args); dt = (time.perf_counter() - t0) * 1e3
            if dt > 3.0: slow.append(("C:" + n, round(dt, 1)))
            return r
        return g
_lib._lib = LibProxy(lib)
wrap(torch.Tensor, "to"); wrap(torch, "full"); wrap(torch, "cat"); wrap(torch, "empty"); wrap(torch, "zeros")
wrap(torch.cuda.Event, "record", "Event.record"); wrap(torch.cuda.Stream, "wait_event", "Stream.wait_event")
wrap(torch.Tensor, "record_stream"); wrap(torch.Tensor, "contiguous")
def step():
    return pipeline.poses_from_vertex(m_h, vh, model, K, round_hyp_num=512, bbox_xy=bbox, rate=rate, sync_rng=False, chunks=4)
for _ in range(3): step()
torch.cuda.synchronize(); slow.clear()
t0 = time.perf_counter(); per = []
for i in range(60):
    s0 = time.perf_counter(); step(); per.append((time.perf_counter() - s0) * 1e3)
torch.cuda.synchronize()
print("total %.1f ms; step cpu max %.1f at %d" % ((time.perf_counter() - t0) * 1e3, max(per), int(np.argmax(per))))
print("slow calls:", slow[:20])
