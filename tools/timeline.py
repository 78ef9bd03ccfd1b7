"""Diagnostic (not product): event timeline of the chunked host-input pipeline."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from esa_pose_estimation_b200 import _lib, pipeline, ransac_voting_gpu as rv
from tests.synth import ESA_K

class A: batch=64; size=256; vn=11; hn=512; fg=0.25
a = A()
dev = torch.device("cuda", 0)
mask_np, vertex_np, model_np, geom_np, _ = bench.make_batch_numpy(a, 11, 8)
mask_np, vertex_np, geom_np = (bench.tile_to(x, 64) for x in (mask_np, vertex_np, geom_np))
m_h, v_h = torch.from_numpy(mask_np).pin_memory(), torch.from_numpy(vertex_np).pin_memory()
model = torch.from_numpy(model_np).to(dev); K = torch.from_numpy(ESA_K).to(dev)
bbox = torch.from_numpy(np.ascontiguousarray(geom_np[:, :2])).to(dev); rate = torch.from_numpy(np.ascontiguousarray(geom_np[:, 2])).to(dev)
vh = rv.vertex_layer_reshape(v_h)
chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 4

# instrument: wrap _run to record events
marks = []
orig = rv._run
def wrapped(*args, **kw):
    st = kw.get("stage", 0)
    s = torch.cuda.current_stream()
    e0 = torch.cuda.Event(enable_timing=True); e0.record(s)
    r = orig(*args, **kw)
    e1 = torch.cuda.Event(enable_timing=True); e1.record(s)
    marks.append((("all", "gather", "vote")[st], e0, e1))
    return r
rv._run = wrapped
if os.environ.get("EPB_NO_POSE"):
    def _nopose(kpts, *a, **k):
        return {"pose7": kpts.new_zeros((kpts.shape[0], 7))}
    pipeline.poses_from_keypoints = _nopose
def step():
    return pipeline.poses_from_vertex(m_h, vh, model, K, round_hyp_num=512, bbox_xy=bbox, rate=rate, sync_rng=False, chunks=chunks)
for _ in range(3): step()
torch.cuda.synchronize(); marks.clear()
def _cg():
    out = {}
    for f in ("/sys/fs/cgroup/cpu.stat", "/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.stat", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try: out[f] = open(f).read().strip().replace("\n", "; ")
        except OSError: pass
    return out
print("affinity", len(os.sched_getaffinity(0)), "cpu_count", os.cpu_count(), "torch threads", torch.get_num_threads())
cg0 = _cg()
import gc
if os.environ.get("EPB_GC_FREEZE"):
    gc.collect(); gc.freeze()
t0 = torch.cuda.Event(enable_timing=True); t0.record()
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import time
cpu0 = time.perf_counter(); cpu = []
for _ in range(nsteps):
    step(); cpu.append((time.perf_counter() - cpu0) * 1e3)
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append(("step_end", e, e))
torch.cuda.synchronize()
if nsteps > 4:
    ends = [t0.elapsed_time(e0) for n, e0, e1 in marks if n == "step_end"]
    cg1 = _cg()
    for k in cg0: print(k, "\n   before:", cg0[k], "\n   after: ", cg1[k])
    import numpy as _np
    gd, cd = _np.diff([0] + ends), _np.diff([0] + cpu)
    print("gpu step: mean %.2f max %.2f (at %d) | cpu enqueue/step: mean %.2f max %.2f (at %d) | total %.1f ms for %d steps"
          % (gd.mean(), gd.max(), gd.argmax(), cd.mean(), cd.max(), cd.argmax(), ends[-1], nsteps))
    if len(sys.argv) > 3:
        print("gpu step ends:", " ".join("%.2f" % x for x in ends))
        print("cpu enqueue  :", " ".join("%.2f" % x for x in cpu))
    g = [e0.elapsed_time(e1) for n, e0, e1 in marks if n == "gather"]; v = [e0.elapsed_time(e1) for n, e0, e1 in marks if n == "vote"]
    print("gather dur mean %.3f max %.3f | vote dur mean %.3f max %.3f" % (_np.mean(g), _np.max(g), _np.mean(v), _np.max(v)))
    sys.exit(0)
for name, e0, e1 in marks:
    print("%-9s start %7.3f  end %7.3f  dur %6.3f" % (name, t0.elapsed_time(e0), t0.elapsed_time(e1), e0.elapsed_time(e1)))
