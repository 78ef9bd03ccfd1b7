"""Diagnostic (not product): what slows vote_count when something else is resident on the SMs?"""
import ctypes, os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "micro", "libcorun.so"))
lib.launch_spin.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_longlong, ctypes.c_void_p]
lib.launch_zc.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
class A: batch=16; size=256; vn=11; hn=512; fg=0.25
dev = torch.device("cuda", 0)
mask_np, vertex_np, *_ = bench.make_batch_numpy(A(), 11, 8)
mask_np, vertex_np = bench.tile_to(mask_np, 16), bench.tile_to(vertex_np, 16)
m = torch.from_numpy(mask_np).to(dev); v = rv.vertex_layer_reshape(torch.from_numpy(vertex_np).to(dev))
host = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
side = torch.cuda.Stream(priority=-1)
def vote_ms(n=10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): rv.ransac_voting_layer_v3(m, v, 512, sync_rng=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
if len(sys.argv) > 1:
    print("carveout preference of the co-running kernels:", sys.argv[1], lib.set_carveout(int(sys.argv[1])))
for _ in range(3): vote_ms(2)
print("vote chunk (16 images) alone: %.3f ms" % vote_ms())
for name, fn in [("spin 148x256 thr", lambda: lib.launch_spin(148, 256, 40_000_000, ctypes.c_void_p(side.cuda_stream))),
                 ("spin 148x64 thr", lambda: lib.launch_spin(148, 64, 40_000_000, ctypes.c_void_p(side.cuda_stream))),
                 ("zero-copy read 148x256", lambda: lib.launch_zc(ctypes.c_void_p(host.data_ptr()), host.numel() // 4, 148, 256, 4, ctypes.c_void_p(side.cuda_stream))),
                 ("zero-copy read 37x256", lambda: lib.launch_zc(ctypes.c_void_p(host.data_ptr()), host.numel() // 4, 37, 256, 4, ctypes.c_void_p(side.cuda_stream)))]:
    torch.cuda.synchronize(); fn()
    t = vote_ms()
    torch.cuda.synchronize()
    print("vote chunk with %-26s: %.3f ms" % (name, t))
