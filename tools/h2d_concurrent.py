#!/usr/bin/env python
"""Host -> device delivery when 1, 2, 4, 8 GPUs of one box pull their inputs at once (e2e scaling
0.61 at N = 4 and 0.47 at N = 8).  Launch under torch.distributed.run with N ranks; every rank, concurrently:

  dma_payload   cudaMemcpyAsync of the 96 MB that the zero-copy gather actually moves (mask + foreground of the field)
  dma_field     cudaMemcpyAsync of the whole pinned field (373 MB), what a plain "copy the inputs" path moves
  zero_copy     the library's gather stage reading the foreground of the pinned field in place over PCIe (what e2e uses)

Each leg: barrier, 10 repetitions timed with CUDA events, max over ranks.  Prints one JSON line (rank 0): per-GPU and
aggregate GB/s.  If `dma_payload` per GPU at N = 8 is not materially above `zero_copy`, no access path from inside the
process beats the host fabric, and the e2e scaling is the box's.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/h2d_concurrent.py
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import bench
    from esa_pose_estimation_b200 import _lib, ransac_voting_gpu as rv
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    numa = bench.bind_to_gpu_numa_node(local)
    a = SimpleNamespace(size=256, vn=11, fg=0.25, batch=64, hn=512)
    mask_np, vertex_np, _, _, _ = bench.make_batch_numpy(a, 11 + rank, 8)
    mask_np, vertex_np = bench.tile_to(mask_np, a.batch), bench.tile_to(vertex_np, a.batch)
    mask_h, vertex_h = torch.from_numpy(mask_np).pin_memory(), torch.from_numpy(vertex_np).pin_memory()
    fg_px = int(mask_np.astype(bool).sum())
    payload = mask_h.numel() + fg_px * a.vn * 8
    pay_h = torch.empty((payload,), dtype=torch.uint8).pin_memory()
    pay_d = torch.empty((payload,), dtype=torch.uint8, device=dev)
    field_d = torch.empty_like(vertex_h, device=dev)
    mask_d = mask_h.to(dev)
    ws = torch.empty((rv.workspace_bytes(a.batch, a.size, a.size, a.vn, a.hn, max_num=30000),), dtype=torch.uint8, device=dev)
    vview = rv.vertex_layer_reshape(vertex_h)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps=10):
        for _ in range(2):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / reps
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    legs = {
        "dma_payload": (lambda: pay_d.copy_(pay_h, non_blocking=True), payload),
        "dma_field": (lambda: field_d.copy_(vertex_h, non_blocking=True), vertex_h.numel() * 4),
        "zero_copy": (lambda: rv._run(_lib.VOTE_V3, mask_d, vview, a.hn, 1, 0.999, 5, 30000, stage=_lib.STAGE_GATHER,
                                      workspace=ws, sync_rng=False), payload),
    }
    out = {"n_gpus": world, "host_numa_binding": numa, "payload_bytes": payload, "field_bytes": vertex_h.numel() * 4}
    for name, (fn, nbytes) in legs.items():
        ms = timed(fn)
        out[name] = {"ms": ms, "gbs_per_gpu": nbytes / (ms * 1e-3) / 1e9, "gbs_aggregate": world * nbytes / (ms * 1e-3) / 1e9}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
