#!/usr/bin/env python
"""BASELINE config[2]: the synthetic SPEED validation set -- 3000 frames, 11 x 384x384 heatmaps each --
sharded by frame across the GPUs of one box, one NCCL all_gather of the [3000, 7] poses.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \\
        tools/bench_c3_multi.py
Each rank holds its 375 frames (2.4 GB of heatmaps) in HBM and runs them as calls of EPB_C3_PER_CALL frames (default:
all 375 in one call -- the pose solve is latency-bound, so one call of 375 costs little more than one of 125; decode ->
select / un-crop -> EPnP-RANSAC -> LM), then the poses are gathered; timed with CUDA events between barriers,
maximum over the ranks; rank 0 prints one JSON line.  One pass = the whole 3000-frame set."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import tools.bench_configs as bc
    bc.DEV = dev
    from esa_pose_estimation_b200 import pipeline
    from tests.synth import ESA_K
    total, per_call = 3000, int(os.environ.get("EPB_C3_PER_CALL", "375"))
    s, e = pipeline.shard_range(total, rank, world)
    n = e - s
    hm, bbox, rate, model = bc.heatmap_batch(n, 11, 384, 100 + rank)
    K = torch.from_numpy(ESA_K).to(dev)

    def one_pass():
        outs = [pipeline.poses_from_heatmaps(hm[i:i + per_call], bbox[i:i + per_call], rate[i:i + per_call], model, K, min_k=8)["pose7"]
                for i in range(0, n, per_call)]
        return pipeline.gather_poses(torch.cat(outs, 0), total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(3):
        poses = one_pass()
    barrier()
    assert poses.shape == (total, 7) and torch.isfinite(poses).all()
    passes = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(passes):
        one_pass()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / passes
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank == 0:
        print(json.dumps({"config": "C3: 3000 frames x 11 x 384x384 heatmaps sharded across %d GPU(s), calls of %d frames, "
                                    "one NCCL all_gather of the poses" % (world, per_call),
                          "metric": "poses/sec (heatmaps -> refined pose)", "unit": "poses/s", "n_gpus": world,
                          "value": total / (ms * 1e-3), "ms_per_3000_frames": ms,
                          "heatmap_bytes_per_gpu": int(hm.numel() * 4), "gpu": torch.cuda.get_device_name(local)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
