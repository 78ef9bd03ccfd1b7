"""Launched by tests/test_nccl_gpu.py under torch.distributed.run (one rank per GPU, NCCL): every rank computes poses
for ITS shard of one common batch through the public API, the ranks all_gather them (pipeline.gather_poses), and the
gathered [N,7] block must equal the single-process result on the whole batch, row for row, on every rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from esa_pose_estimation_b200 import pipeline
    from tests.synth import ESA_K, make_pose_case, tango_model
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 37                                            # not divisible by the world size: ragged shards
    model = tango_model(11, seed=9)
    kp = np.stack([make_pose_case(4200 + i, 11, 0.4, i % 2, model=model)["p2d"] for i in range(n)]).astype(np.float32)
    m_t, K_t = torch.from_numpy(model).to(dev), torch.from_numpy(ESA_K).to(dev)
    whole = pipeline.poses_from_keypoints(torch.from_numpy(kp).to(dev), m_t, K_t)["pose7"]
    s, e = pipeline.shard_range(n, rank, world)
    local_p = pipeline.poses_from_keypoints(torch.from_numpy(kp[s:e]).to(dev), m_t, K_t)["pose7"]
    got = pipeline.gather_poses(local_p, n)
    assert got.shape == (n, 7), got.shape
    assert torch.equal(got[s:e], local_p), "rank %d: own slice differs" % rank
    assert torch.equal(got, whole), "rank %d: gathered poses differ from the single-process result" % rank
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("NCCL_GATHER_OK world=%d" % world)


if __name__ == "__main__":
    main()
