"""End-to-end post-network pose path, batched and sharded by image across the GPUs of one box.

  poses_from_heatmaps   val.py:146-228   heatmaps -> argmax + sub-pixel -> select/un-crop -> EPnP-RANSAC -> LM
  poses_from_vertex     PVNet path       mask + vector field -> RANSAC voting -> EPnP-RANSAC -> LM
  shard_range / gather_poses             images shard by index; the only collective is one
                                         all_gather of the [n,7] float32 poses
Everything between the network output and the pose stays on the device; the kernels are launched
on torch's current stream and nothing synchronises the host.
"""
import torch

from . import inference, pnp as _pnp, ransac_voting_gpu as _voting


def poses_from_heatmaps(hms, bbox_xy, rate, p3d_model, K, min_k=24, sel_thresh=0.8, weighted=True):
    """hms [B,Kp,H,W] CUDA f32; bbox_xy [B,2]; rate [B]; p3d_model [Kp,3]; K [3,3].
    -> dict(pose7 [B,7] f32 (qw,qx,qy,qz,tx,ty,tz), rt6, epnp_rt34, status, xy, maxval)."""
    xy, maxval, _ = inference.decode_heatmaps(hms, refine=True)
    out = _pnp.pose_pipeline(xy, maxval, bbox_xy, rate, p3d_model, K, min_k=min_k, sel_thresh=sel_thresh,
                             weighted=weighted)
    out["xy"], out["maxval"] = xy, maxval
    return out


def poses_from_keypoints(kpts, p3d_model, K, bbox_xy=None, rate=None, weights=None):
    """kpts [B,vn,2] f32 (crop px).  Unit (or given) weights; all keypoints are used."""
    b, vn = kpts.shape[0], kpts.shape[1]
    dev = kpts.device
    if bbox_xy is None:
        bbox_xy = torch.zeros((b, 2), dtype=torch.float64, device=dev)
    if rate is None:
        rate = torch.ones((b,), dtype=torch.float64, device=dev)
    weighted = weights is not None
    if weights is None:
        weights = torch.ones((b, vn), dtype=torch.float32, device=dev)
    # min_k = vn selects every keypoint (large_k = max(#(w > thresh), vn) capped at vn)
    return _pnp.pose_pipeline(kpts, weights, bbox_xy, rate, p3d_model, K, min_k=vn, sel_thresh=float("inf"),
                              weighted=weighted)


def poses_from_vertex(mask, vertex, p3d_model, K, round_hyp_num=512, inlier_thresh=0.999, min_num=5,
                      max_num=30000, bbox_xy=None, rate=None, **kw):
    """mask [B,H,W], vertex [B,H,W,vn,2] (e.g. vertex_layer_reshape of the NCHW network output).
    -> dict(pose7, rt6, epnp_rt34, status, kpts)."""
    kpts = _voting.ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=inlier_thresh,
                                          min_num=min_num, max_num=max_num, **kw)
    out = poses_from_keypoints(kpts, p3d_model, K, bbox_xy=bbox_xy, rate=rate)
    out["kpts"] = kpts
    return out


# ------------------------------------------------------------------------------- multi-GPU
def shard_range(n, rank, world_size):
    """Contiguous block of image indices owned by `rank` (blocks differ by at most one image)."""
    base, rem = divmod(n, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_poses(pose7_local, n_total, group=None):
    """All-gather of the per-rank [n_r,7] float32 poses into [n_total,7] in image order.
    One collective (NCCL over NVLink on GPUs, gloo in the CPU tests); payload 28 B/pose."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return pose7_local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_total + world - 1) // world
    buf = torch.zeros((per, 7), dtype=pose7_local.dtype, device=pose7_local.device)
    s, e = shard_range(n_total, rank, world)
    assert pose7_local.shape[0] == e - s, "local shard does not match shard_range"
    buf[: e - s] = pose7_local
    out = torch.empty((world * per, 7), dtype=pose7_local.dtype, device=pose7_local.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    parts = []
    for r in range(world):
        rs, re_ = shard_range(n_total, r, world)
        parts.append(out[r * per: r * per + (re_ - rs)])
    return torch.cat(parts, 0)
