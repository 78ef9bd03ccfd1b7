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


_pose_streams = {}


def poses_from_heatmaps(hms, bbox_xy, rate, p3d_model, K, min_k=24, sel_thresh=0.8, weighted=True, chunk_frames=384):
    """hms [B,Kp,H,W] CUDA f32; bbox_xy [B,2]; rate [B]; p3d_model [Kp,3]; K [3,3].
    -> dict(pose7 [B,7] f32 (qw,qx,qy,qz,tx,ty,tz), rt6, epnp_rt34, status, xy, maxval).

    Sets of more than `chunk_frames` frames go through in pieces: the decode of piece i+1 (HBM-bound, the caller's
    stream) runs while a second, high-priority stream solves the poses of piece i (latency-bound, a few small CTAs
    per SM), so a long set costs little more than reading its heatmaps once.  Results are valid on the caller's
    stream and identical to the one-piece call."""
    b = hms.shape[0]
    if b <= chunk_frames or not hms.is_cuda or torch.cuda.is_current_stream_capturing():
        xy, maxval, _ = inference.decode_heatmaps(hms, refine=True)
        out = _pnp.pose_pipeline(xy, maxval, bbox_xy, rate, p3d_model, K, min_k=min_k, sel_thresh=sel_thresh,
                                 weighted=weighted)
        out["xy"], out["maxval"] = xy, maxval
        return out
    dev = hms.device
    cur = torch.cuda.current_stream(dev)
    st = _pose_streams.get((dev.index, cur.cuda_stream))
    if st is None:
        st = _pose_streams[(dev.index, cur.cuda_stream)] = {
            "ps": torch.cuda.Stream(device=dev, priority=-1), "done": [None, None], "turn": 0}
    ps = st["ps"]
    # at most two chunked calls in flight per caller stream: a caller that never consumes its results would otherwise
    # run ahead until a launch queue is full, and the driver then parks the thread for 50-100 ms (see _HostPipe)
    if st["done"][st["turn"]] is not None:
        st["done"][st["turn"]].synchronize()
    ready = torch.cuda.Event()
    ready.record(cur)                        # bbox / rate / model / K of the caller's stream
    ps.wait_event(ready)
    parts = []
    for s in range(0, b, chunk_frames):
        e = min(b, s + chunk_frames)
        xy, maxval, _ = inference.decode_heatmaps(hms[s:e], refine=True)
        decoded = torch.cuda.Event()
        decoded.record(cur)
        ps.wait_event(decoded)
        with torch.cuda.stream(ps):
            o = _pnp.pose_pipeline(xy, maxval, bbox_xy[s:e], rate[s:e], p3d_model, K, min_k=min_k, sel_thresh=sel_thresh,
                                   weighted=weighted)
        xy.record_stream(ps); maxval.record_stream(ps)
        o["xy"], o["maxval"] = xy, maxval
        parts.append(o)
    with torch.cuda.stream(ps):
        out = {k: torch.cat([p[k] for p in parts], 0) for k in parts[0]}
        done = torch.cuda.Event()
        done.record(ps)
    cur.wait_event(done)
    st["done"][st["turn"]] = done
    st["turn"] ^= 1
    for t in out.values():
        t.record_stream(cur)
    return out


class GraphedHeatmapPose:
    """poses_from_heatmaps for ONE fixed shape, captured once as a CUDA graph and replayed: the call is a short
    chain of kernels (decode -> sub-pixel refine -> pose), so at small batches the launches themselves are a
    visible part of the latency.  Nothing in the chain draws random numbers or synchronises, so a replay is
    bit-identical to the eager call.

        g = GraphedHeatmapPose(B, Kp, H, W, p3d_model, K, device)
        out = g(hms, bbox_xy, rate)        # copies into the static inputs, replays; dict of STATIC output tensors

    The returned tensors are reused by the next call (clone what must outlive it)."""

    def __init__(self, b, kp, h, w, p3d_model, K, device, min_k=24, sel_thresh=0.8, weighted=True):
        dev = torch.device(device)
        self.hms = torch.zeros((b, kp, h, w), dtype=torch.float32, device=dev)
        self.bbox = torch.zeros((b, 2), dtype=torch.float64, device=dev)
        self.rate = torch.ones((b,), dtype=torch.float64, device=dev)
        self.model = p3d_model.to(device=dev, dtype=torch.float64).contiguous()
        self.K = K.to(device=dev, dtype=torch.float64).contiguous()
        self.args = dict(min_k=min_k, sel_thresh=sel_thresh, weighted=weighted)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):        # warm-up outside the capture: lazy one-time initialisation of the library
            for _ in range(2):
                poses_from_heatmaps(self.hms, self.bbox, self.rate, self.model, self.K, **self.args)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = poses_from_heatmaps(self.hms, self.bbox, self.rate, self.model, self.K, **self.args)

    def __call__(self, hms, bbox_xy, rate, copy_inputs=True):
        if copy_inputs:
            self.hms.copy_(hms, non_blocking=True)
            self.bbox.copy_(bbox_xy, non_blocking=True)
            self.rate.copy_(rate, non_blocking=True)
        self.graph.replay()
        return self.out


_unit_weights = {}


def poses_from_keypoints(kpts, p3d_model, K, bbox_xy=None, rate=None, weights=None):
    """kpts [B,vn,2] f32 (crop px).  Unit (or given) weights; all keypoints are used."""
    b, vn = kpts.shape[0], kpts.shape[1]
    dev = kpts.device
    if bbox_xy is None:
        bbox_xy = torch.zeros((b, 2), dtype=torch.float64, device=dev)
    if rate is None:
        rate = torch.ones((b,), dtype=torch.float64, device=dev)
    weighted = weights is not None
    if weights is None:                      # read-only constant: one tensor per shape instead of a fill kernel per call
        key = (dev.index, b, vn)
        weights = _unit_weights.get(key)
        if weights is None:
            if len(_unit_weights) > 64:
                _unit_weights.clear()
            weights = _unit_weights[key] = torch.ones((b, vn), dtype=torch.float32, device=dev)
    # min_k = vn selects every keypoint (large_k = max(#(w > thresh), vn) capped at vn)
    return _pnp.pose_pipeline(kpts, weights, bbox_xy, rate, p3d_model, K, min_k=vn, sel_thresh=float("inf"),
                              weighted=weighted)


def poses_from_vertex(mask, vertex, p3d_model, K, round_hyp_num=512, inlier_thresh=0.999, min_num=5,
                      max_num=30000, bbox_xy=None, rate=None, chunks=None, pipelined=None, host_inputs_ready=False, **kw):
    """mask [B,H,W], vertex [B,H,W,vn,2] (e.g. vertex_layer_reshape of the NCHW network output).
    -> dict(pose7, rt6, epnp_rt34, status, kpts).

    `vertex` (and `mask`) may live in PINNED host memory.  The batch is then cut into `chunks`
    pieces (default 2 when B >= 16; measured 2.29 ms per 64-image call with 2, 2.35 with 1, 2.49 with 3) and flows through three streams: a high-priority stream compacts
    the foreground and reads the field of piece i+1 in place over PCIe while a second stream votes
    piece i; a third (high-priority) stream solves the poses once the keypoints are there, so the
    latency-bound pose solve of one call runs under the voting of the next.  The caller's stream waits
    for the poses: results are valid on the caller's stream, as usual.  Pinned host inputs are READ BY THE
    GPU after this call returns: leave them untouched until the caller's stream has passed the returned
    tensors (e.g. `torch.cuda.current_stream().synchronize()` or an event recorded after the call).

    Ordering of the inputs.  By default the gather stream is ordered after everything queued on the caller's
    stream (a `copy_(non_blocking=True)` that fills the pinned buffers, the producer of device-resident
    inputs).  The caller's stream also carries the wait for the PREVIOUS call's poses, so that ordering
    serialises consecutive calls.  `host_inputs_ready=True` states that pinned host inputs were complete
    before the call (written by the host, or by a copy the host has synchronised): the gather of this call
    then starts at once, under the voting of the previous call (27.6 k instead of 17.6 k poses/s on
    config[1]).  It is ignored for device-resident `mask` / `vertex` / `idxs`, which are always ordered.

    `pipelined=True` sends DEVICE-resident inputs down the same route (one piece).  Within one caller
    stream that changes nothing (the inputs of call n+1 are ordered after the results of call n), but
    calls issued on several caller streams then overlap: each (device, caller stream) pair owns its own
    workspaces, so several batches can be in flight (bench.py: 2.10 -> 1.9 ms per 64-image call with three).
    The default for device inputs keeps every kernel on the caller's stream, in ONE shared workspace."""
    host = isinstance(vertex, torch.Tensor) and not vertex.is_cuda
    b = vertex.shape[0]
    if chunks is None:
        chunks = 2 if (host and b >= 16) else 1
    chunks = max(1, min(int(chunks), b))
    if pipelined is None:
        pipelined = host
    if not host and not pipelined:
        kpts = _voting.ransac_voting_layer_v3(mask, vertex, round_hyp_num, inlier_thresh=inlier_thresh,
                                              min_num=min_num, max_num=max_num, **kw)
        out = poses_from_keypoints(kpts, p3d_model, K, bbox_xy=bbox_xy, rate=rate)
        out["kpts"] = kpts
        return out
    return _poses_from_host_vertex(mask, vertex, p3d_model, K, round_hyp_num, inlier_thresh, min_num, max_num,
                                   bbox_xy, rate, chunks, kw, host_inputs_ready)


class _HostPipe:
    """Per (device, stream) state of the host-input pipeline: the high-priority gather stream, the voting
    stream and a ring of DEPTH workspace sets, so that the gather of call n+1 (and the host-side enqueue of
    call n+2) proceed while call n still votes."""
    DEPTH = 2

    def __init__(self, dev):
        self.gs = torch.cuda.Stream(device=dev, priority=-1)
        self.vs = torch.cuda.Stream(device=dev)              # voting
        self.ps = torch.cuda.Stream(device=dev, priority=-1)  # pose solve: its few CTAs go ahead of queued vote CTAs
        self.sets = [None] * self.DEPTH       # each: dict(ws=[tensors], done=[events], need=int)
        self.finished = [None] * self.DEPTH   # event after the pose solve of the call that last used the set
        self.turn = 0


def poses_from_vertex_uncertainty(mask, vertex, p3d_model, K, round_hyp_num=256, min_hyp_num=4096, topk=128,
                                  inlier_thresh=0.99, min_num=5, max_num=30000, **kw):
    """PVNet's uncertainty-driven path (estimate_voting_distribution -> inv(sqrtm(cov)) weights ->
    uncertainty PnP; ransac_voting_gpu.py:263-331, evaluation_utils.py:165-188), batched on the device.
    -> dict(rt34 [B,3,4], rt6, pose7, mean, cov, weights)."""
    mean, cov = _voting.estimate_voting_distribution(mask, vertex, round_hyp_num=round_hyp_num,
                                                     min_hyp_num=min_hyp_num, topk=topk, inlier_thresh=inlier_thresh,
                                                     min_num=min_num, max_num=max_num, **kw)
    rt34, rt6, iters, cost = _pnp.uncertainty_pnp_batch(mean, cov, p3d_model, K, return_info=True)
    pose7, _ = _pnp.pose_pack(rt6)
    return dict(rt34=rt34, rt6=rt6, pose7=pose7, mean=mean, cov=cov, weights=_pnp.cov_to_weights(cov),
                lm_iters=iters, lm_cost=cost)


_host_pipes = {}


def _poses_from_host_vertex(mask, vertex, p3d_model, K, hn, thresh, min_num, max_num, bbox_xy, rate, chunks, kw,
                            host_inputs_ready=False):
    from . import _lib
    dev = p3d_model.device if p3d_model.is_cuda else torch.device("cuda", torch.cuda.current_device())
    b, h, w, vn, _ = vertex.shape
    cur = torch.cuda.current_stream(dev)
    key = (dev.index, cur.cuda_stream)
    pipe = _host_pipes.get(key)
    if pipe is None:
        pipe = _host_pipes[key] = _HostPipe(dev)
    gs, vs = pipe.gs, pipe.vs
    bounds = [shard_range(b, i, chunks) for i in range(chunks)]
    per = max(e - s for s, e in bounds)
    need = _voting.workspace_bytes(per, h, w, vn, hn, max_num=max_num, philox=kw.get("idxs") is None)
    turn = pipe.turn
    # At most DEPTH calls in flight: the host waits (spinning on an event) for the call that used this
    # workspace set DEPTH turns ago.  Without it a host that enqueues faster than the GPU drains fills the
    # gather stream's launch queue, and the driver then parks the enqueueing thread for up to ~100 ms
    # (measured: tools/stallhunt.py), long enough to starve the GPU.
    if pipe.finished[turn] is not None:
        pipe.finished[turn].synchronize()
    st = pipe.sets[turn]
    fresh_ws = False
    if st is None or len(st["ws"]) < chunks or st["need"] < need:
        st = pipe.sets[turn] = dict(
            ws=[torch.empty((need,), dtype=torch.uint8, device=dev) for _ in range(chunks)],
            done=[None] * chunks, need=need)
        fresh_ws = True          # blocks from the caller stream's pool: order the side streams after it once
    pipe.turn = (turn + 1) % pipe.DEPTH
    wss = st["ws"]
    kw = dict(kw)
    sync_rng = kw.pop("sync_rng", True)
    per_image = {k: kw.pop(k) for k in ("idxs", "selection") if kw.get(k) is not None}
    philox = "idxs" not in per_image
    gen = rng_state = None
    if philox:                                # one torch generator stream across the chunks
        gen = torch.cuda.default_generators[dev.index]
        start = gen.get_offset()
        with torch.cuda.stream(gs):           # lives on the gather stream: no ordering against `cur` needed
            rng_state = torch.full((1,), start, dtype=torch.int64, device=dev)
        kw["rng_state"] = rng_state

    def chunk_kw(s, e):
        d = dict(kw)
        for k, t in per_image.items():
            d[k] = t[s:e]
        return d

    # What the side streams touch is ordered after the caller's stream: device-resident arguments were produced
    # there, pinned HOST buffers may still be the target of a non_blocking copy_ queued there, and freshly
    # allocated workspaces come from the caching allocator's pool of that stream (a block it just freed may have
    # work pending on `cur`).  `cur` also waits for the previous call's poses, so this ordering serialises
    # consecutive calls; `host_inputs_ready=True` (pinned inputs complete before the call) skips it for host inputs.
    ready = torch.cuda.Event()
    ready.record(cur)
    if mask.is_cuda or vertex.is_cuda or per_image or fresh_ws or not host_inputs_ready:
        gs.wait_event(ready)
        vs.wait_event(ready)
    events, masks_d = [], []
    with torch.cuda.stream(gs):
        for i, (s, e) in enumerate(bounds):
            if st["done"][i] is not None:
                gs.wait_event(st["done"][i])          # the call DEPTH turns ago has consumed this workspace
            m = mask[s:e]
            m_d = m if m.is_cuda else m.to(dev, non_blocking=True)
            _voting._run(_lib.VOTE_V3, m_d, vertex[s:e], hn, 1, thresh, min_num, max_num,
                         stage=_lib.STAGE_GATHER, workspace=wss[i], **chunk_kw(s, e))
            ev = torch.cuda.Event()
            ev.record(gs)
            events.append(ev)
            masks_d.append(m_d)
    kps = []
    with torch.cuda.stream(vs):
        for i, (s, e) in enumerate(bounds):
            vs.wait_event(events[i])
            kps.append(_voting._run(_lib.VOTE_V3, masks_d[i], vertex[s:e], hn, 1, thresh, min_num, max_num,
                                    stage=_lib.STAGE_VOTE, workspace=wss[i], **chunk_kw(s, e))["pts"])
            done = torch.cuda.Event()
            done.record(vs)
            st["done"][i] = done
        kpts = torch.cat(kps, 0)
        voted = torch.cuda.Event()
        voted.record(vs)
    for m_d in masks_d:
        m_d.record_stream(vs)
    if vertex.is_cuda:                        # device-resident field: read by the gather stream only
        vertex.record_stream(gs)
    if mask.is_cuda:
        mask.record_stream(gs)
    # the pose solve is latency-bound (one warp per image): one launch over the whole batch, on the caller's stream
    ps = pipe.ps
    if ready is not None:
        ps.wait_event(ready)                  # model / K / bbox / rate of the caller's stream
    ps.wait_event(voted)
    with torch.cuda.stream(ps):
        out = poses_from_keypoints(kpts, p3d_model, K, bbox_xy=bbox_xy, rate=rate)
        fin = torch.cuda.Event()
        fin.record(ps)
    kpts.record_stream(ps)
    out["kpts"] = kpts
    cur.wait_event(fin)
    for t in out.values():
        t.record_stream(cur)
    pipe.finished[turn] = fin
    if philox:
        if sync_rng:                          # exact: what the reference's loop would have consumed
            cur.wait_stream(gs)
            gen.set_offset(int(rng_state.item()))
        else:                                 # data-independent upper bound, no synchronisation
            props = torch.cuda.get_device_properties(dev)
            inc = _voting._rng_layout(hn * vn * 2, props) + (_voting._rng_layout(h * w, props) if h * w > max_num else 0)
            gen.set_offset(start + b * inc)
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
