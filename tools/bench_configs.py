#!/usr/bin/env python
"""Secondary configurations of BASELINE.json (bench.py measures configs[1], the contract line):

  C1  single SPEED frame, 11-keypoint 128x128 heatmaps: decode + EPnP-RANSAC + LM   (latency)
  C3  heatmap path at 384x384, the per-GPU share of the 3000-frame validation set     (poses/s, decode GB/s)
  C4  768x768 vector fields, 2048 hypotheses/keypoint, voting distribution + uncertainty PnP
  C5  batched LM refinement sweep, 1e3..1e6 poses x 11 points

One JSON line per configuration (same conventions as bench.py: CUDA events on the launching
stream, >= 3 warm-ups, inputs resident in HBM and larger than L2 or rotated, no host sync inside).
    python tools/bench_configs.py [c1 c3 c4 c5]
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from esa_pose_estimation_b200 import _lib, inference, pipeline, pnp as gp, ransac_voting_gpu as rv  # noqa: E402
from tests.synth import ESA_K, make_pose_case, make_vertex_field, tango_model  # noqa: E402

DEV = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
PEAK = 6552.3
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timed(fn, steps, warmup=3):
    import gc
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gc.disable()          # like timeit: a generation-2 collection stalls the enqueueing thread for tens of ms
    try:
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
    finally:
        gc.enable()
    return e0.elapsed_time(e1) / steps


def prof(cls):
    tot, n = _lib.c_double(0), _lib.c_int(0)
    _lib.load().epb_profile_read(cls, tot, n)
    return tot.value, n.value


def heatmap_batch(n, kp, s, seed):
    """Heatmaps of n frames whose peaks are the projections of the model under random poses (device-side)."""
    model = tango_model(kp, seed=9)
    crop = np.zeros((n, kp, 2), np.float32); bbox = np.zeros((n, 2)); rate = np.zeros(n)
    for i in range(n):
        c = make_pose_case(seed * 7919 + i, kp, 0.0, 0, model=model)
        lo, hi = c["p2d"].min(0), c["p2d"].max(0)
        size = (hi - lo).max() * 1.3 + 8
        bbox[i] = (lo + hi) / 2 - size / 2
        rate[i] = s / size
        crop[i] = (c["p2d"] - bbox[i]) * rate[i]
    ys = torch.arange(s, device=DEV, dtype=torch.float32).view(1, 1, s, 1)
    xs = torch.arange(s, device=DEV, dtype=torch.float32).view(1, 1, 1, s)
    c_t = torch.from_numpy(crop).to(DEV)
    g = torch.Generator(device=DEV); g.manual_seed(seed)
    hm = torch.empty((n, kp, s, s), device=DEV)
    for i0 in range(0, n, 25):
        cc = c_t[i0:i0 + 25]
        d2 = (xs - cc[:, :, 0, None, None]) ** 2 + (ys - cc[:, :, 1, None, None]) ** 2
        hm[i0:i0 + 25] = torch.exp(-d2 / 8.0) * 0.9 + 0.01 * torch.randn(d2.shape, device=DEV, generator=g)
    return hm, torch.from_numpy(bbox).to(DEV), torch.from_numpy(rate).to(DEV), torch.from_numpy(model).to(DEV)


def c1():
    hm, bbox, rate, model = heatmap_batch(8, 11, 128, 1)
    K = torch.from_numpy(ESA_K).to(DEV)
    i = [0]

    def step():
        k = i[0] % 8; i[0] += 1
        return pipeline.poses_from_heatmaps(hm[k:k + 1], bbox[k:k + 1], rate[k:k + 1], model, K, min_k=8)
    ms = timed(step, 50)
    # the same call captured once as a CUDA graph and replayed on static buffers (pipeline.GraphedHeatmapPose)
    g = pipeline.GraphedHeatmapPose(1, 11, 128, 128, model, K, DEV, min_k=8)
    eager = pipeline.poses_from_heatmaps(hm[:1], bbox[:1], rate[:1], model, K, min_k=8)["pose7"].clone()
    assert torch.equal(g(hm[:1], bbox[:1], rate[:1])["pose7"], eager)
    ms_graph = timed(lambda: g(None, None, None, copy_inputs=False), 50)
    j = [0]

    def step_graph_copy():
        k = j[0] % 8; j[0] += 1
        return g(hm[k:k + 1], bbox[k:k + 1], rate[k:k + 1])
    ms_graph_copy = timed(step_graph_copy, 50)
    return {"config": "C1 single frame 11x128x128 heatmaps -> decode + EPnP-RANSAC + LM (device resident, no sync inside)",
            "metric": "latency per frame", "value": ms * 1e3, "unit": "us", "higher_is_better": False,
            "cuda_graph_replay_us": ms_graph * 1e3, "cuda_graph_replay_with_input_copies_us": ms_graph_copy * 1e3,
            "note": "reference val.py path on the CPU: ~0.8 ms decode + 0.5 ms cv2 EPnP + LM per frame (SURVEY 8d)"}


def c3():
    n, kp, s = 125, 11, 384                      # 0.81 GB of heatmaps per call (3 calls = the 375-frame GPU share)
    hm, bbox, rate, model = heatmap_batch(n, kp, s, 2)
    K = torch.from_numpy(ESA_K).to(DEV)
    lib = _lib.load()
    lib.epb_profile_enable(1)
    steps = 10
    ms = timed(lambda: pipeline.poses_from_heatmaps(hm, bbox, rate, model, K, min_k=8), steps)
    dec_ms, dec_n = prof(5)
    pose_ms, pose_n = prof(4)
    lib.epb_profile_enable(0)
    per = dec_ms / dec_n
    bytes_ = hm.numel() * 4 + n * kp * 16
    return {"config": "C3 heatmap path, 125 frames x 11 x 384x384 per call (per-GPU share of the 3000-frame set = 3 calls)",
            "metric": "poses/sec (heatmaps -> refined pose)", "value": n / (ms * 1e-3), "unit": "poses/s",
            "ms_per_call": ms, "kernel_ms": {"decode": per, "pose": pose_ms / pose_n},
            "roofline": {"kernel": "decode_kernel", "bound": "hbm", "achieved": bytes_ / (per * 1e-3) / 1e9, "peak": PEAK,
                         "unit": "GB/s", "frac": bytes_ / (per * 1e-3) / 1e9 / PEAK, "algorithmic_bytes_per_launch": bytes_,
                         "l2_policy": "input (0.81 GB) larger than L2"}}


def c4():
    b, s, vn, hn, rounds = 8, 768, 11, 256, 8
    mask, vertex, _ = make_vertex_field(3, 2, s, s, vn, 0.25, noise_deg=2.0)
    mask = np.tile(mask, (b // 2, 1, 1)); vertex = np.tile(vertex, (b // 2, 1, 1, 1))
    m_t = torch.from_numpy(mask).to(DEV)
    v_t = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(DEV))
    lib = _lib.load()
    lib.epb_profile_enable(1)
    torch.manual_seed(3)
    ms = timed(lambda: rv.estimate_voting_distribution(m_t, v_t, round_hyp_num=hn, min_hyp_num=hn * rounds, topk=128,
                                                       sync_rng=False), 5)
    vote_ms, vote_n = prof(2)
    lib.epb_profile_enable(0)
    tn = float(np.minimum(mask.reshape(b, -1).sum(1), 30000).mean())
    tests = b * hn * rounds * vn * tn
    return {"config": "C4 8 x 768x768 fields, 11 keypoints, 8 rounds x 256 = 2048 hypotheses, voting mean/covariance "
                      "(foreground 0.25 -> max_num subsample to ~30000 px)",
            "metric": "images/sec (vector field -> voting distribution)", "value": b / (ms * 1e-3), "unit": "images/s",
            "ms_per_call": ms, "kernel_ms": {"vote_count": vote_ms / vote_n},
            "roofline": {"kernel": "vote_count_kernel", "bound": "fp32-issue", "pair_tests_per_s": tests / (vote_ms / vote_n * 1e-3),
                         "alu_frac": 6 * tests / (vote_ms / vote_n * 1e-3) / (148 * 128 * 1.965e9)}}


def c5():
    out = []
    model = tango_model(11, seed=9)
    K = torch.from_numpy(ESA_K).to(DEV)
    rng = np.random.default_rng(5)
    base = 1000
    p2d = np.zeros((base, 11, 2)); init = np.zeros((base, 6))
    for i in range(base):
        c = make_pose_case(9000 + i, 11, 0.5, 0, model=model)
        p2d[i] = c["p2d"]
        init[i, :3] = c["rvec"] + rng.normal(0, np.deg2rad(2.0) / np.sqrt(3), 3)      # ~2 deg off
        init[i, 3:] = c["t"] * (1 + rng.normal(0, 0.02, 3))                             # 2 % off
    for n in (1000, 10000, 100000, 1000000):
        reps = n // base
        p2 = torch.from_numpy(np.tile(p2d, (reps, 1, 1))).to(DEV)
        it = torch.from_numpy(np.tile(init, (reps, 1))).to(DEV)
        w = torch.ones((n, 11, 3), dtype=torch.float64, device=DEV); w[:, :, 1] = 0
        m = torch.from_numpy(model).to(DEV)
        ms = timed(lambda: gp.lm_refine_batch(p2, m, w, K, it), 5 if n >= 100000 else 20)
        _, iters, _ = gp.lm_refine_batch(p2, m, w, K, it, return_info=True)
        out.append({"poses": n, "ms": ms, "poses_per_s": n / (ms * 1e-3), "mean_lm_iterations": float(iters.float().mean())})
    return {"config": "C5 batched LM refinement sweep, 11-point model, start 2 deg / 2 % off", "metric": "poses/sec (LM only)",
            "unit": "poses/s", "value": out[-1]["poses_per_s"], "sweep": out,
            "note": "reference analogue (uncertainty_pnp.cpp over TinySolver, 1 core): ~3.8e4 poses/s (SURVEY 8d)"}


def c2s():
    """config[1]'s shape where voting is NOT FP32-bound (below ~700 voting pixels per image):
    (a) sparse masks, foreground 0.01 (tn ~ 655); (b) the reference's v5 defaults, max_num = 100 (every image
    subsampled to ~100 voting pixels, ransac_voting_gpu.py:763-764).  Field and mask resident in HBM."""
    b, s, vn, hn = 64, 256, 11, 512
    out = []
    K = torch.from_numpy(ESA_K).to(DEV)
    lib = _lib.load()
    from types import SimpleNamespace
    from bench import make_batch_numpy                     # structured crops: keypoints = projections of the model
    for tag, keep, v5 in (("sparse masks (1 % of the crop votes: every 25th pixel of a 0.25 blob), v3", 0.04, False),
                          ("v5 defaults (max_num 100), foreground 0.25", 1.0, True)):
        mask, vertex, model_np, geom, _ = make_batch_numpy(SimpleNamespace(size=s, vn=vn, fg=0.25), 5, 4)
        if keep < 1.0:
            mask = (mask * (np.random.default_rng(1).random(mask.shape) < keep)).astype(np.uint8)
        mask = np.tile(mask, (b // 4, 1, 1)); vertex = np.tile(vertex, (b // 4, 1, 1, 1)); geom = np.tile(geom, (b // 4, 1))
        m_t = torch.from_numpy(mask).to(DEV)
        v_t = rv.vertex_layer_reshape(torch.from_numpy(vertex).to(DEV))
        model = torch.from_numpy(model_np).to(DEV)
        bbox = torch.from_numpy(np.ascontiguousarray(geom[:, :2])).to(DEV)
        rate = torch.from_numpy(np.ascontiguousarray(geom[:, 2])).to(DEV)
        torch.manual_seed(7)

        def step():
            if v5:
                kpts, _ = rv.ransac_voting_layer_v5(m_t, v_t, hn, sync_rng=False)
                return pipeline.poses_from_keypoints(kpts, model, K, bbox_xy=bbox, rate=rate)["pose7"]
            return pipeline.poses_from_vertex(m_t, v_t, model, K, round_hyp_num=hn, bbox_xy=bbox, rate=rate,
                                              sync_rng=False)["pose7"]
        lib.epb_profile_enable(1)
        ms = timed(step, 50)
        names = ("compaction", "hypothesis", "vote_count", "winner_refine", "pose")
        km = {}
        for cls, name in enumerate(names):
            tot, n = prof(cls)
            km[name] = tot / max(n, 1)
        lib.epb_profile_enable(0)
        finite = float(torch.isfinite(step()).all(1).float().mean())
        tn = float(mask.reshape(b, -1).sum(1).mean())
        touched = b * (s * s + (8 * vn + 4) * min(tn, 100.0 if v5 else tn) * 2)     # mask read + compacted field read and written
        out.append({"case": tag, "tn_mean_before_subsample": tn, "ms_per_call": ms, "poses_per_s": b / (ms * 1e-3),
                    "kernel_ms": km, "finite_pose_fraction": finite, "bytes_touched_per_call": touched,
                    "field_bytes_per_call_SURVEY_8d": b * s * s * (8 * vn + 1)})
    return {"config": "C2 shape (64 x 256x256, 11 keypoints, 512 hypotheses) outside the FP32-bound regime", "metric": "poses/sec",
            "unit": "poses/s", "value": out[0]["poses_per_s"], "cases": out,
            "note": "with a few hundred voting pixels per image the call is a chain of short kernels (launch/latency bound): "
                    "the field bytes SURVEY 8d counts are never read, only the mask and the foreground's field"}


# ----------------------------------------------------------------------------- pieces bench.py times itself
def make_c3_step(n_frames, seed=2, s=384, kp=11):
    """config[2]: `n_frames` frames of the 3000-frame validation set (this rank's share), one call."""
    hm, bbox, rate, model = heatmap_batch(n_frames, kp, s, seed)
    K = torch.from_numpy(ESA_K).to(DEV)

    def step():
        return pipeline.poses_from_heatmaps(hm, bbox, rate, model, K, min_k=8)["pose7"]
    return step, {"frames": n_frames, "heatmap_bytes": hm.numel() * 4, "hm": hm}


def make_c5_step(n_poses, seed=5):
    """config[4]: LM refinement of `n_poses` poses (this rank's share), 11-point model, start 2 deg / 2 % off."""
    model = tango_model(11, seed=9)
    K = torch.from_numpy(ESA_K).to(DEV)
    rng = np.random.default_rng(seed)
    base = min(1000, n_poses)
    p2d = np.zeros((base, 11, 2)); init = np.zeros((base, 6))
    for i in range(base):
        c = make_pose_case(9000 + i, 11, 0.5, 0, model=model)
        p2d[i] = c["p2d"]
        init[i, :3] = c["rvec"] + rng.normal(0, np.deg2rad(2.0) / np.sqrt(3), 3)
        init[i, 3:] = c["t"] * (1 + rng.normal(0, 0.02, 3))
    reps = (n_poses + base - 1) // base
    p2 = torch.from_numpy(np.tile(p2d, (reps, 1, 1))[:n_poses]).to(DEV)
    it = torch.from_numpy(np.tile(init, (reps, 1))[:n_poses]).to(DEV)
    w = torch.ones((n_poses, 11, 3), dtype=torch.float64, device=DEV); w[:, :, 1] = 0
    m = torch.from_numpy(model).to(DEV)

    def step():
        return gp.lm_refine_batch(p2, m, w, K, it)
    return step, {"poses": n_poses, "p2d": p2d, "init": init, "model": model}


def pose_stress():
    """Pose solve under RANSAC pressure: 64 frames x 11 keypoints, clean / one outlier / two outliers /
    no consensus at all (random keypoints: all 100 RANSAC iterations run and fail)."""
    model = tango_model(11, seed=9)
    K = torch.from_numpy(ESA_K).to(DEV)
    m = torch.from_numpy(model).to(DEV)
    out = {}
    rng = np.random.default_rng(4)
    for tag, n_out in (("clean", 0), ("one_outlier", 1), ("two_outliers", 2), ("no_consensus", -1)):
        if n_out >= 0:
            kp = np.stack([make_pose_case(6000 + i, 11, 0.5, n_out, model=model)["p2d"] for i in range(64)])
        else:
            kp = rng.uniform(0, 1900, (64, 11, 2))
        k_t = torch.from_numpy(kp.astype(np.float32)).to(DEV)
        ms = timed(lambda: pipeline.poses_from_keypoints(k_t, m, K)["pose7"], 20)
        st = pipeline.poses_from_keypoints(k_t, m, K)["status"]
        # the batch ends with its slowest frame; the single-frame call times of the first 16 frames say what a
        # typical frame of the kind costs
        one = sorted(timed(lambda i=i: pipeline.poses_from_keypoints(k_t[i:i + 1], m, K)["pose7"], 10) * 1e3
                     for i in range(16))
        out[tag] = {"ms_per_64_frames": ms, "failed": int((st != 0).sum()),
                    "single_frame_us": {"median": one[8], "max": one[-1]}}
    return out


def cpu_heatmap_path(n_frames, kp=11, s=128, cores=None):
    """CPU leg of the heatmap path (configs 0 and 2): the reference's per-frame loop (val.py:151-228) through the
    oracle port (oracle/decode.py = inference.py's get_max_preds / get_final restated and pinned bit-equal to them;
    oracle/pose.py = pnp.py's cv2 call + the C LM), one process per core.  -> frames/s."""
    import multiprocessing as mp
    import time
    cores = cores or (os.cpu_count() or 1)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_heatmap_frame, [(i, kp, s) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_heatmap_frame, [(i, kp, s) for i in range(n_frames)], chunksize=max(1, n_frames // (4 * cores)))
        dt = time.perf_counter() - t0
    return n_frames / dt, dt, cores


def _cpu_heatmap_frame(args):
    i, kp, s = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import decode as odec, pose as opose
    model = tango_model(kp, seed=9)
    c = make_pose_case(777 + i, kp, 0.0, 0, model=model)
    lo, hi = c["p2d"].min(0), c["p2d"].max(0)
    size = (hi - lo).max() * 1.3 + 8
    bbox = (lo + hi) / 2 - size / 2
    rate = s / size
    crop = (c["p2d"] - bbox) * rate
    ys, xs = np.mgrid[0:s, 0:s].astype(np.float32)
    hm = np.exp(-((xs[None] - crop[:, 0, None, None]) ** 2 + (ys[None] - crop[:, 1, None, None]) ** 2) / 8.0).astype(np.float32)[None] * 0.9
    preds, maxvals, _ = odec.decode_frame(hm)
    return opose.frame_pose(preds, maxvals, bbox, rate, model, ESA_K, min_k=8)["t"]


def cpu_lm_sweep(n_poses, cores=None):
    """CPU leg of config[4]: the C restatement of uncertainty_pnp.cpp (oracle/lm_oracle.c; stand-in for cpnp, whose binary
    and source are absent), one process per core.  -> poses/s."""
    import multiprocessing as mp
    import time
    cores = cores or (os.cpu_count() or 1)
    per = max(1, n_poses // cores)
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_lm_chunk, [(i, 50) for i in range(cores)])
        t0 = time.perf_counter()
        pool.map(_cpu_lm_chunk, [(i, per) for i in range(cores)])
        dt = time.perf_counter() - t0
    return per * cores / dt, dt, cores


def _cpu_lm_chunk(args):
    seed, n = args
    from oracle import pose as opose
    model = tango_model(11, seed=9)
    rng = np.random.default_rng(seed)
    cases = [make_pose_case(9000 + (seed * 131 + i) % 1000, 11, 0.5, 0, model=model) for i in range(min(n, 64))]
    w = np.ones((11, 3)); w[:, 1] = 0
    inits = [np.concatenate([c["rvec"] + rng.normal(0, np.deg2rad(2.0) / np.sqrt(3), 3), c["t"] * (1 + rng.normal(0, 0.02, 3))]) for c in cases]
    for i in range(n):
        c = cases[i % len(cases)]
        opose.lm_refine(c["p2d"], model, w, ESA_K, inits[i % len(cases)])
    return n


if __name__ == "__main__":
    which = sys.argv[1:] or ["c1", "c3", "c4", "c5", "c2s"]
    for name in which:
        r = {"c1": c1, "c3": c3, "c4": c4, "c5": c5, "c2s": c2s}[name]()
        r["gpu"] = torch.cuda.get_device_name(0)
        print(json.dumps(r))
        sys.stdout.flush()
