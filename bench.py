#!/usr/bin/env python
"""bench.py -- poses/s of the post-network pose hot path on BASELINE.json config[1]:
batch of 64 crops 256x256, 11-keypoint vector fields, 512 RANSAC hypotheses, voting (v3) +
EPnP-RANSAC + LM on each B200; images shard by batch across GPUs (weak scaling), one NCCL
all_gather of the [64,7] poses per step.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
  python bench.py --impl reference ...                            the reference algorithm on host cores
Prints ONE JSON line (rank 0).  A "step" = one pass of the hot path over one batch per GPU.
`value` is timed with inputs resident in HBM; `e2e` goes through the public Python API with
pinned HOST buffers (H2D of mask + field and D2H of the poses inside the timed region).
Only the `cpu_baseline` / `--impl reference` legs touch oracle/.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TRAFFIC_PROFILE = "profiles/r2_vote_count_full.md"
METRIC = "poses/sec (vector fields -> refined pose)"
UNIT = "poses/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--vn", type=int, default=11)
    ap.add_argument("--hn", type=int, default=512)
    ap.add_argument("--fg", type=float, default=0.25, help="foreground fraction of each crop")
    ap.add_argument("--cpu-sample", type=int, default=0, help="images in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--e2e-chunks", type=int, default=int(os.environ.get("EPB_E2E_CHUNKS", "0")),
                    help="batch pieces of the host-input pipeline (0 = library default)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the host-side baseline (profiling runs)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer e2e leg (ncu launch lists of the device step)")
    ap.add_argument("--no-configs", action="store_true", help="skip the informational legs (other BASELINE configs, reference_gpu)")
    return ap.parse_args()


def config_dict(a, world):
    """The `config` object of BOTH arms (the driver compares them key by key)."""
    return {"workload": workload_name(a), "batch_per_gpu": a.batch, "image": "%dx%d" % (a.size, a.size), "keypoints": a.vn,
            "hypotheses": a.hn, "foreground_fraction": a.fg, "tn_mean": float(int(a.fg * a.size * a.size)),
            "rng": "philox (torch layout); the torch generator is advanced by the data-independent upper bound "
                   "(sync_rng=False: no host sync), not by what the reference would have consumed",
            "l2_policy": "inputs larger than L2 (%.0f MB of vector field per step)" % (a.batch * 2 * a.vn * a.size * a.size * 4 / 1e6),
            "parallelism": "images sharded by batch, 1 NCCL all_gather of poses" if world > 1 else "single GPU"}


def workload_name(a):
    return ("config[1]: batch %d crops %dx%d, %d-keypoint vector fields, %d RANSAC hypotheses, "
            "voting(v3)+EPnP-RANSAC+LM, foreground %.2f (tn~%d px/img)"
            % (a.batch, a.size, a.size, a.vn, a.hn, a.fg, int(a.fg * a.size * a.size)))


# ----------------------------------------------------------------------------- synthetic data
def make_batch_numpy(a, seed, n_images):
    """Structured SPEED-like crops (tests/synth.py: make_pose_field): the keypoints are the projections of the
    Tango model under a random pose, scaled into the crop; field = unit vectors towards them + 2 deg noise."""
    from tests.synth import make_pose_field
    mask, vertex, model, geom, kcrop, _, _ = make_pose_field(seed, n_images, a.size, a.vn, a.fg, noise_deg=2.0)
    return mask, vertex, model, geom, kcrop


def tile_to(arr, n):
    reps = (n + arr.shape[0] - 1) // arr.shape[0]
    return np.concatenate([arr] * reps, 0)[:n]


# ----------------------------------------------------------------------------- reference arm
def _cpu_one(args):
    """One image through the reference algorithm on one core (oracle = CPU restatement)."""
    mask, vertex_hwvn2, model, geom, hn, seed = args
    import cv2
    cv2.setNumThreads(1)
    from oracle import pose as opose, voting as ov
    from tests.synth import ESA_K
    k = ov.ransac_voting_layer_v3(mask[None], vertex_hwvn2[None], hn, idxs_fn=ov.default_idxs_fn(seed))[0]
    p2d = k.astype(np.float64) * (1.0 / geom[2]) + geom[:2]           # val.py:180
    rt = opose.pnp(model, p2d, ESA_K, cv2.SOLVEPNP_EPNP)
    r_exp, _ = cv2.Rodrigues(rt[:, :3])
    cam = np.concatenate([r_exp.reshape(3), rt[:, 3]])
    cam = opose.cpnp(model, p2d, ESA_K, cam)
    return cam


def cpu_run(a, n_images, cores, seed=1):
    """Times n_images of the workload through the oracle on `cores` processes -> (poses/s, seconds)."""
    import multiprocessing as mp
    from tests.synth import vertex_hwvn2
    base = min(n_images, 4)
    mask, vertex, model, geom, _ = make_batch_numpy(a, seed, base)
    vx = vertex_hwvn2(vertex)
    jobs = [(mask[i % base], vx[i % base], model, geom[i % base], a.hn, i) for i in range(n_images)]
    from oracle import _lib as olib
    olib.oracle_lib()
    if cores <= 1:
        t0 = time.perf_counter()
        for j in jobs:
            _cpu_one(j)
        dt = time.perf_counter() - t0
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            pool.map(_cpu_one, jobs[:cores])          # warm the workers
            t0 = time.perf_counter()
            pool.map(_cpu_one, jobs, chunksize=1)
            dt = time.perf_counter() - t0
    return n_images / dt, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = a.cpu_sample or max(4 * cores, 16)
    for _ in range(min(a.warmup, 1)):
        cpu_run(a, cores, cores)
    t_all, n_all = 0.0, 0
    for _ in range(a.steps):
        _, dt = cpu_run(a, per_step, cores)
        t_all += dt; n_all += per_step
    value = n_all / t_all
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus,
            "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * t_all / max(a.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": config_dict(a, max(a.gpus, 1)), "sample_images_per_step": per_step,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d images/step of the same workload through oracle/ (C voting restatement, "
                                       "cv2 EPnP-RANSAC, LM restatement), one process per core" % per_step},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ----------------------------------------------------------------------------- our arm
def bind_to_gpu_numa_node(local):
    """Pin this rank to the CPUs next to its GPU BEFORE the pinned host buffers are allocated (first
    touch places them on that NUMA node): the zero-copy gather of the e2e path then reads host memory
    that is local to the GPU's PCIe root instead of crossing the socket interconnect.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "%d cpus of %s" % (len(cpus), path.split("/")[-2])
    except Exception as e:                      # no sysfs / no permission: keep the default placement
        return "unbound (%s)" % type(e).__name__
    return "unbound"


class ClockSampler:
    """One `nvidia-smi -lms` process for the whole run, started BEFORE the warm-up (its start-up takes
    driver locks for a few hundred ms and must not fall into a timed region); samples carry wall-clock
    timestamps and are attributed to the timed regions afterwards."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None
        self.rows = None

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.02)

    def stop(self):
        if self.p is None:
            self.rows = []
            return
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        import datetime
        rows = []
        for ln in self.f.read().strip().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(c[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                rows.append((ts, float(c[1]), float(c[2]), [n for n, v in zip(
                    ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8])
                    if v.lower().startswith("active")]))
            except ValueError:
                continue
        self.rows = rows
        try:
            os.unlink(self.f.name)
        except OSError:
            pass

    def region(self, t0, t1):
        """Clocks line for the samples taken in [t0, t1] (wall clock); falls back to the nearest sample
        when the region is shorter than the sampling period."""
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.rows:
            return out
        sel = [r for r in self.rows if t0 - 0.005 <= r[0] <= t1 + 0.005]
        if not sel:
            mid = 0.5 * (t0 + t1)
            sel = [min(self.rows, key=lambda r: abs(r[0] - mid))]
            out["note"] = "region shorter than the 20 ms sampling period: nearest sample"
        out.update(sm_mhz=float(np.median([r[1] for r in sel])), sm_max_mhz=float(max(r[2] for r in sel)),
                   reasons=sorted({x for r in sel for x in r[3]}), samples=len(sel))
        return out


def reference_gpu_voting(a, mask_d, vertex_d, n_images, reps=3):
    """Informational: the REFERENCE's own CUDA kernels on the same data and box -- ransac_voting_kernel.cu:11-126
    compiled unmodified for sm_100a (oracle/_ref/libref_voting.so) and driven like ransac_voting_gpu.py:523-598, one
    image at a time: torch.nonzero / masked_select compaction, random_ indices, generate_hypothesis, the zeroed
    [hn,vn,tn] byte tensor, voting_for_hypothesis, torch.sum, max, re-vote of the winners and the 2x2 normal
    equations.  (Its pybind module cannot be built against torch 2.11; the launchers are called through ctypes on
    raw device pointers, everything on the legacy default stream like the original.)  -> ms per image, or None."""
    import ctypes
    import torch
    try:
        from oracle import _lib as olib
        lib = olib.ref_voting_lib(required=False)
    except Exception:
        lib = None
    if lib is None:
        return None
    vp = ctypes.c_void_p
    hn, vn, thresh = a.hn, a.vn, ctypes.c_float(0.999)

    def one(bi):
        cur_mask = mask_d[bi].byte()
        coords = torch.nonzero(cur_mask).float()[:, [1, 0]].contiguous()
        tn = coords.shape[0]
        direct = vertex_d[bi].masked_select(cur_mask[:, :, None, None].bool()).view(tn, vn, 2).contiguous()
        idxs = torch.zeros([hn, vn, 2], dtype=torch.int32, device=mask_d.device).random_(0, tn)
        hyp = torch.zeros([hn, vn, 2], dtype=torch.float32, device=mask_d.device)
        lib.ref_generate_hypothesis(vp(direct.data_ptr()), vp(coords.data_ptr()), vp(idxs.data_ptr()), vp(hyp.data_ptr()), tn, vn, hn)
        inlier = torch.zeros([hn, vn, tn], dtype=torch.uint8, device=mask_d.device)
        lib.ref_voting_for_hypothesis(vp(direct.data_ptr()), vp(coords.data_ptr()), vp(hyp.data_ptr()), vp(inlier.data_ptr()),
                                      tn, vn, hn, thresh)
        counts = torch.sum(inlier, 2)
        _, win = torch.max(counts, 0)
        win_pts = hyp[win, torch.arange(vn, device=hyp.device)].unsqueeze(0).contiguous()
        inl2 = torch.zeros([1, vn, tn], dtype=torch.uint8, device=mask_d.device)
        lib.ref_voting_for_hypothesis(vp(direct.data_ptr()), vp(coords.data_ptr()), vp(win_pts.data_ptr()), vp(inl2.data_ptr()),
                                      tn, vn, 1, thresh)
        w = inl2[0].float().t().unsqueeze(2)                                   # [tn,vn,1]
        normal = torch.stack([direct[:, :, 1], -direct[:, :, 0]], 2) * w       # [tn,vn,2]
        b_ = torch.sum(normal * coords.unsqueeze(1), 2, keepdim=True)          # [tn,vn,1]
        ata = torch.einsum("tvi,tvj->vij", normal, normal)
        atb = torch.einsum("tvi,tvk->vik", normal, b_)
        return torch.linalg.solve(ata, atb)[:, :, 0]

    for bi in range(min(2, n_images)):
        one(bi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for bi in range(n_images):
            one(bi)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * n_images)


def secondary_configs(a, world, rank, dev, timed, barrier):
    """BASELINE.json configs 0, 2, 3, 4 (+ the pose solve under RANSAC pressure), measured with the same protocol as the
    contract line: inputs resident in HBM, CUDA events, max over ranks.  configs 2 and 4 shard by image / pose across the
    ranks (weak work split of a FIXED total: 3000 frames, 1e6 poses); 0 and 3 are single-GPU cases (rank 0, N = 1 only).
    CPU legs (rank 0, N = 1, bounded samples) run the reference's per-frame loop through the oracle port."""
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_configs as bc
    from esa_pose_estimation_b200 import _lib
    lib = _lib.load()
    out = {}
    peak = bc.PEAK

    def try_(name, fn):
        try:
            out[name] = fn()
        except Exception as e:               # informational legs never take the contract line down
            out[name] = {"failed": repr(e)}

    # config[2]: 3000 frames of 11 x 384x384 heatmaps, sharded by frame, one call per rank, one all_gather of the poses
    def c2():
        from esa_pose_estimation_b200 import pipeline
        n_total = 3000
        s_, e_ = pipeline.shard_range(n_total, rank, world)
        step, meta = bc.make_c3_step(e_ - s_, seed=2 + rank)

        def full():
            p = step()
            return pipeline.gather_poses(p, n_total) if world > 1 else p
        for _ in range(5):
            full()
        torch.cuda.synchronize()
        # Every pass is timed on its own (a barrier and a synchronisation per pass; max over ranks per pass) and the
        # MEDIAN pass is the figure.  This leg is ~40 short launches plus a dozen event records / waits per 3.3 ms
        # pass; timed as N back-to-back passes without a synchronisation, the enqueueing thread runs ahead until a
        # launch queue is full, and the driver then parks it for 50-100 ms (the effect poses_from_vertex throttles
        # against with its DEPTH ring): about one pass in fifty, enough to turn a 5-pass mean into 140-370 k poses/s
        # in three of twenty runs.  A caller that consumes each call's result -- the per-pass synchronisation here --
        # never gets there.  The mean and every pass are reported next to the median.
        import gc
        steps = 10
        lib.epb_profile_enable(1)             # (kernel times of the same passes: two CUDA events per library call)
        gc.disable()
        each = []
        for i in range(steps):
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            full()
            e1.record()
            torch.cuda.synchronize()
            t_ = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t_, op=dist.ReduceOp.MAX)
            each.append(float(t_.item()))
        gc.enable()
        dec = _lib.c_double(0); n_ = _lib.c_int(0)
        lib.epb_profile_read(5, dec, n_)
        pose_t = _lib.c_double(0); pose_n = _lib.c_int(0)
        lib.epb_profile_read(4, pose_t, pose_n)
        lib.epb_profile_enable(0)
        ms = sorted(each)[len(each) // 2]
        nan_poses = int(torch.isnan(step()).any(1).sum().item())
        per = dec.value / steps               # decode time of one pass (all its launches: the set goes through in pieces)
        r = {"workload": "3000 frames x 11 x 384x384 heatmaps -> decode + EPnP-RANSAC + LM, %d frames per GPU, NCCL pose gather" % (e_ - s_),
             "poses_per_s": n_total / (ms * 1e-3), "ms_per_pass": ms, "ms_per_pass_is": "median of %d passes timed one by one" % steps,
             "ms_per_pass_mean": sum(each) / len(each), "ms_each_pass": each, "n_gpus": world,
             "roofline": {"kernel": "decode_kernel", "bound": "hbm", "achieved": meta["heatmap_bytes"] / (per * 1e-3) / 1e9 if per > 0 else None,
                          "peak": peak, "unit": "GB/s", "frac": meta["heatmap_bytes"] / (per * 1e-3) / 1e9 / peak if per > 0 else None,
                          "decode_ms_per_pass": per, "launches_per_pass": n_.value / steps,
                          "pose_kernel_ms_per_pass": pose_t.value / steps,
                          "failed_poses": nan_poses,
                          "l2_policy": "inputs larger than L2 (%.1f GB per GPU)" % (meta["heatmap_bytes"] / 1e9)}}
        del meta, step
        torch.cuda.empty_cache()
        return r
    try_("config2_val_set_3000_frames", c2)

    # config[4]: LM refinement sweep; the 1e6-pose point is sharded across the ranks
    def c4():
        sweep = []
        for n_total in ((1000, 10000, 100000, 1000000) if world == 1 else (1000000,)):
            per_rank = n_total // world
            step, meta = bc.make_c5_step(per_rank, seed=5 + rank)
            for _ in range(3):
                step()
            steps = 5 if n_total >= 100000 else 20
            ms = timed(step, steps) / steps
            sweep.append({"poses": per_rank * world, "ms": ms, "poses_per_s": per_rank * world / (ms * 1e-3)})
            del step, meta
        torch.cuda.empty_cache()
        return {"workload": "batched LM refinement, 11-point model, start 2 deg / 2 % off, poses sharded across the GPUs",
                "n_gpus": world, "sweep": sweep, "poses_per_s": sweep[-1]["poses_per_s"]}
    try_("config4_lm_sweep", c4)

    if world == 1:
        try_("config0_single_frame", lambda: {k: v for k, v in bc.c1().items() if k != "config"})
        try_("config3_768_fields_2048_hypotheses_covariance", lambda: {k: v for k, v in bc.c4().items() if k != "config"})
        try_("pose_solve_under_ransac_pressure_64_frames", bc.pose_stress)
        if not a.no_cpu_baseline:
            def cpu_legs():
                cores = os.cpu_count() or 1
                v0, dt0, _ = bc.cpu_heatmap_path(max(16 * cores, 64), kp=11, s=128, cores=cores)
                v2, dt2, _ = bc.cpu_heatmap_path(max(4 * cores, 32), kp=11, s=384, cores=cores)
                v4, dt4, _ = bc.cpu_lm_sweep(20000, cores=cores)
                return {"kind": "port", "cores": cores,
                        "config0_frames_per_s": v0, "config2_frames_per_s": v2, "config4_poses_per_s": v4,
                        "sample": "%d frames at 128x128 in %.1f s, %d frames at 384x384 in %.1f s, 20000 LM solves in %.1f s; "
                                  "reference per-frame loop (val.py:151-228: numpy argmax + log-Taylor refine, cv2 EPnP-RANSAC, LM) "
                                  "through oracle/, one process per core" % (max(16 * cores, 64), dt0, max(4 * cores, 32), dt2, dt4)}
            try_("cpu_baseline", cpu_legs)
    if world > 1:
        barrier()
    return out if rank == 0 else None


def parse_traffic(path):
    """dram__bytes_read.sum + dram__bytes_write.sum of the first launch in an ncu summary under profiles/ (tools/ncu_summary.py
    format: `| `metric` | value | unit |`) -> bytes, or None."""
    try:
        tot, seen = 0.0, 0
        for ln in open(path):
            for key in ("dram__bytes_read.sum`", "dram__bytes_write.sum`"):
                if key in ln and seen < 2:
                    cells = [c.strip() for c in ln.split("|")]
                    val, unit = float(cells[2]), cells[3].lower()
                    tot += val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
                    seen += 1
        return int(tot) if seen == 2 else None
    except Exception:
        return None


def run_ours(a):
    import torch
    import torch.distributed as dist
    from esa_pose_estimation_b200 import _lib, pipeline, ransac_voting_gpu as rv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device and no CPU fallback (use --impl reference for the host arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    numa = bind_to_gpu_numa_node(local) if not os.environ.get("EPB_NO_NUMA_BIND") else "disabled"

    # --- data: 8 distinct structured crops per rank, tiled to the batch (content repeats, the
    # working set per step -- 369 MB of field at the default size -- is larger than the 126 MB L2)
    base = min(a.batch, 8)
    from tests.synth import ESA_K
    mask_np, vertex_np, model_np, geom_np, kcrop_np = make_batch_numpy(a, 11 + rank, base)
    mask_np, vertex_np, geom_np = tile_to(mask_np, a.batch), tile_to(vertex_np, a.batch), tile_to(geom_np, a.batch)
    kcrop_np = tile_to(kcrop_np, a.batch)
    mask_h = torch.from_numpy(mask_np).pin_memory()
    vertex_h = torch.from_numpy(vertex_np).pin_memory()
    mask_d, vertex_d = mask_h.to(dev), vertex_h.to(dev)
    model_d = torch.from_numpy(model_np).to(dev)
    K_d = torch.from_numpy(ESA_K).to(dev)
    bbox_d = torch.from_numpy(np.ascontiguousarray(geom_np[:, :2])).to(dev)
    rate_d = torch.from_numpy(np.ascontiguousarray(geom_np[:, 2])).to(dev)
    pose_h = torch.empty((a.batch, 7), dtype=torch.float32).pin_memory()
    torch.manual_seed(1234 + rank)

    def step_device():
        out = pipeline.poses_from_vertex(mask_d, rv.vertex_layer_reshape(vertex_d), model_d, K_d,
                                         round_hyp_num=a.hn, bbox_xy=bbox_d, rate=rate_d, sync_rng=False)
        step_device.kpts = out["kpts"]
        return pipeline.gather_poses(out["pose7"], a.batch * world) if world > 1 else out["pose7"]

    # informational second leg: the same K device-resident steps issued round-robin on three caller streams
    # (three batches in flight through the library's gather / vote / pose streams): the latency-bound kernels of
    # one call run under the voting of the others
    n_streams = 3
    dev_streams = [torch.cuda.Stream(device=dev) for _ in range(n_streams)]

    def step_streams():
        st = dev_streams[step_streams.n % n_streams]
        step_streams.n += 1
        with torch.cuda.stream(st):
            out = pipeline.poses_from_vertex(mask_d, rv.vertex_layer_reshape(vertex_d), model_d, K_d, round_hyp_num=a.hn,
                                             bbox_xy=bbox_d, rate=rate_d, sync_rng=False, pipelined=True)
            return pipeline.gather_poses(out["pose7"], a.batch * world) if world > 1 else out["pose7"]
    step_streams.n = 0

    def step_e2e():
        # public API on HOST buffers: the mask is copied H2D, the pinned field is read in place over
        # PCIe by the gather kernel (foreground pixels only), the poses are copied D2H
        # (host_inputs_ready: the pinned buffers were filled by the host before the loop and are never rewritten)
        out = pipeline.poses_from_vertex(mask_h, rv.vertex_layer_reshape(vertex_h), model_d, K_d, round_hyp_num=a.hn,
                                         bbox_xy=bbox_d, rate=rate_d, sync_rng=False, chunks=a.e2e_chunks or None,
                                         host_inputs_ready=True)
        p = pipeline.gather_poses(out["pose7"], a.batch * world) if world > 1 else out["pose7"]
        pose_h.copy_(p[rank * a.batch:(rank + 1) * a.batch] if world > 1 else p, non_blocking=True)
        return p

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        # like timeit: no cyclic garbage collection inside a timed region (a generation-2 pass of this process stalls
        # the enqueueing thread for 10-90 ms -- tools measurement: 35 ms in 300 passes of the heatmap path -- and a
        # short region has no queued work to hide that behind)
        import gc
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gc.disable()                          # (no gc.collect() here: tens of ms of idle GPU right before a short region)
        barrier()
        ev0.record()
        for _ in range(steps):
            fn()
        if fn is step_streams:                # join the caller streams before the closing event
            for st in dev_streams:
                torch.cuda.current_stream(dev).wait_stream(st)
        ev1.record()
        barrier()
        ms = ev0.elapsed_time(ev1)
        gc.enable()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("EPB_NO_SAMPLER") else None
    if sampler:
        sampler.wait_first_sample()
    for _ in range(max(a.warmup, 3)):
        last = step_device()
    barrier()
    # sanity: poses finite, keypoints recovered (not part of the timed region)
    assert torch.isfinite(last).all(), "non-finite poses"
    kerr = float((step_device.kpts.double().cpu() - torch.from_numpy(kcrop_np)).abs().max())
    # 2 deg of direction noise at up to ~size px from the keypoint: the refined point lands within ~1 px at 256
    assert kerr < 1.5 * max(a.size, 256) / 256.0, "voting did not recover the planted keypoints (max err %.3f px)" % kerr

    if world > 1:
        # the gathered [N*B,7] block must hold every rank's poses in image order: this rank's slice == its local poses
        out_chk = pipeline.poses_from_vertex(mask_d, rv.vertex_layer_reshape(vertex_d), model_d, K_d, round_hyp_num=a.hn,
                                             bbox_xy=bbox_d, rate=rate_d, sync_rng=False)
        g_chk = pipeline.gather_poses(out_chk["pose7"], a.batch * world)
        assert torch.equal(g_chk[rank * a.batch:(rank + 1) * a.batch], out_chk["pose7"]), "NCCL gather: local slice differs"
        cs = g_chk.double().sum(1)
        ref_cs = [torch.empty_like(cs) for _ in range(world)]
        dist.all_gather(ref_cs, cs)
        assert all(torch.equal(ref_cs[0], r_) for r_ in ref_cs), "NCCL gather: ranks hold different gathered poses"
        barrier()

    launches0 = lib.epb_launch_count()
    t_dev0 = time.time()
    ms_dev = timed(step_device, a.steps)
    t_dev1 = time.time()
    launches = lib.epb_launch_count() - launches0
    # per-kernel-class times from a second, profiled set of steps: the library's profiler brackets every class with a
    # pair of CUDA events it creates on the spot, which does not belong inside the timed region of `value`
    lib.epb_profile_enable(1)
    prof_steps = min(a.steps, 20)
    timed(step_device, prof_steps)
    prof = {}
    for cls, name in ((0, "compaction"), (1, "hypothesis"), (2, "vote_count"), (3, "winner_refine"), (4, "pose")):
        tot, n = _lib.c_double(0), _lib.c_int(0)
        lib.epb_profile_read(cls, tot, n)
        prof[name] = (tot.value, n.value)
    lib.epb_profile_enable(0)

    ms_streams = float("nan")
    if not a.no_e2e:
        for st in dev_streams:
            st.wait_stream(torch.cuda.current_stream(dev))
        for _ in range(2 * n_streams):
            step_streams()
        ms_streams = timed(step_streams, a.steps)
    t_e0 = time.time()
    ms_e2e = float("nan")
    if not a.no_e2e:
        for _ in range(max(a.warmup, 3)):
            step_e2e()
        t_e0 = time.time()
        ms_e2e = timed(step_e2e, a.steps)
    t_e1 = time.time()
    # what the bus gives a plain cudaMemcpyAsync of the same number of bytes from pinned memory on this box: the
    # denominator of the e2e leg's own bound (`e2e.pcie`), measured after its timed region
    dma_gbs = None
    if not a.no_e2e:
        try:
            n_pay = int(mask_np.size + int(mask_np.astype(bool).sum()) * a.vn * 8)
            src = torch.empty(n_pay, dtype=torch.uint8).pin_memory()
            dst = torch.empty(n_pay, dtype=torch.uint8, device=dev)
            src.fill_(1)                                     # touch the pages before the DMA engines do
            for _ in range(3):
                dst.copy_(src, non_blocking=True)
            dma_ms = min(timed(lambda: dst.copy_(src, non_blocking=True), 10) / 10 for _ in range(3))   # a peak: best of 3
            dma_gbs = n_pay / (dma_ms * 1e-3) / 1e9
            del src, dst
        except Exception:
            dma_gbs = None
    clocks = clocks_e2e = None
    if sampler:
        sampler.stop()
        clocks, clocks_e2e = sampler.region(t_dev0, t_dev1), sampler.region(t_e0, t_e1)

    # --- the other BASELINE.json configurations (informational; `value` above is the contract line)
    configs = None if a.no_configs else secondary_configs(a, world, rank, dev, timed, barrier)
    ref_gpu_ms = None
    if rank == 0 and world == 1 and not a.no_configs:
        try:
            ref_gpu_ms = reference_gpu_voting(a, mask_d, rv.vertex_layer_reshape(vertex_d), min(a.batch, 8))
        except Exception as e:           # informational leg: never takes the contract line down
            ref_gpu_ms = "failed: %r" % (e,)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    poses_per_step = a.batch * world
    value = poses_per_step * a.steps / (ms_dev * 1e-3)
    e2e = poses_per_step * a.steps / (ms_e2e * 1e-3)
    # bytes that cross PCIe host->device per step: the mask plus the foreground part of the field
    # (the host tensors themselves are mask + field = `host_input_bytes_per_step`)
    fg_px = int(mask_np.astype(bool).sum())
    h2d = int(mask_h.numel() + fg_px * a.vn * 8) * world
    host_bytes = int(mask_h.numel() + vertex_h.numel() * 4) * world
    d2h = int(pose_h.numel() * 4) * world

    # --- roofline of the dominant kernel (vote_count), measured live with CUDA events
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    vote_ms, vote_n = prof["vote_count"]
    per_launch_s = vote_ms * 1e-3 / max(vote_n, 1)
    alg_bytes = a.batch * a.size * a.size * (8 * a.vn + 1)              # H*W*(8vn+1) per image
    achieved = alg_bytes / per_launch_s / 1e9 if per_launch_s > 0 else 0.0
    tn = float(mask_np.reshape(a.batch, -1).sum(1).mean())
    evals = a.batch * a.hn * a.vn * tn
    sm_count, clk = _lib.c_int(0), _lib.c_int(0)
    lib.epb_device_info(sm_count, clk, None)
    sm_mhz = (clocks or {}).get("sm_mhz") or clk.value / 1e3
    lane_ops_peak = sm_count.value * 128 * sm_mhz * 1e6                  # FP32 lane-ops/s at the observed clock
    step_ms = ms_dev / a.steps
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(a, world), "tn_mean_measured": tn,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "pipeline.poses_from_vertex(pinned host mask, pinned host field, ..., host_inputs_ready=True) + pose D2H",
                "ms_per_step": ms_e2e / a.steps, "host_input_bytes_per_step": host_bytes, "clocks": clocks_e2e, "host_numa_binding": numa,
                "transfer": "mask cudaMemcpyAsync from pinned memory; field read zero-copy from pinned memory by the "
                            "compaction kernel (foreground pixel groups only); poses D2H into pinned memory",
                # the e2e leg is bound by the bus, not by the kernels: bytes that must cross per step over the time of
                # a step, against a plain pinned-memory DMA copy of the same size measured in this run
                "pcie": None if not dma_gbs or not (ms_e2e == ms_e2e) else {
                    "bound": "pcie", "achieved": h2d / world / (ms_e2e / a.steps * 1e-3) / 1e9, "peak": dma_gbs,
                    "unit": "GB/s per GPU", "frac": h2d / world / (ms_e2e / a.steps * 1e-3) / 1e9 / dma_gbs,
                    "peak_source": "cudaMemcpyAsync of one rank's h2d bytes from pinned memory, all ranks at once, "
                                   "this run (rank 0's figure)"}},
        "gpu_launches": int(launches),
        "host_gc": "Python's cyclic GC is disabled inside every timed region (timeit's convention)",

        "roofline": {"kernel": "vote_count_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     # dram__bytes_read.sum + dram__bytes_write.sum of one launch at the default workload, from the
                     # ncu --set full capture summarised in profiles/r1_vote_count_full.md (100.9 MB + 5.4 MB): well
                     # BELOW the algorithmic bytes, because only the foreground of the field is ever touched
                     "traffic": parse_traffic(os.path.join(ROOT, TRAFFIC_PROFILE)) if (a.batch, a.size, a.vn, a.hn, a.fg) == (64, 256, 11, 512, 0.25) else None,
                     "traffic_source": TRAFFIC_PROFILE + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch)",
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes, "ms_per_launch": per_launch_s * 1e3,
                     "share_of_step": (vote_ms / prof_steps) / step_ms if step_ms > 0 else None,
                     "note": "FP32-ALU-bound at this foreground (SURVEY 8d): see alu",
                     # the kernel's real bound: FP32 issue.  Algorithmic cost of one (hypothesis, pixel) test in the
                     # affine formulation: 2 + 2 FMAs for the two forms, 1 subtraction, 1 band FMA
                     "alu": {"pair_tests_per_launch": evals, "pair_tests_per_s": evals / per_launch_s if per_launch_s else 0,
                             "algorithmic_fp32_lane_ops_per_test": 6,
                             "achieved_lane_ops_per_s": 6 * evals / per_launch_s if per_launch_s else 0,
                             "fp32_lane_ops_peak_per_s": lane_ops_peak,
                             "frac": (6 * evals / per_launch_s) / lane_ops_peak if per_launch_s and lane_ops_peak else None,
                             "lane_op_slots_per_test": lane_ops_peak * per_launch_s / evals if evals else None}},
        "kernel_ms_per_step": {k: v[0] / prof_steps for k, v in prof.items()},   # (from the profiled set of steps)
        "clocks": clocks,
        # not the contract line: `value` above is one call after the other on one stream
        "value_batches_in_flight": {"caller_streams": n_streams, "value": poses_per_step * a.steps / (ms_streams * 1e-3),
                                    "ms_per_step": ms_streams / a.steps, "unit": UNIT},
    }
    if configs is not None:
        line["configs"] = configs
    voting_ms = sum(prof[k][0] for k in ("compaction", "hypothesis", "vote_count", "winner_refine")) / prof_steps
    line["reference_gpu"] = {
        "what": "the reference's own voting kernels (ransac_voting_kernel.cu compiled unmodified for sm_100a) driven like "
                "ransac_voting_gpu.py:523-598 (per image: torch compaction, random_, generate_hypothesis, [hn,vn,tn] byte tensor, "
                "voting_for_hypothesis, torch.sum, max, re-vote, normal equations), same data, same GPU; keypoints only (no PnP / LM)",
        "ms_per_image": ref_gpu_ms, "ours_voting_ms_per_image": voting_ms / a.batch,
        "speedup_voting": (ref_gpu_ms / (voting_ms / a.batch)) if isinstance(ref_gpu_ms, float) and voting_ms > 0 else None}
    if a.no_e2e:                              # profiling runs: keep the line valid JSON (no NaN)
        line["e2e"].update(value=None, ms_per_step=None)
        line["value_batches_in_flight"].update(value=None, ms_per_step=None)
    # --- CPU baseline (oracle port) on a bounded sample, rank 0, N=1 only
    if world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        # bounded sample sized for ~10-20 s of host work (about 0.7 s per image and core)
        n_img = a.cpu_sample or max(20 * cores, 32)
        try:
            v, dt = cpu_run(a, n_img, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "%d images of the same workload through oracle/ in %.1f s" % (n_img, dt)}
        except Exception as e:  # the baseline must never take the GPU number down with it
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": cores, "kind": "port", "sample": "failed: %r" % (e,)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
