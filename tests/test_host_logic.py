"""Host-side logic that needs no GPU: the result sink against the reference's own CSV, image
sharding, and the pose gather over a 2-rank gloo group."""
import io
import os
import contextlib

import numpy as np
import torch
import torch.multiprocessing as mp

from esa_pose_estimation_b200 import pipeline
from esa_pose_estimation_b200.submission import SubmissionWriter


def test_submission_writer_matches_reference_csv(golden_dir, tmp_path):
    g = np.load(os.path.join(golden_dir, "submission_ref.npz"))
    w = SubmissionWriter()
    for name, q, r, real in zip(g["names"], g["q"], g["r"], g["real"]):
        (w.append_real_test if real else w.append_test)(str(name), q, r)
    with contextlib.redirect_stdout(io.StringIO()):
        path = w.export(out_dir=str(tmp_path), suffix="t")
    assert open(path).read() == str(g["csv"])


def test_submission_append_batch(tmp_path):
    w = SubmissionWriter()
    pose7 = torch.arange(14, dtype=torch.float32).reshape(2, 7)
    w.append_batch(["b.jpg", "a.jpg"], pose7)
    with contextlib.redirect_stdout(io.StringIO()):
        path = w.export(out_dir=str(tmp_path), suffix="b")
    rows = open(path).read().strip().split("\n")
    assert rows[0].startswith("a.jpg,7.0,8.0") and rows[1].startswith("b.jpg,0.0,1.0")


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 3000):
        for world in (1, 2, 3, 8):
            spans = [pipeline.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, n_total, q):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    s, e = pipeline.shard_range(n_total, rank, world)
    local = torch.arange(s, e, dtype=torch.float32)[:, None].repeat(1, 7) + torch.arange(7) * 0.125
    out = pipeline.gather_poses(local, n_total)
    q.put((rank, out.numpy()))
    dist.destroy_process_group()


def test_gather_poses_world2_gloo():
    n_total, world = 7, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = np.arange(n_total, dtype=np.float32)[:, None].repeat(7, 1) + np.arange(7) * 0.125
    for r in range(world):
        np.testing.assert_array_equal(res[r], expect)


def test_gather_poses_without_process_group_is_identity():
    x = torch.zeros((3, 7))
    assert pipeline.gather_poses(x, 3) is x


def test_camera_constants_match_reference_values():
    """utils.py:24-39 / base_utils.py:250-252: fx/ppx = 3003.41 px, principal point at the image centre."""
    from esa_pose_estimation_b200.camera import INTRINSICS, Camera
    assert abs(Camera.K[0, 0] - 0.0176 / 5.86e-6) < 1e-9 and Camera.K[0, 2] == 960 and Camera.K[1, 2] == 600
    assert np.allclose(Camera.K, INTRINSICS["esa"], atol=1e-4)


def test_crop_window_matches_the_literal_restatement():
    """camera.crop_window (vectorised) against data_load_val.py:125-176 restated literally (oracle/decode.py), on boxes
    that hit every clamp: off the left / top edge, past the right / bottom edge, larger than the frame, square = scale."""
    from esa_pose_estimation_b200.camera import crop_window
    from oracle.decode import crop_window as crop_ref
    rng = np.random.default_rng(3)
    boxes = []
    for _ in range(400):
        cx, cy = rng.uniform(-50, 1970), rng.uniform(-50, 1250)
        hw, hh = rng.uniform(5, 900), rng.uniform(5, 700)
        b = [cx - hw, cy - hh, cx + hw, cy + hh]
        boxes.append([int(v) for v in b] if rng.random() < 0.5 else b)
    boxes += [[0, 0, 1920, 1200], [-30, -40, 100, 90], [1800, 1100, 2000, 1300], [400, 300, 765, 600], [10, 10, 14, 14]]     # (a box under 2 px divides by zero in the reference)
    win, size, rate = crop_window(boxes)
    for i, b in enumerate(boxes):
        w_ref, s_ref, r_ref = crop_ref(b)
        assert list(win[i]) == w_ref and int(size[i]) == s_ref and rate[i] == r_ref, (b, win[i], w_ref)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the oracle port on the host cores): one JSON line with the contract's keys."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "4"], capture_output=True, text=True, timeout=300, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "config[1]" in d["config"]["workload"]
    # both arms print the SAME config object (the driver compares them key by key)
    sys.path.insert(0, root)
    import bench
    a = bench.parse.__globals__["argparse"].Namespace(batch=64, size=256, vn=11, hn=512, fg=0.25, gpus=1)
    assert d["config"] == bench.config_dict(a, 1)
