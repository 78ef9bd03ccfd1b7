"""End-to-end GPU paths: heatmaps -> pose (val.py flow) and vector field -> voting -> pose, against
the oracle on identical inputs, plus __graft_entry__.smoke()."""
import numpy as np
import pytest
import torch

from oracle import decode as odec
from oracle import pose as opose
from oracle import voting as ov
from tests.synth import ESA_K, make_heatmaps, make_pose_case, make_vertex_field, rodrigues, tango_model, vertex_hwvn2

pytestmark = pytest.mark.gpu


def _ang(r1, r2):
    c = (np.trace(r1 @ r2.T) - 1) / 2
    return np.degrees(np.arccos(np.clip(c, -1, 1)))


def test_heatmaps_to_pose_matches_oracle(cuda_dev):
    from esa_pose_estimation_b200 import pipeline
    B, kp, S = 6, 11, 128
    model = tango_model(kp, seed=9)
    hms = np.zeros((B, kp, S, S), np.float32); bbox = np.zeros((B, 2)); rate = np.zeros(B)
    ys, xs = np.mgrid[0:S, 0:S].astype(np.float64)
    rng = np.random.default_rng(3)
    for i in range(B):
        c = make_pose_case(9000 + i, kp, 0.0, 0, model=model)
        lo, hi = c["p2d"].min(0), c["p2d"].max(0)
        size = (hi - lo).max() * 1.3 + 8
        bbox[i] = (lo + hi) / 2 - size / 2
        rate[i] = S / size
        crop = (c["p2d"] - bbox[i]) * rate[i]
        for k in range(kp):
            g = np.exp(-((xs - crop[k, 0]) ** 2 + (ys - crop[k, 1]) ** 2) / 8.0) * rng.uniform(0.6, 1.0)
            hms[i, k] = (g + rng.normal(0, 0.01, (S, S))).astype(np.float32)
    out = pipeline.poses_from_heatmaps(torch.from_numpy(hms).to(cuda_dev), torch.from_numpy(bbox).to(cuda_dev),
                                       torch.from_numpy(rate).to(cuda_dev), torch.from_numpy(model).to(cuda_dev),
                                       torch.from_numpy(ESA_K).to(cuda_dev), min_k=8)
    rt6 = out["rt6"].cpu().numpy(); pose7 = out["pose7"].cpu().numpy()
    for i in range(B):
        preds, maxvals, _ = odec.decode_frame(hms[i:i + 1])
        o = opose.frame_pose(preds, maxvals, bbox[i], rate[i], model, ESA_K, min_k=8)
        assert _ang(rodrigues(rt6[i, :3]), o["pose34"][:, :3]) < 1e-3
        assert np.linalg.norm(rt6[i, 3:] - o["t"]) / np.linalg.norm(o["t"]) < 1e-4
        assert min(np.abs(pose7[i, :4] - o["q"]).max(), np.abs(pose7[i, :4] + o["q"]).max()) < 1e-5


def test_vertex_to_pose_matches_oracle(cuda_dev):
    """mask + field -> voting -> un-crop -> EPnP-RANSAC -> LM, keypoints AND poses against the oracle
    (rotation 1e-3 deg, translation 1e-4 relative: north_star's tolerances)."""
    import cv2
    from esa_pose_estimation_b200 import pipeline, ransac_voting_gpu as rv
    from tests.synth import make_pose_field
    B, vn, S, hn = 4, 11, 96, 256
    mask, vertex, model, geom, kcrop, _, _ = make_pose_field(61, B, S, vn, 0.35, noise_deg=1.5)
    vx = vertex_hwvn2(vertex)
    fn = ov.default_idxs_fn(3)
    idxs = np.zeros((B, 1, hn, vn, 2), np.int32)
    for bi in range(B):
        idxs[bi, 0] = fn(bi, 0, hn, vn, int((mask[bi] != 0).sum()))
    out = pipeline.poses_from_vertex(torch.from_numpy(mask).to(cuda_dev),
                                     rv.vertex_layer_reshape(torch.from_numpy(vertex).to(cuda_dev)),
                                     torch.from_numpy(model).to(cuda_dev), torch.from_numpy(ESA_K).to(cuda_dev),
                                     round_hyp_num=hn, idxs=torch.from_numpy(idxs).to(cuda_dev),
                                     bbox_xy=torch.from_numpy(np.ascontiguousarray(geom[:, :2])).to(cuda_dev),
                                     rate=torch.from_numpy(np.ascontiguousarray(geom[:, 2])).to(cuda_dev))
    k_o = ov.ransac_voting_layer_v3(mask, vx, hn, idxs_fn=fn)
    np.testing.assert_allclose(out["kpts"].cpu().numpy(), k_o, atol=1e-3)
    assert np.abs(k_o - kcrop).max() < 2.0
    assert out["pose7"].shape == (B, 7) and int(out["status"].abs().sum()) == 0
    rt6 = out["rt6"].cpu().numpy()
    for bi in range(B):
        p2d = k_o[bi].astype(np.float64) * (1.0 / geom[bi, 2]) + geom[bi, :2]
        rt = opose.pnp(model, p2d, ESA_K, cv2.SOLVEPNP_EPNP)
        r_exp, _ = cv2.Rodrigues(rt[:, :3])
        cam = opose.cpnp(model, p2d, ESA_K, np.concatenate([r_exp.reshape(3), rt[:, 3]]))
        assert _ang(rodrigues(rt6[bi, :3]), rodrigues(cam[:3])) < 1e-3
        assert np.linalg.norm(rt6[bi, 3:] - cam[3:]) / np.linalg.norm(cam[3:]) < 1e-4


def test_smoke_entry(cuda_dev):
    import __graft_entry__ as g
    g.smoke()


def test_host_field_chunked_pipeline_equals_device_run(cuda_dev):
    """Pinned host field, batch cut in chunks on two streams (gather of chunk i+1 under the voting
    of chunk i): keypoints, poses and the torch generator end state are identical to one device run."""
    from esa_pose_estimation_b200 import pipeline, ransac_voting_gpu as rv
    B, vn, S, hn = 20, 5, 64, 128
    mask, vertex, _ = make_vertex_field(71, B, S, S, vn, 0.4, noise_deg=1.5)
    mask[3] = 0                                                          # a skipped image inside a chunk
    model = torch.from_numpy(tango_model(vn, seed=9)).to(cuda_dev)
    K = torch.from_numpy(np.array([[120.0, 0, 32], [0, 120.0, 32], [0, 0, 1]])).to(cuda_dev)
    m_h, v_h = torch.from_numpy(mask).pin_memory(), torch.from_numpy(vertex).pin_memory()
    torch.manual_seed(5)
    ref = pipeline.poses_from_vertex(m_h.to(cuda_dev), rv.vertex_layer_reshape(v_h.to(cuda_dev)), model, K, round_hyp_num=hn)
    off_ref = torch.cuda.default_generators[cuda_dev.index or 0].get_offset()
    # device-resident inputs through the three-stream route, calls in flight on two caller streams
    m_d, v_d = m_h.to(cuda_dev), v_h.to(cuda_dev)
    idxs = torch.randint(0, 1 << 30, (B, 1, hn, vn, 2), dtype=torch.int32, device=cuda_dev)
    ref_i = pipeline.poses_from_vertex(m_d, rv.vertex_layer_reshape(v_d), model, K, round_hyp_num=hn, idxs=idxs, raw32=True)
    torch.cuda.synchronize()
    streams = [torch.cuda.Stream(device=cuda_dev) for _ in range(2)]
    outs = []
    for i in range(6):
        with torch.cuda.stream(streams[i % 2]):
            outs.append(pipeline.poses_from_vertex(m_d, rv.vertex_layer_reshape(v_d), model, K, round_hyp_num=hn, idxs=idxs,
                                                   raw32=True, pipelined=True))
    torch.cuda.synchronize()
    live = torch.arange(B) != 3
    for o in outs:
        assert torch.equal(o["kpts"].view(torch.int32), ref_i["kpts"].view(torch.int32))
        assert torch.equal(o["pose7"][live].view(torch.int32), ref_i["pose7"][live].view(torch.int32))
    for chunks in (4, 3, 1):
        torch.manual_seed(5)
        out = pipeline.poses_from_vertex(m_h, rv.vertex_layer_reshape(v_h), model, K, round_hyp_num=hn, chunks=chunks)
        assert torch.cuda.default_generators[cuda_dev.index or 0].get_offset() == off_ref
        assert torch.equal(out["kpts"].view(torch.int32), ref["kpts"].view(torch.int32)), chunks
        live = torch.ones(B, dtype=torch.bool); live[3] = False
        assert torch.equal(out["pose7"][live].view(torch.int32), ref["pose7"][live].view(torch.int32)), chunks


def test_host_pipeline_with_many_hypotheses(cuda_dev):
    """Split (gather / vote) runs with > 512 hypotheses per keypoint (the 8-hypotheses-per-thread kernel,
    whose occupancy leaves the gather stage room without the shared-memory padding)."""
    from esa_pose_estimation_b200 import pipeline, ransac_voting_gpu as rv
    B, vn, S, hn = 16, 3, 64, 1024
    mask, vertex, _ = make_vertex_field(73, B, S, S, vn, 0.4, noise_deg=1.5)
    model = torch.from_numpy(tango_model(vn + 3, seed=9)[:vn]).to(cuda_dev)
    K = torch.from_numpy(np.array([[120.0, 0, 32], [0, 120.0, 32], [0, 0, 1]])).to(cuda_dev)
    m_h, v_h = torch.from_numpy(mask).pin_memory(), torch.from_numpy(vertex).pin_memory()
    torch.manual_seed(9)
    ref = rv.ransac_voting_layer_v3(m_h.to(cuda_dev), rv.vertex_layer_reshape(v_h.to(cuda_dev)), hn)
    torch.manual_seed(9)
    out = pipeline.poses_from_vertex(m_h, rv.vertex_layer_reshape(v_h), model, K, round_hyp_num=hn, chunks=4)
    assert torch.equal(out["kpts"].view(torch.int32), ref.view(torch.int32))


def test_heatmap_path_replays_as_a_cuda_graph(cuda_dev):
    """The heatmap chain (decode, refine, pose) neither synchronises nor allocates inside the C library, so it can
    be captured once and replayed: the replay is bit-identical to the eager call, also after the inputs change."""
    from esa_pose_estimation_b200 import pipeline
    B, kp, S = 3, 11, 64
    model = torch.from_numpy(tango_model(kp, seed=9)).to(cuda_dev)
    K = torch.from_numpy(np.array([[120.0, 0, 32], [0, 120.0, 32], [0, 0, 1]])).to(cuda_dev)
    g = pipeline.GraphedHeatmapPose(B, kp, S, S, model, K, cuda_dev, min_k=8)
    for seed in (5, 6):
        hm, _ = make_heatmaps(seed, B, kp, S, S, "gauss")
        hm_t = torch.from_numpy(hm).to(cuda_dev)
        bbox = torch.zeros((B, 2), dtype=torch.float64, device=cuda_dev)
        rate = torch.ones((B,), dtype=torch.float64, device=cuda_dev)
        eager = pipeline.poses_from_heatmaps(hm_t, bbox, rate, model, K, min_k=8)
        out = g(hm_t, bbox, rate)
        torch.cuda.synchronize()
        for k in ("pose7", "rt6", "xy", "maxval"):
            a, b = out[k], eager[k]
            assert torch.equal(a.view(torch.uint8), b.view(torch.uint8)), k


def test_heatmap_path_in_pieces_equals_one_piece(cuda_dev):
    """A long set is decoded piece by piece on the caller's stream while a second stream solves the poses of the
    previous piece: same bits as the one-piece call."""
    from esa_pose_estimation_b200 import pipeline
    B, kp, S = 23, 11, 48
    hm, _ = make_heatmaps(15, B, kp, S, S, "gauss")
    hm_t = torch.from_numpy(hm).to(cuda_dev)
    model = torch.from_numpy(tango_model(kp, seed=9)).to(cuda_dev)
    K = torch.from_numpy(np.array([[120.0, 0, 24], [0, 120.0, 24], [0, 0, 1]])).to(cuda_dev)
    bbox = torch.rand((B, 2), dtype=torch.float64, device=cuda_dev) * 10
    rate = torch.rand((B,), dtype=torch.float64, device=cuda_dev) + 0.5
    one = pipeline.poses_from_heatmaps(hm_t, bbox, rate, model, K, min_k=8, chunk_frames=1000)
    for cf in (5, 8, 22):
        many = pipeline.poses_from_heatmaps(hm_t, bbox, rate, model, K, min_k=8, chunk_frames=cf)
        torch.cuda.synchronize()
        for k in one:
            assert torch.equal(many[k].view(torch.uint8), one[k].view(torch.uint8)), (cf, k)
