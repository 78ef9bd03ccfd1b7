"""GPU parity: csrc/decode.cu (through the C ABI) against the committed reference vectors and
the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import decode as odec
from tests.synth import make_heatmaps

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["gauss11", "gauss30", "randinit11", "edges11"])
def test_decode_matches_reference_vectors(cuda_dev, golden_dir, name):
    from esa_pose_estimation_b200 import inference
    g = np.load(os.path.join(golden_dir, "decode_%s.npz" % name))
    hm = g["hm"]
    preds, maxvals = inference.get_max_preds(hm.copy())
    np.testing.assert_array_equal(preds, g["preds"])                 # argmax bit-exact
    np.testing.assert_array_equal(maxvals, g["maxvals"])
    co = [g["preds"][0, i].copy() for i in range(hm.shape[1])]
    final = inference.get_final(hm.copy(), co)
    np.testing.assert_allclose(final, g["final"], rtol=0, atol=1e-3)  # north-star tolerance: 1e-3 px
    assert np.abs(final - g["final"]).max() < 2e-5                    # in practice float32 rounding
    np.testing.assert_array_equal(np.asarray(co), final)              # coords mutated in place
    gp, gm = inference.getPrediction(torch.from_numpy(hm).to(cuda_dev))
    np.testing.assert_array_equal(gp.cpu().numpy(), g["getpred"])
    np.testing.assert_array_equal(gm.cpu().numpy(), g["getpred_max"])


@pytest.mark.parametrize("name", ["gauss11", "gauss30", "randinit11", "edges11"])
def test_dark_decode_matches_reference_vectors(cuda_dev, golden_dir, name):
    """inference.py:154 get_final2 (11x11 Gaussian blur, renormalise, float32 log, 2x2 Hessian Newton step)
    against the vectors produced by the reference's own function."""
    from esa_pose_estimation_b200 import inference
    g = np.load(os.path.join(golden_dir, "decode_%s.npz" % name))
    hm = g["hm"]
    co = [g["preds"][0, i].copy() for i in range(hm.shape[1])]
    final2 = inference.get_final2(hm.copy(), co)
    err = np.abs(final2 - g["final2"])
    step = np.abs(g["final2"] - g["preds"][0])
    # north-star tolerance 1e-3 px; random-init maps have near-singular Hessians whose Newton steps are
    # hundreds of pixels long and float32-noise dominated in the reference itself: relative there
    assert (err <= 1e-3 + 1e-3 * step).all(), float(err.max())
    np.testing.assert_array_equal(np.asarray(co), final2)


def test_dark_decode_batched_large_map(cuda_dev):
    """Row-band streaming of a map larger than one shared-memory band, against the oracle."""
    from esa_pose_estimation_b200 import inference
    hm, _ = make_heatmaps(77, 2, 3, 384, 384, "gauss")
    xy, _, _ = inference.decode_heatmaps(torch.from_numpy(hm).to(cuda_dev), refine=False)
    out = inference.refine_dark(torch.from_numpy(hm).to(cuda_dev), xy).cpu().numpy()
    for bi in range(2):
        co = [xy[bi, i].cpu().numpy().copy() for i in range(3)]
        ref = odec.get_final2(hm[bi:bi + 1].copy(), co)
        np.testing.assert_allclose(out[bi], ref, rtol=0, atol=1e-3)


@pytest.mark.parametrize("shape", [(3, 11, 128, 128), (2, 30, 64, 64), (2, 5, 37, 53), (1, 11, 384, 384), (1, 2, 7, 9)])
def test_decode_fused_matches_oracle(cuda_dev, shape):
    from esa_pose_estimation_b200 import inference
    b, k, h, w = shape
    for kind in ("gauss", "randinit", "edges"):
        hm, _ = make_heatmaps(b * 7 + h, b, k, h, w, kind)
        xy, mv, idx = inference.decode_heatmaps(torch.from_numpy(hm).to(cuda_dev))
        xy, mv, idx = xy.cpu().numpy(), mv.cpu().numpy(), idx.cpu().numpy()
        for bi in range(b):
            preds, maxvals, flat = odec.decode_frame(hm[bi:bi + 1])
            np.testing.assert_array_equal(idx[bi], flat)
            np.testing.assert_array_equal(mv[bi], maxvals.astype(np.float32))
            np.testing.assert_allclose(xy[bi], preds, rtol=0, atol=1e-3)


def test_decode_nan_and_full_size(cuda_dev):
    from esa_pose_estimation_b200 import inference
    hm = np.zeros((1, 2, 16, 16), np.float32)
    hm[0, 0, 3, 4] = np.nan
    hm[0, 0, 9, 9] = 5.0
    hm[0, 1] = -np.inf
    xy, mv, idx = inference.decode_heatmaps(torch.from_numpy(hm).to(cuda_dev), refine=False)
    assert idx.cpu().tolist() == [[3 * 16 + 4, 0]]                    # NaN is the maximum; all -inf -> 0
    assert np.isnan(mv[0, 0].item())
    # SPEED-sized frame: argmax position planted, checked without an oracle pass
    big = torch.full((1, 3, 1200, 1920), -1.0, device=cuda_dev)
    pos = [(17, 1919), (1199, 0), (600, 960)]
    for kk, (y, x) in enumerate(pos):
        big[0, kk, y, x] = 2.0 + kk
    xy, mv, idx = inference.decode_heatmaps(big, refine=False)
    assert [tuple(v) for v in xy[0].cpu().long().tolist()] == [(x, y) for (y, x) in pos]
    assert idx[0].cpu().tolist() == [y * 1920 + x for (y, x) in pos]
