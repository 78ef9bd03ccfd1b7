"""Device metrics (csrc/metrics.cu) against the reference's own Evaluator methods (golden), the oracle, and --
for the nearest-point search -- the reference's nearest_neighborhood.cu compiled unmodified (oracle/_ref)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import _lib as olib
from oracle import metrics as om

pytestmark = pytest.mark.gpu


def test_pose_metrics_match_the_reference(cuda_dev, golden_dir):
    from esa_pose_estimation_b200 import evaluation as ev
    g = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    K = g["K"]
    for tag in ("a", "b"):
        model, diameter, pred, gt = g[tag + "_model"], float(g[tag + "_diameter"]), g[tag + "_pred"], g[tag + "_gt"]
        o = {k: v.cpu().numpy() for k, v in ev.pose_metrics(pred, gt, model, K).items()}
        s = {k: v.cpu().numpy() for k, v in ev.pose_metrics(pred, gt, model, K, sym=True, want=("proj", "add")).items()}
        np.testing.assert_allclose(o["proj"], g[tag + "_proj"], rtol=1e-10)
        np.testing.assert_allclose(o["add"], g[tag + "_add"], rtol=1e-10)
        np.testing.assert_allclose(o["cm"], g[tag + "_cm"], rtol=1e-12)
        # arccos near 1 amplifies the last bit of the trace: absolute tolerance in degrees
        np.testing.assert_allclose(o["deg"], g[tag + "_deg"], rtol=1e-9, atol=1e-5, equal_nan=True)
        np.testing.assert_allclose(s["proj"], g[tag + "_proj_sym"], rtol=1e-10)
        np.testing.assert_allclose(s["add"], g[tag + "_add_sym"], rtol=1e-10)
        # flags through the Evaluator mirror, pose by pose and as one block
        e1, e2, es = ev.Evaluator(), ev.Evaluator(), ev.Evaluator()
        for i in range(pred.shape[0]):
            e1.projection_2d(pred[i], gt[i], model, K)
            e1.add_metric(pred[i], gt[i], model, diameter)
            e1.cm_degree_5_metric(pred[i], gt[i])
            es.projection_2d_sym(pred[i], gt[i], model, K)
            es.add_metric_sym(pred[i], gt[i], model, diameter)
        e2.evaluate_batch(pred, gt, model, K, diameter)
        for e in (e1, e2):
            assert e.projection_2d_recorder == g[tag + "_proj_ok"].tolist()
            assert e.add_recorder == g[tag + "_add_ok"].tolist()
            assert e.cm_degree_5_recorder == g[tag + "_cm5_ok"].tolist()
        assert es.projection_2d_recorder == g[tag + "_proj_sym_ok"].tolist()
        assert es.add_recorder == g[tag + "_add_sym_ok"].tolist()


@pytest.mark.parametrize("dim,pn1,pn2", [(3, 1531, 977), (2, 4000, 333), (3, 5, 2049)])
def test_nearest_idx_bitwise_vs_reference_kernel(cuda_dev, dim, pn1, pn2):
    """our kernel == the C restatement == the reference's findNearestPointIdxLauncher (its own host launcher and
    kernels, compiled unmodified for sm_100a)."""
    from esa_pose_estimation_b200 import evaluation as ev
    rng = np.random.default_rng(dim * 1000 + pn1)
    ref = rng.normal(size=(pn1, dim)).astype(np.float32)
    que = rng.normal(size=(pn2, dim)).astype(np.float32)
    que[: min(pn2, pn1) // 2] = ref[: min(pn2, pn1) // 2] + rng.normal(size=(min(pn2, pn1) // 2, dim)).astype(np.float32) * 1e-3
    ref[pn1 // 2] = ref[0]                              # exact duplicate: first index must win
    ours = ev.find_nearest_point_idx(ref, que)
    orc = om.nearest_idx(ref, que)
    assert np.array_equal(ours, orc)
    lib = olib.ref_nearest_lib(required=False)
    if lib is None:
        pytest.skip("oracle/_ref/libref_nearest.so not built")
    idx = np.zeros(pn2, np.int32)
    lib.findNearestPointIdxLauncher(olib.fptr(ref), olib.fptr(que), olib.iptr(idx), ctypes.c_int(1), ctypes.c_int(pn1),
                                    ctypes.c_int(pn2), ctypes.c_int(dim), ctypes.c_int(0))
    assert np.array_equal(ours, idx)


def test_metrics_large_model_properties(cuda_dev):
    """n_model = 8192 (ADD-S is 6.7e7 distance evaluations per pose): identical poses give 0; ADD-S <= ADD; a sample
    of poses equals the oracle."""
    from esa_pose_estimation_b200 import evaluation as ev
    from tests.synth import rodrigues
    rng = np.random.default_rng(5)
    model = rng.uniform(-0.1, 0.1, (8192, 3))
    K = np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]])
    n = 24
    gt = np.zeros((n, 3, 4)); pred = np.zeros((n, 3, 4))
    for i in range(n):
        gt[i, :, :3] = rodrigues(rng.normal(size=3)); gt[i, :, 3] = (0.0, 0.0, 1.0)
        pred[i, :, :3] = rodrigues(rng.normal(size=3) * 0.05) @ gt[i, :, :3]; pred[i, :, 3] = gt[i, :, 3] + rng.normal(size=3) * 0.005
    z = ev.pose_metrics(gt, gt, model, K, sym=True)
    assert float(z["add"].abs().max()) == 0.0 and float(z["proj"].abs().max()) == 0.0
    a = ev.pose_metrics(pred, gt, model, K)
    s = ev.pose_metrics(pred, gt, model, K, sym=True)
    assert bool((s["add"] <= a["add"] + 1e-12).all()) and bool((s["proj"] <= a["proj"] + 1e-9).all())
    for i in (0, 7, 23):
        assert abs(float(s["add"][i]) - om.add_metric(pred[i], gt[i], model, 1.0, sym=True)[0]) < 1e-12
        assert abs(float(a["add"][i]) - om.add_metric(pred[i], gt[i], model, 1.0)[0]) < 1e-12
        assert abs(float(s["proj"][i]) - om.projection_2d(pred[i], gt[i], model, K, sym=True)[0]) < 1e-8
