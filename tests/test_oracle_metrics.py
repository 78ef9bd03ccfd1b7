"""oracle/metrics.py (CPU restatement of evaluation.py:340-411) against the reference's own Evaluator
methods (tests/golden/metrics_ref.npz, produced by tests/golden/make_golden_metrics.py)."""
import os

import numpy as np

from oracle import metrics as om


def test_oracle_metrics_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "metrics_ref.npz"))
    K = g["K"]
    for tag in ("a", "b"):
        model, diameter, pred, gt = g[tag + "_model"], float(g[tag + "_diameter"]), g[tag + "_pred"], g[tag + "_gt"]
        for i in range(pred.shape[0]):
            d, ok = om.projection_2d(pred[i], gt[i], model, K)
            assert d == g[tag + "_proj"][i] and ok == g[tag + "_proj_ok"][i]
            d, ok = om.add_metric(pred[i], gt[i], model, diameter)
            assert d == g[tag + "_add"][i] and ok == g[tag + "_add_ok"][i]
            cm, deg, ok = om.cm_degree_5_metric(pred[i], gt[i])
            assert cm == g[tag + "_cm"][i] and ok == g[tag + "_cm5_ok"][i]
            assert (np.isnan(deg) and np.isnan(g[tag + "_deg"][i])) or deg == g[tag + "_deg"][i]
            d, ok = om.projection_2d(pred[i], gt[i], model, K, sym=True)
            assert d == g[tag + "_proj_sym"][i] and ok == g[tag + "_proj_sym_ok"][i]
            d, ok = om.add_metric(pred[i], gt[i], model, diameter, sym=True)
            assert d == g[tag + "_add_sym"][i] and ok == g[tag + "_add_sym_ok"][i]
        # the fixtures cover passing and failing poses for every flag
        for k in ("_proj_ok", "_add_ok", "_cm5_ok"):
            assert g[tag + k].any() and not g[tag + k].all()


def test_nearest_idx_is_first_minimum_in_float32():
    rng = np.random.default_rng(0)
    ref = rng.normal(size=(300, 3)).astype(np.float32)
    ref[17] = ref[5]                                   # exact tie: the lower index wins
    que = np.concatenate([ref[5:6] + 1e-4, rng.normal(size=(50, 3)).astype(np.float32)])
    idx = om.nearest_idx(ref, que)
    assert idx[0] == 5
    d = ((ref[None].astype(np.float64) - que[:, None].astype(np.float64)) ** 2).sum(-1)
    # float64 argmin agrees except where two distances differ by less than float32 resolution
    agree = idx == d.argmin(1)
    assert agree.mean() > 0.95
    for q in np.nonzero(~agree)[0]:
        assert abs(d[q, idx[q]] - d[q].min()) <= 1e-5 * d[q].min()
