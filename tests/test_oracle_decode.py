"""Oracle (oracle/decode.py) pinned against vectors produced by the reference's inference.py."""
import glob
import os

import numpy as np
import pytest

from oracle import decode as odec


def _cases(golden_dir):
    return sorted(glob.glob(os.path.join(golden_dir, "decode_*.npz")))


def test_golden_files_present(golden_dir):
    assert len(_cases(golden_dir)) >= 4


@pytest.mark.parametrize("name", ["gauss11", "gauss30", "randinit11", "edges11"])
def test_oracle_matches_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "decode_%s.npz" % name))
    hm = g["hm"]
    preds, maxvals = odec.get_max_preds(hm.copy())
    np.testing.assert_array_equal(preds, g["preds"])
    np.testing.assert_array_equal(maxvals, g["maxvals"])
    co, mv = odec.two_stage_argmax(hm)
    np.testing.assert_array_equal(np.asarray(co), g["preds"][0])          # two-stage == flat argmax
    final = odec.get_final(hm.copy(), co)
    np.testing.assert_array_equal(np.asarray(final, np.float32), g["final"])  # same FP64 math.log path
    gp, gm = odec.get_prediction(hm)
    np.testing.assert_array_equal(gp, g["getpred"])
    np.testing.assert_array_equal(gm, g["getpred_max"])
    # DARK-style decode (inference.py:154-170): np.matrix in the reference, plain arrays here
    co2 = [g["preds"][0, i].copy() for i in range(hm.shape[1])]
    final2 = odec.get_final2(hm.copy(), co2)
    np.testing.assert_allclose(np.asarray(final2, np.float32), g["final2"], rtol=0, atol=1e-5)


def test_select_keypoints_ties_and_floor():
    mv = [0.9, 0.5, 0.9, 0.85, 0.1]
    assert odec.select_keypoints(mv, min_k=2) == [0, 2, 3]        # 3 above 0.8, ties keep lower index first
    assert odec.select_keypoints(mv, min_k=4) == [0, 2, 3, 1]
    assert len(odec.select_keypoints([0.1] * 30, min_k=24)) == 24
