"""Replay helpers for tests/golden/voting_drivers.npz (made by tests/golden/make_golden_voting.py from the
reference's own ransac_voting_gpu.py): the random draws the reference made, in its call order."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "voting_drivers.npz")


def load():
    return np.load(PATH)


class Replay:
    """idxs_fn / selection_fn for oracle/voting.py that hand out the recorded draws in order."""

    def __init__(self, g, name):
        self.idxs = list(g[name + "_idxs"]) if int(g[name + "_n_idxs"]) else []
        self.sel = list(g[name + "_sel"]) if int(g[name + "_n_sel"]) else []
        self.i = self.s = 0

    def idxs_fn(self, bi, round_idx, hn, vn, tn):
        a = self.idxs[self.i]
        self.i += 1
        assert a.shape == (hn, vn, 2) and a.max() < tn
        return a

    def selection_fn(self, bi, h, w):
        a = self.sel[self.s]
        self.s += 1
        assert a.shape == (h, w)
        return a

    def exhausted(self):
        return self.i == len(self.idxs) and self.s == len(self.sel)


def dense_draws(g, name, live, rounds=1, shape_hw=None):
    """The recorded draws laid out the way the CUDA wrapper takes explicit inputs:
    idxs [n_items, rounds, hn, vn, 2] int32 (zeros for items the reference skipped as degenerate) and,
    if any were drawn, selection [n_items, h, w] float32.  `live` flags the items (images, or
    (image, class) pairs in loop order) that reached the random draws."""
    idxs = g[name + "_idxs"]
    hn, vn = idxs.shape[1], idxs.shape[2]
    n = len(live)
    out = np.zeros((n, rounds, hn, vn, 2), np.int32)
    k = 0
    for i in range(n):
        if live[i]:
            out[i] = idxs[k:k + rounds]
            k += rounds
    assert k == idxs.shape[0]
    sel = None
    if int(g[name + "_n_sel"]):
        s = g[name + "_sel"]
        sel = np.ones((n,) + s.shape[1:], np.float32)
        k = 0
        for i in range(n):
            if live[i]:
                sel[i] = s[k]
                k += 1
        assert k == s.shape[0]
    return out, sel
