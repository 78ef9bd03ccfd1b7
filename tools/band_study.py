#!/usr/bin/env python
"""CPU study (numpy, float64) for DESIGN.md section 12: how many (hypothesis, pixel) pairs of config[1] fall inside the
undecided band of vote_count, and in how many warp-trips (2 records x 32 lanes x 4 hypotheses) at least one does, for
  (a) the band the kernel computes per pair (w = gamma a' + EH[h]), and
  (b) a band that is constant per thread over a 256-record tile (unit-normalised records, |h - c| bounded by the
      distance to the tile's bounding circle), which would save the per-pair band FMA.
Result on the bench data: (a) 4.3e-4 of the pairs, 10.5 % of the warp-trips; (b) 1.0e-3 and 23 %.
    python tools/band_study.py"""
import numpy as np, sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
from bench import make_batch_numpy
from oracle import voting as ov
from tests.synth import vertex_hwvn2
a = SimpleNamespace(size=256, vn=11, fg=0.25)
mask, vertex, model, geom, kcrop = make_batch_numpy(a, 11, 1)
vx = vertex_hwvn2(vertex)
_, coords, direct = ov.compact(mask[0] != 0, vx[0], 30000, ov.default_selection_fn(0), 0)
tn = coords.shape[0]; hn = 512
rng = np.random.default_rng(0)
T = 0.999; u = 2.0**-24
theta = np.arccos(T); k = np.tan(theta)
beta = 10.5*u/(T*np.sin(0.75*theta)); gamma = beta*(1/k+1)*1.0001
es = 3*(6.2*k+4.1)*u*2; ea = es*512 + 3.5e-6*(1+k)
res = []
for v in range(3):
    idxs = rng.integers(0, tn, (hn, 11, 2)).astype(np.int32)
    hyp = ov.generate_hypothesis(direct, coords, idxs)[:, v].astype(np.float64)     # [hn,2]
    n = direct[:, v].astype(np.float64)                                              # [tn,2]
    nm = np.maximum(np.abs(n[:,0]), np.abs(n[:,1])); sc = 2.0**(-np.floor(np.log2(nm)))
    ns = n*sc[:,None]                                                                 # max-norm in [1,2)
    c = coords.astype(np.float64)
    d = hyp[None,:,:] - c[:,None,:]                                                   # [tn,hn,2]
    ap = k*(ns[:,None,0]*d[...,0] + ns[:,None,1]*d[...,1])
    pp = -ns[:,None,1]*d[...,0] + ns[:,None,0]*d[...,1]
    m = ap - np.abs(pp)
    EH = es*np.abs(hyp).sum(1) + ea
    w_pair = gamma*ap + EH[None,:]
    und_pair = np.abs(m) <= w_pair
    # tile-constant band: tiles of 256 consecutive pixels, thread = 4 hypotheses {2t,2t+1,2t+256,2t+257}
    n2 = np.linalg.norm(ns, axis=1)
    und_tile = np.zeros_like(und_pair)
    hyp_of_thread = np.stack([2*np.arange(128), 2*np.arange(128)+1, 2*np.arange(128)+256, 2*np.arange(128)+257], 1)  # [128,4]
    for t0 in range(0, tn, 256):
        cc = c[t0:t0+256]; c0 = 0.5*(cc.min(0)+cc.max(0)); r = np.linalg.norm(cc - c0, axis=1).max()
        Dh = np.linalg.norm(hyp - c0, axis=1) + r                                       # bound on |h - c| over the tile
        Dt = Dh[hyp_of_thread].max(1)                                                   # per thread
        EHt = EH[hyp_of_thread].max(1)
        nmax = n2[t0:t0+256].max()                                                      # |n|_2 bound of the tile (or normalise records)
        for mode, nb in (("maxnorm", nmax), ("unit", 1.0)):
            pass
        # unit-normalised records: a' scales by 1/|n|_2, so does m; band W = gamma*k*D + EH (EH scaled conservatively by 1)
        W = gamma*k*Dt + EHt                                                            # [128]
        Wh = np.zeros(hn); Wh[hyp_of_thread] = W[:,None]
        m_unit = m[t0:t0+256]/n2[t0:t0+256,None]
        und_tile[t0:t0+256] = np.abs(m_unit) <= Wh[None,:]
    def warp_trip_rate(und):
        # warp-trip = 2 consecutive records x (32 lanes x 4 hyps): lanes 32w..32w+31 <-> hyps via hyp_of_thread
        tn2 = (tn//2)*2
        u2 = und[:tn2].reshape(tn2//2, 2, hn).any(1)                                    # [trips, hn]
        per_thread = u2[:, hyp_of_thread].any(2)                                        # [trips,128]
        per_warp = per_thread.reshape(-1, 4, 32).any(2)                                 # [trips,4]
        return per_warp.mean()
    res.append((und_pair.mean(), warp_trip_rate(und_pair), und_tile.mean(), warp_trip_rate(und_tile)))
    print("keypoint", v, "pair band: undecided %.2e, warp-trips with an undecided pair %.3f | tile band: %.2e, %.3f" % res[-1])
