"""Generates tests/golden/metrics_ref.npz by running the REFERENCE's own Evaluator methods
(/root/reference/evaluation.py:340-411, imported unmodified) on seeded synthetic poses and models.

What the reference needs and this image lacks is supplied from outside the file:
  * `plyfile`, `matplotlib.pyplot`: imported at module level, unused by the five metric methods -> empty stubs;
  * `lib.utils.extend_utils.extend_utils.find_nearest_point_idx`: a cffi wrapper around the CUDA launcher
    findNearestPointIdxLauncher whose `lib, ffi` import is commented out in the shipped file (and there is no GPU
    here) -> bound to oracle.metrics.nearest_idx, the C restatement of nearest_neighborhood.cu:48-121 that
    tests/test_metrics_gpu.py pins bitwise against the reference file compiled unmodified (oracle/_ref/libref_nearest.so);
  * `Evaluator.__init__` opens the LINEMOD model database: the methods are called unbound on a bare namespace that
    carries the recorders and the reference's own Projector.
Run from the repo root in the build container:  python tests/golden/make_golden_metrics.py"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import metrics as om  # noqa: E402
from tests.synth import random_pose, rodrigues  # noqa: E402


def load_reference():
    for name in ("plyfile", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            if name == "plyfile":
                m.PlyData = object
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, REF)
    ext = types.ModuleType("lib.utils.extend_utils.extend_utils")
    ext.find_nearest_point_idx = om.nearest_idx
    for pkg in ("lib", "lib.utils", "lib.utils.extend_utils"):
        if pkg not in sys.modules:
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF, *pkg.split("."))]
            sys.modules[pkg] = m
    sys.modules["lib.utils.extend_utils.extend_utils"] = ext
    import evaluation                                     # /root/reference/evaluation.py, unmodified
    return evaluation


def cases(seed, n_poses, n_model):
    rng = np.random.default_rng(seed)
    model = rng.uniform(-0.1, 0.1, (n_model, 3)) * np.array([1.0, 0.6, 0.3])
    diameter = float(np.max(np.linalg.norm(model[:, None] - model[None], axis=-1)))
    pred, gt = np.zeros((n_poses, 3, 4)), np.zeros((n_poses, 3, 4))
    for i in range(n_poses):
        rv, t = random_pose(rng)
        t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.1, 0.1), rng.uniform(0.5, 1.5)])
        gt[i, :, :3], gt[i, :, 3] = rodrigues(rv), t
        scale = [1e-4, 1e-2, 3e-2, 0.2, 3.0][i % 5]        # from sub-threshold to grossly wrong (incl. > 90 deg)
        drv = rng.normal(size=3) * scale
        pred[i, :, :3] = rodrigues(drv) @ gt[i, :, :3]
        pred[i, :, 3] = t + rng.normal(size=3) * scale * 0.05
    return model, diameter, pred, gt


def main():
    ev = load_reference()
    K = ev.Projector.intrinsic_matrix["linemod"].copy()
    out = {"K": K}
    for tag, (seed, n_poses, n_model) in {"a": (1, 20, 500), "b": (2, 10, 1531)}.items():
        model, diameter, pred, gt = cases(seed, n_poses, n_model)
        self = types.SimpleNamespace(projector=ev.Projector(), projection_2d_recorder=[], add_recorder=[],
                                     cm_degree_5_recorder=[], proj_mean_diffs=[], add_dists=[], cm=[], degree=[])
        sym = types.SimpleNamespace(projector=ev.Projector(), projection_2d_recorder=[], add_recorder=[],
                                    proj_mean_diffs=[], add_dists=[])
        deg_all = []
        for i in range(n_poses):
            ev.Evaluator.projection_2d(self, pred[i], gt[i], model, K)
            ev.Evaluator.add_metric(self, pred[i], gt[i], model, diameter)
            n_before = len(self.degree)
            ev.Evaluator.cm_degree_5_metric(self, pred[i], gt[i])
            deg_all.append(self.degree[-1] if len(self.degree) > n_before else np.nan)   # NaN angles are not recorded (:408)
            ev.Evaluator.projection_2d_sym(sym, pred[i], gt[i], model, K)
            ev.Evaluator.add_metric_sym(sym, pred[i], gt[i], model, diameter)
        out.update({
            tag + "_model": model, tag + "_diameter": diameter, tag + "_pred": pred, tag + "_gt": gt,
            tag + "_proj": np.array(self.proj_mean_diffs), tag + "_proj_ok": np.array(self.projection_2d_recorder),
            tag + "_add": np.array(self.add_dists), tag + "_add_ok": np.array(self.add_recorder),
            tag + "_cm": np.array(self.cm), tag + "_deg": np.array(deg_all), tag + "_cm5_ok": np.array(self.cm_degree_5_recorder),
            tag + "_proj_sym": np.array(sym.proj_mean_diffs), tag + "_proj_sym_ok": np.array(sym.projection_2d_recorder),
            tag + "_add_sym": np.array(sym.add_dists), tag + "_add_sym_ok": np.array(sym.add_recorder)})
    path = os.path.join(ROOT, "tests", "golden", "metrics_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
