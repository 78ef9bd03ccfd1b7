"""LINEMOD-style pose metrics on the device -- mirror of the reference's ``evaluation.py`` metric methods
(/root/reference/evaluation.py:340-411; twins in lib/utils/evaluation_utils.py:75-141).

  pose_metrics(pose_pred[N,3,4], pose_targets[N,3,4], model[n,3], K=None, sym=False)   batched, one launch
  find_nearest_point_idx(ref_pts, que_pts)      lib/utils/extend_utils/extend_utils.py:40 (same call, device search)
  Evaluator                                     the reference class's recorders and method signatures:
      projection_2d / projection_2d_sym (:340, :348), add_metric / add_metric_sym (:356, :385),
      cm_degree_5_metric (:399), plus evaluate_batch() for a whole [N,3,4] block at once.
The reference's ``Evaluator.__init__`` opens the LINEMOD model database; here the model points and diameter are
arguments, exactly as in its metric methods.
"""
import numpy as np
import torch

from . import _lib


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def _f64(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    return torch.as_tensor(np.ascontiguousarray(x, np.float64), device=dev)


def pose_metrics(pose_pred, pose_targets, model=None, K=None, sym=False, want=("proj", "add", "cm", "deg")):
    """-> dict of float64 [N] device tensors: proj (needs model and K), add (needs model), cm, deg."""
    dev = pose_pred.device if isinstance(pose_pred, torch.Tensor) and pose_pred.is_cuda else _dev()
    pred, gt = _f64(pose_pred, dev).reshape(-1, 3, 4), _f64(pose_targets, dev).reshape(-1, 3, 4)
    n_pose = pred.shape[0]
    assert gt.shape[0] == n_pose
    model_d = _f64(model, dev).reshape(-1, 3) if model is not None else None
    K_d = _f64(K, dev).reshape(-1)[:9].contiguous() if K is not None else None
    out = {}
    for k in want:
        if (k == "proj" and (model_d is None or K_d is None)) or (k == "add" and model_d is None):
            continue
        out[k] = torch.empty((n_pose,), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().epb_pose_metrics(_lib.ptr(pred), _lib.ptr(gt), n_pose, _lib.ptr(model_d),
                                          0 if model_d is None else model_d.shape[0], _lib.ptr(K_d), int(bool(sym)),
                                          _lib.ptr(out.get("proj")), _lib.ptr(out.get("add")), _lib.ptr(out.get("cm")),
                                          _lib.ptr(out.get("deg")), _lib.stream_ptr())
    _lib.check(st, "epb_pose_metrics")
    return out


def find_nearest_point_idx(ref_pts, que_pts):
    """for every point in que_pts the index of the nearest point in ref_pts (extend_utils.py:40-61).
    numpy in -> numpy int32 out like the reference; CUDA tensors in -> CUDA tensor out."""
    as_numpy = not isinstance(ref_pts, torch.Tensor)
    dev = ref_pts.device if not as_numpy and ref_pts.is_cuda else _dev()
    ref = torch.as_tensor(np.ascontiguousarray(ref_pts, np.float32) if as_numpy else ref_pts).to(device=dev, dtype=torch.float32).contiguous()
    que = torch.as_tensor(np.ascontiguousarray(que_pts, np.float32) if not isinstance(que_pts, torch.Tensor) else que_pts).to(
        device=dev, dtype=torch.float32).contiguous()
    assert ref.shape[1] == que.shape[1] and 1 < que.shape[1] <= 3
    idx = torch.empty((que.shape[0],), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        st = _lib.load().epb_nearest_point_idx(_lib.ptr(ref), _lib.ptr(que), _lib.ptr(idx), ref.shape[0], que.shape[0],
                                               ref.shape[1], _lib.stream_ptr())
    _lib.check(st, "epb_nearest_point_idx")
    return idx.cpu().numpy() if as_numpy else idx


class Evaluator(object):
    """Recorders and metric methods of the reference's Evaluator (evaluation.py:324-411)."""

    def __init__(self, name=None):
        self.name = name
        self.projection_2d_recorder = []
        self.add_recorder = []
        self.cm_degree_5_recorder = []
        self.proj_mean_diffs = []
        self.add_dists = []
        self.cm = []
        self.degree = []

    def _one(self, pose_pred, pose_targets, model, K, sym, want):
        o = pose_metrics(np.asarray(pose_pred)[None], np.asarray(pose_targets)[None], model, K, sym=sym, want=want)
        return {k: float(v.item()) for k, v in o.items()}

    def projection_2d(self, pose_pred, pose_targets, model, K, threshold=5):
        d = self._one(pose_pred, pose_targets, model, K, False, ("proj",))["proj"]
        self.proj_mean_diffs.append(d)
        self.projection_2d_recorder.append(d < threshold)

    def projection_2d_sym(self, pose_pred, pose_targets, model, K, threshold=5):
        d = self._one(pose_pred, pose_targets, model, K, True, ("proj",))["proj"]
        self.proj_mean_diffs.append(d)
        self.projection_2d_recorder.append(d < threshold)

    def add_metric(self, pose_pred, pose_targets, model, diameter, percentage=0.1):
        d = self._one(pose_pred, pose_targets, model, None, False, ("add",))["add"]
        self.add_recorder.append(d < diameter * percentage)
        self.add_dists.append(d)

    def add_metric_sym(self, pose_pred, pose_targets, model, diameter, percentage=0.1):
        d = self._one(pose_pred, pose_targets, model, None, True, ("add",))["add"]
        self.add_recorder.append(d < diameter * percentage)
        self.add_dists.append(d)

    def cm_degree_5_metric(self, pose_pred, pose_targets):
        o = self._one(pose_pred, pose_targets, None, None, False, ("cm", "deg"))
        self.cm.append(o["cm"])
        if not np.isnan(o["deg"]):
            self.degree.append(o["deg"])
        self.cm_degree_5_recorder.append(o["cm"] < 5 and o["deg"] < 5)

    def evaluate_batch(self, pose_pred, pose_targets, model, K, diameter, sym=False, threshold=5, percentage=0.1):
        """All three metrics for a block of poses in one launch; recorders extended in pose order."""
        o = {k: v.cpu().numpy() for k, v in pose_metrics(pose_pred, pose_targets, model, K, sym=sym).items()}
        self.proj_mean_diffs += o["proj"].tolist()
        self.projection_2d_recorder += (o["proj"] < threshold).tolist()
        self.add_dists += o["add"].tolist()
        self.add_recorder += (o["add"] < diameter * percentage).tolist()
        self.cm += o["cm"].tolist()
        self.degree += o["deg"][~np.isnan(o["deg"])].tolist()
        self.cm_degree_5_recorder += ((o["cm"] < 5) & (o["deg"] < 5)).tolist()
        return o

    def average_precision(self):
        """means of the recorders, as the reference prints them (evaluation.py:455-470)."""
        return dict(projection_2d=float(np.mean(self.projection_2d_recorder)) if self.projection_2d_recorder else float("nan"),
                    add=float(np.mean(self.add_recorder)) if self.add_recorder else float("nan"),
                    cm_degree_5=float(np.mean(self.cm_degree_5_recorder)) if self.cm_degree_5_recorder else float("nan"))
