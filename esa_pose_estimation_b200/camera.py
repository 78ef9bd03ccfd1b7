"""Camera intrinsics of the reference, as plain constants (no compute: nothing here touches the GPU).

  Camera      /root/reference/utils.py:24-39            SPEED / Tango camera (fx = fy = 0.0176 m, 5.86 um pixels,
                                                         1920 x 1200) -> K in pixels
  INTRINSICS  /root/reference/lib/utils/base_utils.py:240-253  Projector.intrinsic_matrix table
  crop_window /root/reference/data_load_val.py:125-176   detector box -> square crop window, padded size and the
                                                         `rate` that val.py:180 un-crops with (integer host arithmetic)
"""
import numpy as np


class Camera:
    """Utility class for accessing camera parameters (same attribute names as the reference)."""
    fx = 0.0176       # focal length [m]
    fy = 0.0176
    nu = 1920         # horizontal pixels
    nv = 1200         # vertical pixels
    ppx = 5.86e-6     # pixel pitch [m / pixel]
    ppy = ppx
    fpx = fx / ppx    # focal length [pixels]
    fpy = fy / ppy
    k = [[fpx, 0, nu / 2],
         [0, fpy, nv / 2],
         [0, 0, 1]]
    K = np.array(k)


INTRINSICS = {
    "linemod": np.array([[572.4114, 0., 325.2611], [0., 573.57043, 242.04899], [0., 0., 1.]]),
    "blender": np.array([[700., 0., 320.], [0., 700., 240.], [0., 0., 1.]]),
    "pascal": np.asarray([[-3000.0, 0.0, 0.0], [0.0, 3000.0, 0.0], [0.0, 0.0, 1.0]]),
    "esa": np.asarray([[3003.41297, 0.0, 960.0], [0.0, 3003.41297, 600.0], [0.0, 0.0, 1.0]]),
}


def crop_window(bbox, img_w=1920, img_h=1200, scale=384, k=1.05):
    """data_load_val.py:125-176 for a batch of detector boxes (x, y, w, h) = (left, top, right, bottom):
    -> (window [B,4] int64 = x_new, y_new, w_new, h_new; size [B] int64 (side of the edge-padded square);
        rate [B] float64 = scale / size, 1.0 where size == scale).
    `window[:, :2]` and `rate` are the `bbox_xy` / `rate` arguments of pipeline.poses_from_heatmaps.
    Python's int() truncates toward zero, hence np.trunc."""
    b = np.asarray(bbox, np.float64).reshape(-1, 4)
    x, y, w, h = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    c0 = np.trunc((x + w) / 2)
    c1 = np.trunc((y + h) / 2)
    half = np.trunc(np.maximum(w - x, h - y) / 2)
    x_new, y_new = np.trunc(c0 - k * half), np.trunc(c1 - k * half)
    w_new, h_new = np.trunc(c0 + k * half), np.trunc(c1 + k * half)
    neg = x_new < 0
    w_new = np.where(neg, w_new - x_new, w_new); x_new = np.where(neg, 0, x_new)
    neg = y_new < 0
    h_new = np.where(neg, h_new - y_new, h_new); y_new = np.where(neg, 0, y_new)
    over = w_new > img_w
    x_new = np.where(over, np.maximum(x_new + img_w - w_new, 0), x_new); w_new = np.where(over, img_w, w_new)
    over = h_new > img_h
    y_new = np.where(over, np.maximum(y_new + img_h - h_new, 0), y_new); h_new = np.where(over, img_h, h_new)
    window = np.stack([x_new, y_new, w_new, h_new], 1).astype(np.int64)
    size = np.maximum(window[:, 2] - window[:, 0], window[:, 3] - window[:, 1])
    rate = np.where(size != scale, scale / size.astype(np.float64), 1.0)
    return window, size, rate
