"""Camera intrinsics of the reference, as plain constants (no compute: nothing here touches the GPU).

  Camera      /root/reference/utils.py:24-39            SPEED / Tango camera (fx = fy = 0.0176 m, 5.86 um pixels,
                                                         1920 x 1200) -> K in pixels
  INTRINSICS  /root/reference/lib/utils/base_utils.py:240-253  Projector.intrinsic_matrix table
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
