"""Deterministic synthetic SPEED-shaped inputs shared by the tests, the golden generator and
bench.py: Gaussian / random-init heatmaps, PVNet vector fields + masks,
structured PnP cases with the ESA camera.  numpy only."""
import numpy as np

# ESA/SPEED intrinsics: /root/reference/lib/utils/base_utils.py:250-252, utils.py:28-39
ESA_K = np.array([[0.0176 / 5.86e-6, 0, 960.0], [0, 0.0176 / 5.86e-6, 600.0], [0, 0, 1.0]])


def rodrigues(rvec):
    rvec = np.asarray(rvec, np.float64)
    th = np.linalg.norm(rvec)
    if th < 1e-16:
        return np.eye(3)
    r = rvec / th
    kx = np.array([[0, -r[2], r[1]], [r[2], 0, -r[0]], [-r[1], r[0], 0]])
    return np.cos(th) * np.eye(3) + (1 - np.cos(th)) * np.outer(r, r) + np.sin(th) * kx


def tango_model(n=11, seed=7):
    """The 11-point Tango model is not in the reference repo: synthesised in a +-0.4 m box."""
    rng = np.random.default_rng(seed)
    return rng.uniform(-0.4, 0.4, (n, 3))


def random_pose(rng):
    rv = rng.normal(size=3)
    rv *= rng.uniform(0.05, 3.0) / np.linalg.norm(rv)
    t = np.array([rng.uniform(-0.2, 0.2), rng.uniform(-0.2, 0.2), rng.uniform(3.0, 40.0)])
    return rv, t


def project(p3d, rvec, t, K=ESA_K):
    pc = p3d @ rodrigues(rvec).T + t
    return np.stack([K[0, 0] * pc[:, 0] / pc[:, 2] + K[0, 2], K[1, 1] * pc[:, 1] / pc[:, 2] + K[1, 2]], 1)


def make_pose_case(seed, n=11, noise=0.5, n_outliers=0, model=None):
    rng = np.random.default_rng(seed)
    p3d = tango_model(n, seed=seed + 1000) if model is None else np.asarray(model)
    rvec, t = random_pose(rng)
    p2d = project(p3d, rvec, t) + rng.normal(0, 1.0, (n, 2)) * noise
    out_idx = rng.choice(n, n_outliers, replace=False) if n_outliers else np.zeros(0, int)
    if n_outliers:
        p2d[out_idx] += rng.choice([-1, 1], (n_outliers, 2)) * rng.uniform(40, 200, (n_outliers, 2))
    return dict(p3d=p3d, p2d=p2d, rvec=rvec, t=t, outliers=out_idx)


def make_heatmaps(seed, b, k, h, w, kind="gauss", sigma=2.0):
    """-> (hm [b,k,h,w] f32, centres [b,k,2] (x,y))."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    hm = np.zeros((b, k, h, w), np.float32)
    cen = np.zeros((b, k, 2))
    for bi in range(b):
        for ki in range(k):
            if kind == "randinit":
                hm[bi, ki] = rng.normal(0, 0.05, (h, w)).astype(np.float32)
                continue
            if kind == "edges" and ki % 3 == 0:
                cx, cy = rng.choice([0.3, 1.2, w - 1.6, w - 2.4]), rng.uniform(0, h - 1)
            elif kind == "edges" and ki % 3 == 1:
                cx, cy = rng.uniform(0, w - 1), rng.choice([0.0, 1.7, h - 1.0, h - 2.2])
            else:
                cx, cy = rng.uniform(3, w - 4), rng.uniform(3, h - 4)
            cen[bi, ki] = (cx, cy)
            g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * sigma ** 2)) * rng.uniform(0.5, 1.0)
            hm[bi, ki] = (g + rng.normal(0, 0.01, (h, w))).astype(np.float32)
    if kind == "edges":
        # exact ties (first maximum must win), an all-negative map and a constant map
        hm[0, -1] = -np.abs(hm[0, -1]) - 0.1
        if k > 2:
            m = hm[0, 2].max()
            hm[0, 2, h // 2, w // 3] = m
            hm[0, 2, h // 2, w // 3 + 5] = m
            hm[0, 2, h // 3, w // 2] = m
        if k > 4:
            hm[0, 4] = 0.25
    return hm, cen


def ellipse_mask(h, w, frac, rng=None, jitter=0.0):
    """Filled ellipse centred in the crop with foreground fraction ~frac (1.0 -> everything)."""
    if frac >= 1.0:
        return np.ones((h, w), np.uint8)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    cy, cx = (h - 1) / 2.0, (w - 1) / 2.0
    if rng is not None and jitter:
        cy += rng.uniform(-jitter, jitter) * h
        cx += rng.uniform(-jitter, jitter) * w
    # area = pi a b = frac h w, with a/b = w/h
    a = np.sqrt(frac * w * w / np.pi)
    bb = np.sqrt(frac * h * h / np.pi)
    return ((((xs - cx) / a) ** 2 + ((ys - cy) / bb) ** 2) <= 1.0).astype(np.uint8)


def make_vertex_field(seed, b, h, w, vn, fg_frac=0.25, kind="structured", noise_deg=2.0, spread=0.35):
    """PVNet-style network output: mask [b,h,w] u8, vertex NCHW [b,2vn,h,w] f32 (x,y interleaved per
    keypoint: channel 2v = dx, 2v+1 = dy), keypoints [b,vn,2].
    structured: unit vectors towards the keypoint (linemod_dataset.py:69-82) + angular noise;
    randinit:   N(0, 0.05^2) per element, like a random-init head."""
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    mask = np.zeros((b, h, w), np.uint8)
    vertex = np.zeros((b, 2 * vn, h, w), np.float32)
    kpts = np.zeros((b, vn, 2))
    for bi in range(b):
        mask[bi] = ellipse_mask(h, w, fg_frac, rng, jitter=0.05)
        if kind == "randinit":
            vertex[bi] = rng.normal(0, 0.05, (2 * vn, h, w)).astype(np.float32)
            continue
        for v in range(vn):
            kx = w / 2 + rng.uniform(-spread, spread) * w
            ky = h / 2 + rng.uniform(-spread, spread) * h
            kpts[bi, v] = (kx, ky)
            dx, dy = kx - xs, ky - ys
            ang = np.arctan2(dy, dx) + np.deg2rad(noise_deg) * rng.normal(size=(h, w))
            vertex[bi, 2 * v] = np.cos(ang)
            vertex[bi, 2 * v + 1] = np.sin(ang)
    return mask, vertex, kpts


def vertex_hwvn2(vertex_nchw):
    """numpy form of vertex_layer_reshape (base_utils.py:311-316): [b,2vn,h,w] -> [b,h,w,vn,2]."""
    b, c, h, w = vertex_nchw.shape
    return np.ascontiguousarray(vertex_nchw.transpose(0, 2, 3, 1)).reshape(b, h, w, c // 2, 2)


def make_pose_field(seed, n_images, size=256, vn=11, fg=0.25, noise_deg=2.0, model=None, mask_seed=None):
    """Structured SPEED-like crops whose keypoints are CONSISTENT with a 3-D model (so that the pose solve
    behind the voting has a consensus): the vn keypoints are the projections of the Tango model under a random
    pose, scaled into the crop; field = unit vectors towards them + angular noise; mask = centred ellipse.
    -> mask [n,s,s] u8, vertex NCHW [n,2vn,s,s] f32, model [vn,3], geom [n,3] = (bbox x, bbox y, rate) for the
    un-crop of val.py:180 (ori = pred / rate + (x, y)), kcrop [n,vn,2] planted keypoints in crop pixels, and the
    planted poses (rvec [n,3], t [n,3])."""
    rng = np.random.default_rng(seed if mask_seed is None else mask_seed)
    s = size
    model = tango_model(vn, seed=9) if model is None else np.asarray(model)
    ys, xs = np.mgrid[0:s, 0:s].astype(np.float32)
    mask = np.zeros((n_images, s, s), np.uint8)
    vertex = np.zeros((n_images, 2 * vn, s, s), np.float32)
    kcrop = np.zeros((n_images, vn, 2))
    geom = np.zeros((n_images, 3))
    rvecs, ts = np.zeros((n_images, 3)), np.zeros((n_images, 3))
    for i in range(n_images):
        c = make_pose_case(seed * 100003 + i, vn, 0.0, 0, model=model)
        lo, hi = c["p2d"].min(0), c["p2d"].max(0)
        width = (hi - lo).max() * 1.6 + 8
        org = (lo + hi) / 2 - width / 2
        rate = s / width
        kc = (c["p2d"] - org) * rate
        kcrop[i] = kc
        geom[i] = (org[0], org[1], rate)
        rvecs[i], ts[i] = c["rvec"], c["t"]
        mask[i] = ellipse_mask(s, s, fg, rng, jitter=0.03)
        for v in range(vn):
            ang = np.arctan2(kc[v, 1] - ys, kc[v, 0] - xs) + np.float32(np.deg2rad(noise_deg)) * rng.standard_normal((s, s), dtype=np.float32)
            vertex[i, 2 * v] = np.cos(ang)
            vertex[i, 2 * v + 1] = np.sin(ang)
    return mask, vertex, model, geom, kcrop, rvecs, ts
