import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import numpy as np, torch, cv2
from synth import make_pose_case, ESA_K
from esa_pose_estimation_b200 import pnp as P
dev = torch.device("cuda:0")
cases = [make_pose_case(5000 + i, 11 if i % 2 else 24, 0.7, i % 3) for i in range(48)]
nmax = 24
p3 = np.zeros((48, nmax, 3)); p2 = np.zeros((48, nmax, 2)); npts = np.zeros(48, np.int32)
for i, c in enumerate(cases):
    n = len(c["p3d"]); npts[i] = n
    p3[i, :n], p2[i, :n] = c["p3d"], c["p2d"]
rt, mask, status = P.pnp_batch(torch.from_numpy(p3).to(dev), torch.from_numpy(p2).to(dev), torch.from_numpy(ESA_K).to(dev),
                               npts=torch.from_numpy(npts).to(dev), return_status=True)
mask = mask.cpu().numpy(); rt = rt.cpu().numpy()
for i, c in enumerate(cases):
    ok, rv, tv, inl = cv2.solvePnPRansac(c["p3d"][None], c["p2d"][None], ESA_K, np.zeros((8, 1)), reprojectionError=5.0, flags=cv2.SOLVEPNP_EPNP)
    m = 0
    for k in inl.ravel(): m |= 1 << int(k)
    if int(mask[i]) != m:
        # reprojection errors of cv2's pose for the differing points
        R, _ = cv2.Rodrigues(rv)
        pr = (ESA_K @ (R @ c["p3d"].T + tv)).T; pr = pr[:, :2] / pr[:, 2:]
        err = np.linalg.norm(pr - c["p2d"], axis=1)
        print("case", i, "n", npts[i], "outliers", sorted(c["outliers"]), "gpu mask %x cv2 %x diff %x" % (int(mask[i]), m, int(mask[i]) ^ m), "errs", np.round(err, 2))
print("done")
