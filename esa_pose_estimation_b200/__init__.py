"""B200-native post-network pose hot path of bonjour-l/esa-pose-estimation.

Heatmap decode (``inference``), RANSAC keypoint voting (``ransac_voting``, ``ransac_voting_gpu``),
EPnP-RANSAC + LM pose solve (``pnp``, ``cpnp``), result sink (``submission``) and the batched /
sharded glue (``pipeline``), behind the reference's own Python call signatures.  All compute is
hand-written CUDA for sm_100a reached through the C ABI in include/esa_pose_b200.h
(``libesa_pose_b200.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
