/* esa_pose_b200.h -- C ABI of the B200-native post-network pose hot path.
 *
 * Drop-in boundary for bonjour-l/esa-pose-estimation.  Every entry
 * point takes plain DEVICE pointers and sizes plus a CUDA stream (cudaStream_t passed as
 * void*), never allocates, never synchronises the host, and returns an int status
 * (EPB_OK == 0) instead of the reference's exit()-on-error (cuda_common.h:17-26).
 * Reference interfaces replaced (paths relative to /root/reference):
 *
 *   epb_decode_heatmaps              inference.py:22 get_max_preds, :136 get_final (+ :75 my_taylor),
 *                                    :171 getPrediction; val.py:151-164 two-stage torch.max
 *   epb_refine_keypoints_dark        inference.py:154 get_final2 (:96 gaussian_blur, :54 taylor)
 *   epb_generate_hypothesis          lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:20
 *                                    (pybind ransac_voting.generate_hypothesis -> kernel .cu:11)
 *   epb_voting_for_hypothesis        ransac_voting.cpp:41 (-> kernel .cu:88)
 *   epb_generate_hypothesis_vanishing_point / epb_voting_for_hypothesis_vanishing_point
 *                                    ransac_voting.cpp:64,85 (-> kernels .cu:170,268)
 *   epb_voting_workspace_bytes, epb_voting_run
 *                                    ransac_voting_gpu.py:514 ransac_voting_layer_v3, :669 _v4,
 *                                    :763 _v5, :10 ransac_voting_layer, :99 _v2 (multi-class),
 *                                    :218 ransac_voting_hypothesis,
 *                                    :263 estimate_voting_distribution, :333 ..._with_mean
 *                                    (whole Python driver, batched, one stream-ordered call)
 *   epb_pnp_epnp_ransac              pnp.py:46 pnp() == cv2.solvePnPRansac(EPNP, 5 px) + Rodrigues
 *   epb_lm_refine                    lib/utils/extend_utils/src/utils_python_binding.h:23-31
 *                                    uncertainty_pnp(); val.py:200-202 cpnp.cpnp / cpnp.cpnp_m
 *   epb_pose_pack                    val.py:203-224 Rodrigues + scipy as_quat -> (w,x,y,z), t
 *   epb_pose_pipeline                val.py:172-228 per-frame glue, batched (select, un-crop,
 *                                    EPnP-RANSAC, LM, quaternion)
 *   epb_esa_score                    demo.py:295-310
 *   epb_pose_metrics                 evaluation.py:340-411 projection_2d[_sym], add_metric[_sym], cm_degree_5_metric
 *   epb_nearest_point_idx            lib/utils/extend_utils/extend_utils.py:40 find_nearest_point_idx
 *                                    (-> src/nearest_neighborhood.cu:48-121)
 *   epb_cov_to_weights               lib/utils/evaluation_utils.py:170-181 cov -> inv(sqrtm(cov)) weights
 *   epb_p3p                          lib/utils/extend_utils/extend_utils.py:83-95 (and :147-156): cv2.solvePnP(P3P)
 *                                    on the four best-weighted correspondences
 */
#ifndef ESA_POSE_B200_H_
#define ESA_POSE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EPB_VERSION 100

enum {
  EPB_OK = 0,
  EPB_ERR_INVALID = 1,    /* bad argument (null pointer, non-positive size, unsupported mode) */
  EPB_ERR_CUDA = 2,       /* a CUDA launch failed; see epb_last_cuda_error() */
  EPB_ERR_WORKSPACE = 3,  /* workspace smaller than epb_*_workspace_bytes() */
  EPB_ERR_NO_DEVICE = 4
};

int epb_version(void);
/* CUDA error code of the last failing launch on this thread (0 if none) and its string. */
int epb_last_cuda_error(void);
const char* epb_last_cuda_error_string(void);
/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
unsigned long long epb_launch_count(void);
/* SM count / clock of the current device (roofline reporting). */
int epb_device_info(int* sm_count, int* sm_clock_khz, size_t* l2_bytes);

/* Optional timing of kernel classes with CUDA events on the launching stream (bench.py roofline).
 * Classes: 0 compaction, 1 hypothesis, 2 vote_count, 3 refine/distribution, 4 pose, 5 decode.
 * epb_profile_enable(1) starts recording; epb_profile_read() synchronises the recorded events and
 * returns their summed duration / count; epb_profile_enable(0) stops and frees them. */
int epb_profile_enable(int on);
int epb_profile_read(int cls, double* total_ms, int* launches);

/* ------------------------------------------------------------------ heatmap decode (a1,a2) */
enum {
  EPB_DECODE_REFINE = 1,        /* log-domain sub-pixel step of inference.py:75-94 */
  EPB_DECODE_ZERO_NONPOS = 2    /* zero the coordinates when maxval <= 0 (inference.py:183-184) */
};
/* hm [n_maps,H,W] f32 contiguous (n_maps = B*K).  xy [n_maps,2] f32 (x = column, y = row),
 * maxval [n_maps] f32, idx [n_maps] i32 (flat row-major argmax; first maximum; NaN counts as
 * the maximum, like torch.max / np.argmax).  Any output pointer may be NULL. */
int epb_decode_heatmaps(const float* hm, int n_maps, int H, int W, int flags, float* xy,
                        float* maxval, int32_t* idx, void* stream);

/* inference.py:136 get_final on caller-supplied peaks: xy [n_maps,2] f32 in/out (the integer part
 * of each coordinate selects the stencil centre, exactly like int(coord[0]) at inference.py:79). */
int epb_refine_keypoints(const float* hm, int n_maps, int H, int W, float* xy, void* stream);

/* inference.py:154 get_final2 (DARK-style decode: 11x11 Gaussian blur of the whole map, renormalised to
 * its maximum, float32 log, full 2x2 Hessian Newton step of inference.py:54 taylor) on caller-supplied
 * peaks: xy [n_maps,2] f32 in/out. */
int epb_refine_keypoints_dark(const float* hm, int n_maps, int H, int W, float* xy, void* stream);

/* ------------------------------------------------- pybind-level voting primitives (a9,a10,a21) */
/* direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hypo [hn,vn,2] f32.
 * Unlike the reference (which returns at::zeros and skips degenerate pairs) the kernel writes
 * the zeros itself, so `hypo` need not be pre-zeroed. */
int epb_generate_hypothesis(const float* direct, const float* coords, const int32_t* idxs,
                            float* hypo, int tn, int vn, int hn, void* stream);
/* Writes 1 where the pixel votes for the hypothesis; other bytes untouched (caller zeroes,
 * ransac_voting_gpu.py:557).  inliers [hn,vn,tn] u8. */
int epb_voting_for_hypothesis(const float* direct, const float* coords, const float* hypo,
                              uint8_t* inliers, int tn, int vn, int hn, float thresh, void* stream);
int epb_generate_hypothesis_vanishing_point(const float* direct, const float* coords,
                                            const int32_t* idxs, float* hypo3, int tn, int vn,
                                            int hn, void* stream);
int epb_voting_for_hypothesis_vanishing_point(const float* direct, const float* coords,
                                              const float* hypo3, uint8_t* inliers, int tn, int vn,
                                              int hn, float thresh, void* stream);

/* ------------------------------------------------------- fused batched voting (a6-a14) */
enum {
  EPB_VOTE_V3 = 0,           /* -> pts */
  EPB_VOTE_V4 = 1,           /* -> pts, var */
  EPB_VOTE_V5 = 2,           /* -> pts, conf (re-vote at 0.999) */
  EPB_VOTE_HYPOTHESIS = 3,   /* -> hyp [B,hn,vn,2], counts [B,hn,vn] */
  EPB_VOTE_DISTRIBUTION = 4, /* -> mean, cov (top-k) */
  EPB_VOTE_DISTRIBUTION_WITH_MEAN = 5,
  EPB_VOTE_V1 = 6,           /* ransac_voting_layer    (:10)  multi-class, winners -> pts [B,classes,vn,2] */
  EPB_VOTE_V2 = 7,           /* ransac_voting_layer_v2 (:99)  multi-class, pinverse refinement x refine_iters */
  EPB_VOTE_MOTION = 8        /* ransac_motion_voting   (:960) mean of (vector + pixel) over the foreground -> pts;
                                min_num = 1, max_num = INT_MAX, no random numbers */
};
enum {
  EPB_MASK_NONZERO = 0, /* v3/v4/v5: mask.byte() != 0 */
  EPB_MASK_EQ1 = 1,     /* hypothesis / distribution: mask == 1 */
  EPB_MASK_CLASS = 2    /* v1/v2: mask == class, class = 1..classes (one virtual image per pair) */
};
enum {
  EPB_RNG_IDXS = 0,   /* hypothesis indices supplied: idxs [B,rounds,hn,vn,2] i32 (values < tn) */
  EPB_RNG_RAW32 = 1,  /* raw 32-bit draws supplied in `idxs`; index = draw % tn on device */
  EPB_RNG_PHILOX = 2  /* Philox4x32-10 in the layout of torch's CUDA random_() / uniform_() */
};

enum {
  EPB_STAGE_ALL = 0,    /* the whole run */
  EPB_STAGE_GATHER = 1, /* only: foreground compaction + gather of the field into the workspace.  This
                           is the only stage that reads `mask` and `vertex`; run it on a separate
                           (high-priority) stream to overlap the PCIe read of a host-resident field
                           of batch chunk i+1 with the voting of chunk i */
  EPB_STAGE_VOTE = 2    /* only: hypotheses, inlier counts, winners / distribution, on a workspace
                           that EPB_STAGE_GATHER filled with the same parameters */
};

typedef struct {
  int mode;             /* EPB_VOTE_* */
  int B, H, W, vn;
  int hn;               /* round_hyp_num */
  int rounds;           /* 1 for v3/v4/v5/hypothesis; ceil(min_hyp_num/round_hyp_num) otherwise */
  float inlier_thresh;
  int min_num, max_num;
  int topk;             /* distribution only */
  int mask_mode;        /* EPB_MASK_* */
  /* vertex element strides (in floats) so that both the physical NCHW output of the network,
   * viewed as [b,h,w,vn,2] by vertex_layer_reshape (base_utils.py:311-316), and a contiguous
   * [b,h,w,vn,2] tensor are consumed without a copy: element (b,y,x,v,c) is at
   * vertex[b*sb + y*sy + x*sx + v*sv + c*sc]. */
  long long sb, sy, sx, sv, sc;
  int rng_mode;         /* EPB_RNG_* */
  unsigned long long philox_seed;
  unsigned long long philox_offset; /* torch generator offset before the call */
  int philox_sm_count;  /* multiProcessorCount torch would see (grid clamp of its RNG kernels) */
  int philox_threads_per_sm; /* maxThreadsPerMultiProcessor */
  int stage;            /* EPB_STAGE_* */
  int classes;          /* v1/v2: class_num - 1 foreground classes (0 or 1 otherwise).  Outputs and the
                           optional idxs / selection / tn_out / status arrays then have B*classes leading
                           entries, ordered (image, class) like the reference's loops */
  int refine_iters;     /* v2: refine_iter_num */
} epb_voting_params;

size_t epb_voting_workspace_bytes(const epb_voting_params* p);

typedef struct {
  const uint8_t* mask;      /* [B,H,W] u8 */
  const float* vertex;      /* strided, see params.  Read exactly once, foreground pixels only, by
                               one gather kernel: it may be DEVICE memory or page-locked HOST memory
                               mapped into the device address space (cudaHostAlloc / cudaHostRegister
                               under UVA), in which case only the foreground part of the field crosses
                               PCIe (zero-copy) and no staging copy of the field is needed */
  const int32_t* idxs;      /* EPB_RNG_IDXS / RAW32: [B,rounds,hn,vn,2]; else NULL */
  const float* selection;   /* optional [B,H,W] f32 uniform draws for the max_num subsample
                               (EPB_RNG_IDXS / RAW32); NULL -> Philox or no subsample possible */
  const float* mean_in;     /* DISTRIBUTION_WITH_MEAN: [B,vn,2] */
  float* pts;               /* [B,vn,2]   v3/v4/v5 */
  float* var_or_conf;       /* [B,vn]     v4 / v5 */
  float* hyp;               /* [B,rounds*hn,vn,2]  HYPOTHESIS (optional otherwise) */
  int32_t* counts;          /* [B,rounds*hn,vn]    HYPOTHESIS (optional otherwise) */
  float* mean;              /* [B,vn,2]   DISTRIBUTION* */
  float* cov;               /* [B,vn,2,2] DISTRIBUTION* */
  int32_t* tn_out;          /* optional [B]: foreground count after the subsample */
  int32_t* status;          /* optional [B,vn]: 0 ok, 1 image skipped (< min_num), 2 singular refine */
  unsigned long long* philox_consumed; /* optional [1]: generator offset increment (EPB_RNG_PHILOX) */
  unsigned long long* philox_state;    /* optional [1], in/out: when given, the generator offset is read
                                          from here instead of params.philox_offset and the offset after
                                          the run is written back, so consecutive runs (batch chunks)
                                          continue one torch generator stream without a host round trip */
} epb_voting_io;

int epb_voting_run(const epb_voting_params* p, const epb_voting_io* io, void* workspace,
                   size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------ pose (a15-a19) */
/* All pose entry points are batched, one warp per image, FP64.
 * p3d [B,n_max,3] (or [n_max,3] if p3d_batched == 0), p2d [B,n_max,2], K [B,9] or [9] row-major
 * (fx=K[0], fy=K[4], cx=K[2], cy=K[5]), npts [B] (NULL -> all n_max).  n_max <= 32 (one lane per
 * correspondence; the reference uses 11..30 keypoints), else EPB_ERR_INVALID. */
enum {
  EPB_POSE_OK = 0,
  EPB_POSE_FAILED = 1,      /* no consensus (cv2 returns ok=False and stale memory; we return NaN) */
  EPB_POSE_TOO_FEW = 2
};
/* pnp.pnp(): rt34 [B,3,4] f64 row-major [R|t]; inlier_mask [B] u64 bitmask (optional);
 * status [B] (optional). */
int epb_pnp_epnp_ransac(const double* p3d, int p3d_batched, const double* p2d, const double* K,
                        int K_batched, const int32_t* npts, int B, int n_max,
                        double reproj_err, int max_iters, double confidence, double* rt34,
                        unsigned long long* inlier_mask, int32_t* status, void* stream);
/* uncertainty_pnp C entry, batched: w2d [B,n_max,3] = (wxx,wxy,wyy); init_rt/result_rt [B,6]
 * (angle-axis, t).  iters/final_cost optional. */
int epb_lm_refine(const double* p2d, const double* p3d, int p3d_batched, const double* w2d,
                  const double* K, int K_batched, const double* init_rt, const int32_t* npts, int B,
                  int n_max, double* result_rt, int32_t* iters, double* final_cost, void* stream);
/* rt6 [B,6] -> pose7 [B,7] f32 (qw,qx,qy,qz,tx,ty,tz) and rt34 [B,3,4] f64 (either optional). */
int epb_pose_pack(const double* rt6, int B, float* pose7, double* rt34, void* stream);
/* R [B,3,3] (inside rt34 [B,3,4]) -> angle-axis like cv2.Rodrigues; rt6 [B,6]. */
int epb_rt34_to_rt6(const double* rt34, int B, double* rt6, void* stream);

/* Voting covariances cov [n,2,2] f32 -> LM weights w2d [n,3] f64 = (wxx,wxy,wyy).
 *   EPB_WEIGHTS_INV_SQRTM    lib/utils/evaluation_utils.py:170-181 (Evaluator.evaluate_uncertainty): inv(sqrtm(cov));
 *                            zeros where cov[0][0] < 1e-6, an entry is NaN, or cov is not positive definite
 *   EPB_WEIGHTS_INV_MAX_EIG  lib/utils/extend_utils/extend_utils.py:133-141 (uncertainty_pnp_v2): wxx = wyy =
 *                            1 / largest eigenvalue, wxy = 0; zeros where cov[0][0] < 1e-5 */
enum { EPB_WEIGHTS_INV_SQRTM = 0, EPB_WEIGHTS_INV_MAX_EIG = 1 };
int epb_cov_to_weights(const float* cov, int n, int mode, double* w2d, void* stream);

/* extend_utils.py:83-95: the pose through three correspondences that reprojects a fourth best
 * (cv2.solvePnP(flags=SOLVEPNP_P3P) on 4 points) -- uncertainty_pnp's LM initialiser, and its whole answer
 * when pn == 4.  p3d [n,3] (or [B,n,3] if p3d_batched), p2d [B,n,2], K [9] (or [B,9]), all f64.
 * w2d == NULL: n must be 4, points used in the given order (0..2 exact, 3 disambiguates).
 * w2d [B,n,3]: the four correspondences with the largest wxx + wxy, in ascending order of that key
 * (np.argsort(w[:,0] + w[:,1])[-4:]).  -> rt34 [B,3,4] f64 (NaN + EPB_POSE_FAILED in status when no
 * candidate has positive depths). */
int epb_p3p(const double* p3d, int p3d_batched, const double* p2d, const double* w2d, const double* K,
            int K_batched, int B, int n, double* rt34, int32_t* status, void* stream);

/* val.py:172-228 batched.  preds [B,K,2] f32 crop px, maxvals [B,K] f32, bbox_xy [B,2] f64,
 * rate [B] f64, p3d_model [K,3] f64 (shared) , Kmat [9] f64.
 * Selects large_k = max(#(maxval > sel_thresh), min_k) best keypoints (ties: lower index),
 * un-crops, EPnP-RANSAC, LM -> pose7 [B,7] f32, rt6 [B,6].  weighted: 0 unit weights (cpnp.cpnp), 1 wxx = wyy =
 * maxval (cpnp.cpnp_m as ASSUMED in SURVEY.md 8c -- the cpnp binary and source are absent from the reference),
 * 2 wxx = wyy = sqrt(maxval) (the other plausible reading, for users who know their cpnp build). */
int epb_pose_pipeline(const float* preds, const float* maxvals, const double* bbox_xy,
                      const double* rate, const double* p3d_model, const double* Kmat, int B, int K,
                      int min_k, double sel_thresh, int weighted, float* pose7, double* rt6,
                      double* epnp_rt34, int32_t* status, void* stream);

/* demo.py:295-310: pose7 pred/gt [B,7] f32 -> score_t [B], score_r [B] f64. */
int epb_esa_score(const float* pose7_pred, const float* pose7_gt, int B, double* score_t,
                  double* score_r, void* stream);

/* ------------------------------------------------------- LINEMOD-style metrics (8f.3) */
/* evaluation.py:340-411 (projection_2d[_sym], add_metric[_sym], cm_degree_5_metric), batched: pred / gt poses
 * [N,3,4] f64 row-major [R|t], model [n_model,3] f64, K [9] f64 (needed for proj2d).  symmetric != 0 selects the
 * nearest-point variants (ADD-S, projection_2d_sym: float32 search like nearest_neighborhood.cu, FP64 distances).
 * Outputs [N] f64, any may be NULL: proj2d = mean 2-D reprojection distance (px), add = mean 3-D distance,
 * cm = translation error * 100, deg = rotation error in degrees (NaN where trace < -1, like np.arccos).
 * The pass flags of the reference (proj2d < threshold, add < diameter * percentage, cm < 5 and deg < 5) are
 * comparisons on these values; the Python mirror applies them. */
int epb_pose_metrics(const double* pred_rt34, const double* gt_rt34, int N, const double* model, int n_model,
                     const double* K, int symmetric, double* proj2d, double* add, double* cm, double* deg,
                     void* stream);
/* lib/utils/extend_utils/extend_utils.py:40-61 find_nearest_point_idx (-> nearest_neighborhood.cu:48-121) on
 * DEVICE buffers: ref_pts [pn1,dim], que_pts [pn2,dim] f32 (dim 2 or 3) -> idxs [pn2] i32, index of the nearest
 * reference point (first minimum). */
int epb_nearest_point_idx(const float* ref_pts, const float* que_pts, int32_t* idxs, int pn1, int pn2, int dim,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ESA_POSE_B200_H_ */
