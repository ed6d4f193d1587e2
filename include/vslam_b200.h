/*
 * vslam_b200.h -- C-ABI of the B200-native stereo-VO + sliding-window-BA hot path.
 *
 * Drop-in boundary for shangzhouye/stereo-visual-slam.  The reference has no FFI or
 * plugin registry: its boundary is the C++ signatures of visual_odometry.hpp / optimization.hpp.
 * The C++ host layer in stereo-visual-slam_b200/host/ keeps those signatures and calls the entry
 * points below; every entry point cites the reference code it replaces.  Plain pointers and sizes
 * only -- no torch, OpenCV, Eigen or Sophus types cross this line.
 *
 * Conventions
 *   - every call returns an int status (VSLAM_OK == 0, negative = error) and never throws;
 *   - `*_dev` entry points take DEVICE pointers, enqueue on the context's stream and return without
 *     synchronising; all other entry points take HOST pointers and are synchronous on return;
 *   - one context = one device + one stream; a context is not thread-safe;
 *   - there is no CPU fallback: without a usable sm_100-class device vslam_ctx_create fails.
 */
#ifndef VSLAM_B200_H
#define VSLAM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VSLAM_ABI_VERSION 1

/* status codes (reference convention: 0 ok / -1 image missing, visual_odometry.cpp:53-57,73-77) */
enum {
    VSLAM_OK = 0,
    VSLAM_E_INVALID = -1,   /* null pointer / bad size (the reference's "-1: could not open or find the image") */
    VSLAM_E_CAPACITY = -2,  /* input exceeds the capacity the context was created with */
    VSLAM_E_CUDA = -3,      /* CUDA runtime error; see vslam_last_error() */
    VSLAM_E_NODEVICE = -4,  /* no CUDA device / not an sm_100-class device */
    VSLAM_E_NUMERIC = -5,   /* linear solve failed on every LM trial etc. */
    VSLAM_E_OVERFLOW = -6   /* a device-side work list overflowed (raise the capacity) */
};

typedef struct vslam_ctx vslam_ctx;

/* POD mirror of cv::KeyPoint (28 bytes: pt.x, pt.y, size, angle, response, octave, class_id). */
typedef struct vslam_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} vslam_keypoint;

/* POD mirror of cv::DMatch (16 bytes). distance holds an integer 0..256 as float. */
typedef struct vslam_dmatch {
    int32_t queryIdx;
    int32_t trainIdx;
    int32_t imgIdx;
    float distance;
} vslam_dmatch;

/* Capacities fixed at context creation; all device memory is owned by the context. */
typedef struct vslam_config {
    int32_t device;        /* CUDA device ordinal */
    int32_t max_images;    /* images per batch (a stereo pair is 2 images) */
    int32_t max_width;     /* level-0 width  */
    int32_t max_height;    /* level-0 height */
    int32_t max_keypoints; /* per image (nfeatures upper bound) */
    int32_t max_ba_poses;
    int32_t max_ba_points;
    int32_t max_ba_obs;
} vslam_config;

int vslam_abi_version(void);
const char* vslam_status_string(int status);
/* last CUDA error text seen by this context (empty string if none) */
const char* vslam_last_error(const vslam_ctx* ctx);

int vslam_ctx_create(const vslam_config* cfg, vslam_ctx** out);
void vslam_ctx_destroy(vslam_ctx* ctx);

/* Pinned (page-locked) host memory from a process-wide pool, for the image-sized buffers the reference keeps in
 * cv::Mat -- the frames cv::imread returns (visual_odometry.cpp:42-51) and the disparity image (:163-168): uploads and
 * downloads of such buffers are direct DMAs.  Freed blocks are recycled (a fresh cudaHostAlloc per frame would cost
 * more than it saves).  vslam_host_alloc returns NULL when no CUDA device can pin memory; vslam_host_free ignores
 * pointers it did not hand out. */
void* vslam_host_alloc(size_t bytes);
void vslam_host_free(void* p);
/* run on a caller-owned stream (a cudaStream_t, e.g. torch's current stream); NULL = context's own */
int vslam_ctx_set_stream(vslam_ctx* ctx, void* cuda_stream);
int vslam_ctx_synchronize(vslam_ctx* ctx);
/* Batched entry points cut large batches into chunks that alternate between the context stream and an internal second
 * stream (disjoint scratch, joined back into the context stream before the call returns / the stream continues), so
 * that one chunk's launch tails overlap the other chunk's kernels.  on = 0 keeps every launch on the context stream,
 * e.g. to time kernels in isolation.  Default: on. */
int vslam_ctx_set_concurrency(vslam_ctx* ctx, int on);
/* number of kernels this context has launched since creation (for bench.py's gpu_launches) */
int64_t vslam_ctx_launch_count(const vslam_ctx* ctx);

/* Optional per-launch timing: when enabled every kernel launch of this context is bracketed by CUDA
 * events on the launching stream; vslam_ctx_timing_read synchronises and returns the accumulated device
 * time and launch count of kernel `id` (0 <= id < vslam_kernel_count()) since the last enable. */
int vslam_kernel_count(void);
const char* vslam_kernel_name(int id);
int vslam_ctx_timing_enable(vslam_ctx* ctx, int on);
int vslam_ctx_timing_read(vslam_ctx* ctx, int id, double* total_ms, int64_t* launches);

/* ------------------------------------------------------------------------------------------------
 * K10  brute-force Hamming matching + mutual cross-check + distance gate.
 * Replaces cv::BFMatcher(NORM_HAMMING, crossCheck=true)::match and the gate of VO::feature_matching
 * (visual_odometry.cpp:24,33 and :219-251).  query = descriptors_1, train = descriptors_2, rows are
 * 32-byte ORB descriptors.  Output is ordered by ascending queryIdx, imgIdx = 0.
 *   cross_check != 0 : strict mutual nearest neighbour, first minimum wins ties (cv2 4.13 semantics)
 *   gate_rel >= 0    : keep distance <= max(gate_rel * min_distance, gate_abs)
 *                      (reference: gate_rel = 2.0, gate_abs = 30.0 * frame_gap); gate_rel < 0 = no gate
 * nq == 0 or nt == 0 yields *n_out = 0 (the reference dereferences an empty range there: UB).
 * nq, nt <= 65535.  `out` must hold min(nq, nt) entries when cross_check, nq otherwise.
 * ---------------------------------------------------------------------------------------------- */
int vslam_match_hamming(vslam_ctx* ctx, const uint8_t* query, int nq, const uint8_t* train, int nt,
                        int cross_check, double gate_rel, double gate_abs,
                        vslam_dmatch* out, int* n_out);

/* Device-resident batched form.  Pair b matches rows d_query + b*q_stride_rows*32 (d_nq[b] rows) against
 * d_train + b*t_stride_rows*32 (d_nt[b] rows); counts are read on the device (no host sync).  Matches
 * of pair b are written to d_out + b*out_stride, their number to d_n_out[b]. */
int vslam_match_hamming_batch_dev(vslam_ctx* ctx, const uint8_t* d_query, const int32_t* d_nq, int q_stride_rows,
                                  const uint8_t* d_train, const int32_t* d_nt, int t_stride_rows,
                                  int batch, int max_rows, int cross_check, double gate_rel, double gate_abs,
                                  vslam_dmatch* d_out, int out_stride, int32_t* d_n_out);

/* ------------------------------------------------------------------------------------------------
 * K1-K9  ORB feature detection: pyramid + FAST-9 + NMS + Harris ranking + ANMS + orientation + rBRIEF.
 * Replaces detector_->detect / adaptive_non_maximal_suppresion / descriptor_->compute inside
 * VO::feature_detection (visual_odometry.cpp:70-94; ORB::create(3000) at :22,:31; ANMS :96-157).
 * Bit-exact against cv2 4.13.0 cv::ORB (8 levels, scale 1.2, edge 31, HARRIS_SCORE, patch 31, FAST 20).
 *   nfeatures  : cv::ORB nfeatures (reference: 3000)
 *   anms_keep  : ANMS target (reference: 500); <= 0 disables ANMS.  As in the reference, ANMS is a
 *                no-op when fewer than anms_keep keypoints were detected, and keeps radius ties.
 *   anms_c     : robustness coefficient (reference: 1.11f)
 * Output order is canonical: octave ascending, response descending, then y, x (what cv::ORB::compute's
 * stable by-octave regrouping yields from the response-sorted ANMS list).  Because cv::ORB keeps
 * response ties, up to a few more than nfeatures keypoints can be returned; image i's keypoints
 * start at kp_out + i*cap and desc_out + i*cap*32 where cap = vslam_orb_keypoint_capacity(ctx).
 * Returns VSLAM_E_INVALID for a null image (the reference's -1, visual_odometry.cpp:73-77).
 * ---------------------------------------------------------------------------------------------- */
int vslam_orb_keypoint_capacity(const vslam_ctx* ctx);
int vslam_orb_detect_compute(vslam_ctx* ctx, const uint8_t* image, int width, int height, int row_pitch,
                             int nfeatures, int anms_keep, float anms_c,
                             vslam_keypoint* kp_out, uint8_t* desc_out, int32_t* n_out);
int vslam_orb_detect_compute_batch(vslam_ctx* ctx, const uint8_t* images, int n_images, int width, int height,
                                   int row_pitch, long long image_stride, int nfeatures, int anms_keep,
                                   float anms_c, vslam_keypoint* kp_out, uint8_t* desc_out, int32_t* n_out);
/* device-resident form: images, outputs and counts stay in HBM; asynchronous on the context stream */
int vslam_orb_detect_compute_batch_dev(vslam_ctx* ctx, const uint8_t* d_images, int n_images, int width,
                                       int height, int row_pitch, long long image_stride, int nfeatures,
                                       int anms_keep, float anms_c, vslam_keypoint* d_kp, uint8_t* d_desc,
                                       int32_t* d_n);
/* synchronise and report VSLAM_E_OVERFLOW if any of the last n_images raised a work-list overflow */
int vslam_orb_last_flags(vslam_ctx* ctx, int n_images);
/* test tap: intermediate state of image `img` after the last ORB call (any output pointer may be NULL).
 * cand_xy_score receives pairs {x | y << 16, FAST score} in arbitrary order. */
int vslam_orb_debug_read(vslam_ctx* ctx, int img, int level, uint8_t* level_pixels, uint8_t* blurred_pixels,
                         int* w_out, int* h_out, uint32_t* cand_xy_score, int cand_cap, int* n_cand);

/* ------------------------------------------------------------------------------------------------
 * K11  per-match triangulation (DLT) with the reference's depth gates.
 * Takes the role of VO::disparity_map + Frame::find_3d + VO::set_ref_3d_position
 * (visual_odometry.cpp:159-217, types_def.cpp:9-18) in sparse-stereo form; arithmetic oracle is
 * cv::triangulatePoints.  xl / xr are n x 2 float32 pixel coordinates, P1 / P2 row-major 3x4
 * projection matrices, T_c_w a row-major 3x4 [R|t] (NULL = identity).  xyz_world (n x 3 float32) is
 * T_c_w^-1 * p for EVERY input (no compaction); flags bit0 = usable (10 < Z < 400, visual_odometry.cpp:194),
 * bit1 = reliable depth (Z < 40, visual_odometry.cpp:201), Z measured in the left camera frame.
 * ---------------------------------------------------------------------------------------------- */
int vslam_triangulate(vslam_ctx* ctx, const float* xl, const float* xr, int n, const double* P1, const double* P2,
                      const double* T_c_w, float* xyz_world, uint8_t* flags);
/* device form over match lists: pair b triangulates d_matches[b*match_stride + i], i < d_n_matches[b], between
 * d_kp_left[b*kp_stride + queryIdx] and d_kp_right[b*kp_stride + trainIdx]; P1/P2 are HOST pointers (passed by
 * value to the kernel), d_T_c_w is batch x 12 doubles on the device or NULL. */
int vslam_triangulate_matches_batch_dev(vslam_ctx* ctx, const vslam_keypoint* d_kp_left,
                                        const vslam_keypoint* d_kp_right, int kp_stride,
                                        const vslam_dmatch* d_matches, const int32_t* d_n_matches, int match_stride,
                                        int batch, const double* P1, const double* P2, const double* d_T_c_w,
                                        float* d_xyz, uint8_t* d_flags);

/* ------------------------------------------------------------------------------------------------
 * Batched stereo frontend = the north star's "detect + match + triangulate" on n_pairs stereo pairs:
 * vslam_orb_detect_compute on every left and right image, VO::feature_matching (left = query,
 * right = train; gate_rel / gate_abs as in vslam_match_hamming) and vslam_triangulate per match.
 * Strides: cap = vslam_orb_keypoint_capacity(ctx).  kp [2*n_pairs][cap] and desc [2*n_pairs][cap][32]
 * hold the left images first, then the right images; n_kp [2*n_pairs]; matches [n_pairs][cap];
 * n_matches [n_pairs]; xyz [n_pairs][cap][3]; flags [n_pairs][cap].
 * The _dev form takes device pointers (P1/P2 stay host pointers) and does not synchronise.
 * ---------------------------------------------------------------------------------------------- */
int vslam_stereo_frontend_batch(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs, int width,
                                int height, int row_pitch, long long image_stride, int nfeatures, int anms_keep,
                                float anms_c, double gate_rel, double gate_abs, const double* P1, const double* P2,
                                const double* T_c_w, vslam_keypoint* kp, uint8_t* desc, int32_t* n_kp,
                                vslam_dmatch* matches, int32_t* n_matches, float* xyz, uint8_t* flags);
/* The host-buffer call in two halves: _begin enqueues uploads, kernels and downloads of the batch and returns, _end
 * waits and returns VSLAM_OK / VSLAM_E_OVERFLOW.  One batch per context at a time; inputs and outputs must stay valid
 * (and be pinned for the copies to be asynchronous) until _end.  Two contexts alternating keep two batches in flight
 * -- the streaming form of the reference's frame loop (run_vslam.cpp:44-75 reads, processes and stores one frame after
 * the other). */
int vslam_stereo_frontend_batch_begin(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs, int width,
                                      int height, int row_pitch, long long image_stride, int nfeatures, int anms_keep,
                                      float anms_c, double gate_rel, double gate_abs, const double* P1, const double* P2,
                                      const double* T_c_w, vslam_keypoint* kp, uint8_t* desc, int32_t* n_kp,
                                      vslam_dmatch* matches, int32_t* n_matches, float* xyz, uint8_t* flags);
int vslam_stereo_frontend_batch_end(vslam_ctx* ctx);
int vslam_stereo_frontend_batch_dev(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int n_pairs,
                                    int width, int height, int row_pitch, long long image_stride, int nfeatures,
                                    int anms_keep, float anms_c, double gate_rel, double gate_abs, const double* P1,
                                    const double* P2, const double* d_T_c_w, vslam_keypoint* d_kp, uint8_t* d_desc,
                                    int32_t* d_n_kp, vslam_dmatch* d_matches, int32_t* d_n_matches, float* d_xyz,
                                    uint8_t* d_flags);

/* ------------------------------------------------------------------------------------------------
 * K13-K16  sliding-window bundle adjustment: Levenberg-Marquardt with Schur complement and Huber kernel.
 * Replaces the arithmetic of optimize_map (optimization.cpp:103-288) and optimize_pose_only
 * (optimization.cpp:290-436): g2o's OptimizationAlgorithmLevenberg + BlockSolver<6,3> driving the
 * reference's VertexPose / VertexXYZ / EdgeProjection / PoseOnlyEdgeProjection (optimization.cpp:26-101),
 * followed by the adaptive chi2 relabel loop (optimization.cpp:224-266 / 382-424).
 *   poses        n_poses x 12 doubles, row-major 3x4 [R|t] of T_c_w, updated in place (no vertex is fixed)
 *   points       n_points x 3 doubles (Landmark::pt_3d_ widened), updated in place unless pose_only
 *   obs_*        one entry per graph edge in INSERTION order: pose index, point index, measured pixel
 *   Kmat         row-major 3x3 intrinsics
 *   chi2_per_obs (optional, n_obs) edge->chi2() as the relabel loop sees it
 *   point_inlier (optional, n_points, in/out) Landmark::is_inlier after relabelling; the LAST edge of a
 *                landmark in insertion order decides (the reference iterates a pointer-keyed std::map,
 *                optimization.cpp:156,254-266); points without edges keep the value passed in
 * The caller applies if_update_map / if_update_landmark (optimization.cpp:272-287) by choosing what to
 * copy back.  The selection of edges (is_inlier / reliable_depth_, optimization.cpp:160,334) is the
 * caller's graph-building job, as in the reference.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vslam_ba_options {
    double huber_delta;    /* 5.991 (optimization.cpp:154,205) */
    double chi2_threshold; /* 5.991 initial relabel threshold */
    int32_t num_iterations; /* optimizer.optimize(num_ite) */
    int32_t pose_only;     /* 0 = optimize_map, 1 = optimize_pose_only */
    int32_t max_trials;    /* g2o maxTrialsAfterFailure, 10 */
    int32_t reserved;
    double tau;            /* g2o initial-lambda factor, 1e-5 */
} vslam_ba_options;

typedef struct vslam_ba_result {
    int32_t iterations;    /* outer LM iterations executed */
    int32_t trials;        /* LM trials = linear solves */
    int32_t accepted;
    int32_t reserved;
    double chi2_initial, chi2_final; /* robustified */
    double lambda_final;
    double chi2_threshold; /* after the adaptive doubling */
    int32_t n_inlier_obs, n_outlier_obs;
} vslam_ba_result;

/* device-side phase profile of the last vslam_ba_optimize call: nanoseconds (GPU globaltimer) spent in
 * [zero, build, schur_init, schur, schur_reduce, solve, update, trial_err] incl. their grid-wide barriers */
int vslam_ba_last_phase_ns(vslam_ctx* ctx, uint64_t* ns8);

int vslam_ba_optimize(vslam_ctx* ctx, int n_poses, double* poses, int n_points, double* points, int n_obs,
                      const int32_t* obs_pose, const int32_t* obs_point, const double* obs_uv, const double* Kmat,
                      const vslam_ba_options* opt, vslam_ba_result* res, double* chi2_per_obs,
                      uint8_t* point_inlier);

/* ------------------------------------------------------------------------------------------------
 * K12  PnP-RANSAC + inlier refit.
 * Replaces cv::solvePnPRansac(pts3d, pts2d, K, noDist, rvec, tvec, false, 100, 4.0, 0.99, inliers) in
 * VO::motion_estimation (visual_odometry.cpp:253-314, call at :277).  xyz: n x 3 float32 world points
 * (Landmark::pt_3d_), uv: n x 2 float32 pixels, Kmat row-major 3x3 (no distortion).  Outputs: rvec
 * (Rodrigues) and tvec as cv::solvePnPRansac returns them, optionally the same pose as a 3x4 [R|t]
 * (T_c_w, may be NULL), and the ascending inlier index list.
 * The RANSAC is OpenCV 4.13's, sample for sample: cv::RNG((uint64)-1) sample stream with duplicate rejection,
 * 5-point EPnP minimal solver in OpenCV's exact fp64 arithmetic, float32 reprojection errors against
 * (float)(reproj_err^2) without a cheirality test, "goodCount > max(best, 4)" update rule and the
 * confidence-driven shrinking of the iteration budget (RANSACUpdateNumIters) -- so the inlier list equals
 * cv2's index for index on any input.  The returned pose is the least-squares optimum of the reprojection
 * error over those inliers (OpenCV: solvePnP(SOLVEPNP_ITERATIVE) on them; here Gauss-Newton to convergence).
 *   n == 5      OpenCV skips RANSAC: EPnP pose of the five points, all five reported as inliers
 *   n  < 5      *n_inliers = 0 (OpenCV asserts n >= 4 and uses P3P for n == 4; the reference rejects
 *               every frame with fewer than 10 inliers, visual_odometry.cpp:316-346)
 *   iters       <= 512 (VSLAM_E_CAPACITY above); 0 < confidence < 1 (VSLAM_E_INVALID otherwise)
 * `inliers` must hold n entries; the first *n_inliers are the ascending inlier indices, the rest is unspecified.
 * ---------------------------------------------------------------------------------------------- */
int vslam_pnp_ransac(vslam_ctx* ctx, const float* xyz, const float* uv, int n, const double* Kmat, int iters,
                     float reproj_err, double confidence, double* rvec, double* tvec, double* T_c_w,
                     int32_t* inliers, int32_t* n_inliers);
/* test tap: the model (row-major R then t, 12 doubles) and the inlier count of RANSAC sample `it` of the last
 * vslam_pnp_ransac call, and the number of iterations OpenCV's loop executes before its budget runs out */
int vslam_pnp_debug_read(vslam_ctx* ctx, int it, double* R_t12, int32_t* count, int32_t* executed);

/* K7 stand-alone: VO::adaptive_non_maximal_suppresion (visual_odometry.cpp:96-157) on caller keypoints.
 * keep_idx receives the indices (ascending, into `keypoints`) of the survivors: radius >= num-th largest
 * radius, ties kept; identity when n < num (visual_odometry.cpp:100). */
int vslam_anms(vslam_ctx* ctx, const vslam_keypoint* keypoints, int n, int num, float c_robust,
               int32_t* keep_idx, int32_t* n_keep);

/* ------------------------------------------------------------------------------------------------
 * K18-K23  dense semi-global stereo matching: the reference's own depth source.
 * Replaces cv::StereoSGBM::create(0, 96, 9, 8*9*9, 32*9*9, 1, 63, 10, 100, 32)->compute(left, right, disp)
 * and the convertTo(CV_32F, 1/16) that follows it in VO::disparity_map (visual_odometry.cpp:159-174).
 * Bit-exact against cv2 4.13.0 StereoSGBM (MODE_SGBM): Birchfield-Tomasi cost on the x-Sobel and raw planes,
 * 9x9 block sum, five aggregation paths, uniqueness / LR check / sub-pixel fit, medianBlur(3), filterSpeckles.
 *   params   NULL = the reference's constants (vslam_sgbm_default_params).  The kernels are specialised for
 *            min_disparity 0, num_disparities 96, block_size 9 (anything else: VSLAM_E_INVALID); P1, P2,
 *            disp12_max_diff, pre_filter_cap, uniqueness_ratio, speckle_* are free.
 *   disp16   n_pairs x height x width int16, disparity * 16, -16 = invalid (CV_16S as cv::StereoSGBM writes it)
 *   disp_f32 n_pairs x height x width float, disparity in pixels, -1 = invalid (the reference's Frame::disparity_)
 *            either output may be NULL.
 * width - 96 must exceed 4 (OpenCV raises for narrower images); returns VSLAM_E_INVALID otherwise.
 * ---------------------------------------------------------------------------------------------- */
typedef struct vslam_sgbm_params {
    int32_t min_disparity, num_disparities, block_size, P1, P2, disp12_max_diff, pre_filter_cap, uniqueness_ratio,
        speckle_window_size, speckle_range;
} vslam_sgbm_params;
void vslam_sgbm_default_params(vslam_sgbm_params* p);
int vslam_sgbm_compute(vslam_ctx* ctx, const uint8_t* left, const uint8_t* right, int n_pairs, int width, int height,
                       int row_pitch, long long image_stride, const vslam_sgbm_params* params, int16_t* disp16,
                       float* disp_f32);
/* device-resident form: asynchronous on the context stream */
int vslam_sgbm_compute_dev(vslam_ctx* ctx, const uint8_t* d_left, const uint8_t* d_right, int n_pairs, int width,
                           int height, int row_pitch, long long image_stride, const vslam_sgbm_params* params,
                           int16_t* d_disp16, float* d_disp_f32);
/* test taps: stage 0 = cost volume C [h][w-96][96] u16, 1..3 = path volumes (1 holds S4 after a full run),
 * 4 = disparity before the median, 5 = after the median; stop_after 1 = stop after the vertical sweep */
int vslam_sgbm_debug_read(vslam_ctx* ctx, int pair, int stage, void* host_out, size_t bytes);
int vslam_sgbm_debug_stop_after(vslam_ctx* ctx, int stage);

/* ------------------------------------------------------------------------------------------------
 * K17  landmark-sharded BA session for one-process-per-GPU runs (no reference equivalent: the reference
 * is single-process).  Rank r owns the landmarks [shard_begin, shard_end) with all their observations,
 * poses are replicated.  The CALLER owns the LM control flow (same schedule as vslam_ba_optimize) and
 * performs the collectives on three DEVICE reduce buffers between the phases:
 *   r1 = [Hpp 36K | bp 6K | chi2 | max diag Hll]   all-reduce(sum) after BUILD (max for the last entry)
 *   r2 = [S 6Kx6K | bs 6K]                          all-reduce(sum) after SCHUR  -- one per LM trial
 *   r3 = [chi2_trial, scale, ok, -, cnt_le[0..5]]   all-reduce(sum) after SOLVE_UPDATE / RELABEL_COUNT
 * Phases (vslam_ba_session_phase): 1 BUILD, 2 IMPORT_BUILD (after reducing r1), 3 SCHUR(value = lambda),
 * 4 SOLVE_UPDATE(value = lambda; after reducing r2), 5 RELABEL_COUNT, 6 RELABEL_APPLY(value = threshold).
 * vslam_ba_session_trial_done(accept) records the LM verdict.  vslam_ba_session_end downloads the poses and
 * this rank's shard of points / per-edge chi2 / inlier flags (zeros elsewhere: sum over ranks assembles).
 * stereo-visual-slam_b200/sharding.py is the reference driver (torch.distributed, NCCL or gloo).
 * STREAM CONTRACT: every phase is enqueued on the context stream (vslam_ctx_set_stream) without synchronising; the
 * caller's collectives on r1/r2/r3 and its host reads of them must be ordered on that SAME stream (with
 * torch.distributed: make the context stream torch's current stream before vslam_ba_session_begin, as
 * ffi.GpuBaSession does) -- otherwise the phases race with the all-reduces.
 * For a single-process, multi-device caller use vslam_ba_optimize_multi below: it needs no collective library.
 * ---------------------------------------------------------------------------------------------- */
int vslam_ba_reduce_sizes(int n_poses, int* r1_doubles, int* r2_doubles, int* r3_doubles);
int vslam_ba_session_begin(vslam_ctx* ctx, int n_poses, const double* poses, int n_points, const double* points,
                           int n_obs, const int32_t* obs_pose, const int32_t* obs_point, const double* obs_uv,
                           const double* Kmat, const vslam_ba_options* opt, int shard_begin, int shard_end,
                           double* d_r1, double* d_r2, double* d_r3);
int vslam_ba_session_phase(vslam_ctx* ctx, int phase, double value);
int vslam_ba_session_trial_done(vslam_ctx* ctx, int accept);
int vslam_ba_session_end(vslam_ctx* ctx, double* poses, double* points, double* chi2_per_obs, uint8_t* point_inlier);
/* ------------------------------------------------------------------------------------------------
 * K17b  ONE window on n_dev GPUs of this process (strong scaling), same arguments and results as
 * vslam_ba_optimize.  ctxs[r] must live on distinct devices with peer access (NVLink / NVSwitch); rank r owns a
 * contiguous, observation-balanced landmark range, poses are replicated.  The whole Levenberg-Marquardt loop runs
 * device-side in one persistent kernel per GPU; per LM trial the ranks exchange their partial reduced camera system
 * [S | bs | bp | chi2] by direct peer stores + release/acquire flag words (no host round trip, no collective library)
 * and sum it in rank order, so every rank solves the bit-identical 6K x 6K system; the accept/reject scalars, the
 * lambda initialisation and the relabel counts ride on the same flag protocol.  res->reserved returns the number of
 * exchanges executed.  Replaces the same reference code as vslam_ba_optimize (optimization.cpp:103-436); n_dev == 1
 * forwards to it.  Host-synchronous; every context's stream is used.
 * ---------------------------------------------------------------------------------------------- */
int vslam_ba_optimize_multi(vslam_ctx* const* ctxs, int n_dev, int n_poses, double* poses, int n_points, double* points,
                            int n_obs, const int32_t* obs_pose, const int32_t* obs_point, const double* obs_uv,
                            const double* Kmat, const vslam_ba_options* opt, vslam_ba_result* res, double* chi2_per_obs,
                            uint8_t* point_inlier);

/* device-side profile of the last vslam_ba_optimize_multi call as seen by rank 0 (ctxs[0]): nanoseconds (GPU
 * globaltimer) [kernel start .. first exchange complete (includes the launch skew between the GPUs), first exchange
 * complete .. end, inside the per-trial system exchanges, publish + wait for the slowest peer over all exchanges] */
int vslam_ba_multi_last_profile_ns(vslam_ctx* ctx, uint64_t* ns4);

/* Measurement probe (the north star's "tensor cores only for the dense J^T J camera-block GEMM"): after phase SCHUR of a
 * session that covers ALL landmarks, form the same product -(Hpl Hll^-1 Hpl^T) as ONE dense fp64 SYRK on the tensor
 * cores (DMMA m8n8k4) into d_S_dense (n x n device doubles, n = 6 * n_poses, upper triangle written) and report the
 * device times of the dense-operand fill and of the SYRK.  The session's sparse result is in r2 for comparison.
 * optimize_map's BlockSolver_6_3 (optimization.cpp:111-120) is the call both variants accelerate. */
int vslam_ba_session_schur_dense(vslam_ctx* ctx, double* d_S_dense, float* ms_fill, float* ms_syrk);

#ifdef __cplusplus
}
#endif
#endif /* VSLAM_B200_H */
