/*
 * xrsfm_b200.h — C ABI of the B200-native engine behind XRSfM's two hot paths.
 *
 * Plain C: pointers + sizes only, no C++/torch types.  Every entry point names the
 * reference interface it replaces (paths relative to the openxrlab/xrsfm tree).
 *
 *   path M (matching):  xrsfm::FeatureMatching -> SiftMatch(uint8) -> SiftMatchGPU
 *                       src/feature/feature_processing.cc:118-154,222-308
 *                       3rdparty/SiftGPU/SiftGPU.h:277-372, SiftMatchCU.cpp:55-215
 *   path B (bundle adj): xrsfm::BASolver::{GBA,KGBA,LBA} -> ceres::Solve
 *                       src/optimization/ba_solver.h:14-30, ba_solver.cc:330-391,523-678
 *
 * All functions return XRB_OK (0) or a negative xrb_status unless stated otherwise;
 * xrb_last_error() gives a thread-local message for the last failure.
 * The library fails loudly (XRB_ERR_NO_DEVICE) when no sm_100 GPU is usable:
 * there is no CPU fallback behind this ABI.
 */
#ifndef XRSFM_B200_H_
#define XRSFM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XRB_ABI_VERSION 2

typedef enum xrb_status {
    XRB_OK = 0,
    XRB_ERR_INVALID = -1,   /* bad argument */
    XRB_ERR_NO_DEVICE = -2, /* no CUDA device / wrong architecture */
    XRB_ERR_CUDA = -3,      /* CUDA runtime error (SiftMatchCU.cpp:209-212 returns -1) */
    XRB_ERR_CAPACITY = -4,  /* caller buffer too small; sizes were still reported */
    XRB_ERR_NUMERIC = -5,   /* BA: non-finite step / factorisation breakdown */
    XRB_ERR_COMM = -6       /* BA multi-GPU: exchange hook failed */
} xrb_status;

int xrb_abi_version(void);
const char *xrb_last_error(void);
/* number of kernels this library has launched in the calling process so far */
uint64_t xrb_kernel_launch_count(void);

/* ------------------------------------------------------------------ */
/* Path M — SIFT descriptor matching                                   */
/* ------------------------------------------------------------------ */

typedef struct xrb_matcher xrb_matcher;

/* Replaces SiftMatchGPU::SiftMatchGPU(int) + SetLanguage(CUDA_DEVICE0+device) +
 * VerifyContextGL() + Allocate(max_features, mbm)  (SiftGPU.h:307-324,
 * feature_processing.cc:53-88).  max_features is rounded up to a multiple of 32
 * like SiftMatchCU::SetMaxSift (SiftMatchCU.cpp:84-87).  Returns NULL on failure. */
xrb_matcher *xrb_match_create(int max_features, int device);
void xrb_match_destroy(xrb_matcher *m);
int xrb_match_max_features(const xrb_matcher *m);

/* Per-pair compatibility surface ------------------------------------ */

/* Replaces SiftMatchGPU::SetDescriptors(int index, int num, const unsigned char*, int id)
 * (SiftGPU.h:331-334, SiftMatchCU.cpp:100-118).  `desc` is a HOST pointer to
 * num x 128 row-major uint8 (src/base/types.h:9-10).  index is clamped to {0,1},
 * num is clamped to max_features, id != -1 && id == previous id skips the upload. */
int xrb_match_set_descriptors(xrb_matcher *m, int index, int num, const uint8_t *desc,
                              int id);

/* Replaces SiftMatchGPU::GetSiftMatch (SiftGPU.h:337-343, SiftMatchCU.cpp:175-215).
 * Returns the number of matches written to match_buffer (ascending first index,
 * truncated at max_match), 0 if a set is empty, or -1 on a CUDA error — the
 * reference's own convention. */
int xrb_match_get(xrb_matcher *m, int max_match, uint32_t (*match_buffer)[2],
                  float distmax, float ratiomax, int mutual_best_match);

/* Batched surface (what a replacement FeatureMatching uses) ---------- */

/* Make the descriptors of n_images images resident in HBM.  counts[i] features for
 * image i (clamped to max_features), descs[i] = HOST pointer to counts[i] x 128 uint8.
 * Replaces the per-pair H2D copies of feature_processing.cc:136-137. */
int xrb_match_upload_images(xrb_matcher *m, int n_images, const int32_t *counts,
                            const uint8_t *const *descs);

/* Same, from one contiguous HOST block: image i starts at row offsets[i] (rows of
 * 128 bytes), offsets has n_images+1 entries. */
int xrb_match_upload_packed(xrb_matcher *m, int n_images, const int64_t *row_offsets,
                            const uint8_t *desc_block);

/* Same, but the block is already DEVICE memory (no copy is made; the caller keeps it
 * alive).  Used to measure the HBM-resident path. */
int xrb_match_attach_device(xrb_matcher *m, int n_images, const int64_t *row_offsets_host,
                            const uint8_t *desc_block_device);

/* Match n_pairs image pairs (indices into the uploaded set), HOST in / HOST out.
 * out_offsets has n_pairs+1 entries (prefix sums of per-pair match counts);
 * matches of pair p are out[out_offsets[p] .. out_offsets[p+1]) as (idx1, idx2),
 * each list exactly what SiftMatchCU::GetBestMatch would have produced for that pair
 * (SiftMatchCU.cpp:186-215).  out_capacity is in matches; on XRB_ERR_CAPACITY
 * out_offsets is still valid.  Semantics per pair: feature_processing.cc:118-154. */
int xrb_match_pairs(xrb_matcher *m, int n_pairs, const int32_t (*pairs)[2], float distmax,
                    float ratiomax, int mutual_best_match, int max_match,
                    int64_t *out_offsets, uint32_t (*out)[2], int64_t out_capacity);

/* Device-resident variant: pairs_dev, counts_dev[n_pairs], out_dev are DEVICE pointers;
 * matches of pair p go to out_dev[p*out_stride ...]; work is enqueued on `stream`
 * (a cudaStream_t, may be NULL) and NOT synchronised. */
int xrb_match_pairs_device(xrb_matcher *m, int n_pairs, const int32_t (*pairs_dev)[2],
                           float distmax, float ratiomax, int mutual_best_match,
                           int max_match, int32_t *counts_dev, uint32_t (*out_dev)[2],
                           int out_stride, void *stream);

/* Test hook (no reference counterpart): float(acos(double(min(float(v)*2^-18,1)))) for
 * v = 0..n-1 as CUDA computes it, so the CPU oracle's libm can be checked over the whole
 * domain of ProgramCU.cu:1830-1831. out is a HOST array. */
int xrb_match_debug_dist_table(float *out_host, int n);

/* Which kernel generation a matcher runs: 0 = auto (best available), 1 = dp4a tiles,
 * 2 = tcgen05 + global top-2 state, 3 = tcgen05 with the pair's state in shared memory and
 * the finalize step fused in (falls back to 2 for images above 4096 descriptors).
 * Returns the variant now in force. */
int xrb_match_set_variant(xrb_matcher *m, int variant);

/* ------------------------------------------------------------------ */
/* Path B — bundle adjustment                                          */
/* ------------------------------------------------------------------ */

typedef struct xrb_ba_solver xrb_ba_solver;

/* Flat (SoA) restatement of what BASolver::SetUp builds per observation
 * (ba_solver.cc:330-356): residual block (q[4], t[3], X[3], intrinsics[K]). */
typedef struct xrb_ba_problem {
    int32_t n_cams, n_pts, n_obs, n_intr;
    double *cam_q;             /* [4*n_cams] Eigen coeffs order x,y,z,w; in/out   */
    double *cam_t;             /* [3*n_cams] in/out                               */
    double *pts;               /* [3*n_pts]  in/out                               */
    const double *intr;        /* [8*n_intr] camera params, padded to 8           */
    const int32_t *intr_model; /* [n_intr] model id 0..4 (camera_model.hpp:93-210)*/
    const int32_t *cam_intr;   /* [n_cams] -> intrinsics index                    */
    const int32_t *obs_cam;    /* [n_obs]                                         */
    const int32_t *obs_pt;     /* [n_obs]                                         */
    const double *obs_uv;      /* [2*n_obs] measured pixel                        */
    const uint8_t *cam_q_fixed; /* [n_cams] or NULL: SetParameterBlockConstant(q) */
    const uint8_t *cam_t_fixed; /* [n_cams] or NULL: ba_solver.cc:611-614         */
    const uint8_t *pt_fixed;    /* [n_pts]  or NULL: ba_solver.cc:380-382         */
} xrb_ba_problem;

/* ceres::Solver::Options fields the reference sets (ba_solver.cc:70-77,624-634,
 * 665-670) plus the cost-functor constants (cost_factor_ceres.h:29-31, ba_solver.cc:343). */
typedef struct xrb_ba_options {
    int32_t max_iterations;     /* 50 GBA accurate / 20 otherwise / 5 LBA */
    double function_tolerance;  /* 1e-5 / 1e-4                            */
    double parameter_tolerance; /* 1e-6 / 1e-5                            */
    double gradient_tolerance;  /* 1e-10 (Ceres default)                  */
    double initial_radius;      /* 1e4 (Ceres default) / 1e6 KGBA         */
    double huber_a;             /* 5.99                                   */
    double min_depth;           /* 1e-2                                   */
    double neg_depth_residual;  /* 12.0                                   */
    int32_t verbose;            /* minimizer_progress_to_stdout           */
    int32_t fixed_iterations;   /* !=0: run exactly max_iterations LM iterations,
                                   ignoring the tolerance tests (bench only) */
} xrb_ba_options;

enum {
    XRB_BA_CONVERGENCE = 0,    /* ceres::CONVERGENCE    */
    XRB_BA_NO_CONVERGENCE = 1, /* ceres::NO_CONVERGENCE */
    XRB_BA_FAILURE = 2         /* ceres::FAILURE        */
};

typedef struct xrb_ba_iteration {
    int32_t iteration;
    int32_t step_is_valid, step_is_successful;
    double cost;               /* cost after this iteration (incl. fixed cost) */
    double cost_change;        /* x_cost - candidate_cost                      */
    double gradient_max_norm;
    double step_norm;
    double relative_decrease;  /* rho                                          */
    double trust_region_radius;
    double model_cost_change;
} xrb_ba_iteration;

/* The fields PrintSolverSummary reads (ba_solver.cc:14-68). */
typedef struct xrb_ba_summary {
    int32_t num_residuals_reduced;
    int32_t num_effective_parameters_reduced;
    int32_t num_successful_steps, num_unsuccessful_steps;
    int32_t termination_type;
    double initial_cost, final_cost, fixed_cost;
    double total_time_in_seconds;
    double linear_solver_seconds, residual_seconds; /* device-event split */
    int32_t n_iterations_logged; /* entries valid in `iterations` (incl. iteration 0);
                                    like Ceres, an iteration that ends on a tolerance test
                                    is not logged and its step is discarded */
    int32_t num_lm_iterations;   /* passes through the LM loop executed = linear solves,
                                    incl. the terminating one: the unit of the metric */
    xrb_ba_iteration iterations[128];
} xrb_ba_summary;

void xrb_ba_default_options(xrb_ba_options *opt); /* Ceres defaults + xrsfm constants */

xrb_ba_solver *xrb_ba_create(int device);
void xrb_ba_destroy(xrb_ba_solver *s);

/* Multi-GPU.  With world > 1 each rank owns a shard of the points (all their observations) and
 * a full replica of the cameras; once per linear solve the packed
 * [reduced camera system | rhs | camera blocks | gradient | scalars] buffer is summed over ranks
 * ("ncclAllReduce(ncclDouble, ncclSum) of the camera-block contributions").
 *
 * Native exchange: the library owns an NCCL communicator and issues ncclAllReduce on the
 * solver's own stream — no host synchronisation, nothing of the caller's in the loop; the
 * solved points come back through one grouped ncclBroadcast (an all-gather of unequal shards).
 *   rank 0:     xrb_nccl_unique_id(id)  -> ship the 128 bytes to every rank by any means
 *   every rank: xrb_ba_comm_init(s, id, rank, world)      (collective: ncclCommInitRank)
 * NCCL is bound at run time (libnccl.so.2, the copy already mapped in the process if any);
 * XRB_ERR_COMM when it is missing. */
#define XRB_NCCL_ID_BYTES 128
int xrb_nccl_unique_id(uint8_t id[XRB_NCCL_ID_BYTES]);
int xrb_ba_comm_init(xrb_ba_solver *s, const uint8_t id[XRB_NCCL_ID_BYTES], int rank, int world);

/* Alternative: a caller-supplied hook (used where the caller already owns a communicator).
 * `allreduce(buf_dev, count, stream, user)` must perform an in-place SUM over ranks of `count`
 * doubles in DEVICE memory ORDERED ON `stream` (a cudaStream_t): it may return as soon as the
 * reduction is enqueued there — e.g. ncclAllReduce(buf, buf, count, ncclDouble, ncclSum, comm,
 * stream) — and the solver launches its next kernel on the same stream without synchronising.
 * A hook that reduces on another stream must make `stream` wait for it (event) before it
 * returns.  Must return 0.  A communicator set by xrb_ba_comm_init takes precedence. */
typedef int (*xrb_allreduce_fn)(void *buf_dev, size_t count, void *stream, void *user);
int xrb_ba_set_exchange(xrb_ba_solver *s, int rank, int world, xrb_allreduce_fn fn,
                        void *user);

/* Host-only helper (needs no GPU): the contiguous point range [*lo, *hi) rank `rank` of
 * `world` owns, given the number of observations of every point.  Shards are balanced by
 * the Schur-complement work sum_p (k_p^2 + 4 k_p).  xrb_ba_load uses exactly this split. */
int xrb_ba_shard_range(int32_t n_pts, const int32_t *obs_per_point, int rank, int world,
                       int32_t *lo, int32_t *hi);

/* Replaces ceres::Solve at ba_solver.cc:591,636,672 (problem build included): HOST
 * arrays in, poses/points updated in place, summary filled.  In multi-GPU mode every
 * rank passes the FULL problem and the solver keeps only its shard of the points
 * (balanced by sum k_p^2); all ranks return the same poses and points. */
int xrb_ba_solve(xrb_ba_solver *s, const xrb_ba_problem *prob, const xrb_ba_options *opt,
                 xrb_ba_summary *summary);

/* Split form, used to time the HBM-resident inner loop: load() uploads and builds the
 * point-major CSR, run() iterates on the device state (may be called repeatedly after
 * reset()), fetch() copies poses/points back into the problem's host arrays. */
int xrb_ba_load(xrb_ba_solver *s, const xrb_ba_problem *prob);
int xrb_ba_reset(xrb_ba_solver *s); /* restore the state uploaded by load() */
int xrb_ba_run(xrb_ba_solver *s, const xrb_ba_options *opt, xrb_ba_summary *summary,
               void *stream);
int xrb_ba_fetch(xrb_ba_solver *s, xrb_ba_problem *prob);

/* Per-observation residuals (after the depth branch, before the loss) at the current
 * device state, in the caller's observation order: out[2*n_obs] HOST.  This is
 * ReProjectionCost::operator() (cost_factor_ceres.h:19-40) over the whole problem. */
int xrb_ba_residuals(xrb_ba_solver *s, double *out_residuals);

/* Batched local bundle adjustment: n_problems independent problems (the <= 8-frame windows BASolver::LBA
 * builds, ba_solver.cc:523-591, called once per registered frame at src/mapper/incremental_mapper.cc:71; or
 * pose refinements) solved concurrently on one device by n_workers engines (<= 0: 8), each with its own
 * streams, so that the per-solve launch and synchronisation latencies overlap.  Every problem gets exactly
 * what xrb_ba_solve would give it (same options for all; states updated in place, one summary each).
 * Returns the first failure, after all problems were attempted. */
int xrb_ba_solve_batch(int device, int n_problems, const xrb_ba_problem *problems, const xrb_ba_options *opt,
                       xrb_ba_summary *summaries, int n_workers);

/* Batched pose refinement: the Ceres block of RegisterImage (src/geometry/pnp.cc:38-71) for n_poses frames in
 * ONE kernel launch (one CTA per pose, the whole trust-region loop on the device).  Per pose: the 2D-3D
 * correspondences [offsets[p], offsets[p+1]) of uv (frame.points[p2d_id], pixels) and xyz (the tracks' world
 * points, constant), an optional inlier mask (SolvePnP_colmap's inlier_mask, pnp.cc:44-46; NULL = all), the
 * camera (8 padded parameters + model id 0..4, constant) and the pose q (Eigen coeffs x,y,z,w) / t, refined in
 * place.  Same cost functor, Huber loss, quaternion parameterisation and minimiser semantics as xrb_ba_run with one
 * variable camera and constant points; xrb_pose_default_options gives what pnp.cc:56-57 sets (Ceres defaults,
 * max_num_iterations = 10).  final_cost / initial_cost are 1/2 sum rho like ceres::Solver::Summary (pnp.cc:60-67
 * prints sqrt(cost / num_residuals)). */
typedef struct xrb_pose_summary {
    int32_t num_residuals;
    int32_t num_lm_iterations; /* linear solves executed */
    int32_t num_successful_steps, num_unsuccessful_steps; /* iteration 0 counts as successful (Ceres) */
    int32_t termination_type;  /* XRB_BA_CONVERGENCE / NO_CONVERGENCE / FAILURE */
    int32_t reserved;
    double initial_cost, final_cost;
} xrb_pose_summary;
void xrb_pose_default_options(xrb_ba_options *opt);
int xrb_pose_refine_batch(int device, int n_poses, const int64_t *offsets, const double *uv, const double *xyz,
                          const uint8_t *inlier_mask, const double *intr, const int32_t *intr_model, double *q,
                          double *t, const xrb_ba_options *opt, xrb_pose_summary *summaries);
/* Device time (ms, CUDA events) of the kernel of the last xrb_pose_refine_batch on `device`; < 0 if none yet. */
double xrb_pose_last_kernel_ms(int device);

/* Post-BA point filter on the solver's CURRENT state (after xrb_ba_run / xrb_ba_solve; single GPU).
 * Replaces Point3dProcessor::FilterPoints3d(map, max_re, deg) (src/geometry/track_processor.cc:321-349,
 * FilterPoint3d :279-319, UpdateTrackAngle :253-277, Reprojection_Error :19-26), which the mapper calls
 * after every KGBA (src/mapper/incremental_mapper.cc:83-85) — the poses and points it needs are still
 * resident.  HOST outputs, indexed like the loaded problem:
 *   keep_obs[n_obs]    0 = the reference would DeleteObservation / drop it with its track
 *   pt_outlier[n_pts]  1 = SetTrackOutlier
 *   pt_error[n_pts]    track.error (mean reprojection error of the kept observations), 0 if not computed
 *   pt_angle[n_pts]    track.angle_ as the early-exit scan leaves it (radians), 0 if not computed
 *   counts[2]          num_filtered1 (observations), num_filtered2 (tracks by angle)
 * A track's observations are visited in ascending camera index (the reference iterates a std::map keyed
 * by frame id: flatten the frames in id order, as the BA flattening does). */
int xrb_ba_filter_points3d(xrb_ba_solver *s, double max_re, double deg, uint8_t *keep_obs, uint8_t *pt_outlier,
                           double *pt_error, double *pt_angle, int32_t counts[2]);

/* Last-run device timings in milliseconds: [0] linearise+Schur, [1] reduced system
 * factor+solve, [2] back-substitution+update, [3] cost evaluation, [4] exchange,
 * [5] whole run; and launches of each (same indices). */
int xrb_ba_profile(const xrb_ba_solver *s, double ms[6], int64_t launches[6]);

/* Debug hooks (no reference counterpart).
 * xrb_debug_chol_trace: enable != 0 arms a recorder in the factorisation's chain CTA; enable == 0
 * stops it and copies, for the last factorisation, the SM clock after each 64-column block column
 * (up to cap values); returns the number copied (or a negative status).
 * xrb_debug_tile_solve: solve A x = rhs for a symmetric positive definite A given dense (n x n
 * row-major, lower triangle read, entries further than bw from the diagonal ignored) with the
 * reduced-camera-system solver (tile Cholesky + substitutions) on the current device; the best
 * device time of `reps` runs goes to *ms_out.  HOST pointers. */
int xrb_debug_chol_trace(int enable, int64_t *out, int cap_records);
int xrb_debug_tile_solve(int n, int bw, const double *A, const double *rhs, double *x_out, int reps,
                         double *ms_out);
/* ---- Geometric verification of matched pairs (the step that follows path M) --------------------
 * Replaces the loop body of FeatureMatching that calls SolveFundamnetalCOLMAP
 * (src/feature/feature_processing.cc:256-296, src/geometry/epipolar_geometry.hpp:10-27):
 * colmap::LORANSAC<7-point, 8-point> with the options below (defaults = the reference's), the squared
 * Sampson error as residual, inlier count / residual sum as support.  One launch verifies a whole
 * batch of pairs, one CTA per pair.  Each pair starts from a freshly seeded generator (std::mt19937,
 * seed 0, restated bit for bit together with libstdc++'s uniform_int_distribution): the sample
 * sequence is the one the reference draws for the first pair an OpenMP thread processes. */
typedef struct xrb_fm_options {  /* colmap::RANSACOptions as epipolar_geometry.hpp:13-18 sets them */
    double max_error, min_inlier_ratio, confidence;
    int64_t min_num_trials, max_num_trials;
} xrb_fm_options;
typedef struct xrb_fm_report {
    int32_t success, best_is_local;  /* best_is_local: the model came from the local optimisation */
    int64_t num_trials, num_inliers;
    double residual_sum;
    double F[9];                     /* row-major, x2^T F x1 = 0 */
} xrb_fm_report;
void xrb_fm_default_options(xrb_fm_options *o);
/* offsets[n_pairs+1]: matches of pair p are rows offsets[p] .. offsets[p+1] of pts1 / pts2 (x, y pairs of
 * frame1.points[match.id1] / frame2.points[match.id2]); inlier_mask has offsets[n_pairs] entries.
 * HOST pointers; a pair with fewer than 7 matches reports success = 0. */
int xrb_fm_loransac_batch(int device, int n_pairs, const int64_t *offsets, const double *pts1, const double *pts2,
                          const xrb_fm_options *opt, xrb_fm_report *reports, char *inlier_mask);
/* host-only debug hook: the first `trials` samples (7 indices each) a fresh generator draws for n matches */
int xrb_debug_fm_samples(int n, int trials, int32_t *out);

/* Host-only debug hooks (no device needed): the symbolic plan (tile slots, task lists with the flag
 * values they wait for) of a tile pattern pat[nt x nt] (lower part), and the column order of a band.
 * counts[16] = nt, n_tiles, n_tiles_orig, n_f, n_w, n_b, n_wb, n_far, n_chain_f, n_chain_b, depth_f,
 * depth_b; output arrays may be null and are filled up to their capacity in int32 elements. */
int xrb_debug_chol_plan(int nt, const uint8_t *pat, int32_t *counts, int32_t *tab, int32_t *ftasks, int cap_f,
                        int32_t *wtasks, int cap_w, int32_t *btasks, int cap_b, int32_t *wbtasks, int cap_wb,
                        int32_t *far_rows, int32_t *far_slots, int cap_far);
int xrb_debug_column_order(int n_cams, const int32_t *widths, int bw, int allow_nd, int32_t *start, int32_t *n_pad,
                           int32_t *parts);

/* Finer split of the last run, out[n >= 8]: [0..2] total ms of k_lin (+ memsets), k_gather,
 * k_cam_blocks; [3] linear solves executed; [4] off-diagonal 6x6 blocks of the reduced camera
 * system; [5] (block, point) incidences the gather walks; [6] reduced system dimension (variable camera
 * columns, without padding); [7] its half bandwidth in the natural column order.
 * n >= 16 adds the plan of the sparse tile Cholesky: [8] independent interiors of the column order (1 =
 * natural), [9] tile columns, [10] tiles incl. fill, [11] original tiles, [12] flops executed per solve,
 * [13] / [14] longest dependency paths (tasks) of factorisation / back-substitution, [15] chain CTAs.
 * n >= 20: [16] CTAs of the fused windowed Schur kernel (0 = gather path), [17] its camera stride,
 * [18] largest camera span of a point, [19] longest track. */
int xrb_ba_profile_detail(const xrb_ba_solver *s, double *out, int n);

/* ------------------------------------------------------------------ */
/* Wire formats either side of the two paths     */
/* ------------------------------------------------------------------ */
/* All files are the reference's raw little-endian dumps (src/utility/io_base.hpp:13-87).
 * Every reader is two-pass: *_scan reports the sizes, the caller allocates, *_read fills
 * flat arrays — the layouts the hot paths take, without the reference's AoS objects. */

/* ftr.bin — ReadFeatures / SaveFeatures (src/utility/io_feature.hpp:37-100):
 *   int32 n_frames; per frame: name\0, int32 n_points, n_points x {float x, y, size, angle},
 *   n_points x 128 uint8.  names_bytes counts the terminating NULs. */
int xrb_ftr_scan(const char *path, int32_t *n_frames, int64_t *total_points, int64_t *names_bytes);
/* row_offsets[n_frames+1] (rows of 128 bytes), desc_block[total_points*128]; optional (may be
 * NULL): keypoints[total_points*4], names[names_bytes] (NUL-separated), name_offsets[n_frames+1]. */
int xrb_ftr_read(const char *path, int32_t n_frames, int64_t *row_offsets, uint8_t *desc_block,
                 float *keypoints, char *names, int64_t *name_offsets);
/* SaveFeatures(file, frames, with_descs = true) (run_matching.cc:31).  keypoints NULL writes
 * zeros; names NULL writes empty names. */
int xrb_ftr_write(const char *path, int32_t n_frames, const int64_t *row_offsets,
                  const uint8_t *desc_block, const float *keypoints, const char *names,
                  const int64_t *name_offsets);
/* ftr.bin -> HBM: the descriptors of every frame become the matcher's resident images (as
 * xrb_match_upload_packed would leave them), streamed through a small pinned staging ring; the
 * keypoints are skipped.  Frames above max_features are clamped like SetDescriptors
 * (SiftMatchCU.cpp:102). */
int xrb_match_upload_ftr(xrb_matcher *m, const char *path);

/* fp.bin — ReadFramePairs / SaveFramePairs (io_feature.hpp:102-147):
 *   uint64 n_pairs; per pair: int32 id1, id2, uint64 n_matches, n_matches x Match{int32 id1,
 *   int32 id2, float64 distance}, float64 E[9] (column-major), int32 inlier_num,
 *   n_matches x char inlier_mask.  Like the reference's reader (:120-126), pairs with
 *   id1 == id2 are dropped: scan and read do not count them. */
int xrb_fp_scan(const char *path, int64_t *n_pairs, int64_t *total_matches);
/* ids[n_pairs][2], offsets[n_pairs+1], matches[total][2]; optional: distances[total],
 * E[9*n_pairs], inlier_num[n_pairs], inlier_mask[total]. */
int xrb_fp_read(const char *path, int64_t n_pairs, int32_t (*ids)[2], int64_t *offsets,
                int32_t (*matches)[2], double *distances, double *E, int32_t *inlier_num,
                char *inlier_mask);
/* matches are the (i, j) lists xrb_match_pairs returns (same 32-bit pattern).  distances NULL
 * -> 0.0 (Match's default, types.h:15); E NULL -> zeros (the reference leaves E unassigned on
 * the fundamental-matrix path); inlier_mask NULL -> all 1; inlier_num NULL -> number of 1s. */
int xrb_fp_write(const char *path, int64_t n_pairs, const int32_t (*ids)[2], const int64_t *offsets,
                 const int32_t (*matches)[2], const double *distances, const double *E,
                 const int32_t *inlier_num, const char *inlier_mask);

/* COLMAP-style model — ReadColMapDataBinary / WriteColMapDataBinary (src/utility/io_ecim.cc:9-87,
 * 145-232): cameras.bin, images.bin, points3D.bin under dir (dir must end with '/', the
 * reference concatenates).  The model is flattened straight into an xrb_ba_problem exactly as
 * BASolver::SetUp would walk it (ba_solver.cc:330-356): one observation per (frame, p2d) whose
 * track id is present in points3D.bin, frames in file order, p2d in order. */
typedef struct xrb_colmap_sizes {
    int32_t n_cameras, n_frames, n_points;
    int64_t n_p2d;    /* 2-D points over all frames (tracked or not) */
    int64_t n_obs;    /* observations the BA problem will hold */
} xrb_colmap_sizes;
int xrb_colmap_scan(const char *dir, xrb_colmap_sizes *sizes);
/* prob's arrays must be allocated for the scanned sizes (n_cams = n_frames, n_pts = n_points,
 * n_obs, n_intr = n_cameras; the *_fixed arrays are left untouched).  Optional outputs:
 * frame_ids[n_frames], camera_ids[n_cameras], track_ids[n_points], obs_p2d[n_obs]. */
int xrb_colmap_read_problem(const char *dir, const xrb_colmap_sizes *sizes, xrb_ba_problem *prob,
                            int32_t *frame_ids, int32_t *camera_ids, uint64_t *track_ids,
                            int32_t *obs_p2d);
/* Copy the model dir_in -> dir_out with the poses and points of prob (same order as
 * xrb_colmap_read_problem produced): what WriteColMapDataBinary would write after the BA. */
int xrb_colmap_write_updated(const char *dir_in, const char *dir_out, const xrb_colmap_sizes *sizes,
                             const xrb_ba_problem *prob);

#ifdef __cplusplus
}
#endif
#endif /* XRSFM_B200_H_ */
