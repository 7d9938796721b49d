// ba_kernels.cuh — device-side contract of the bundle-adjustment engine (path B).
//
// What runs where (one LM iteration = one pass through steps 2-5, TrustRegionMinimizer +
// LevenbergMarquardtStrategy + SchurEliminator<2,3,3> semantics):
//   1  k_colnorm        iteration 0 only: Jacobi column scaling 1/(1+||J_col||)
//   2  k_schur          fused residual + Jacobian + Huber + block Hessians + Schur
//                       complement: builds the reduced camera system S, rhs (point-major,
//                       one warp per point, FP64 atomics into L2-resident S)
//   3  k_cam_diag + chol_* (ba_chol.cu)  S += U + D^2, blocked FP64 Cholesky, solve
//   4  k_backsub + k_cam_update          point steps, candidate state, model cost change
//   5  k_cost           cost at the candidate
// Reference entry points replaced: ceres::Solve at ba_solver.cc:591,636,672 for problems
// built by BASolver::SetUp (ba_solver.cc:330-356) from ReProjectionCost
// (cost_factor_ceres.h:19-40) and the camera models (camera_model.hpp:93-210).
#pragma once
#include <cstdint>
#include <vector>

#include "ba_plan.cuh"
#include "common.cuh"

namespace xrb {

struct BAConsts {  // cost-functor / loss constants (xrb_ba_options)
    double huber_a, huber_b;  // a, a^2
    double min_depth, neg_depth_residual;
};

// Read-only problem description on the device (point-major CSR).
struct BAProblemDev {
    int n_cams, n_pts_local, n_obs_local, nc;  // nc = reduced camera-system dimension
    const double *intr;                         // [8 * n_intr]
    const int32_t *intr_model;                  // [n_intr]
    const int32_t *cam_intr;                    // [n_cams]
    const int32_t *colq, *colt;                 // [n_cams] first reduced column or -1
    const int32_t *pt_ptr;                      // [n_pts_local + 1]
    const int32_t *obs_cam;                     // [n_obs_local] (point-major order)
    const double *obs_uv;                       // [2 * n_obs_local]
    const uint8_t *pt_var;                      // [n_pts_local]
    // generation-2 Schur structure (ba_struct.cu); n_blocks < 0 selects generation 1
    const int32_t *obs_pt;                      // [n_obs_local] local point of each observation
    const int32_t *cam_ptr, *cam_obs;           // camera-major CSR of point-major obs indices
    int n_blocks;                               // off-diagonal 6x6 blocks of S with >= 1 point
    long long n_inc;                            // (block, point) incidences in `inc`
    const int32_t *blk_ptr;                     // [n_blocks + 1] offsets into inc
    const int2 *blk_cams;                       // [n_blocks] (cam_a, cam_b), cols(a) >= cols(b)
    const int2 *inc;                            // (obs_i of cam_a, obs_j of cam_b), same point
};

// Mutable state: a set of (q, t, X) arrays.
struct BAStateDev {
    double *q, *t, *X;  // [4C], [3C], [3 * n_pts_local]
};

// Exchange buffer layout (one SUM all-reduce per linear solve in multi-GPU mode):
//   rhs  : nt x 64         right-hand side (padded); becomes y = L^-1 rhs during the factorisation
//   U    : nc x 6          rows of the camera block-diagonal J_c^T J_c
//   Ud   : nc x 6          rows of the diagonal blocks of W V^-1 W^T (subtracted by k_cam_diag)
//   gc   : nc              J_c^T r (scaled), for the gradient norm
//   S    : n_tiles_orig x 4096  the structurally non-zero tiles of the lower triangle (TileMap);
//                          the fill tiles follow and are not exchanged
//   n2c  : nc              squared column norms (iteration 0 only, exchanged on its own)
struct BALinSys {
    double *S;
    TileMap tm;
    double *rhs;
    double *U, *Ud, *gc, *n2c;
    double *Vinv, *gp;  // per local point: V^-1 (6), g_p (3)
    double *sc, *sp;    // Jacobi scaling: camera columns [nc], point columns [3 * n_pts_local]
    double *Tt;         // generation 2: per observation W (V + D^2)^-1/2  (6 x 3 = 18 doubles)
    double *h;          // generation 2: per local point (V + D^2)^-1 g_p  (3 doubles)
};

// Fused Schur complement of sequence-like scenes (ba_kernels.cu §2b): windows of kWinCams consecutive
// cameras, stride = kWinCams - (camera span of a point); every window split over `parts` CTAs.
constexpr int kWinCams = 24, kWinBlocks = kWinCams * (kWinCams - 1) / 2, kWinMaxSpan = 16, kWinMaxTrack = 32;
struct WindowPlanDev {
    const int32_t *win_pts = nullptr;   // local point ids sorted by first camera
    const int32_t *cta_ptr = nullptr;   // [n_ctas + 1] ranges into win_pts
    const int32_t *cta_cam0 = nullptr;  // [n_ctas] first camera of the CTA's window
    double *pS = nullptr, *pC = nullptr;  // partial windows: [n_ctas][kWinBlocks][36], [n_ctas][kWinCams][54]
    int n_ctas = 0, n_win = 0, parts = 1, stride = 8;
};

// scalar slots (doubles) produced on the device, read by the host controller
enum {
    SC_MODEL_CHANGE = 0,  // -(J s)^T (r + J s / 2)
    SC_CAND_COST = 1,
    SC_STEP_NORM2 = 2,    // ||x - x_candidate||^2 (ambient)
    SC_XNORM2 = 3,        // ||x_candidate||^2 over variable blocks
    SC_COST = 4,          // cost at the current state (iteration 0 / fixed cost)
    SC_GRAD_MAX_PT = 5,   // max |g_p / scale| over local variable points (as double bits)
    SC_GRAD_MAX_CAM = 6,
    SC_FAIL = 7,          // != 0: singular point block / non-finite value
    SC_COUNT = 8
};

int ba_launch_colnorm(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                      const BALinSys &L, cudaStream_t st);
// unit Jacobi scales + zeroed column norms for iteration 0, and ||x||^2 over the variable blocks -> *out
int ba_launch_start_norm(const BAProblemDev &P, const BAStateDev &x, const BALinSys &L, int with_cams, double *out,
                         cudaStream_t st);
int ba_launch_finish_scaling(const BAProblemDev &P, const BALinSys &L, cudaStream_t st);
// generation 2: linearise (per-observation records, no atomics) -> gather per block ->
// camera-major diagonal blocks / gradient / rhs
int ba_launch_lin(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                  const BALinSys &L, double inv_radius, double *scalars, cudaStream_t st);
int ba_launch_gather(const BAProblemDev &P, const BALinSys &L, cudaStream_t st);
int ba_launch_cam_blocks(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                         const BALinSys &L, cudaStream_t st);
int ba_launch_schur_window(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, const BALinSys &L, double inv_radius,
                           double *scalars, const WindowPlanDev &W, cudaStream_t st);
// anchor[p] = first camera of local point p; stats[3] = max camera span, longest track, a camera twice on a point
int ba_launch_point_anchor(const BAProblemDev &P, int32_t *anchor, int32_t *stats, cudaStream_t st);
int ba_launch_cam_diag(const BAProblemDev &P, const BAStateDev &x, const BALinSys &L,
                       double inv_radius, double *scalars, cudaStream_t st);
int ba_launch_backsub(const BAProblemDev &P, const BAStateDev &x, const BAStateDev &cand,
                      const BAConsts &k, const BALinSys &L, const double *yc, double *step_p,
                      double *scalars, cudaStream_t st);
int ba_launch_cam_update(const BAProblemDev &P, const BAStateDev &x, const BAStateDev &cand,
                         const BALinSys &L, const double *yc, double *scalars, int add_norms,
                         cudaStream_t st);
// mode 0: active residual blocks (>=1 variable block); mode 1: all-constant blocks (fixed cost)
int ba_launch_cost(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, int mode,
                   double *scalar_out, cudaStream_t st);
int ba_launch_residuals(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                        const int32_t *obs_orig, double *out, cudaStream_t st);

// ba_filter.cu — post-BA point filter (FilterPoints3d, track_processor.cc:321-349) over the resident CSR
int ba_launch_filter(const BAProblemDev &P, const BAStateDev &st, const int32_t *obs_orig, double *ctr, int32_t *order,
                     uint8_t *flag, double max_re, double deg, uint8_t *keep_obs, uint8_t *pt_outlier, double *pt_error,
                     double *pt_angle, int32_t *counts, cudaStream_t stream);

// ba_tilechol.cu — sparse tile Cholesky of S in one persistent kernel (task DAG over resident CTAs,
// flags in global memory), forward substitution folded in (rhs -> y), back-substitution in a second
// kernel.  dinv: nt x 4 x 256 doubles (inverses of the 16 x 16 diagonal blocks of L).
struct CholWorkspace {  // per solver: dependency flags and partial sums of one factorisation in flight
    DevBuf flags, wpart;
    void release();
};
int ba_launch_tile_cholesky_solve(const CholPlanDev &plan, CholWorkspace &ws, double *tiles, double *rhs, double *dinv,
                                  double *x_out, double *fail_flag, cudaStream_t st, int64_t *launches);
// identity rows of the padding columns (tile-aligned parts, matrix end): S[c][c] = 1 for the listed columns
int ba_launch_set_holes(const BALinSys &L, const int32_t *holes, int n_holes, cudaStream_t st);
int ba_tile_cholesky_aborted(const CholWorkspace &ws, int *aborted);
// debug: clock64 of the chain CTA after each block column of the last factorisation
int ba_tile_cholesky_trace(int enable, long long *out, int cap);

}  // namespace xrb
