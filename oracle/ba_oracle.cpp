/*
 * oracle/ba_oracle.cpp — CPU restatement of XRSfM's bundle adjustment (path B), FP64.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under xrsfm_b200/ (the product) may link, import or
 * execute this file; it is the checker used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: "parity unpinned".  The reference has no BA arithmetic of its own: it
 * builds a ceres::Problem and calls ceres::Solve (src/optimization/ba_solver.cc:591,636,672).
 * Ceres Solver is an external, un-vendored dependency (CMakeLists.txt:15-24: XRPrimer
 * bundle or system Ceres <= 2.1, no pinned version) that is absent from /root/reference and
 * from this image, and the reference ships no tests or golden vectors.
 * This file therefore restates (a) the reference's own cost functor and camera models and
 * (b) the published Ceres 2.0/2.1 algorithm for the options the reference selects; it is
 * cross-checked by an independent numpy/scipy implementation in tests/test_ba_oracle.py
 * (finite-difference Jacobians, full damped normal equations instead of Schur elimination).
 *
 * Reference lines restated:
 *   ReProjectionCost::operator()      src/optimization/cost_factor_ceres.h:19-40
 *   camera models / WorldToImage      src/base/camera_model.hpp:57-68,93-210
 *   problem structure (SetUp)         src/optimization/ba_solver.cc:330-356
 *   constant blocks / gauge           ba_solver.cc:602-621, 655-663, 380-389
 *   solver options                    ba_solver.cc:70-77, 624-634, 665-670
 *   summary fields printed            ba_solver.cc:14-68
 *   pose refinement after PnP         src/geometry/pnp.cc:38-71 (the same functor, loss and parameterisation with one
 *                                     variable camera and constant points: xro_ba_solve on that problem is the checker
 *                                     of xrsfm_b200/csrc/pose_refine.cu, tests/test_pose_gpu.py)
 * Ceres (external) semantics restated, Ceres 2.1 file names for orientation:
 *   HuberLoss / Corrector             loss_function.cc, corrector.cc
 *   EigenQuaternionParameterization   local_parameterization.cc (x,y,z,w; x+ = dq * x)
 *   TrustRegionMinimizer              trust_region_minimizer.cc
 *   LevenbergMarquardtStrategy        levenberg_marquardt_strategy.cc
 *   SchurEliminator<2,3,3> + Cholesky schur_eliminator_impl.h, schur_complement_solver.cc
 */
#include <omp.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../include/xrsfm_b200.h"

namespace {

// ----------------------------------------------------------------------------------------
// Per-observation model
// ----------------------------------------------------------------------------------------

/* WorldToImage + d(uv)/d(xy) for the five models (camera_model.hpp:93-210).
 * Note the reference quirk for ids 0/1: Distortion() returns duv = xy, so u = 2 f x + c. */
inline void world_to_image(int model, const double *p, double x, double y, double uv[2],
                           double D[4]) {
    switch (model) {
        case 0: {  // SimplePinhole: f, cx, cy
            uv[0] = p[0] * (x + x) + p[1];
            uv[1] = p[0] * (y + y) + p[2];
            D[0] = 2 * p[0], D[1] = 0, D[2] = 0, D[3] = 2 * p[0];
            break;
        }
        case 1: {  // Pinhole: fx, fy, cx, cy
            uv[0] = p[0] * (x + x) + p[2];
            uv[1] = p[1] * (y + y) + p[3];
            D[0] = 2 * p[0], D[1] = 0, D[2] = 0, D[3] = 2 * p[1];
            break;
        }
        case 2:    // SimpleRadial: f, cx, cy, k
        case 3: {  // Radial (xrsfm flavour): fx, fy, cx, cy, k
            const double fx = p[0], fy = model == 2 ? p[0] : p[1];
            const double cx = model == 2 ? p[1] : p[2], cy = model == 2 ? p[2] : p[3];
            const double k = model == 2 ? p[3] : p[4];
            const double r2 = x * x + y * y, radial = k * r2;
            uv[0] = fx * (x + x * radial) + cx;
            uv[1] = fy * (y + y * radial) + cy;
            D[0] = fx * (1 + radial + 2 * k * x * x), D[1] = fx * (2 * k * x * y);
            D[2] = fy * (2 * k * x * y), D[3] = fy * (1 + radial + 2 * k * y * y);
            break;
        }
        default: {  // OpenCV: fx, fy, cx, cy, k1, k2, p1, p2
            const double fx = p[0], fy = p[1], cx = p[2], cy = p[3];
            const double k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
            const double x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
            const double radial = k1 * r2 + k2 * r2 * r2;
            const double du = x * radial + 2 * p1 * xy + p2 * (r2 + 2 * x2);
            const double dv = y * radial + 2 * p2 * xy + p1 * (r2 + 2 * y2);
            uv[0] = fx * (x + du) + cx;
            uv[1] = fy * (y + dv) + cy;
            const double drad_dx = (k1 + 2 * k2 * r2) * 2 * x, drad_dy = (k1 + 2 * k2 * r2) * 2 * y;
            const double ddu_dx = radial + x * drad_dx + 2 * p1 * y + p2 * (2 * x + 4 * x);
            const double ddu_dy = x * drad_dy + 2 * p1 * x + p2 * (2 * y);
            const double ddv_dx = y * drad_dx + 2 * p2 * y + p1 * (2 * x);
            const double ddv_dy = radial + y * drad_dy + 2 * p2 * x + p1 * (2 * y + 4 * y);
            D[0] = fx * (1 + ddu_dx), D[1] = fx * ddu_dy;
            D[2] = fy * ddv_dx, D[3] = fy * (1 + ddv_dy);
        }
    }
}

struct ObsEval {
    double r[2];      // residual (raw or robustified)
    double Jd[6];     // 2x3 wrt quaternion tangent (row-major)
    double Jt[6];     // 2x3 wrt translation
    double JX[6];     // 2x3 wrt point
    double rho0;      // loss value rho(s)
    int depth_branch; // 1 when pc.z < min_depth (constant residual, zero Jacobian)
};

/* ReProjectionCost (cost_factor_ceres.h:19-40) + autodiff Jacobians in closed form +
 * EigenQuaternionParameterization::ComputeJacobian + HuberLoss/Corrector.
 * want_jac = false leaves the Jacobians untouched. robustify = false returns raw r, J. */
inline void eval_obs(const double q[4], const double t[3], const double X[3], int model,
                     const double *intr, const double uv_meas[2], const xrb_ba_options &o,
                     bool want_jac, bool robustify, ObsEval &e) {
    const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
    // Eigen: uv = 2 * (u x v); pc = v + w*uv + u x uv   (Quaternion::_transformVector)
    double cx_ = 2 * (uy * X[2] - uz * X[1]);
    double cy_ = 2 * (uz * X[0] - ux * X[2]);
    double cz_ = 2 * (ux * X[1] - uy * X[0]);
    const double px = X[0] + w * cx_ + (uy * cz_ - uz * cy_);
    const double py = X[1] + w * cy_ + (uz * cx_ - ux * cz_);
    const double pz = X[2] + w * cz_ + (ux * cy_ - uy * cx_);
    const double pcx = px + t[0], pcy = py + t[1], pcz = pz + t[2];
    e.depth_branch = pcz < o.min_depth;
    if (e.depth_branch) {  // cost_factor_ceres.h:29-31
        e.r[0] = e.r[1] = o.neg_depth_residual;
        if (want_jac)
            for (int i = 0; i < 6; ++i) e.Jd[i] = e.Jt[i] = e.JX[i] = 0.0;
    } else {
        const double iz = 1.0 / pcz, x = pcx * iz, y = pcy * iz;
        double uvp[2], D[4];
        world_to_image(model, intr, x, y, uvp, D);
        e.r[0] = uvp[0] - uv_meas[0];
        e.r[1] = uvp[1] - uv_meas[1];
        if (want_jac) {
            // A = D * [[1/z,0,-x/z],[0,1/z,-y/z]]
            double A[6];
            A[0] = D[0] * iz, A[1] = D[1] * iz, A[2] = -(D[0] * x + D[1] * y) * iz;
            A[3] = D[2] * iz, A[4] = D[3] * iz, A[5] = -(D[2] * x + D[3] * y) * iz;
            for (int i = 0; i < 6; ++i) e.Jt[i] = A[i];
            // d pc / d X = I + 2w[u]x + 2[u]x[u]x  (== R(q) for unit q)
            const double uu = ux * ux + uy * uy + uz * uz;
            double M[9] = {
                1 + 2 * (ux * ux - uu), 2 * (ux * uy - w * uz), 2 * (ux * uz + w * uy),
                2 * (ux * uy + w * uz), 1 + 2 * (uy * uy - uu), 2 * (uy * uz - w * ux),
                2 * (ux * uz - w * uy), 2 * (uy * uz + w * ux), 1 + 2 * (uz * uz - uu)};
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 3; ++c)
                    e.JX[r * 3 + c] = A[r * 3] * M[c] + A[r * 3 + 1] * M[3 + c] + A[r * 3 + 2] * M[6 + c];
            // d pc / d u = -2w[v]x + 2((u.v) I + u v^T - 2 v u^T);  d pc / d w = 2 (u x v)
            const double vx = X[0], vy = X[1], vz = X[2];
            const double udv = ux * vx + uy * vy + uz * vz;
            double G[12];  // 3x4, columns (x,y,z,w)
            const double u[3] = {ux, uy, uz}, v[3] = {vx, vy, vz};
            const double vxm[9] = {0, -vz, vy, vz, 0, -vx, -vy, vx, 0};
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c)
                    G[r * 4 + c] = -2 * w * vxm[r * 3 + c] +
                                   2 * ((r == c ? udv : 0.0) + u[r] * v[c] - 2 * v[r] * u[c]);
            G[3] = cx_, G[7] = cy_, G[11] = cz_;
            double Jq[8];  // 2x4
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 4; ++c)
                    Jq[r * 4 + c] = A[r * 3] * G[c] + A[r * 3 + 1] * G[4 + c] + A[r * 3 + 2] * G[8 + c];
            // plus-Jacobian 4x3 (row-major) for x = (x,y,z,w):
            //   [ w, z,-y; -z, w, x;  y,-x, w; -x,-y,-z ]
            const double Pj[12] = {w, uz, -uy, -uz, w, ux, uy, -ux, w, -ux, -uy, -uz};
            for (int r = 0; r < 2; ++r)
                for (int c = 0; c < 3; ++c)
                    e.Jd[r * 3 + c] = Jq[r * 4] * Pj[c] + Jq[r * 4 + 1] * Pj[3 + c] +
                                      Jq[r * 4 + 2] * Pj[6 + c] + Jq[r * 4 + 3] * Pj[9 + c];
        }
    }
    // HuberLoss(a) + Corrector: rho2 <= 0 always -> plain sqrt(rho1) scaling
    const double s = e.r[0] * e.r[0] + e.r[1] * e.r[1];
    const double b = o.huber_a * o.huber_a;
    double rho1 = 1.0;
    if (s > b) {
        const double rt = std::sqrt(s);
        e.rho0 = 2.0 * o.huber_a * rt - b;
        rho1 = std::max(DBL_MIN, o.huber_a / rt);
    } else {
        e.rho0 = s;
    }
    if (robustify && rho1 != 1.0) {
        const double sc = std::sqrt(rho1);
        e.r[0] *= sc, e.r[1] *= sc;
        if (want_jac)
            for (int i = 0; i < 6; ++i) e.Jd[i] *= sc, e.Jt[i] *= sc, e.JX[i] *= sc;
    }
}

/* EigenQuaternionParameterization::Plus: x+ = dq (x) x, dq = (sin|d|/|d| d, cos|d|). */
inline void quat_plus(const double q[4], const double d[3], double out[4]) {
    const double n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (n > 0.0) {
        const double sbd = std::sin(n) / n;
        const double ax = sbd * d[0], ay = sbd * d[1], az = sbd * d[2], aw = std::cos(n);
        const double bx = q[0], by = q[1], bz = q[2], bw = q[3];
        out[0] = aw * bx + ax * bw + ay * bz - az * by;
        out[1] = aw * by - ax * bz + ay * bw + az * bx;
        out[2] = aw * bz + ax * by - ay * bx + az * bw;
        out[3] = aw * bw - ax * bx - ay * by - az * bz;
    } else {
        for (int i = 0; i < 4; ++i) out[i] = q[i];
    }
}

// ----------------------------------------------------------------------------------------
// Dense / banded Cholesky of the reduced camera system (lower triangle, row-major n x n)
// ----------------------------------------------------------------------------------------
constexpr int NB = 48;

/* In-place blocked Cholesky restricted to a half-bandwidth `bw` (in scalars; bw >= n means
 * dense).  Fill-in of a banded SPD matrix stays inside the band.  Returns false on a
 * non-positive pivot. */
bool band_cholesky(double *S, int n, int bw) {
    for (int k0 = 0; k0 < n; k0 += NB) {
        const int kb = std::min(NB, n - k0);
        // diagonal block
        for (int j = k0; j < k0 + kb; ++j) {
            double d = S[(size_t)j * n + j];
            for (int p = k0; p < j; ++p) d -= S[(size_t)j * n + p] * S[(size_t)j * n + p];
            if (!(d > 0.0) || !std::isfinite(d)) return false;
            d = std::sqrt(d);
            S[(size_t)j * n + j] = d;
            const double inv = 1.0 / d;
            for (int i = j + 1; i < k0 + kb; ++i) {
                double v = S[(size_t)i * n + j];
                for (int p = k0; p < j; ++p) v -= S[(size_t)i * n + p] * S[(size_t)j * n + p];
                S[(size_t)i * n + j] = v * inv;
            }
        }
        const int iend = std::min(n, k0 + kb + bw);
        // panel: rows below solve X * Lkk^T = A
#pragma omp parallel for schedule(dynamic, 16)
        for (int i = k0 + kb; i < iend; ++i) {
            double *ri = S + (size_t)i * n;
            for (int j = k0; j < k0 + kb; ++j) {
                const double *rj = S + (size_t)j * n;
                double v = ri[j];
                for (int p = k0; p < j; ++p) v -= ri[p] * rj[p];
                ri[j] = v / rj[j];
            }
        }
        // trailing update: A[i][j] -= sum_p L[i][p] L[j][p], lower part inside the band
#pragma omp parallel for schedule(dynamic, 8)
        for (int i = k0 + kb; i < iend; ++i) {
            const double *li = S + (size_t)i * n + k0;
            double *ri = S + (size_t)i * n;
            for (int j = k0 + kb; j <= i; ++j) {
                const double *lj = S + (size_t)j * n + k0;
                double acc = 0.0;
                for (int p = 0; p < kb; ++p) acc += li[p] * lj[p];
                ri[j] -= acc;
            }
        }
    }
    return true;
}

void chol_solve(const double *L, int n, int bw, double *x) {
    for (int i = 0; i < n; ++i) {  // L y = b
        double v = x[i];
        const double *ri = L + (size_t)i * n;
        for (int p = std::max(0, i - bw - NB); p < i; ++p) v -= ri[p] * x[p];
        x[i] = v / ri[i];
    }
    for (int i = n - 1; i >= 0; --i) {  // L^T x = y
        double v = x[i];
        const int pend = std::min(n, i + bw + NB + 1);
        for (int p = i + 1; p < pend; ++p) v -= L[(size_t)p * n + i] * x[p];
        x[i] = v / L[(size_t)i * n + i];
    }
}

inline bool invert_sym3(const double V[6], double Vi[6]) {
    // V = [a b c; b d e; c e f] stored (a,b,c,d,e,f)
    const double a = V[0], b = V[1], c = V[2], d = V[3], e = V[4], f = V[5];
    const double A = d * f - e * e, B = c * e - b * f, Cc = b * e - c * d;
    const double det = a * A + b * B + c * Cc;
    if (!(std::fabs(det) > 0.0) || !std::isfinite(det)) return false;
    const double id = 1.0 / det;
    Vi[0] = A * id, Vi[1] = B * id, Vi[2] = Cc * id;
    Vi[3] = (a * f - c * c) * id, Vi[4] = (b * c - a * e) * id, Vi[5] = (a * d - b * b) * id;
    return true;
}

// ----------------------------------------------------------------------------------------
// The minimizer
// ----------------------------------------------------------------------------------------
struct Solver {
    const xrb_ba_problem &P;
    xrb_ba_options O;
    int C, NP, NO;
    std::vector<int> pt_ptr, pt_obs;  // CSR by point -> caller observation index
    std::vector<int> colq, colt;      // camera -> first reduced column of its q / t block, -1 fixed
    std::vector<uint8_t> pt_var;      // point is a variable of the reduced program
    std::vector<uint8_t> obs_active;  // residual block has at least one variable block
    int nc = 0;                       // reduced camera-system dimension
    int bw = 0;                       // half bandwidth of S in scalars
    int n_var_pts = 0, n_var_q = 0, n_var_t = 0, n_res_blocks = 0;
    double fixed_cost = 0.0;

    std::vector<double> q, t, X, cq, ct, cX;  // state and candidate
    std::vector<double> sc, sp;               // Jacobi scaling (camera cols, point cols)
    std::vector<double> S, rhs, yc;           // reduced system
    std::vector<double> Vinv, gp;             // per point: V^-1 (6) and g_p (3)
    std::vector<double> step_c, step_p;       // trust-region step (scaled space)
    std::vector<double> gc_unscaled;          // gradient wrt camera tangent (unscaled J)
    double radius, decrease_factor = 2.0;

    explicit Solver(const xrb_ba_problem &p, const xrb_ba_options &o)
        : P(p), O(o), C(p.n_cams), NP(p.n_pts), NO(p.n_obs) {}

    const double *intr_of(int cam) const { return P.intr + 8 * (size_t)P.cam_intr[cam]; }
    int model_of(int cam) const { return P.intr_model[P.cam_intr[cam]]; }

    void setup() {
        q.assign(P.cam_q, P.cam_q + 4 * (size_t)C);
        t.assign(P.cam_t, P.cam_t + 3 * (size_t)C);
        X.assign(P.pts, P.pts + 3 * (size_t)NP);
        // CSR by point (stable in caller order)
        pt_ptr.assign(NP + 1, 0);
        for (int o = 0; o < NO; ++o) pt_ptr[P.obs_pt[o] + 1]++;
        for (int p = 0; p < NP; ++p) pt_ptr[p + 1] += pt_ptr[p];
        pt_obs.resize(NO);
        {
            std::vector<int> cur(pt_ptr.begin(), pt_ptr.end() - 1);
            for (int o = 0; o < NO; ++o) pt_obs[cur[P.obs_pt[o]]++] = o;
        }
        // variable blocks = not constant AND referenced by at least one residual block
        std::vector<int> cam_obs(C, 0);
        for (int o = 0; o < NO; ++o) cam_obs[P.obs_cam[o]]++;
        colq.assign(C, -1), colt.assign(C, -1);
        nc = 0;
        for (int c = 0; c < C; ++c) {
            if (!cam_obs[c]) continue;
            if (!(P.cam_q_fixed && P.cam_q_fixed[c])) colq[c] = nc, nc += 3, n_var_q++;
            if (!(P.cam_t_fixed && P.cam_t_fixed[c])) colt[c] = nc, nc += 3, n_var_t++;
        }
        pt_var.assign(NP, 0);
        for (int p = 0; p < NP; ++p)
            if (pt_ptr[p + 1] > pt_ptr[p] && !(P.pt_fixed && P.pt_fixed[p])) pt_var[p] = 1, n_var_pts++;
        obs_active.assign(NO, 0);
        for (int o = 0; o < NO; ++o) {
            const int c = P.obs_cam[o];
            obs_active[o] = colq[c] >= 0 || colt[c] >= 0 || pt_var[P.obs_pt[o]];
            n_res_blocks += obs_active[o];
        }
        // half bandwidth of S (scalars): max column distance between co-observing cameras
        bw = 0;
        for (int p = 0; p < NP; ++p) {
            if (!pt_var[p]) continue;
            int lo = INT32_MAX, hi = -1;
            for (int k = pt_ptr[p]; k < pt_ptr[p + 1]; ++k) {
                const int c = P.obs_cam[pt_obs[k]];
                for (int col : {colq[c], colt[c]})
                    if (col >= 0) lo = std::min(lo, col), hi = std::max(hi, col + 2);
            }
            if (hi >= 0) bw = std::max(bw, hi - lo);
        }
        bw = std::min(std::max(bw, 5), std::max(nc - 1, 0));
        sc.assign(nc, 1.0), sp.assign(3 * (size_t)NP, 1.0);
        S.assign((size_t)nc * nc, 0.0);
        rhs.assign(nc, 0.0), yc.assign(nc, 0.0);
        Vinv.assign(6 * (size_t)NP, 0.0), gp.assign(3 * (size_t)NP, 0.0);
        step_c.assign(nc, 0.0), step_p.assign(3 * (size_t)NP, 0.0);
        gc_unscaled.assign(nc, 0.0);
        cq = q, ct = t, cX = X;
        // cost of residual blocks whose parameter blocks are all constant (Ceres fixed_cost)
        fixed_cost = 0.0;
        for (int o = 0; o < NO; ++o)
            if (!obs_active[o]) {
                ObsEval e;
                const int c = P.obs_cam[o], p = P.obs_pt[o];
                eval_obs(&q[4 * c], &t[3 * c], &X[3 * p], model_of(c), intr_of(c), P.obs_uv + 2 * o,
                         O, false, false, e);
                fixed_cost += 0.5 * e.rho0;
            }
    }

    /* cost = 1/2 sum rho over active residual blocks (deterministic: per-point partial sums) */
    double cost_at(const std::vector<double> &qq, const std::vector<double> &tt,
                   const std::vector<double> &XX) const {
        std::vector<double> part(NP, 0.0);
#pragma omp parallel for schedule(static, 256)
        for (int p = 0; p < NP; ++p) {
            double s = 0.0;
            for (int k = pt_ptr[p]; k < pt_ptr[p + 1]; ++k) {
                const int o = pt_obs[k];
                if (!obs_active[o]) continue;
                const int c = P.obs_cam[o];
                ObsEval e;
                eval_obs(&qq[4 * c], &tt[3 * c], &XX[3 * p], model_of(c), intr_of(c),
                         P.obs_uv + 2 * o, O, false, false, e);
                s += 0.5 * e.rho0;
            }
            part[p] = s;
        }
        double total = 0.0;
        for (int p = 0; p < NP; ++p) total += part[p];
        return total;
    }

    /* scaled, robustified per-observation blocks at the current state */
    inline void lin_obs(int o, ObsEval &e, int &c, double Jc[12], int cols[6]) const {
        c = P.obs_cam[o];
        const int p = P.obs_pt[o];
        eval_obs(&q[4 * c], &t[3 * c], &X[3 * p], model_of(c), intr_of(c), P.obs_uv + 2 * o, O, true,
                 true, e);
        for (int k = 0; k < 3; ++k) {
            cols[k] = colq[c] >= 0 ? colq[c] + k : -1;
            cols[3 + k] = colt[c] >= 0 ? colt[c] + k : -1;
        }
        for (int r = 0; r < 2; ++r)
            for (int k = 0; k < 3; ++k) {
                Jc[r * 6 + k] = cols[k] >= 0 ? e.Jd[r * 3 + k] * sc[cols[k]] : 0.0;
                Jc[r * 6 + 3 + k] = cols[3 + k] >= 0 ? e.Jt[r * 3 + k] * sc[cols[3 + k]] : 0.0;
            }
        if (pt_var[p])
            for (int r = 0; r < 2; ++r)
                for (int k = 0; k < 3; ++k) e.JX[r * 3 + k] *= sp[3 * (size_t)p + k];
        else
            for (int i = 0; i < 6; ++i) e.JX[i] = 0.0;
    }

    /* Jacobi scaling: 1 / (1 + ||J_col||) from the robustified iteration-0 Jacobian. */
    void compute_jacobi_scaling() {
        std::vector<double> n2c(nc, 0.0);
        for (int p = 0; p < NP; ++p) {
            double n2p[3] = {0, 0, 0};
            for (int k = pt_ptr[p]; k < pt_ptr[p + 1]; ++k) {
                const int o = pt_obs[k];
                if (!obs_active[o]) continue;
                ObsEval e;
                int c, cols[6];
                double Jc[12];
                lin_obs(o, e, c, Jc, cols);  // sc == sp == 1 here
                for (int j = 0; j < 6; ++j)
                    if (cols[j] >= 0) n2c[cols[j]] += Jc[j] * Jc[j] + Jc[6 + j] * Jc[6 + j];
                for (int j = 0; j < 3; ++j) n2p[j] += e.JX[j] * e.JX[j] + e.JX[3 + j] * e.JX[3 + j];
            }
            if (pt_var[p])
                for (int j = 0; j < 3; ++j) sp[3 * (size_t)p + j] = 1.0 / (1.0 + std::sqrt(n2p[j]));
        }
        for (int j = 0; j < nc; ++j) sc[j] = 1.0 / (1.0 + std::sqrt(n2c[j]));
    }

    /* One LM linear solve: builds S, rhs by Schur elimination of the points, solves, back-
     * substitutes.  On return step_c/step_p hold the step (scaled space, sign applied) and
     * model_cost_change is filled.  Returns false on a linear-solver failure. */
    bool compute_step(double &model_cost_change, double &gradient_max_norm, bool want_gradient) {
        std::fill(S.begin(), S.end(), 0.0);
        std::fill(rhs.begin(), rhs.end(), 0.0);
        std::vector<double> U((size_t)nc * 6, 0.0);  // rows of the camera block-diagonal (6 wide)
        std::vector<double> gc(nc, 0.0);
        const int nlocks = 1024;
        std::vector<omp_lock_t> locks(nlocks);
        for (auto &l : locks) omp_init_lock(&l);
        const int n = nc;
        const double inv_radius = 1.0 / radius;
        bool ok = true;

#pragma omp parallel
        {
            std::vector<double> Jc, W, Tm;
            std::vector<int> colsv, camv;
            std::vector<double> rv;
#pragma omp for schedule(dynamic, 64)
            for (int p = 0; p < NP; ++p) {
                const int k0 = pt_ptr[p], kn = pt_ptr[p + 1] - k0;
                if (kn == 0) continue;
                Jc.assign((size_t)kn * 12, 0.0), W.assign((size_t)kn * 18, 0.0);
                Tm.assign((size_t)kn * 18, 0.0), colsv.assign((size_t)kn * 6, -1);
                camv.assign(kn, -1), rv.assign((size_t)kn * 2, 0.0);
                double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
                std::vector<double> JXs((size_t)kn * 6, 0.0);
                int nact = 0;
                for (int k = 0; k < kn; ++k) {
                    const int o = pt_obs[k0 + k];
                    if (!obs_active[o]) continue;
                    ObsEval e;
                    int c;
                    lin_obs(o, e, c, &Jc[(size_t)k * 12], &colsv[(size_t)k * 6]);
                    camv[k] = c, rv[2 * k] = e.r[0], rv[2 * k + 1] = e.r[1];
                    for (int i = 0; i < 6; ++i) JXs[(size_t)k * 6 + i] = e.JX[i];
                    const double *J = e.JX;
                    V[0] += J[0] * J[0] + J[3] * J[3], V[1] += J[0] * J[1] + J[3] * J[4];
                    V[2] += J[0] * J[2] + J[3] * J[5], V[3] += J[1] * J[1] + J[4] * J[4];
                    V[4] += J[1] * J[2] + J[4] * J[5], V[5] += J[2] * J[2] + J[5] * J[5];
                    for (int j = 0; j < 3; ++j) g[j] += J[j] * e.r[0] + J[3 + j] * e.r[1];
                    nact++;
                }
                if (!nact) continue;
                double Vi[6] = {0, 0, 0, 0, 0, 0};
                if (pt_var[p]) {
                    // LM diagonal on the point block: clamp(diag(J^T J), 1e-6, 1e32) / radius
                    const double d0 = std::min(std::max(V[0], 1e-6), 1e32) * inv_radius;
                    const double d1 = std::min(std::max(V[3], 1e-6), 1e32) * inv_radius;
                    const double d2 = std::min(std::max(V[5], 1e-6), 1e32) * inv_radius;
                    double Vd[6] = {V[0] + d0, V[1], V[2], V[3] + d1, V[4], V[5] + d2};
                    if (!invert_sym3(Vd, Vi)) {
#pragma omp atomic write
                        ok = false;
                        continue;
                    }
                    for (int i = 0; i < 6; ++i) Vinv[6 * (size_t)p + i] = Vi[i];
                    for (int j = 0; j < 3; ++j) gp[3 * (size_t)p + j] = g[j];
                }
                const double Vf[9] = {Vi[0], Vi[1], Vi[2], Vi[1], Vi[3], Vi[4], Vi[2], Vi[4], Vi[5]};
                // per observation: U_c += Jc^T Jc, g_c += Jc^T r, W = Jc^T JX, T = W V^-1
                for (int k = 0; k < kn; ++k) {
                    if (camv[k] < 0) continue;
                    const double *J = &Jc[(size_t)k * 12];
                    const int *cols = &colsv[(size_t)k * 6];
                    const double *JX = &JXs[(size_t)k * 6];
                    double *Wk = &W[(size_t)k * 18], *Tk = &Tm[(size_t)k * 18];
                    for (int a = 0; a < 6; ++a)
                        for (int b = 0; b < 3; ++b) Wk[a * 3 + b] = J[a] * JX[b] + J[6 + a] * JX[3 + b];
                    for (int a = 0; a < 6; ++a)
                        for (int b = 0; b < 3; ++b)
                            Tk[a * 3 + b] = Wk[a * 3] * Vf[b] + Wk[a * 3 + 1] * Vf[3 + b] + Wk[a * 3 + 2] * Vf[6 + b];
                    omp_lock_t *lk = &locks[(size_t)camv[k] % nlocks];
                    omp_set_lock(lk);
                    for (int a = 0; a < 6; ++a) {
                        if (cols[a] < 0) continue;
                        const int base = cols[a] - (a < 3 ? a : a - 3);  // first col of this 3-block
                        (void)base;
                        for (int b = 0; b < 6; ++b)
                            U[(size_t)cols[a] * 6 + b] += J[a] * J[b] + J[6 + a] * J[6 + b];
                        const double gr = J[a] * rv[2 * k] + J[6 + a] * rv[2 * k + 1];
                        gc[cols[a]] += gr;
                        rhs[cols[a]] += gr - (Tk[a * 3] * g[0] + Tk[a * 3 + 1] * g[1] + Tk[a * 3 + 2] * g[2]);
                    }
                    omp_unset_lock(lk);
                }
                if (!pt_var[p]) continue;
                // S[c_i, c_j] -= T_i W_j^T  (lower triangle of the scalar matrix only)
                for (int i = 0; i < kn; ++i) {
                    if (camv[i] < 0) continue;
                    for (int j = 0; j < kn; ++j) {
                        if (camv[j] < 0) continue;
                        const int *ci = &colsv[(size_t)i * 6], *cj = &colsv[(size_t)j * 6];
                        const double *Ti = &Tm[(size_t)i * 18], *Wj = &W[(size_t)j * 18];
                        omp_lock_t *lk = &locks[((size_t)camv[i] * 131 + camv[j]) % nlocks];
                        omp_set_lock(lk);
                        for (int a = 0; a < 6; ++a) {
                            if (ci[a] < 0) continue;
                            for (int b = 0; b < 6; ++b) {
                                if (cj[b] < 0 || cj[b] > ci[a]) continue;
                                S[(size_t)ci[a] * n + cj[b]] -=
                                    Ti[a * 3] * Wj[b * 3] + Ti[a * 3 + 1] * Wj[b * 3 + 1] + Ti[a * 3 + 2] * Wj[b * 3 + 2];
                            }
                        }
                        omp_unset_lock(lk);
                    }
                }
            }
        }
        for (auto &l : locks) omp_destroy_lock(&l);
        if (!ok) return false;
        // camera block-diagonal: S += U + D_c^2, D_c^2 = clamp(diag U)/radius
        for (int c = 0; c < C; ++c) {
            int cols[6];
            for (int k = 0; k < 3; ++k) cols[k] = colq[c] >= 0 ? colq[c] + k : -1, cols[3 + k] = colt[c] >= 0 ? colt[c] + k : -1;
            for (int a = 0; a < 6; ++a) {
                if (cols[a] < 0) continue;
                for (int b = 0; b < 6; ++b) {
                    if (cols[b] < 0 || cols[b] > cols[a]) continue;
                    S[(size_t)cols[a] * n + cols[b]] += U[(size_t)cols[a] * 6 + b];
                }
                const double dg = std::min(std::max(U[(size_t)cols[a] * 6 + a], 1e-6), 1e32);
                S[(size_t)cols[a] * n + cols[a]] += dg * inv_radius;
            }
        }
        if (want_gradient) {
            // gradient of the unscaled problem: g = J^T r = g_scaled / scale
            double gmax = 0.0;
            for (int c = 0; c < C; ++c) {
                if (colq[c] >= 0) {
                    double d[3], qn[4];
                    for (int k = 0; k < 3; ++k) d[k] = -gc[colq[c] + k] / sc[colq[c] + k];
                    quat_plus(&q[4 * c], d, qn);
                    for (int k = 0; k < 4; ++k) gmax = std::max(gmax, std::fabs(q[4 * c + k] - qn[k]));
                }
                if (colt[c] >= 0)
                    for (int k = 0; k < 3; ++k) gmax = std::max(gmax, std::fabs(gc[colt[c] + k] / sc[colt[c] + k]));
            }
            for (int p = 0; p < NP; ++p)
                if (pt_var[p])
                    for (int k = 0; k < 3; ++k)
                        gmax = std::max(gmax, std::fabs(gp[3 * (size_t)p + k] / sp[3 * (size_t)p + k]));
            gradient_max_norm = gmax;
        }
        // reduced system
        yc = rhs;
        if (nc > 0) {
            if (!band_cholesky(S.data(), nc, bw)) return false;
            chol_solve(S.data(), nc, bw, yc.data());
        }
        for (int j = 0; j < nc; ++j)
            if (!std::isfinite(yc[j])) return false;
        // back-substitution + model cost change, per point
        std::vector<double> mpart(NP, 0.0);
        bool finite = true;
#pragma omp parallel for schedule(dynamic, 64)
        for (int p = 0; p < NP; ++p) {
            const int k0 = pt_ptr[p], kn = pt_ptr[p + 1] - k0;
            double yp[3] = {0, 0, 0};
            std::vector<double> Jcs((size_t)kn * 12), JXs((size_t)kn * 6), rs((size_t)kn * 2);
            std::vector<int> cl((size_t)kn * 6);
            std::vector<uint8_t> act(kn, 0);
            double acc[3] = {0, 0, 0};
            for (int k = 0; k < kn; ++k) {
                const int o = pt_obs[k0 + k];
                if (!obs_active[o]) continue;
                ObsEval e;
                int c;
                lin_obs(o, e, c, &Jcs[(size_t)k * 12], &cl[(size_t)k * 6]);
                act[k] = 1;
                for (int i = 0; i < 6; ++i) JXs[(size_t)k * 6 + i] = e.JX[i];
                rs[2 * k] = e.r[0], rs[2 * k + 1] = e.r[1];
                // W^T y_c = JX^T (Jc y_c)
                double jy0 = 0, jy1 = 0;
                for (int a = 0; a < 6; ++a)
                    if (cl[(size_t)k * 6 + a] >= 0) {
                        jy0 += Jcs[(size_t)k * 12 + a] * yc[cl[(size_t)k * 6 + a]];
                        jy1 += Jcs[(size_t)k * 12 + 6 + a] * yc[cl[(size_t)k * 6 + a]];
                    }
                for (int j = 0; j < 3; ++j) acc[j] += e.JX[j] * jy0 + e.JX[3 + j] * jy1;
            }
            if (pt_var[p]) {
                const double *Vi = &Vinv[6 * (size_t)p];
                const double b0 = gp[3 * (size_t)p] - acc[0], b1 = gp[3 * (size_t)p + 1] - acc[1],
                             b2 = gp[3 * (size_t)p + 2] - acc[2];
                yp[0] = Vi[0] * b0 + Vi[1] * b1 + Vi[2] * b2;
                yp[1] = Vi[1] * b0 + Vi[3] * b1 + Vi[4] * b2;
                yp[2] = Vi[2] * b0 + Vi[4] * b1 + Vi[5] * b2;
                for (int j = 0; j < 3; ++j) {
                    step_p[3 * (size_t)p + j] = -yp[j];
                    if (!std::isfinite(yp[j])) {
#pragma omp atomic write
                        finite = false;
                    }
                }
            }
            // model residual m = J * step ; contribution -(m . (r + m/2))
            double s = 0.0;
            for (int k = 0; k < kn; ++k) {
                if (!act[k]) continue;
                double m0 = 0, m1 = 0;
                for (int a = 0; a < 6; ++a)
                    if (cl[(size_t)k * 6 + a] >= 0) {
                        m0 -= Jcs[(size_t)k * 12 + a] * yc[cl[(size_t)k * 6 + a]];
                        m1 -= Jcs[(size_t)k * 12 + 6 + a] * yc[cl[(size_t)k * 6 + a]];
                    }
                for (int j = 0; j < 3; ++j) {
                    m0 -= JXs[(size_t)k * 6 + j] * yp[j];
                    m1 -= JXs[(size_t)k * 6 + 3 + j] * yp[j];
                }
                s -= m0 * (rs[2 * k] + 0.5 * m0) + m1 * (rs[2 * k + 1] + 0.5 * m1);
            }
            mpart[p] = s;
        }
        if (!finite) return false;
        for (int j = 0; j < nc; ++j) step_c[j] = -yc[j];
        model_cost_change = 0.0;
        for (int p = 0; p < NP; ++p) model_cost_change += mpart[p];
        return true;
    }

    /* candidate = Plus(x, step * scale); returns ||x - candidate|| (ambient) */
    double make_candidate() {
        double n2 = 0.0;
        cq = q, ct = t, cX = X;
        for (int c = 0; c < C; ++c) {
            if (colq[c] >= 0) {
                double d[3];
                for (int k = 0; k < 3; ++k) d[k] = step_c[colq[c] + k] * sc[colq[c] + k];
                quat_plus(&q[4 * c], d, &cq[4 * c]);
                for (int k = 0; k < 4; ++k) n2 += (q[4 * c + k] - cq[4 * c + k]) * (q[4 * c + k] - cq[4 * c + k]);
            }
            if (colt[c] >= 0)
                for (int k = 0; k < 3; ++k) {
                    const double d = step_c[colt[c] + k] * sc[colt[c] + k];
                    ct[3 * c + k] = t[3 * c + k] + d;
                    n2 += (t[3 * c + k] - ct[3 * c + k]) * (t[3 * c + k] - ct[3 * c + k]);
                }
        }
        for (int p = 0; p < NP; ++p)
            if (pt_var[p])
                for (int k = 0; k < 3; ++k) {
                    const double d = step_p[3 * (size_t)p + k] * sp[3 * (size_t)p + k];
                    cX[3 * (size_t)p + k] = X[3 * (size_t)p + k] + d;
                    const double df = X[3 * (size_t)p + k] - cX[3 * (size_t)p + k];
                    n2 += df * df;
                }
        return std::sqrt(n2);
    }

    double x_norm() const {
        double n2 = 0.0;
        for (int c = 0; c < C; ++c) {
            if (colq[c] >= 0)
                for (int k = 0; k < 4; ++k) n2 += q[4 * c + k] * q[4 * c + k];
            if (colt[c] >= 0)
                for (int k = 0; k < 3; ++k) n2 += t[3 * c + k] * t[3 * c + k];
        }
        for (int p = 0; p < NP; ++p)
            if (pt_var[p])
                for (int k = 0; k < 3; ++k) n2 += X[3 * (size_t)p + k] * X[3 * (size_t)p + k];
        return std::sqrt(n2);
    }

    /* TrustRegionMinimizer::Minimize with LevenbergMarquardtStrategy (monotonic steps). */
    void minimize(xrb_ba_summary &sum) {
        const auto t0 = std::chrono::steady_clock::now();
        memset(&sum, 0, sizeof sum);
        setup();
        sum.num_residuals_reduced = 2 * n_res_blocks;
        sum.num_effective_parameters_reduced = 3 * (n_var_q + n_var_t + n_var_pts);
        sum.fixed_cost = fixed_cost;
        radius = O.initial_radius;
        decrease_factor = 2.0;
        auto log_iter = [&](const xrb_ba_iteration &it) {
            if (sum.n_iterations_logged < 128) sum.iterations[sum.n_iterations_logged++] = it;
            if (it.step_is_successful) sum.num_successful_steps++; else sum.num_unsuccessful_steps++;
            if (O.verbose)
                printf("%4d % .6e % .2e % .2e % .2e % .2e % .2e\n", it.iteration, it.cost, it.cost_change,
                       it.gradient_max_norm, it.step_norm, it.relative_decrease, it.trust_region_radius);
        };
        sum.termination_type = XRB_BA_NO_CONVERGENCE;
        if (n_res_blocks == 0 || sum.num_effective_parameters_reduced == 0) {
            sum.initial_cost = sum.final_cost = fixed_cost;
            sum.termination_type = XRB_BA_CONVERGENCE;
            return;
        }
        // ---- IterationZero
        double x_cost = cost_at(q, t, X);
        compute_jacobi_scaling();
        double xnorm = x_norm();
        xrb_ba_iteration it;
        memset(&it, 0, sizeof it);
        it.iteration = 0, it.step_is_valid = 1, it.step_is_successful = 1;
        it.cost = x_cost + fixed_cost;
        it.trust_region_radius = radius;
        sum.initial_cost = it.cost;
        double last_grad = 0.0;
        bool have_step = false, step_ok = false;
        double model_cost_change = 0.0;
        int consecutive_invalid = 0;
        // the linear solve for iteration k also yields the gradient at the current point, so
        // iteration 0's gradient norm comes from the first solve below.
        step_ok = compute_step(model_cost_change, last_grad, true);
        have_step = true;
        it.gradient_max_norm = last_grad;
        log_iter(it);
        int iteration = 0;
        for (;;) {
            // ---- FinalizeIterationAndCheckIfMinimizerCanContinue (guards)
            if (iteration >= O.max_iterations) { sum.termination_type = XRB_BA_NO_CONVERGENCE; break; }
            if (!O.fixed_iterations && it.gradient_max_norm <= O.gradient_tolerance) { sum.termination_type = XRB_BA_CONVERGENCE; break; }
            if (radius <= 1e-32) { sum.termination_type = XRB_BA_CONVERGENCE; break; }
            iteration++;
            xrb_ba_iteration cur;
            memset(&cur, 0, sizeof cur);
            cur.iteration = iteration;
            // ---- ComputeTrustRegionStep
            if (!have_step) step_ok = compute_step(model_cost_change, last_grad, false);
            have_step = false;
            sum.num_lm_iterations++;
            cur.model_cost_change = model_cost_change;
            cur.step_is_valid = step_ok && model_cost_change > 0.0;
            if (!cur.step_is_valid) {  // HandleInvalidStep
                if (++consecutive_invalid >= 5) {
                    sum.termination_type = XRB_BA_FAILURE;
                    break;
                }
                radius *= 0.5;  // StepIsInvalid
                cur.cost = x_cost + fixed_cost;
                cur.gradient_max_norm = it.gradient_max_norm;
                cur.trust_region_radius = radius;
                it = cur;
                log_iter(it);
                continue;
            }
            consecutive_invalid = 0;
            // ---- ComputeCandidatePointAndEvaluateCost
            cur.step_norm = make_candidate();
            const double cand_cost = cost_at(cq, ct, cX);
            // ---- ParameterToleranceReached / FunctionToleranceReached (before accept/reject)
            if (!O.fixed_iterations && cur.step_norm <= O.parameter_tolerance * (xnorm + O.parameter_tolerance)) {
                sum.termination_type = XRB_BA_CONVERGENCE;
                break;
            }
            cur.cost_change = x_cost - cand_cost;
            if (!O.fixed_iterations && std::fabs(cur.cost_change) <= O.function_tolerance * x_cost) {
                sum.termination_type = XRB_BA_CONVERGENCE;
                break;
            }
            cur.relative_decrease = cur.cost_change / model_cost_change;
            if (cur.relative_decrease > 1e-3) {  // HandleSuccessfulStep
                q.swap(cq), t.swap(ct), X.swap(cX);
                xnorm = x_norm();
                x_cost = cand_cost;
                cur.step_is_successful = 1;
                // StepAccepted
                radius = radius / std::max(1.0 / 3.0, 1.0 - std::pow(2.0 * cur.relative_decrease - 1.0, 3));
                radius = std::min(1e16, radius);
                decrease_factor = 2.0;
                // EvaluateGradientAndJacobian at the new point: folded into the next solve
                step_ok = compute_step(model_cost_change, last_grad, true);
                have_step = true;
                cur.gradient_max_norm = last_grad;
            } else {  // StepRejected
                radius = radius / decrease_factor;
                decrease_factor *= 2.0;
                cur.gradient_max_norm = it.gradient_max_norm;
            }
            cur.cost = cand_cost + fixed_cost;
            cur.trust_region_radius = radius;
            it = cur;
            log_iter(it);
        }
        // SetSummaryFinalCost: min over logged iteration costs
        sum.final_cost = sum.initial_cost;
        for (int i = 0; i < sum.n_iterations_logged; ++i)
            sum.final_cost = std::min(sum.final_cost, sum.iterations[i].cost);
        sum.total_time_in_seconds =
            std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};

}  // namespace

extern "C" {

void xro_ba_default_options(xrb_ba_options *o) {
    o->max_iterations = 50;  // ceres default (InitSolverOptions sets 100, callers override)
    o->function_tolerance = 1e-6;
    o->parameter_tolerance = 1e-8;
    o->gradient_tolerance = 1e-10;
    o->initial_radius = 1e4;
    o->huber_a = 5.99;
    o->min_depth = 1e-2;
    o->neg_depth_residual = 12.0;
    o->verbose = 0;
    o->fixed_iterations = 0;
}

/* ceres::Solve for the reprojection problem; poses/points are updated in place. */
int xro_ba_solve(xrb_ba_problem *prob, const xrb_ba_options *opt, xrb_ba_summary *summary,
                 int n_threads) {
    if (n_threads > 0) omp_set_num_threads(n_threads);
    Solver s(*prob, *opt);
    s.minimize(*summary);
    memcpy(prob->cam_q, s.q.data(), sizeof(double) * 4 * (size_t)prob->n_cams);
    memcpy(prob->cam_t, s.t.data(), sizeof(double) * 3 * (size_t)prob->n_cams);
    memcpy(prob->pts, s.X.data(), sizeof(double) * 3 * (size_t)prob->n_pts);
    return 0;
}

/* Raw residuals (ReProjectionCost::operator()) in caller observation order. */
int xro_ba_residuals(const xrb_ba_problem *P, const xrb_ba_options *opt, double *out) {
#pragma omp parallel for schedule(static, 1024)
    for (int o = 0; o < P->n_obs; ++o) {
        const int c = P->obs_cam[o], p = P->obs_pt[o];
        ObsEval e;
        eval_obs(P->cam_q + 4 * (size_t)c, P->cam_t + 3 * (size_t)c, P->pts + 3 * (size_t)p,
                 P->intr_model[P->cam_intr[c]], P->intr + 8 * (size_t)P->cam_intr[c],
                 P->obs_uv + 2 * (size_t)o, *opt, false, false, e);
        out[2 * (size_t)o] = e.r[0], out[2 * (size_t)o + 1] = e.r[1];
    }
    return 0;
}

/* One observation: residual[2], Jd[6], Jt[6], JX[6], rho0 — raw (robustify=0) or corrected. */
int xro_ba_eval_obs(const double *q, const double *t, const double *X, int model,
                    const double *intr, const double *uv, const xrb_ba_options *opt,
                    int robustify, double *out21) {
    ObsEval e;
    eval_obs(q, t, X, model, intr, uv, *opt, true, robustify != 0, e);
    out21[0] = e.r[0], out21[1] = e.r[1];
    for (int i = 0; i < 6; ++i) out21[2 + i] = e.Jd[i], out21[8 + i] = e.Jt[i], out21[14 + i] = e.JX[i];
    out21[20] = e.rho0;
    return e.depth_branch;
}

void xro_quat_plus(const double *q, const double *d, double *out) { quat_plus(q, d, out); }

/* cost = 1/2 sum rho at the problem's current state (all observations) */
double xro_ba_cost(const xrb_ba_problem *P, const xrb_ba_options *opt) {
    double total = 0.0;
    for (int o = 0; o < P->n_obs; ++o) {
        const int c = P->obs_cam[o], p = P->obs_pt[o];
        ObsEval e;
        eval_obs(P->cam_q + 4 * (size_t)c, P->cam_t + 3 * (size_t)c, P->pts + 3 * (size_t)p,
                 P->intr_model[P->cam_intr[c]], P->intr + 8 * (size_t)P->cam_intr[c],
                 P->obs_uv + 2 * (size_t)o, *opt, false, false, e);
        total += 0.5 * e.rho0;
    }
    return total;
}
/* Post-BA filtering (the checker of xrsfm_b200/csrc/ba_filter.cu) — restatement of
 * Point3dProcessor::FilterPoints3d / FilterPoint3d / UpdateTrackAngle
 * (src/geometry/track_processor.cc:253-349; Reprojection_Error :19-26; CalculateTriangulationAngle
 * src/geometry/colmap/base/triangulation.cc:124-147; Pose::center src/base/types.h:45) over the
 * same flat problem the BA takes.  A track's observations are visited in ascending camera index
 * (the reference iterates a std::map keyed by frame id: flatten frames in id order).
 *   keep_obs[n_obs]   0 where the observation is deleted (re > max_re or depth outside [1e-3, 1e3],
 *                     or the whole track became an outlier)
 *   pt_outlier[n_pts] 1 = SetTrackOutlier; pt_error = mean reprojection error of the kept
 *                     observations (only written when the track survives the first test);
 *                     pt_angle = track.angle_ as the early-exit scan leaves it
 *   counts[2]         num_filtered1 (observations), num_filtered2 (tracks by angle)
 * Points without observations are left alone (the reference never holds such tracks). */
int xro_filter_points3d(const xrb_ba_problem *P, double max_re, double deg, uint8_t *keep_obs, uint8_t *pt_outlier,
                        double *pt_error, double *pt_angle, int32_t *counts) {
    const double min_tri_angle_rad = deg * 0.0174532925199432954743716805978692718781530857086181640625;
    std::vector<std::vector<int>> obs_of(P->n_pts);
    for (int o = 0; o < P->n_obs; ++o) obs_of[P->obs_pt[o]].push_back(o);
    counts[0] = counts[1] = 0;
    for (int p = 0; p < P->n_pts; ++p) {
        std::vector<int> &obs = obs_of[p];
        std::stable_sort(obs.begin(), obs.end(), [&](int a, int b) { return P->obs_cam[a] < P->obs_cam[b]; });
        pt_outlier[p] = 0;
        if (obs.empty()) continue;
        const double *X = P->pts + 3 * (size_t)p;
        double re_sum = 0.0;
        std::vector<int> del;
        for (const int o : obs) {
            const int c = P->obs_cam[o];
            const double *q = P->cam_q + 4 * (size_t)c, *t = P->cam_t + 3 * (size_t)c;
            const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
            const double cx_ = 2 * (uy * X[2] - uz * X[1]), cy_ = 2 * (uz * X[0] - ux * X[2]), cz_ = 2 * (ux * X[1] - uy * X[0]);
            const double pcx = X[0] + w * cx_ + (uy * cz_ - uz * cy_) + t[0];
            const double pcy = X[1] + w * cy_ + (uz * cx_ - ux * cz_) + t[1];
            const double pcz = X[2] + w * cz_ + (ux * cy_ - uy * cx_) + t[2];
            double uv[2], D[4];
            world_to_image(P->intr_model[P->cam_intr[c]], P->intr + 8 * (size_t)P->cam_intr[c], pcx / pcz, pcy / pcz, uv, D);
            const double dx = uv[0] - P->obs_uv[2 * (size_t)o], dy = uv[1] - P->obs_uv[2 * (size_t)o + 1];
            const double re = std::sqrt(dx * dx + dy * dy);
            keep_obs[o] = 1;
            if (re > max_re || pcz < 1e-3 || pcz > 1e3)
                del.push_back(o);
            else
                re_sum += re;
        }
        if (del.size() >= obs.size() - 1) {  // track_processor.cc:300-303
            counts[0] += (int32_t)obs.size();
            pt_outlier[p] = 1;
            for (const int o : obs) keep_obs[o] = 0;
            continue;
        }
        counts[0] += (int32_t)del.size();
        for (const int o : del) keep_obs[o] = 0;
        const size_t n_keep = obs.size() - del.size();
        pt_error[p] = re_sum / (double)n_keep;
        // UpdateTrackAngle (:253-277): centres of the remaining observations, pairs (i, j > i) in order,
        // early exit at the first running maximum above the threshold
        std::vector<double> ctr;
        for (const int o : obs) {
            if (!keep_obs[o]) continue;
            const int c = P->obs_cam[o];
            const double *q = P->cam_q + 4 * (size_t)c, *t = P->cam_t + 3 * (size_t)c;
            // -(q^-1 * t): Eigen's inverse() = conjugate / squaredNorm
            const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
            const double ux = -q[0] / n2, uy = -q[1] / n2, uz = -q[2] / n2, w = q[3] / n2;
            const double cx_ = 2 * (uy * t[2] - uz * t[1]), cy_ = 2 * (uz * t[0] - ux * t[2]), cz_ = 2 * (ux * t[1] - uy * t[0]);
            ctr.push_back(-(t[0] + w * cx_ + (uy * cz_ - uz * cy_)));
            ctr.push_back(-(t[1] + w * cy_ + (uz * cx_ - ux * cz_)));
            ctr.push_back(-(t[2] + w * cz_ + (ux * cy_ - uy * cx_)));
        }
        const int nc = (int)(ctr.size() / 3);
        double max_angle = 0;
        bool done = false;
        for (int i = 0; i < nc && !done; ++i)
            for (int j = i + 1; j < nc; ++j) {
                const double *a = &ctr[3 * i], *b = &ctr[3 * j];
                const double base2 = (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
                const double r1 = (X[0] - a[0]) * (X[0] - a[0]) + (X[1] - a[1]) * (X[1] - a[1]) + (X[2] - a[2]) * (X[2] - a[2]);
                const double r2 = (X[0] - b[0]) * (X[0] - b[0]) + (X[1] - b[1]) * (X[1] - b[1]) + (X[2] - b[2]) * (X[2] - b[2]);
                const double den = 2.0 * std::sqrt(r1 * r2);
                double angle = 0.0;
                if (den != 0.0) {
                    const double ang = std::fabs(std::acos((r1 + r2 - base2) / den));
                    angle = std::min(ang, M_PI - ang);
                }
                if (angle > max_angle) {
                    max_angle = angle;
                    if (max_angle > min_tri_angle_rad) {
                        done = true;
                        break;
                    }
                }
            }
        pt_angle[p] = max_angle;
        if (max_angle < min_tri_angle_rad) {  // :343-346
            pt_outlier[p] = 1;
            counts[1] += 1;
            for (const int o : obs) keep_obs[o] = 0;
        }
    }
    return 0;
}
}
