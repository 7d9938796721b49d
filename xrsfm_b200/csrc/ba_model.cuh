// ba_model.cuh — the per-observation model shared by the bundle-adjustment kernels (ba_kernels.cu) and the
// batched pose refinement (pose_refine.cu): residual, closed-form Jacobians, Huber corrector, quaternion Plus.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>

#include "ba_kernels.cuh"

namespace xrb {

// =====================================================================================
// Per-observation model: ReProjectionCost::operator() (cost_factor_ceres.h:19-40) with the
// Jacobians Ceres' autodiff + EigenQuaternionParameterization produce, in closed form, and
// HuberLoss(a) + Corrector (rho'' < 0 always, so the correction is a plain sqrt(rho') scale).
// =====================================================================================

// uv = WorldToImage(params, xy) and D = d(uv)/d(xy) (row-major 2x2) for model ids 0..4
// (camera_model.hpp:93-210).  Ids 0/1 keep the reference quirk Distortion() = xy => 2 f x + c.
__device__ __forceinline__ void world_to_image(int model, const double *__restrict__ p, double x,
                                               double y, double &u, double &v, double D[4]) {
    if (model == 2 || model == 3) {
        const double fx = p[0], fy = model == 2 ? p[0] : p[1];
        const double cx = model == 2 ? p[1] : p[2], cy = model == 2 ? p[2] : p[3];
        const double k = model == 2 ? p[3] : p[4];
        const double r2 = x * x + y * y, radial = k * r2;
        u = fx * (x + x * radial) + cx;
        v = fy * (y + y * radial) + cy;
        const double kxy2 = 2.0 * k * x * y;
        D[0] = fx * (1.0 + radial + 2.0 * k * x * x), D[1] = fx * kxy2;
        D[2] = fy * kxy2, D[3] = fy * (1.0 + radial + 2.0 * k * y * y);
    } else if (model == 0) {
        u = p[0] * (x + x) + p[1], v = p[0] * (y + y) + p[2];
        D[0] = 2.0 * p[0], D[1] = 0.0, D[2] = 0.0, D[3] = 2.0 * p[0];
    } else if (model == 1) {
        u = p[0] * (x + x) + p[2], v = p[1] * (y + y) + p[3];
        D[0] = 2.0 * p[0], D[1] = 0.0, D[2] = 0.0, D[3] = 2.0 * p[1];
    } else {  // OpenCV: fx fy cx cy k1 k2 p1 p2
        const double fx = p[0], fy = p[1], k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
        const double x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
        const double radial = k1 * r2 + k2 * r2 * r2;
        const double du = x * radial + 2.0 * p1 * xy + p2 * (r2 + 2.0 * x2);
        const double dv = y * radial + 2.0 * p2 * xy + p1 * (r2 + 2.0 * y2);
        u = fx * (x + du) + p[2], v = fy * (y + dv) + p[3];
        const double dr = k1 + 2.0 * k2 * r2;
        const double ddu_dx = radial + 2.0 * x2 * dr + 2.0 * p1 * y + 6.0 * p2 * x;
        const double ddu_dy = 2.0 * xy * dr + 2.0 * p1 * x + 2.0 * p2 * y;
        const double ddv_dx = 2.0 * xy * dr + 2.0 * p2 * y + 2.0 * p1 * x;
        const double ddv_dy = radial + 2.0 * y2 * dr + 2.0 * p2 * x + 6.0 * p1 * y;
        D[0] = fx * (1.0 + ddu_dx), D[1] = fx * ddu_dy, D[2] = fy * ddv_dx, D[3] = fy * (1.0 + ddv_dy);
    }
}

struct Obs {
    double r0, r1;   // residual (robustified when requested)
    double Jd[6];    // 2x3 d r / d(quaternion tangent)
    double Jt[6];    // 2x3 d r / d t
    double JX[6];    // 2x3 d r / d X
    double rho0;     // rho(s)
};

template <bool kJac>
__device__ __forceinline__ void eval_obs(const double *__restrict__ q, const double *__restrict__ t,
                                         const double *__restrict__ X, int model,
                                         const double *__restrict__ intr, double um, double vm,
                                         const BAConsts &k, bool robustify, Obs &e) {
    const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
    const double X0 = X[0], X1 = X[1], X2 = X[2];
    // Eigen Quaternion::_transformVector: uv = 2 (u x v); pc = v + w uv + u x uv
    const double c0 = 2.0 * (uy * X2 - uz * X1);
    const double c1 = 2.0 * (uz * X0 - ux * X2);
    const double c2 = 2.0 * (ux * X1 - uy * X0);
    const double pcx = X0 + w * c0 + (uy * c2 - uz * c1) + t[0];
    const double pcy = X1 + w * c1 + (uz * c0 - ux * c2) + t[1];
    const double pcz = X2 + w * c2 + (ux * c1 - uy * c0) + t[2];
    if (pcz < k.min_depth) {  // cost_factor_ceres.h:29-31: constant residual, zero Jacobian
        e.r0 = e.r1 = k.neg_depth_residual;
        if (kJac) {
#pragma unroll
            for (int i = 0; i < 6; ++i) e.Jd[i] = e.Jt[i] = e.JX[i] = 0.0;
        }
    } else {
        const double iz = 1.0 / pcz, x = pcx * iz, y = pcy * iz;
        double u, v, D[4];
        world_to_image(model, intr, x, y, u, v, D);
        e.r0 = u - um, e.r1 = v - vm;
        if (kJac) {
            double A[6];
            A[0] = D[0] * iz, A[1] = D[1] * iz, A[2] = -(D[0] * x + D[1] * y) * iz;
            A[3] = D[2] * iz, A[4] = D[3] * iz, A[5] = -(D[2] * x + D[3] * y) * iz;
#pragma unroll
            for (int i = 0; i < 6; ++i) e.Jt[i] = A[i];
            // d pc/dX = I + 2w[u]x + 2[u]x[u]x  (== R(q) on the unit sphere)
            const double uu = ux * ux + uy * uy + uz * uz;
            const double M0 = 1.0 + 2.0 * (ux * ux - uu), M1 = 2.0 * (ux * uy - w * uz), M2 = 2.0 * (ux * uz + w * uy);
            const double M3 = 2.0 * (ux * uy + w * uz), M4 = 1.0 + 2.0 * (uy * uy - uu), M5 = 2.0 * (uy * uz - w * ux);
            const double M6 = 2.0 * (ux * uz - w * uy), M7 = 2.0 * (uy * uz + w * ux), M8 = 1.0 + 2.0 * (uz * uz - uu);
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                e.JX[r * 3 + 0] = A[r * 3] * M0 + A[r * 3 + 1] * M3 + A[r * 3 + 2] * M6;
                e.JX[r * 3 + 1] = A[r * 3] * M1 + A[r * 3 + 1] * M4 + A[r * 3 + 2] * M7;
                e.JX[r * 3 + 2] = A[r * 3] * M2 + A[r * 3 + 1] * M5 + A[r * 3 + 2] * M8;
            }
            // G = d pc / d(x,y,z,w):  d/du = -2w[v]x + 2((u.v)I + u v^T - 2 v u^T), d/dw = 2(u x v)
            const double udv = ux * X0 + uy * X1 + uz * X2;
            double G[12];
            G[0] = 2.0 * (udv + ux * X0 - 2.0 * X0 * ux);
            G[1] = 2.0 * (w * X2 + ux * X1 - 2.0 * X0 * uy);
            G[2] = 2.0 * (-w * X1 + ux * X2 - 2.0 * X0 * uz);
            G[3] = c0;
            G[4] = 2.0 * (-w * X2 + uy * X0 - 2.0 * X1 * ux);
            G[5] = 2.0 * (udv + uy * X1 - 2.0 * X1 * uy);
            G[6] = 2.0 * (w * X0 + uy * X2 - 2.0 * X1 * uz);
            G[7] = c1;
            G[8] = 2.0 * (w * X1 + uz * X0 - 2.0 * X2 * ux);
            G[9] = 2.0 * (-w * X0 + uz * X1 - 2.0 * X2 * uy);
            G[10] = 2.0 * (udv + uz * X2 - 2.0 * X2 * uz);
            G[11] = c2;
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const double j0 = A[r * 3] * G[0] + A[r * 3 + 1] * G[4] + A[r * 3 + 2] * G[8];
                const double j1 = A[r * 3] * G[1] + A[r * 3 + 1] * G[5] + A[r * 3 + 2] * G[9];
                const double j2 = A[r * 3] * G[2] + A[r * 3 + 1] * G[6] + A[r * 3 + 2] * G[10];
                const double j3 = A[r * 3] * G[3] + A[r * 3 + 1] * G[7] + A[r * 3 + 2] * G[11];
                // EigenQuaternionParameterization::ComputeJacobian (4x3, row-major):
                //   [ w, z,-y; -z, w, x;  y,-x, w; -x,-y,-z ]
                e.Jd[r * 3 + 0] = j0 * w - j1 * uz + j2 * uy - j3 * ux;
                e.Jd[r * 3 + 1] = j0 * uz + j1 * w - j2 * ux - j3 * uy;
                e.Jd[r * 3 + 2] = -j0 * uy + j1 * ux + j2 * w - j3 * uz;
            }
        }
    }
    const double s = e.r0 * e.r0 + e.r1 * e.r1;
    double rho1 = 1.0;
    if (s > k.huber_b) {  // HuberLoss::Evaluate
        const double rt = sqrt(s);
        e.rho0 = 2.0 * k.huber_a * rt - k.huber_b;
        rho1 = fmax(DBL_MIN, k.huber_a / rt);
    } else {
        e.rho0 = s;
    }
    if (robustify && rho1 != 1.0) {  // Corrector, alpha == 0 branch
        const double sc = sqrt(rho1);
        e.r0 *= sc, e.r1 *= sc;
        if (kJac) {
#pragma unroll
            for (int i = 0; i < 6; ++i) e.Jd[i] *= sc, e.Jt[i] *= sc, e.JX[i] *= sc;
        }
    }
}

// EigenQuaternionParameterization::Plus: x+ = dq (x) x, dq = (sin|d|/|d| d, cos|d|)
__device__ __forceinline__ void quat_plus(const double *q, double d0, double d1, double d2,
                                          double out[4]) {
    const double n = sqrt(d0 * d0 + d1 * d1 + d2 * d2);
    if (n > 0.0) {
        const double sbd = sin(n) / n;
        const double ax = sbd * d0, ay = sbd * d1, az = sbd * d2, aw = cos(n);
        const double bx = q[0], by = q[1], bz = q[2], bw = q[3];
        out[0] = aw * bx + ax * bw + ay * bz - az * by;
        out[1] = aw * by - ax * bz + ay * bw + az * bx;
        out[2] = aw * bz + ax * by - ay * bx + az * bw;
        out[3] = aw * bw - ax * bx - ay * by - az * bz;
    } else {
        out[0] = q[0], out[1] = q[1], out[2] = q[2], out[3] = q[3];
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    return v;
}

}  // namespace xrb
