// ba_filter.cu — post-BA point filter on the resident problem.
//
// Replaces Point3dProcessor::FilterPoints3d / FilterPoint3d / UpdateTrackAngle
// (src/geometry/track_processor.cc:253-349) with Reprojection_Error (:19-26) and
// CalculateTriangulationAngle (src/geometry/colmap/base/triangulation.cc:124-147), called after every
// KGBA (src/mapper/incremental_mapper.cc:83-85).  The reference walks the map track by track on the
// host; here the poses and points the solver just produced are still in HBM, one thread takes one
// point and walks its observations in the order the reference's std::map would (ascending frame).
//
// This translation unit is compiled with -fmad=false: the decisions are threshold tests on a few
// dozen flops, and without contraction they are the same IEEE operations the host code performs
// (the arc cosine of the angle test is the only library call).
#include <cuda_runtime.h>

#include <cmath>

#include "ba_kernels.cuh"

namespace xrb {

namespace {

__device__ __forceinline__ void project(int model, const double *__restrict__ p, double x, double y, double &u, double &v) {
    if (model == 2 || model == 3) {  // camera_model.hpp:93-210, see ba_kernels.cu
        const double fx = p[0], fy = model == 2 ? p[0] : p[1];
        const double cx = model == 2 ? p[1] : p[2], cy = model == 2 ? p[2] : p[3];
        const double k = model == 2 ? p[3] : p[4];
        const double r2 = x * x + y * y, radial = k * r2;
        u = fx * (x + x * radial) + cx;
        v = fy * (y + y * radial) + cy;
    } else if (model == 0) {
        u = p[0] * (x + x) + p[1], v = p[0] * (y + y) + p[2];
    } else if (model == 1) {
        u = p[0] * (x + x) + p[2], v = p[1] * (y + y) + p[3];
    } else {
        const double fx = p[0], fy = p[1], k1 = p[4], k2 = p[5], p1 = p[6], p2 = p[7];
        const double x2 = x * x, xy = x * y, y2 = y * y, r2 = x2 + y2;
        const double radial = k1 * r2 + k2 * r2 * r2;
        const double du = x * radial + 2 * p1 * xy + p2 * (r2 + 2 * x2);
        const double dv = y * radial + 2 * p2 * xy + p1 * (r2 + 2 * y2);
        u = fx * (x + du) + p[2], v = fy * (y + dv) + p[3];
    }
}

// Pose::center (src/base/types.h:45): -(q^-1 t), Eigen's inverse() = conjugate / squaredNorm
__global__ void k_cam_centres(int n_cams, const double *__restrict__ q_all, const double *__restrict__ t_all,
                              double *__restrict__ ctr) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cams) return;
    const double *q = q_all + 4 * (size_t)c, *t = t_all + 3 * (size_t)c;
    const double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const double ux = -q[0] / n2, uy = -q[1] / n2, uz = -q[2] / n2, w = q[3] / n2;
    const double cx = 2 * (uy * t[2] - uz * t[1]), cy = 2 * (uz * t[0] - ux * t[2]), cz = 2 * (ux * t[1] - uy * t[0]);
    ctr[3 * (size_t)c] = -(t[0] + w * cx + (uy * cz - uz * cy));
    ctr[3 * (size_t)c + 1] = -(t[1] + w * cy + (uz * cx - ux * cz));
    ctr[3 * (size_t)c + 2] = -(t[2] + w * cz + (ux * cy - uy * cx));
}

__global__ void __launch_bounds__(128)
k_filter_points(BAProblemDev P, BAStateDev st, const int32_t *__restrict__ obs_orig, const double *__restrict__ ctr,
                int32_t *__restrict__ order, uint8_t *__restrict__ flag, double max_re, double min_angle,
                uint8_t *__restrict__ keep_obs, uint8_t *__restrict__ pt_outlier, double *__restrict__ pt_error,
                double *__restrict__ pt_angle, int32_t *__restrict__ counts) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_pts_local) return;
    const int k0 = P.pt_ptr[p], k1 = P.pt_ptr[p + 1], n = k1 - k0;
    pt_outlier[p] = 0, pt_error[p] = 0.0, pt_angle[p] = 0.0;
    if (n == 0) return;
    // observations in ascending camera index, position breaking ties (the std::map order of a track)
    int last_cam = -1, last_o = -1;
    for (int s = 0; s < n; ++s) {
        int best = -1, best_cam = 0;
        for (int o = k0; o < k1; ++o) {
            const int c = P.obs_cam[o];
            if ((c > last_cam || (c == last_cam && o > last_o)) && (best < 0 || c < best_cam)) best = o, best_cam = c;
        }
        order[k0 + s] = best, last_cam = best_cam, last_o = best;
    }
    const double X0 = st.X[3 * (size_t)p], X1 = st.X[3 * (size_t)p + 1], X2 = st.X[3 * (size_t)p + 2];
    double re_sum = 0.0;
    int n_del = 0;
    for (int s = 0; s < n; ++s) {
        const int o = order[k0 + s], c = P.obs_cam[o];
        const double *q = st.q + 4 * (size_t)c, *t = st.t + 3 * (size_t)c;
        const double ux = q[0], uy = q[1], uz = q[2], w = q[3];
        const double cx = 2 * (uy * X2 - uz * X1), cy = 2 * (uz * X0 - ux * X2), cz = 2 * (ux * X1 - uy * X0);
        const double pcx = X0 + w * cx + (uy * cz - uz * cy) + t[0];
        const double pcy = X1 + w * cy + (uz * cx - ux * cz) + t[1];
        const double pcz = X2 + w * cz + (ux * cy - uy * cx) + t[2];
        const int ci = P.cam_intr[c];
        double u, v;
        project(P.intr_model[ci], P.intr + 8 * (size_t)ci, pcx / pcz, pcy / pcz, u, v);
        const double dx = u - P.obs_uv[2 * (size_t)o], dy = v - P.obs_uv[2 * (size_t)o + 1];
        const double re = sqrt(dx * dx + dy * dy);
        const bool del = re > max_re || pcz < 1e-3 || pcz > 1e3;  // track_processor.cc:288-296
        flag[o] = del ? 0 : 1;
        if (del) ++n_del; else re_sum += re;
    }
    if (n_del >= n - 1) {  // :300-303: fewer than two observations would remain
        atomicAdd(&counts[0], n);
        pt_outlier[p] = 1;
        for (int o = k0; o < k1; ++o) keep_obs[obs_orig[o]] = 0;
        return;
    }
    atomicAdd(&counts[0], n_del);
    pt_error[p] = re_sum / (double)(n - n_del);
    // UpdateTrackAngle (:253-277): pairs (i, j > i) of the remaining observations in order, early exit at
    // the first running maximum above the threshold
    double max_angle = 0.0;
    bool done = false;
    for (int i = 0; i < n && !done; ++i) {
        const int oi = order[k0 + i];
        if (!flag[oi]) continue;
        const double *a = ctr + 3 * (size_t)P.obs_cam[oi];
        const double r1 = (X0 - a[0]) * (X0 - a[0]) + (X1 - a[1]) * (X1 - a[1]) + (X2 - a[2]) * (X2 - a[2]);
        for (int j = i + 1; j < n; ++j) {
            const int oj = order[k0 + j];
            if (!flag[oj]) continue;
            const double *b = ctr + 3 * (size_t)P.obs_cam[oj];
            const double base2 = (a[0] - b[0]) * (a[0] - b[0]) + (a[1] - b[1]) * (a[1] - b[1]) + (a[2] - b[2]) * (a[2] - b[2]);
            const double r2 = (X0 - b[0]) * (X0 - b[0]) + (X1 - b[1]) * (X1 - b[1]) + (X2 - b[2]) * (X2 - b[2]);
            const double den = 2.0 * sqrt(r1 * r2);
            double angle = 0.0;
            if (den != 0.0) {
                const double ang = fabs(acos((r1 + r2 - base2) / den));
                angle = fmin(ang, M_PI - ang);
            }
            if (angle > max_angle) {
                max_angle = angle;
                if (max_angle > min_angle) {
                    done = true;
                    break;
                }
            }
        }
    }
    pt_angle[p] = max_angle;
    const bool out = max_angle < min_angle;  // :343-346
    if (out) {
        pt_outlier[p] = 1;
        atomicAdd(&counts[1], 1);
    }
    for (int o = k0; o < k1; ++o) keep_obs[obs_orig[o]] = out ? 0 : flag[o];
}

}  // namespace

int ba_launch_filter(const BAProblemDev &P, const BAStateDev &st, const int32_t *obs_orig, double *ctr, int32_t *order,
                     uint8_t *flag, double max_re, double deg, uint8_t *keep_obs, uint8_t *pt_outlier, double *pt_error,
                     double *pt_angle, int32_t *counts, cudaStream_t stream) {
    const double min_angle = deg * 0.0174532925199432954743716805978692718781530857086181640625;
    XRB_CUDA(cudaMemsetAsync(counts, 0, 2 * sizeof(int32_t), stream));
    if (P.n_cams > 0) {
        k_cam_centres<<<(P.n_cams + 127) / 128, 128, 0, stream>>>(P.n_cams, st.q, st.t, ctr);
        XRB_LAUNCHED();
    }
    if (P.n_pts_local > 0) {
        k_filter_points<<<(P.n_pts_local + 127) / 128, 128, 0, stream>>>(P, st, obs_orig, ctr, order, flag, max_re, min_angle,
                                                                          keep_obs, pt_outlier, pt_error, pt_angle, counts);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

}  // namespace xrb
