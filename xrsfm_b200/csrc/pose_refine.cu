// pose_refine.cu — batched pose refinement: thousands of 6-DoF Levenberg-Marquardt problems in ONE launch.
//
// Replaces the Ceres block of RegisterImage (src/geometry/pnp.cc:38-71): after the P3P LORANSAC, the frame's pose
// (Tcw.q, Tcw.t) is refined over the inlier 2D-3D correspondences with ReProjectionCost (cost_factor_ceres.h:19-40),
// HuberLoss(5.99), EigenQuaternionParameterization, points and intrinsics constant, default solver options with
// max_num_iterations = 10.  The mapper does this once per registered frame (incremental_mapper.cc:46); a GPU has
// nothing to do on one such problem (a 6 x 6 system over a few hundred residuals), so the drop-in is batched:
// one CTA per pose, the whole trust-region loop inside the kernel, no host round trip per iteration.
//
// The loop is the one xrb_ba_run drives from the host (ba_api.cu do_run: Ceres' TrustRegionMinimizer with the
// Levenberg-Marquardt strategy, Jacobi scaling from the iteration-0 Jacobian, the same termination tests in the
// same order); with no variable point the reduced camera system IS the 6 x 6 normal matrix, so "Schur complement",
// "Cholesky" and "back-substitution" collapse into a 21-value block reduction and a few hundred flops on one thread.
// All reductions use a fixed order: a pose's result does not depend on the batch it is in.
#include <cuda_runtime.h>

#include <algorithm>
#include <climits>
#include <mutex>

#include "ba_model.cuh"
#include "common.cuh"

namespace xrb {
namespace {

constexpr int kPoseThreads = 128;
constexpr int kPoseWarps = kPoseThreads / 32;
constexpr int kAcc = 28;  // 21 (lower triangle of J^T J) + 6 (J^T r) + 1 (cost)

struct PoseBatch {
    int n;
    const long long *off;
    const double *uv, *xyz;
    const uint8_t *inlier;  // or nullptr
    const double *intr;
    const int32_t *model;
    double *q, *t;
    xrb_pose_summary *sum;
};

struct PoseShared {
    double q[4], t[3], cq[4], ct[3], sc[6];
    double acc[kAcc];
    double red[kPoseWarps][kAcc];
    double step[4];  // model_cost_change, step_norm, cand_xnorm, ok
    double intr[8];
    int count[kPoseWarps];
};

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b

// acc[0..kAcc) <- block sum of v[0..n), fixed order: lanes by butterfly, warps 0..3 in sequence
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N], PoseShared &S) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double s = warp_sum(v[i]);
        if (lane == 0) S.red[warp][i] = s;
    }
    __syncthreads();
    if (threadIdx.x < N) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kPoseWarps; ++w) s += S.red[w][threadIdx.x];
        S.acc[threadIdx.x] = s;
    }
    __syncthreads();
}

// J^T J, J^T r (robustified, columns scaled by S.sc) and the cost 1/2 sum rho at (q, t) -> S.acc
__device__ __forceinline__ void linearise(const PoseBatch &B, PoseShared &S, long long lo, int n, int model,
                                          const BAConsts &k, const double *q, const double *t) {
    double v[kAcc];
#pragma unroll
    for (int i = 0; i < kAcc; ++i) v[i] = 0.0;
    for (int i = threadIdx.x; i < n; i += kPoseThreads) {
        const long long o = lo + i;
        if (B.inlier && !B.inlier[o]) continue;
        Obs e;
        eval_obs<true>(q, t, B.xyz + 3 * o, model, S.intr, B.uv[2 * o], B.uv[2 * o + 1], k, true, e);
        double J[12];
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int j = 0; j < 3; ++j) J[r * 6 + j] = e.Jd[r * 3 + j] * S.sc[j], J[r * 6 + 3 + j] = e.Jt[r * 3 + j] * S.sc[3 + j];
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b) v[tri(a, b)] += J[a] * J[b] + J[6 + a] * J[6 + b];
            v[21 + a] += J[a] * e.r0 + J[6 + a] * e.r1;
        }
        v[27] += 0.5 * e.rho0;
    }
    block_sum(v, S);
}

// cost 1/2 sum rho at (q, t) -> S.acc[27]
__device__ __forceinline__ void cost_only(const PoseBatch &B, PoseShared &S, long long lo, int n, int model,
                                          const BAConsts &k, const double *q, const double *t) {
    double v[1] = {0.0};
    for (int i = threadIdx.x; i < n; i += kPoseThreads) {
        const long long o = lo + i;
        if (B.inlier && !B.inlier[o]) continue;
        Obs e;
        eval_obs<false>(q, t, B.xyz + 3 * o, model, S.intr, B.uv[2 * o], B.uv[2 * o + 1], k, false, e);
        v[0] += 0.5 * e.rho0;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double s = warp_sum(v[0]);
    if (lane == 0) S.red[warp][0] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double c = 0.0;
#pragma unroll
        for (int w = 0; w < kPoseWarps; ++w) c += S.red[w][0];
        S.acc[27] = c;
    }
    __syncthreads();
}

// max-norm of the gradient as Ceres measures it on a manifold: ||x - Plus(x, -g)||_inf (k_cam_diag does the same)
__device__ __forceinline__ double gradient_max_norm(const PoseShared &S) {
    double qn[4], g = 0.0;
    quat_plus(S.q, -S.acc[21] / S.sc[0], -S.acc[22] / S.sc[1], -S.acc[23] / S.sc[2], qn);
#pragma unroll
    for (int j = 0; j < 4; ++j) g = fmax(g, fabs(S.q[j] - qn[j]));
#pragma unroll
    for (int j = 0; j < 3; ++j) g = fmax(g, fabs(S.acc[24 + j] / S.sc[3 + j]));
    return g;
}

// One thread: (H + D^2) y = g by Cholesky with D^2 = clamp(diag H, 1e-6, 1e32) / radius, the candidate pose and the
// quantities the step evaluation needs.
__device__ void solve_step(PoseShared &S, double radius) {
    const double inv_radius = 1.0 / radius;
    double A[21], y[6];
#pragma unroll
    for (int i = 0; i < 21; ++i) A[i] = S.acc[i];
#pragma unroll
    for (int a = 0; a < 6; ++a) A[tri(a, a)] += fmin(fmax(S.acc[tri(a, a)], 1e-6), 1e32) * inv_radius;
    bool bad = false;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = A[tri(j, j)];
#pragma unroll
        for (int m = 0; m < j; ++m) d -= A[tri(j, m)] * A[tri(j, m)];
        if (!(d > 0.0) || !isfinite(d)) bad = true, d = 1.0;
        const double l = sqrt(d);
        A[tri(j, j)] = l;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double s = A[tri(i, j)];
#pragma unroll
            for (int m = 0; m < j; ++m) s -= A[tri(i, m)] * A[tri(j, m)];
            A[tri(i, j)] = s / l;
        }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {  // L z = g
        double s = S.acc[21 + i];
#pragma unroll
        for (int m = 0; m < i; ++m) s -= A[tri(i, m)] * y[m];
        y[i] = s / A[tri(i, i)];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {  // L^T y = z
        double s = y[i];
#pragma unroll
        for (int m = i + 1; m < 6; ++m) s -= A[tri(m, i)] * y[m];
        y[i] = s / A[tri(i, i)];
    }
    // model cost change -(J s)^T (r + J s / 2) with s = -y:  y^T g - y^T H y / 2
    double yg = 0.0, yHy = 0.0;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        yg += y[a] * S.acc[21 + a];
        double row = 0.0;
#pragma unroll
        for (int b = 0; b < 6; ++b) row += S.acc[a >= b ? tri(a, b) : tri(b, a)] * y[b];
        yHy += y[a] * row;
    }
    const double model = yg - 0.5 * yHy;
    double sn2 = 0.0, xn2 = 0.0;
    quat_plus(S.q, -y[0] * S.sc[0], -y[1] * S.sc[1], -y[2] * S.sc[2], S.cq);
#pragma unroll
    for (int j = 0; j < 4; ++j) sn2 += (S.q[j] - S.cq[j]) * (S.q[j] - S.cq[j]), xn2 += S.cq[j] * S.cq[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const double v = S.t[j] - y[3 + j] * S.sc[3 + j];
        S.ct[j] = v;
        sn2 += (S.t[j] - v) * (S.t[j] - v), xn2 += v * v;
    }
    bool finite = isfinite(model) && isfinite(sn2);
#pragma unroll
    for (int a = 0; a < 6; ++a) finite = finite && isfinite(y[a]);
    S.step[0] = model, S.step[1] = sqrt(sn2), S.step[2] = sqrt(xn2), S.step[3] = (!bad && finite) ? 1.0 : 0.0;
}

__global__ void __launch_bounds__(kPoseThreads, 3)
k_pose_refine(PoseBatch B, xrb_ba_options O) {
    __shared__ PoseShared S;
    const BAConsts k{O.huber_a, O.huber_a * O.huber_a, O.min_depth, O.neg_depth_residual};
    const int tid = threadIdx.x;
    for (int p = blockIdx.x; p < B.n; p += gridDim.x) {
        const long long lo = B.off[p];
        const int n = (int)(B.off[p + 1] - lo);
        const int model = B.model[p];
        __syncthreads();  // the previous pose's shared state is dead
        if (tid < 4) S.q[tid] = B.q[4 * (size_t)p + tid];
        if (tid < 3) S.t[tid] = B.t[3 * (size_t)p + tid];
        if (tid < 6) S.sc[tid] = 1.0;
        if (tid < 8) S.intr[tid] = B.intr[8 * (size_t)p + tid];
        int mine = 0;
        for (int i = tid; i < n; i += kPoseThreads) mine += !B.inlier || B.inlier[lo + i];
        mine = __reduce_add_sync(0xFFFFFFFFu, mine);
        if ((tid & 31) == 0) S.count[tid >> 5] = mine;
        __syncthreads();
        int n_res = 0;
#pragma unroll
        for (int w = 0; w < kPoseWarps; ++w) n_res += S.count[w];

        xrb_pose_summary out;
        out.num_residuals = 2 * n_res;
        out.num_lm_iterations = 0, out.num_successful_steps = 0, out.num_unsuccessful_steps = 0;
        out.termination_type = XRB_BA_NO_CONVERGENCE, out.reserved = 0;
        out.initial_cost = out.final_cost = 0.0;
        if (n_res == 0) {  // nothing depends on the pose
            out.termination_type = XRB_BA_CONVERGENCE;
            if (tid == 0) B.sum[p] = out;
            continue;
        }
        // ---- iteration 0: cost, Jacobi scaling 1 / (1 + ||column||) from the robustified Jacobian, gradient
        linearise(B, S, lo, n, model, k, S.q, S.t);
        double x_cost = S.acc[27];
        out.initial_cost = x_cost;
        out.num_successful_steps = 1;  // iteration 0 counts, as in ceres::Solver::Summary and xrb_ba_summary
        double min_cost = x_cost;
        __syncthreads();
        if (tid == 0) {
            double sc[6];
#pragma unroll
            for (int j = 0; j < 6; ++j) sc[j] = 1.0 / (1.0 + sqrt(S.acc[tri(j, j)]));
#pragma unroll
            for (int a = 0; a < 6; ++a) {
#pragma unroll
                for (int b = 0; b <= a; ++b) S.acc[tri(a, b)] *= sc[a] * sc[b];
                S.acc[21 + a] *= sc[a];
                S.sc[a] = sc[a];
            }
        }
        __syncthreads();
        double xnorm = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) xnorm += S.q[j] * S.q[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) xnorm += S.t[j] * S.t[j];
        xnorm = sqrt(xnorm);
        double grad_max = gradient_max_norm(S);
        double radius = O.initial_radius, decrease_factor = 2.0;
        int iteration = 0, consecutive_invalid = 0;
        // every thread runs the same control flow on the same shared values
        for (;;) {
            if (iteration >= O.max_iterations) { out.termination_type = XRB_BA_NO_CONVERGENCE; break; }
            if (!O.fixed_iterations && grad_max <= O.gradient_tolerance) { out.termination_type = XRB_BA_CONVERGENCE; break; }
            if (radius <= 1e-32) { out.termination_type = XRB_BA_CONVERGENCE; break; }
            iteration++;
            __syncthreads();  // everyone has read the previous step's results
            if (tid == 0) solve_step(S, radius);
            __syncthreads();
            out.num_lm_iterations++;
            cost_only(B, S, lo, n, model, k, S.cq, S.ct);
            const double model_change = S.step[0], step_norm = S.step[1], cand_xnorm = S.step[2];
            const double cand_cost = S.acc[27];
            const bool valid = S.step[3] != 0.0 && isfinite(cand_cost) && model_change > 0.0;
            if (!valid) {  // HandleInvalidStep
                if (++consecutive_invalid >= 5) { out.termination_type = XRB_BA_FAILURE; break; }
                radius *= 0.5;
                out.num_unsuccessful_steps++;
                continue;
            }
            consecutive_invalid = 0;
            if (!O.fixed_iterations && step_norm <= O.parameter_tolerance * (xnorm + O.parameter_tolerance)) {
                out.termination_type = XRB_BA_CONVERGENCE;
                break;
            }
            const double cost_change = x_cost - cand_cost;
            if (!O.fixed_iterations && fabs(cost_change) <= O.function_tolerance * x_cost) {
                out.termination_type = XRB_BA_CONVERGENCE;
                break;
            }
            const double rho = cost_change / model_change;
            min_cost = fmin(min_cost, cand_cost);  // the logged cost of the iteration, accepted or not (do_run)
            if (rho > 1e-3) {  // HandleSuccessfulStep
                __syncthreads();
                if (tid < 4) S.q[tid] = S.cq[tid];
                if (tid < 3) S.t[tid] = S.ct[tid];
                __syncthreads();
                xnorm = cand_xnorm, x_cost = cand_cost;
                out.num_successful_steps++;
                radius = radius / fmax(1.0 / 3.0, 1.0 - pow(2.0 * rho - 1.0, 3.0));
                radius = fmin(1e16, radius);
                decrease_factor = 2.0;
                linearise(B, S, lo, n, model, k, S.q, S.t);
                grad_max = gradient_max_norm(S);
            } else {  // StepRejected
                radius = radius / decrease_factor;
                decrease_factor *= 2.0;
                out.num_unsuccessful_steps++;
            }
        }
        out.final_cost = fmin(out.initial_cost, min_cost);
        if (tid < 4) B.q[4 * (size_t)p + tid] = S.q[tid];
        if (tid < 3) B.t[3 * (size_t)p + tid] = S.t[tid];
        if (tid == 0) B.sum[p] = out;
    }
}

constexpr int kMaxDevices = 64;
struct PoseWorkspace {  // grow-only device buffers, a stream and two events per device, kept between calls
    std::mutex mu;
    DevBuf off, uv, xyz, in, intr, model, q, t, sum;
    cudaStream_t st = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int sms = 148;
    double kernel_ms = -1.0;
};
PoseWorkspace g_ws[kMaxDevices];

}  // namespace
}  // namespace xrb

using namespace xrb;

extern "C" {

void xrb_pose_default_options(xrb_ba_options *o) {
    xrb_ba_default_options(o);  // Ceres defaults + the cost-functor constants
    o->max_iterations = 10;      // pnp.cc:57
    o->function_tolerance = 1e-6, o->parameter_tolerance = 1e-8, o->gradient_tolerance = 1e-10;
    o->initial_radius = 1e4;
}

int xrb_pose_refine_batch(int device, int n_poses, const int64_t *offsets, const double *uv, const double *xyz,
                          const uint8_t *inlier_mask, const double *intr, const int32_t *intr_model, double *q, double *t,
                          const xrb_ba_options *opt, xrb_pose_summary *summaries) {
    if (n_poses < 0 || !opt || (n_poses && (!offsets || !intr || !intr_model || !q || !t || !summaries))) {
        set_error("pose_refine_batch: bad arguments");
        return XRB_ERR_INVALID;
    }
    if (n_poses == 0) return XRB_OK;
    if (offsets[0] < 0) {
        set_error("pose_refine_batch: offsets must start at >= 0");
        return XRB_ERR_INVALID;
    }
    for (int p = 0; p < n_poses; ++p) {
        if (offsets[p + 1] < offsets[p] || offsets[p + 1] - offsets[p] > (int64_t)INT32_MAX) {
            set_error("pose_refine_batch: offsets must be non-decreasing");
            return XRB_ERR_INVALID;
        }
        if (intr_model[p] < 0 || intr_model[p] > 4) {
            set_error("pose_refine_batch: pose %d has camera model id %d (0..4 are defined)", p, intr_model[p]);
            return XRB_ERR_INVALID;
        }
    }
    const int64_t total = offsets[n_poses];
    if (total && (!uv || !xyz)) {
        set_error("pose_refine_batch: null correspondence arrays");
        return XRB_ERR_INVALID;
    }
    int rc = select_device(device);
    if (rc) return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    if (device < 0 || device >= kMaxDevices) {
        set_error("pose_refine_batch: device %d out of range", device);
        return XRB_ERR_INVALID;
    }
    PoseWorkspace &W = g_ws[device];
    std::lock_guard<std::mutex> lock(W.mu);  // one batch at a time per device: the workspace is shared
    const size_t tot = (size_t)std::max<int64_t>(total, 1), np = (size_t)n_poses;
    if ((rc = W.off.reserve((np + 1) * 8)) || (rc = W.uv.reserve(tot * 16)) || (rc = W.xyz.reserve(tot * 24)) ||
        (rc = W.in.reserve(tot)) || (rc = W.intr.reserve(np * 64)) || (rc = W.model.reserve(np * 4)) ||
        (rc = W.q.reserve(np * 32)) || (rc = W.t.reserve(np * 24)) || (rc = W.sum.reserve(np * sizeof(xrb_pose_summary))))
        return rc;
    if (!W.st) {
        XRB_CUDA(cudaStreamCreateWithFlags(&W.st, cudaStreamNonBlocking));
        XRB_CUDA(cudaEventCreate(&W.ev[0]));
        XRB_CUDA(cudaEventCreate(&W.ev[1]));
        cudaDeviceGetAttribute(&W.sms, cudaDevAttrMultiProcessorCount, device);
    }
    cudaStream_t st = W.st;
    XRB_CUDA(cudaMemcpyAsync(W.off.p, offsets, (np + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total) {
        XRB_CUDA(cudaMemcpyAsync(W.uv.p, uv, (size_t)total * 16, cudaMemcpyHostToDevice, st));
        XRB_CUDA(cudaMemcpyAsync(W.xyz.p, xyz, (size_t)total * 24, cudaMemcpyHostToDevice, st));
        if (inlier_mask) XRB_CUDA(cudaMemcpyAsync(W.in.p, inlier_mask, (size_t)total, cudaMemcpyHostToDevice, st));
    }
    XRB_CUDA(cudaMemcpyAsync(W.intr.p, intr, np * 64, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaMemcpyAsync(W.model.p, intr_model, np * 4, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaMemcpyAsync(W.q.p, q, np * 32, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaMemcpyAsync(W.t.p, t, np * 24, cudaMemcpyHostToDevice, st));
    const int grid = std::min(n_poses, W.sms * 8);
    PoseBatch B{n_poses, W.off.as<long long>(), W.uv.as<double>(), W.xyz.as<double>(),
                inlier_mask ? W.in.as<uint8_t>() : nullptr, W.intr.as<double>(), W.model.as<int32_t>(),
                W.q.as<double>(), W.t.as<double>(), W.sum.as<xrb_pose_summary>()};
    XRB_CUDA(cudaEventRecord(W.ev[0], st));
    k_pose_refine<<<grid, kPoseThreads, 0, st>>>(B, *opt);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    XRB_CUDA(cudaEventRecord(W.ev[1], st));
    XRB_CUDA(cudaMemcpyAsync(q, W.q.p, np * 32, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaMemcpyAsync(t, W.t.p, np * 24, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaMemcpyAsync(summaries, W.sum.p, np * sizeof(xrb_pose_summary), cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, W.ev[0], W.ev[1]) == cudaSuccess) W.kernel_ms = ms;
    return XRB_OK;
}

double xrb_pose_last_kernel_ms(int device) {
    if (device < 0 || device >= kMaxDevices) return -1.0;
    std::lock_guard<std::mutex> lock(g_ws[device].mu);
    return g_ws[device].kernel_ms;
}

}  // extern "C"
