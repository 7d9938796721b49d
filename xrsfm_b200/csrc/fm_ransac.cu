// fm_ransac.cu — geometric verification of matched image pairs on the GPU: LO-RANSAC fundamental
// matrix, one CTA per pair, every pair of a batch in one launch.
//
// Replaces the loop body of FeatureMatching that calls SolveFundamnetalCOLMAP
// (src/feature/feature_processing.cc:256-296, src/geometry/epipolar_geometry.hpp:10-27):
//   colmap::LORANSAC<FundamentalMatrixSevenPointEstimator, FundamentalMatrixEightPointEstimator>
//   (src/geometry/colmap/optim/loransac.h:96-238, ransac.h:136-167) with RandomSampler
//   (optim/random_sampler.cc:41-62, util/random.h:86-122: std::mt19937 seeded with 0, partial
//   Fisher-Yates on a persistent permutation through std::uniform_int_distribution<uint32_t>),
//   InlierSupportMeasurer (optim/support_measurement.cc:36-62), the 7-point and 8-point estimators and
//   the squared Sampson error (estimators/fundamental_matrix.cc:46-295).
//
// How the sequential algorithm maps to a CTA of 128 threads, 128 trials at a time:
//   1. thread 0 draws the 128 x 7 sample indices — the generator and the distribution are restated
//      bit for bit (mt19937; libstdc++'s multiply-shift rejection sampler), so the sample sequence IS
//      the reference's for a pair that starts from a freshly seeded generator;
//   2. every thread solves one trial: null space of the 7 x 9 system, the cubic in lambda, up to three
//      models, and scores each over all matches in index order (inlier count + residual sum);
//   3. the trials are replayed in order by the whole CTA: better-than-best test, local optimisation
//      (8-point on the inliers: normalisation, 9 x 9 Gram matrix reduced across the CTA, Jacobi
//      eigenvectors, rank-2 projection) and the dynamic termination rule, exactly where the reference
//      would take them.
// The null space comes from pivoted elimination and the singular vectors from Jacobi rotations of the
// Gram matrix instead of Eigen::JacobiSVD: equal up to rounding, and up to the order in which the
// (at most three) models of one trial are visited — the same caveat the CPU oracle states.
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace xrb {

namespace {

constexpr int kFmThreads = 128;   // = trials per batch
constexpr int kMin = 7, kMinLocal = 8;
constexpr int kPermSmem = 8192;   // matches whose permutation fits in shared memory (uint16)

// ---- std::mt19937 + std::uniform_int_distribution<uint32_t> of libstdc++ (bits/uniform_int_dist.h:
// _S_nd, "Fast Random Integer Generation in an Interval") ------------------------------------------
struct Mt19937 {
    uint32_t mt[624];
    int idx;
    __host__ __device__ void seed(uint32_t s) {
        mt[0] = s;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    __host__ __device__ uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7FFFFFFFu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908B0DFu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9D2C5680u;
        y ^= (y << 15) & 0xEFC60000u;
        y ^= y >> 18;
        return y;
    }
    // uniform_int_distribution<uint32_t>(a, b)(gen) for b - a < 2^32 - 1
    __host__ __device__ uint32_t uniform(uint32_t a, uint32_t b) {
        const uint32_t range = b - a + 1u;
        uint64_t product = (uint64_t)next() * (uint64_t)range;
        uint32_t low = (uint32_t)product;
        if (low < range) {
            const uint32_t threshold = (0u - range) % range;
            while (low < threshold) {
                product = (uint64_t)next() * (uint64_t)range;
                low = (uint32_t)product;
            }
        }
        return (uint32_t)(product >> 32) + a;
    }
};

template <class IdxT>
__host__ __device__ void draw_sample(Mt19937 &g, IdxT *perm, uint32_t n, IdxT *out7) {
    const uint32_t last = n - 1;
    for (uint32_t i = 0; i < (uint32_t)kMin; ++i) {  // Shuffle(7, &sample_idxs_): util/random.h:115-122
        const uint32_t j = g.uniform(i, last);
        const IdxT t = perm[i];
        perm[i] = perm[j], perm[j] = t;
    }
    for (int i = 0; i < kMin; ++i) out7[i] = perm[i];
}

// ---- small dense kernels, one thread each ---------------------------------------------------------
// cyclic Jacobi on a symmetric N x N matrix (row-major, destroyed); V columns = eigenvectors
template <int N>
__device__ void jacobi_eig(double *A, double *V) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) V[i * N + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0.0, diag = 0.0;
        for (int i = 0; i < N; ++i) {
            diag += A[i * N + i] * A[i * N + i];
            for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
        for (int p = 0; p < N - 1; ++p)
            for (int q = p + 1; q < N; ++q) {
                const double apq = A[p * N + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * N + q] - A[p * N + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(1.0 + theta * theta));
                const double c = 1.0 / sqrt(1.0 + t * t), s = c * t;
                for (int k = 0; k < N; ++k) {  // columns p, q
                    const double akp = A[k * N + p], akq = A[k * N + q];
                    A[k * N + p] = c * akp - s * akq, A[k * N + q] = s * akp + c * akq;
                }
                for (int k = 0; k < N; ++k) {  // rows p, q
                    const double apk = A[p * N + k], aqk = A[q * N + k];
                    A[p * N + k] = c * apk - s * aqk, A[q * N + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < N; ++k) {
                    const double vkp = V[k * N + p], vkq = V[k * N + q];
                    V[k * N + p] = c * vkp - s * vkq, V[k * N + q] = s * vkp + c * vkq;
                }
            }
    }
}

// real roots of c0 x^3 + c1 x^2 + c2 x + c3 after dropping leading zeros (polynomial.cc:208-275
// semantics: only roots with a zero imaginary part survive the caller's kMaxRootImag test)
__device__ int real_roots(const double *ca, double *re) {
    int lead = 0;
    while (lead < 4 && ca[lead] == 0.0) ++lead;
    const double *c = ca + lead;
    const int d = 3 - lead;
    if (d <= 0) return 0;
    if (d == 1) {
        re[0] = -c[1] / c[0];
        return 1;
    }
    if (d == 2) {
        const double a = c[0], b = c[1], cc = c[2], disc = b * b - 4 * a * cc;
        if (disc < 0) return 0;
        const double sq = sqrt(disc), q = -0.5 * (b + (b >= 0 ? sq : -sq));
        re[0] = q / a, re[1] = q != 0 ? cc / q : 0.0;
        return 2;
    }
    const double a = c[1] / c[0], b = c[2] / c[0], c0 = c[3] / c[0];
    const double p = b - a * a / 3.0, q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + c0;
    const double disc = q * q / 4.0 + p * p * p / 27.0;
    auto polish = [&](double x) {
        for (int it = 0; it < 2; ++it) {
            const double f = ((x + a) * x + b) * x + c0, df = (3.0 * x + 2.0 * a) * x + b;
            if (df != 0.0) x -= f / df;
        }
        return x;
    };
    if (disc > 0) {
        const double sq = sqrt(disc);
        re[0] = polish(cbrt(-q / 2.0 + sq) + cbrt(-q / 2.0 - sq) - a / 3.0);
        return 1;
    }
    const double r = sqrt(fmax(0.0, -p / 3.0));
    double arg = r > 0 ? (-q / 2.0) / (r * r * r) : 0.0;
    arg = fmax(-1.0, fmin(1.0, arg));
    const double phi = acos(arg);
    for (int k = 0; k < 3; ++k) re[k] = polish(2.0 * r * cos((phi - 2.0 * M_PI * k) / 3.0) - a / 3.0);
    return 3;
}

// FundamentalMatrixSevenPointEstimator::Estimate (fundamental_matrix.cc:46-139); models row-major
__device__ int seven_point(const double *x1, const double *x2, double *models) {
    double A[7 * 9];
    for (int i = 0; i < 7; ++i) {
        const double x0 = x1[2 * i], y0 = x1[2 * i + 1], u = x2[2 * i], v = x2[2 * i + 1];
        double *a = A + 9 * i;
        a[0] = u * x0, a[1] = u * y0, a[2] = u, a[3] = v * x0, a[4] = v * y0, a[5] = v, a[6] = x0, a[7] = y0, a[8] = 1;
    }
    // null space by elimination with complete pivoting: 7 pivot columns, 2 free ones
    int colperm[9];
    for (int j = 0; j < 9; ++j) colperm[j] = j;
    for (int s = 0; s < 7; ++s) {
        int pr = s, pc = s;
        double best = -1.0;
        for (int r = s; r < 7; ++r)
            for (int c = s; c < 9; ++c)
                if (fabs(A[r * 9 + c]) > best) best = fabs(A[r * 9 + c]), pr = r, pc = c;
        if (best <= 0.0) return 0;
        if (pr != s)
            for (int c = 0; c < 9; ++c) {
                const double t = A[s * 9 + c];
                A[s * 9 + c] = A[pr * 9 + c], A[pr * 9 + c] = t;
            }
        if (pc != s) {
            for (int r = 0; r < 7; ++r) {
                const double t = A[r * 9 + s];
                A[r * 9 + s] = A[r * 9 + pc], A[r * 9 + pc] = t;
            }
            const int t = colperm[s];
            colperm[s] = colperm[pc], colperm[pc] = t;
        }
        const double inv = 1.0 / A[s * 9 + s];
        for (int c = s; c < 9; ++c) A[s * 9 + c] *= inv;
        for (int r = 0; r < 7; ++r) {
            if (r == s) continue;
            const double f = A[r * 9 + s];
            if (f == 0.0) continue;
            for (int c = s; c < 9; ++c) A[r * 9 + c] -= f * A[s * 9 + c];
        }
    }
    double f1[9], f2[9];
    for (int r = 0; r < 7; ++r) f1[colperm[r]] = -A[r * 9 + 7], f2[colperm[r]] = -A[r * 9 + 8];
    f1[colperm[7]] = 1.0, f1[colperm[8]] = 0.0, f2[colperm[7]] = 0.0, f2[colperm[8]] = 1.0;
    {   // orthonormal basis, like the singular vectors the reference takes (unit scale for the kEps test)
        double n1 = 0, d12 = 0;
        for (int k = 0; k < 9; ++k) n1 += f1[k] * f1[k];
        n1 = 1.0 / sqrt(n1);
        for (int k = 0; k < 9; ++k) f1[k] *= n1;
        for (int k = 0; k < 9; ++k) d12 += f1[k] * f2[k];
        double n2 = 0;
        for (int k = 0; k < 9; ++k) f2[k] -= d12 * f1[k], n2 += f2[k] * f2[k];
        n2 = 1.0 / sqrt(n2);
        for (int k = 0; k < 9; ++k) f2[k] *= n2;
    }
    for (int k = 0; k < 9; ++k) f1[k] -= f2[k];
    const double t0 = f1[4] * f1[8] - f1[5] * f1[7], t1 = f1[3] * f1[8] - f1[5] * f1[6], t2 = f1[3] * f1[7] - f1[4] * f1[6];
    const double t3 = f2[4] * f2[8] - f2[5] * f2[7], t4 = f2[3] * f2[8] - f2[5] * f2[6], t5 = f2[3] * f2[7] - f2[4] * f2[6];
    double co[4];
    co[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
    co[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
            f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
            f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
            f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
    co[2] = f1[0] * t3 - f1[1] * t4 + f1[2] * t5 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
            f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
            f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
            f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
    co[3] = f2[0] * t3 - f2[1] * t4 + f2[2] * t5;
    double re[3];
    const int nr = real_roots(co, re);
    int n_models = 0;
    for (int i = 0; i < nr; ++i) {
        double F[9];
        for (int k = 0; k < 9; ++k) F[k] = re[i] * f1[k] + f2[k];
        if (fabs(F[8]) < 1e-10) continue;  // kEps
        for (int k = 0; k < 9; ++k) models[9 * n_models + k] = F[k] / F[8];
        ++n_models;
    }
    return n_models;
}

// ComputeSquaredSampsonError (fundamental_matrix.cc:201-248) for one match
__device__ __forceinline__ double sampson(const double *E, double x10, double x11, double x20, double x21) {
    const double Ex0 = E[0] * x10 + E[1] * x11 + E[2], Ex1 = E[3] * x10 + E[4] * x11 + E[5], Ex2 = E[6] * x10 + E[7] * x11 + E[8];
    const double Et0 = E[0] * x20 + E[3] * x21 + E[6], Et1 = E[1] * x20 + E[4] * x21 + E[7];
    const double x2tEx1 = x20 * Ex0 + x21 * Ex1 + Ex2;
    return x2tEx1 * x2tEx1 / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1);
}

struct Support {
    long long num_inliers;
    double residual_sum;
};
__device__ __forceinline__ bool better(const Support &a, const Support &b) {
    if (a.num_inliers > b.num_inliers) return true;
    return a.num_inliers == b.num_inliers && a.residual_sum < b.residual_sum;
}
__device__ Support score_serial(const double *E, const double2 *p1, const double2 *p2, int n, double max_residual) {
    Support s{0, 0.0};
    for (int i = 0; i < n; ++i) {
        const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
        const double r = sampson(E, a.x, a.y, b.x, b.y);
        if (r <= max_residual) s.num_inliers += 1, s.residual_sum += r;
    }
    return s;
}

__device__ long long num_trials_for(long long num_inliers, long long num_samples, double confidence, int k_min) {  // ransac.h:151-167
    const double inlier_ratio = (double)num_inliers / (double)num_samples;
    const double nom = 1 - confidence;
    if (nom <= 0) return LLONG_MAX;
    const double denom = 1 - pow(inlier_ratio, (double)k_min);
    if (denom <= 0) return 1;
    const double v = ceil(log(nom) / log(denom));
    return v >= 9.2e18 ? LLONG_MAX : (long long)v;
}

// block-wide sum of `v` in a fixed order (warp shuffles, then the four warp results in order); every
// thread gets the result
__device__ double block_sum(double v, double *red) {
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < kFmThreads / 32; ++w) s += red[w];
    return s;
}

struct FmShared {
    Mt19937 gen;
    unsigned short perm16[kPermSmem];
    int samples[kFmThreads][kMin];
    double models[kFmThreads][3][9];
    int n_models[kFmThreads];
    long long inl[kFmThreads][3];
    double rsum[kFmThreads][3];
    double red[kFmThreads / 32];
    double gram[45], local[9], cand[9];
    double norm[8];   // cx1, cy1, nf1, cx2, cy2, nf2
    int flag;
};

// FundamentalMatrixEightPointEstimator::Estimate (fundamental_matrix.cc:147-192) on the matches whose
// squared Sampson error w.r.t. `model` is within max_residual; the whole CTA works, result in sh.local.
__device__ void eight_point_on_inliers(FmShared &sh, const double *model, const double2 *p1, const double2 *p2, int n,
                                       double max_residual) {
    const int tid = threadIdx.x;
    // CenterAndNormalizeImagePoints (:250-295), both images
    double s[4] = {0, 0, 0, 0};
    double cnt = 0;
    for (int i = tid; i < n; i += kFmThreads) {
        const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
        if (sampson(model, a.x, a.y, b.x, b.y) <= max_residual) s[0] += a.x, s[1] += a.y, s[2] += b.x, s[3] += b.y, cnt += 1;
    }
    const double m = block_sum(cnt, sh.red);
    const double cx1 = block_sum(s[0], sh.red) / m, cy1 = block_sum(s[1], sh.red) / m;
    const double cx2 = block_sum(s[2], sh.red) / m, cy2 = block_sum(s[3], sh.red) / m;
    double r1 = 0, r2 = 0;
    for (int i = tid; i < n; i += kFmThreads) {
        const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
        if (sampson(model, a.x, a.y, b.x, b.y) <= max_residual) {
            r1 += (a.x - cx1) * (a.x - cx1) + (a.y - cy1) * (a.y - cy1);
            r2 += (b.x - cx2) * (b.x - cx2) + (b.y - cy2) * (b.y - cy2);
        }
    }
    const double nf1 = sqrt(2.0) / sqrt(block_sum(r1, sh.red) / m), nf2 = sqrt(2.0) / sqrt(block_sum(r2, sh.red) / m);
    // Gram matrix of the constraint rows c = (x1 x2, y1 x2, x2, x1 y2, y1 y2, y2, x1, y1, 1)
    double g[45];
    for (int k = 0; k < 45; ++k) g[k] = 0.0;
    for (int i = tid; i < n; i += kFmThreads) {
        const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
        if (sampson(model, a.x, a.y, b.x, b.y) > max_residual) continue;
        const double x1 = nf1 * a.x - nf1 * cx1, y1 = nf1 * a.y - nf1 * cy1, x2 = nf2 * b.x - nf2 * cx2, y2 = nf2 * b.y - nf2 * cy2;
        const double c[9] = {x1 * x2, y1 * x2, x2, x1 * y2, y1 * y2, y2, x1, y1, 1.0};
        int w = 0;
        for (int r = 0; r < 9; ++r)
            for (int q = 0; q <= r; ++q) g[w++] += c[r] * c[q];
    }
    for (int k = 0; k < 45; ++k) {
        const double v = block_sum(g[k], sh.red);
        if (tid == 0) sh.gram[k] = v;
    }
    __syncthreads();
    if (tid == 0) {
        double A[81], V[81];
        int w = 0;
        for (int r = 0; r < 9; ++r)
            for (int q = 0; q <= r; ++q, ++w) A[r * 9 + q] = A[q * 9 + r] = sh.gram[w];
        jacobi_eig<9>(A, V);
        int kmin = 0;
        for (int k = 1; k < 9; ++k)
            if (A[k * 9 + k] < A[kmin * 9 + kmin]) kmin = k;
        double E[9];
        for (int k = 0; k < 9; ++k) E[k] = V[k * 9 + kmin];  // E(r, c) = v[r * 3 + c]
        // rank 2: remove the component along the right singular vector of the smallest singular value
        double G[9], V3[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) G[r * 3 + c] = E[r] * E[c] + E[3 + r] * E[3 + c] + E[6 + r] * E[6 + c];  // E^T E
        jacobi_eig<3>(G, V3);
        int k3 = 0;
        for (int k = 1; k < 3; ++k)
            if (G[k * 3 + k] < G[k3 * 3 + k3]) k3 = k;
        const double v2[3] = {V3[k3], V3[3 + k3], V3[6 + k3]};
        double F[9];
        for (int r = 0; r < 3; ++r) {
            const double ev = E[r * 3] * v2[0] + E[r * 3 + 1] * v2[1] + E[r * 3 + 2] * v2[2];
            for (int c = 0; c < 3; ++c) F[r * 3 + c] = E[r * 3 + c] - ev * v2[c];
        }
        // points2_norm_matrix^T * F * points1_norm_matrix
        const double M1[9] = {nf1, 0, -nf1 * cx1, 0, nf1, -nf1 * cy1, 0, 0, 1}, M2[9] = {nf2, 0, -nf2 * cx2, 0, nf2, -nf2 * cy2, 0, 0, 1};
        double Tm[9];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) Tm[r * 3 + c] = M2[0 * 3 + r] * F[0 * 3 + c] + M2[1 * 3 + r] * F[1 * 3 + c] + M2[2 * 3 + r] * F[2 * 3 + c];
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) sh.local[r * 3 + c] = Tm[r * 3] * M1[c] + Tm[r * 3 + 1] * M1[3 + c] + Tm[r * 3 + 2] * M1[6 + c];
    }
    __syncthreads();
}

// support of sh.<model> over all matches by the whole CTA: the count is exact; the residual sum is
// accumulated in the block's fixed reduction order
__device__ Support score_block(FmShared &sh, const double *E, const double2 *p1, const double2 *p2, int n, double max_residual) {
    double cnt = 0, rs = 0;
    for (int i = threadIdx.x; i < n; i += kFmThreads) {
        const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
        const double r = sampson(E, a.x, a.y, b.x, b.y);
        if (r <= max_residual) cnt += 1, rs += r;
    }
    Support s;
    s.num_inliers = (long long)block_sum(cnt, sh.red);
    s.residual_sum = block_sum(rs, sh.red);
    return s;
}

__global__ void __launch_bounds__(kFmThreads)
k_fm_loransac(int n_pairs, const long long *__restrict__ offsets, const double2 *__restrict__ pts1, const double2 *__restrict__ pts2,
              xrb_fm_options opt, int *__restrict__ perm_scratch, xrb_fm_report *__restrict__ reports, char *__restrict__ inlier_mask) {
    extern __shared__ __align__(16) unsigned char fm_smem_raw[];
    FmShared &sh = *reinterpret_cast<FmShared *>(fm_smem_raw);
    const int tid = threadIdx.x;
    for (int pair = blockIdx.x; pair < n_pairs; pair += gridDim.x) {
        const long long o0 = offsets[pair];
        const int n = (int)(offsets[pair + 1] - o0);
        const double2 *p1 = pts1 + o0, *p2 = pts2 + o0;
        xrb_fm_report rep;
        rep.success = 0, rep.best_is_local = 0, rep.num_trials = 0, rep.num_inliers = 0, rep.residual_sum = DBL_MAX;
        for (int k = 0; k < 9; ++k) rep.F[k] = 0.0;
        __syncthreads();  // the previous pair's shared state is no longer read
        if (n < kMin) {
            if (tid == 0) reports[pair] = rep;
            for (int i = tid; i < n; i += kFmThreads) inlier_mask[o0 + i] = 0;
            continue;
        }
        // RANSAC ctor (ransac.h:136-148): the trial cap implied by min_inlier_ratio
        long long max_num_trials = opt.max_num_trials;
        {
            const long long dyn = num_trials_for((long long)(opt.min_inlier_ratio * 100000), 100000, opt.confidence, kMin);
            if (dyn < max_num_trials) max_num_trials = dyn;
        }
        const double max_residual = opt.max_error * opt.max_error;
        const bool perm_in_smem = n <= kPermSmem;
        int *perm32 = perm_scratch + o0;
        if (perm_in_smem)
            for (int i = tid; i < n; i += kFmThreads) sh.perm16[i] = (unsigned short)i;
        else
            for (int i = tid; i < n; i += kFmThreads) perm32[i] = i;
        if (tid == 0) sh.gen.seed(0u);  // util/random.cc:44: the seed is 0 whatever is asked for
        Support best{0, DBL_MAX};
        double best_model[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool best_is_local = false, abort = false;
        long long dyn_max = max_num_trials, num_trials = 0;
        __syncthreads();
        while (num_trials < max_num_trials && !abort) {
            const long long left = max_num_trials - num_trials;
            const int batch = left < kFmThreads ? (int)left : kFmThreads;
            if (tid == 0) {  // the sample sequence is inherently serial (one permutation, one generator)
                for (int t = 0; t < batch; ++t) {
                    if (perm_in_smem) {
                        unsigned short s7[kMin];
                        draw_sample(sh.gen, sh.perm16, (uint32_t)n, s7);
                        for (int i = 0; i < kMin; ++i) sh.samples[t][i] = s7[i];
                    } else {
                        draw_sample(sh.gen, perm32, (uint32_t)n, sh.samples[t]);
                    }
                }
            }
            __syncthreads();
            if (tid < batch) {
                double x1[14], x2[14];
                for (int i = 0; i < kMin; ++i) {
                    const double2 a = __ldg(p1 + sh.samples[tid][i]), b = __ldg(p2 + sh.samples[tid][i]);
                    x1[2 * i] = a.x, x1[2 * i + 1] = a.y, x2[2 * i] = b.x, x2[2 * i + 1] = b.y;
                }
                const int nm = seven_point(x1, x2, &sh.models[tid][0][0]);
                sh.n_models[tid] = nm;
                for (int mi = 0; mi < nm; ++mi) {
                    const Support s = score_serial(sh.models[tid][mi], p1, p2, n, max_residual);
                    sh.inl[tid][mi] = s.num_inliers, sh.rsum[tid][mi] = s.residual_sum;
                }
            }
            __syncthreads();
            // replay in order (uniform control flow: every thread holds the same scalars)
            for (int t = 0; t < batch && !abort; ++t, ++num_trials) {
                const int nm = sh.n_models[t];
                for (int mi = 0; mi < nm; ++mi) {
                    const Support s{sh.inl[t][mi], sh.rsum[t][mi]};
                    if (better(s, best)) {
                        best = s, best_is_local = false;
                        for (int k = 0; k < 9; ++k) best_model[k] = sh.models[t][mi][k];
                        if (s.num_inliers > kMin && s.num_inliers >= kMinLocal) {  // loransac.h:166-199
                            eight_point_on_inliers(sh, best_model, p1, p2, n, max_residual);
                            const Support ls = score_block(sh, sh.local, p1, p2, n, max_residual);
                            if (better(ls, best)) {
                                best = ls, best_is_local = true;
                                for (int k = 0; k < 9; ++k) best_model[k] = sh.local[k];
                            }
                        }
                        dyn_max = num_trials_for(best.num_inliers, n, opt.confidence, kMin);
                    }
                    if (num_trials >= dyn_max && num_trials >= opt.min_num_trials) {
                        abort = true;
                        break;
                    }
                }
            }
            // the reference notices `abort` at the top of the next trial and counts that trial (loransac.h:118-121)
            if (abort && num_trials < max_num_trials) num_trials += 1;
            __syncthreads();
        }
        rep.num_trials = num_trials, rep.num_inliers = best.num_inliers, rep.residual_sum = best.residual_sum;
        rep.best_is_local = best_is_local ? 1 : 0;
        for (int k = 0; k < 9; ++k) rep.F[k] = best_model[k];
        rep.success = best.num_inliers >= kMin ? 1 : 0;
        if (tid == 0) reports[pair] = rep;
        for (int i = tid; i < n; i += kFmThreads) {
            char in = 0;
            if (rep.success) {
                const double2 a = __ldg(p1 + i), b = __ldg(p2 + i);
                in = sampson(best_model, a.x, a.y, b.x, b.y) <= max_residual ? 1 : 0;
            }
            inlier_mask[o0 + i] = in;
        }
    }
}

}  // namespace

}  // namespace xrb

using namespace xrb;

extern "C" {

void xrb_fm_default_options(xrb_fm_options *o) {  // epipolar_geometry.hpp:13-18
    if (!o) return;
    o->max_error = 4.0, o->min_inlier_ratio = 0.25, o->confidence = 0.999;
    o->min_num_trials = 100, o->max_num_trials = 10000;
}

int xrb_fm_loransac_batch(int device, int n_pairs, const int64_t *offsets, const double *pts1, const double *pts2,
                          const xrb_fm_options *opt, xrb_fm_report *reports, char *inlier_mask) {
    if (n_pairs < 0 || !opt || (n_pairs && (!offsets || !reports))) {
        set_error("fm_loransac_batch: bad arguments");
        return XRB_ERR_INVALID;
    }
    if (n_pairs == 0) return XRB_OK;
    const int64_t total = offsets[n_pairs];
    for (int p = 0; p < n_pairs; ++p)
        if (offsets[p + 1] < offsets[p] || offsets[p + 1] - offsets[p] > (int64_t)INT32_MAX) {
            set_error("fm_loransac_batch: offsets must be non-decreasing");
            return XRB_ERR_INVALID;
        }
    if (total && (!pts1 || !pts2 || !inlier_mask)) {
        set_error("fm_loransac_batch: null point arrays");
        return XRB_ERR_INVALID;
    }
    int rc = select_device(device);
    if (rc) return rc;
    static_assert(sizeof(long long) == sizeof(int64_t), "");
    DevBuf d_off, d_p1, d_p2, d_perm, d_rep, d_mask;
    const size_t tot = (size_t)std::max<int64_t>(total, 1);
    if ((rc = d_off.reserve(((size_t)n_pairs + 1) * 8)) || (rc = d_p1.reserve(tot * 16)) || (rc = d_p2.reserve(tot * 16)) ||
        (rc = d_perm.reserve(tot * 4)) || (rc = d_rep.reserve((size_t)n_pairs * sizeof(xrb_fm_report))) ||
        (rc = d_mask.reserve(tot)))
        return rc;
    cudaStream_t st;
    XRB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    XRB_CUDA(cudaMemcpyAsync(d_off.p, offsets, ((size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, st));
    if (total) {
        XRB_CUDA(cudaMemcpyAsync(d_p1.p, pts1, (size_t)total * 16, cudaMemcpyHostToDevice, st));
        XRB_CUDA(cudaMemcpyAsync(d_p2.p, pts2, (size_t)total * 16, cudaMemcpyHostToDevice, st));
    }
    static bool attr_set[64] = {};
    if (device >= 0 && device < 64 && !attr_set[device]) {
        XRB_CUDA(cudaFuncSetAttribute(k_fm_loransac, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FmShared)));
        attr_set[device] = true;
    }
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = std::min(n_pairs, sms * 3);
    k_fm_loransac<<<grid, kFmThreads, sizeof(FmShared), st>>>(n_pairs, d_off.as<long long>(), d_p1.as<double2>(), d_p2.as<double2>(), *opt,
                                                              d_perm.as<int>(), d_rep.as<xrb_fm_report>(), d_mask.as<char>());
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    XRB_CUDA(cudaMemcpyAsync(reports, d_rep.p, (size_t)n_pairs * sizeof(xrb_fm_report), cudaMemcpyDeviceToHost, st));
    if (total) XRB_CUDA(cudaMemcpyAsync(inlier_mask, d_mask.p, (size_t)total, cudaMemcpyDeviceToHost, st));
    const cudaError_t e = cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
    d_off.release(), d_p1.release(), d_p2.release(), d_perm.release(), d_rep.release(), d_mask.release();
    if (e != cudaSuccess) {
        set_error("fm_loransac_batch: %s", cudaGetErrorString(e));
        return XRB_ERR_CUDA;
    }
    return XRB_OK;
}

/* host-only debug hook: the first `trials` samples (7 indices each) a fresh generator draws for n matches */
int xrb_debug_fm_samples(int n, int trials, int32_t *out) {
    if (n < kMin || trials < 0 || !out) return XRB_ERR_INVALID;
    std::vector<int> perm(n);
    for (int i = 0; i < n; ++i) perm[i] = i;
    Mt19937 *g = new Mt19937();
    g->seed(0u);
    for (int t = 0; t < trials; ++t) draw_sample(*g, perm.data(), (uint32_t)n, out + 7 * t);
    delete g;
    return XRB_OK;
}

}  // extern "C"
