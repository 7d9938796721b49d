/*
 * oracle/fmatrix_oracle.cpp — CPU restatement of XRSfM's geometric verification of a matched
 * image pair: LO-RANSAC fundamental matrix (the step that follows path M).
 *
 * TEST INFRASTRUCTURE ONLY: the checker of xrsfm_b200/csrc/fm_ransac.cu (tests/test_fm_gpu.py).
 * Nothing under xrsfm_b200/ may link, import or execute this file.
 *
 * PARITY STATUS: "parity unpinned".  The reference ships no tests or golden vectors for this
 * path, and its own implementation (vendored COLMAP code) needs Eigen, which is not in this
 * image, so it cannot be compiled here.  What differs from the reference by construction:
 * the singular vectors / polynomial roots come from our own Jacobi SVD and closed-form cubic
 * instead of Eigen::JacobiSVD / Eigen::EigenSolver — equal up to rounding (and up to the order
 * in which the up-to-three 7-point models are visited).  The random sample sequence IS the
 * reference's when built against the same libstdc++: std::mt19937 seeded with 0 and
 * std::uniform_int_distribution<uint32_t>, partial Fisher-Yates on a persistent permutation.
 *
 * What is restated (paths relative to the reference tree):
 *   SolveFundamnetalCOLMAP          src/geometry/epipolar_geometry.hpp:10-27 (options: max_error 4,
 *                                   max/min trials 10000/100, confidence 0.999, min_inlier_ratio 0.25)
 *   RANSAC ctor, ComputeNumTrials   src/geometry/colmap/optim/ransac.h:136-167
 *   LORANSAC::Estimate              src/geometry/colmap/optim/loransac.h:96-238
 *   RandomSampler, Shuffle, PRNG    optim/random_sampler.cc:41-62, util/random.h:86-122,
 *                                   util/random.cc:36-50 (the seed is forced to 0 at :44)
 *   InlierSupportMeasurer           optim/support_measurement.cc:36-62
 *   7-point / 8-point estimators    estimators/fundamental_matrix.cc:46-199
 *   ComputeSquaredSampsonError      estimators/fundamental_matrix.cc:201-248
 *   CenterAndNormalizeImagePoints   estimators/fundamental_matrix.cc:250-295
 *   FindPolynomialRootsCompanionMatrix  estimators/polynomial.cc:208-275 (semantics: leading
 *                                   zeros removed, complex roots reported with their imaginary part)
 *   caller-side acceptance          src/feature/feature_processing.cc:225-227,283-296
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <random>
#include <vector>

namespace {

// ---- small dense linear algebra (no Eigen in this image) -----------------------------------
// One-sided (Hestenes) Jacobi: A is m x n row-major, n <= 9.  On return V (n x n, row-major,
// columns = right singular vectors) is sorted by descending singular value, like
// Eigen::JacobiSVD.  Rows are padded with zeros to n when m < n so that V is complete.
void jacobi_right_singular(const double *A, int m, int n, double *V, double *sv) {
    const int mm = std::max(m, n);
    std::vector<double> U((size_t)mm * n, 0.0);
    for (int r = 0; r < m; ++r)
        for (int c = 0; c < n; ++c) U[(size_t)r * n + c] = A[(size_t)r * n + c];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) V[i * n + j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        bool rotated = false;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                double a = 0, b = 0, g = 0;
                for (int r = 0; r < mm; ++r) {
                    const double up = U[(size_t)r * n + p], uq = U[(size_t)r * n + q];
                    a += up * up, b += uq * uq, g += up * uq;
                }
                if (g == 0.0 || std::fabs(g) <= 1e-300 + 2.3e-16 * std::sqrt(a * b)) continue;
                rotated = true;
                const double zeta = (b - a) / (2.0 * g);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
                const double c = 1.0 / std::sqrt(1.0 + t * t), s = c * t;
                for (int r = 0; r < mm; ++r) {
                    const double up = U[(size_t)r * n + p], uq = U[(size_t)r * n + q];
                    U[(size_t)r * n + p] = c * up - s * uq;
                    U[(size_t)r * n + q] = s * up + c * uq;
                }
                for (int r = 0; r < n; ++r) {
                    const double vp = V[r * n + p], vq = V[r * n + q];
                    V[r * n + p] = c * vp - s * vq;
                    V[r * n + q] = s * vp + c * vq;
                }
            }
        if (!rotated) break;
    }
    std::vector<double> norm(n);
    std::vector<int> order(n);
    for (int c = 0; c < n; ++c) {
        double s = 0;
        for (int r = 0; r < mm; ++r) s += U[(size_t)r * n + c] * U[(size_t)r * n + c];
        norm[c] = std::sqrt(s);
    }
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return norm[x] > norm[y]; });
    std::vector<double> Vs((size_t)n * n);
    for (int k = 0; k < n; ++k) {
        sv[k] = norm[order[k]];
        for (int r = 0; r < n; ++r) Vs[r * n + k] = V[r * n + order[k]];
    }
    std::memcpy(V, Vs.data(), sizeof(double) * n * n);
}

// Real and imaginary parts of the roots of c[0] x^d + ... + c[d] (d <= 3 after leading zeros
// are dropped).  Returns the number of roots, or -1 when the polynomial is constant.
int poly_roots(const double *coeffs_all, int n_coeffs, double *re, double *im) {
    int lead = 0;
    while (lead < n_coeffs && coeffs_all[lead] == 0.0) ++lead;  // RemoveLeadingZeros
    const double *c = coeffs_all + lead;
    const int d = n_coeffs - lead - 1;
    if (d <= 0) return -1;
    if (d == 1) {
        re[0] = -c[1] / c[0], im[0] = 0;
        return 1;
    }
    if (d == 2) {
        const double a = c[0], b = c[1], cc = c[2], disc = b * b - 4 * a * cc;
        if (disc >= 0) {
            const double sq = std::sqrt(disc);
            const double q = -0.5 * (b + (b >= 0 ? sq : -sq));
            re[0] = q / a, re[1] = q != 0 ? cc / q : 0.0, im[0] = im[1] = 0;
        } else {
            re[0] = re[1] = -b / (2 * a);
            im[0] = std::sqrt(-disc) / (2 * a), im[1] = -im[0];
        }
        return 2;
    }
    // cubic: x^3 + a x^2 + b x + c0 = 0, depressed by x = y - a/3
    const double a = c[1] / c[0], b = c[2] / c[0], c0 = c[3] / c[0];
    const double p = b - a * a / 3.0, q = 2.0 * a * a * a / 27.0 - a * b / 3.0 + c0;
    const double disc = q * q / 4.0 + p * p * p / 27.0;
    auto polish = [&](double x) {  // two Newton steps on the original cubic
        for (int it = 0; it < 2; ++it) {
            const double f = ((x + a) * x + b) * x + c0, df = (3.0 * x + 2.0 * a) * x + b;
            if (df != 0.0) x -= f / df;
        }
        return x;
    };
    if (disc > 0) {  // one real root, a complex-conjugate pair
        const double sq = std::sqrt(disc);
        const double u = std::cbrt(-q / 2.0 + sq), v = std::cbrt(-q / 2.0 - sq);
        re[0] = polish(u + v - a / 3.0), im[0] = 0;
        re[1] = re[2] = -(u + v) / 2.0 - a / 3.0;
        im[1] = (u - v) * std::sqrt(3.0) / 2.0, im[2] = -im[1];
    } else {  // three real roots (trigonometric form)
        const double r = std::sqrt(std::max(0.0, -p / 3.0));
        double arg = r > 0 ? (-q / 2.0) / (r * r * r) : 0.0;
        arg = std::max(-1.0, std::min(1.0, arg));
        const double phi = std::acos(arg);
        for (int k = 0; k < 3; ++k) {
            re[k] = polish(2.0 * r * std::cos((phi - 2.0 * M_PI * k) / 3.0) - a / 3.0);
            im[k] = 0;
        }
    }
    return 3;
}

struct Mat3 {
    double m[9];  // row-major
};

// FundamentalMatrixSevenPointEstimator::Estimate (fundamental_matrix.cc:46-139)
int seven_point(const double *p1, const double *p2, Mat3 *models) {
    double A[7 * 9];
    for (int i = 0; i < 7; ++i) {
        const double x0 = p1[2 * i], y0 = p1[2 * i + 1], x1 = p2[2 * i], y1 = p2[2 * i + 1];
        double *a = A + 9 * i;
        a[0] = x1 * x0, a[1] = x1 * y0, a[2] = x1, a[3] = y1 * x0, a[4] = y1 * y0, a[5] = y1, a[6] = x0, a[7] = y0, a[8] = 1;
    }
    double V[81], sv[9], f1[9], f2[9];
    jacobi_right_singular(A, 7, 9, V, sv);
    for (int k = 0; k < 9; ++k) f1[k] = V[k * 9 + 7], f2[k] = V[k * 9 + 8];
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
    double re[3], im[3];
    const int nr = poly_roots(co, 4, re, im);
    if (nr < 0) return 0;
    int n_models = 0;
    for (int i = 0; i < nr; ++i) {
        if (std::fabs(im[i]) > 1e-10) continue;  // kMaxRootImag
        const double lambda = re[i];
        double F[9];  // the 9-vector, then "resize(3,3)" of a column-major Eigen matrix and a transpose:
        for (int k = 0; k < 9; ++k) F[k] = lambda * f1[k] + f2[k];  // mu = 1
        // MatrixXd(1x9).resize(3,3) reinterprets column-major: G(r, c) = F[c*3 + r]; the model is
        // G^T, i.e. model(r, c) = F[r*3 + c]; the test is on G(2,2) = F[8]
        if (std::fabs(F[8]) < 1e-10) continue;  // kEps
        for (int k = 0; k < 9; ++k) models[n_models].m[k] = F[k] / F[8];
        ++n_models;
    }
    return n_models;
}

// CenterAndNormalizeImagePoints (fundamental_matrix.cc:250-295)
void center_normalize(const double *p, int n, std::vector<double> *out, double M[9]) {
    double cx = 0, cy = 0;
    for (int i = 0; i < n; ++i) cx += p[2 * i], cy += p[2 * i + 1];
    cx /= n, cy /= n;
    double rms = 0;
    for (int i = 0; i < n; ++i) {
        const double dx = p[2 * i] - cx, dy = p[2 * i + 1] - cy;
        rms += dx * dx + dy * dy;
    }
    rms = std::sqrt(rms / n);
    const double nf = std::sqrt(2.0) / rms;
    const double Mloc[9] = {nf, 0, -nf * cx, 0, nf, -nf * cy, 0, 0, 1};
    std::memcpy(M, Mloc, sizeof(Mloc));
    out->resize((size_t)2 * n);
    for (int i = 0; i < n; ++i) {
        const double x = p[2 * i], y = p[2 * i + 1];
        const double n0 = M[0] * x + M[1] * y + M[2], n1 = M[3] * x + M[4] * y + M[5], n2 = M[6] * x + M[7] * y + M[8];
        const double inv = 1.0 / n2;
        (*out)[2 * i] = n0 * inv, (*out)[2 * i + 1] = n1 * inv;
    }
}

// FundamentalMatrixEightPointEstimator::Estimate (fundamental_matrix.cc:147-192)
void eight_point(const double *p1, const double *p2, int n, Mat3 *model) {
    std::vector<double> q1, q2;
    double M1[9], M2[9];
    center_normalize(p1, n, &q1, M1);
    center_normalize(p2, n, &q2, M2);
    std::vector<double> C((size_t)n * 9);
    for (int i = 0; i < n; ++i) {
        const double x1 = q1[2 * i], y1 = q1[2 * i + 1], x2 = q2[2 * i], y2 = q2[2 * i + 1];
        double *c = C.data() + (size_t)9 * i;
        c[0] = x1 * x2, c[1] = y1 * x2, c[2] = x2, c[3] = x1 * y2, c[4] = y1 * y2, c[5] = y2, c[6] = x1, c[7] = y1, c[8] = 1;
    }
    double V[81], sv[9];
    jacobi_right_singular(C.data(), n, 9, V, sv);
    // nullspace vector viewed as a column-major 3x3 "ematrix_t"; E = ematrix_t^T, so E(r, c) = v[r*3 + c]
    double E[9];
    for (int k = 0; k < 9; ++k) E[k] = V[k * 9 + 8];
    // rank 2: F = U diag(s0, s1, 0) V^T = E - E v2 v2^T with v2 the right singular vector of the
    // smallest singular value
    double V3[9], s3[3];
    jacobi_right_singular(E, 3, 3, V3, s3);
    const double v2[3] = {V3[2], V3[5], V3[8]};
    double F[9];
    for (int r = 0; r < 3; ++r) {
        const double ev = E[r * 3] * v2[0] + E[r * 3 + 1] * v2[1] + E[r * 3 + 2] * v2[2];
        for (int c = 0; c < 3; ++c) F[r * 3 + c] = E[r * 3 + c] - ev * v2[c];
    }
    // points2_norm_matrix^T * F * points1_norm_matrix
    double T[9];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += M2[k * 3 + r] * F[k * 3 + c];
            T[r * 3 + c] = s;
        }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += T[r * 3 + k] * M1[k * 3 + c];
            model->m[r * 3 + c] = s;
        }
}

// ComputeSquaredSampsonError (fundamental_matrix.cc:201-248)
void sampson(const double *p1, const double *p2, int n, const double *E, double *res) {
    for (int i = 0; i < n; ++i) {
        const double x10 = p1[2 * i], x11 = p1[2 * i + 1], x20 = p2[2 * i], x21 = p2[2 * i + 1];
        const double Ex0 = E[0] * x10 + E[1] * x11 + E[2], Ex1 = E[3] * x10 + E[4] * x11 + E[5], Ex2 = E[6] * x10 + E[7] * x11 + E[8];
        const double Et0 = E[0] * x20 + E[3] * x21 + E[6], Et1 = E[1] * x20 + E[4] * x21 + E[7];
        const double x2tEx1 = x20 * Ex0 + x21 * Ex1 + Ex2;
        res[i] = x2tEx1 * x2tEx1 / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1);
    }
}

struct Support {  // support_measurement.h:45-51
    size_t num_inliers = 0;
    double residual_sum = std::numeric_limits<double>::max();
};
Support evaluate(const std::vector<double> &res, double max_residual) {
    Support s;
    s.num_inliers = 0, s.residual_sum = 0;
    for (const double r : res)
        if (r <= max_residual) s.num_inliers += 1, s.residual_sum += r;
    return s;
}
bool better(const Support &a, const Support &b) {
    if (a.num_inliers > b.num_inliers) return true;
    return a.num_inliers == b.num_inliers && a.residual_sum < b.residual_sum;
}

size_t compute_num_trials(size_t num_inliers, size_t num_samples, double confidence, int k_min) {  // ransac.h:151-167
    const double inlier_ratio = num_inliers / static_cast<double>(num_samples);
    const double nom = 1 - confidence;
    if (nom <= 0) return std::numeric_limits<size_t>::max();
    const double denom = 1 - std::pow(inlier_ratio, k_min);
    if (denom <= 0) return 1;
    return static_cast<size_t>(std::ceil(std::log(nom) / std::log(denom)));
}

}  // namespace

extern "C" {

struct xro_fm_options {  // colmap::RANSACOptions as SolveFundamnetalCOLMAP sets them
    double max_error, min_inlier_ratio, confidence;
    int64_t min_num_trials, max_num_trials;
};

struct xro_fm_report {
    int32_t success, best_is_local;
    int64_t num_trials, num_inliers;
    double residual_sum;
    double F[9];  // row-major
};

void xro_fm_default_options(xro_fm_options *o) {  // epipolar_geometry.hpp:13-18
    o->max_error = 4.0, o->max_num_trials = 10000, o->min_num_trials = 100, o->confidence = 0.999, o->min_inlier_ratio = 0.25;
}

// thread_local std::mt19937 of util/random.cc: one per OpenMP thread in the reference, seeded
// with 0 whatever is asked for (:44), advancing across the pairs that thread processes
void *xro_prng_create(void) { return new std::mt19937(0u); }
void xro_prng_destroy(void *p) { delete static_cast<std::mt19937 *>(p); }

int64_t xro_ransac_num_trials(int64_t num_inliers, int64_t num_samples, double confidence, int k_min) {
    const size_t n = compute_num_trials((size_t)num_inliers, (size_t)num_samples, confidence, k_min);
    return n > (size_t)INT64_MAX ? INT64_MAX : (int64_t)n;
}

int xro_fm_seven_point(const double *p1, const double *p2, double *models_out /* [3][9] */) {
    Mat3 m[3];
    const int n = seven_point(p1, p2, m);
    for (int k = 0; k < n; ++k) std::memcpy(models_out + 9 * k, m[k].m, sizeof(m[k].m));
    return n;
}

void xro_fm_eight_point(int n, const double *p1, const double *p2, double *F) {
    Mat3 m;
    eight_point(p1, p2, n, &m);
    std::memcpy(F, m.m, sizeof(m.m));
}

void xro_fm_sampson(int n, const double *p1, const double *p2, const double *F, double *res) { sampson(p1, p2, n, F, res); }

// LORANSAC<SevenPoint, EightPoint>::Estimate (loransac.h:96-238).  samples_out (optional,
// [7 * max_samples_out]) receives the sample indices of the first trials, for tests.
int xro_fm_loransac(void *prng, const xro_fm_options *opt_in, int n, const double *p1, const double *p2,
                    xro_fm_report *rep, char *inlier_mask, int32_t *samples_out, int max_samples_out) {
    constexpr int kMin = 7, kMinLocal = 8;
    std::mt19937 &gen = *static_cast<std::mt19937 *>(prng);
    xro_fm_options opt = *opt_in;
    {   // RANSAC ctor (ransac.h:136-148)
        const size_t kNumSamples = 100000;
        const size_t dyn = compute_num_trials(static_cast<size_t>(opt.min_inlier_ratio * kNumSamples), kNumSamples,
                                              opt.confidence, kMin);
        opt.max_num_trials = (int64_t)std::min<size_t>((size_t)opt.max_num_trials, dyn);
    }
    std::memset(rep, 0, sizeof(*rep));
    rep->residual_sum = std::numeric_limits<double>::max();
    if (n < kMin) return 0;
    const size_t num_samples = (size_t)n;
    Support best;
    Mat3 best_model{};
    bool best_is_local = false, abort = false;
    const double max_residual = opt.max_error * opt.max_error;
    std::vector<double> residuals(num_samples), xin, yin;
    std::vector<size_t> idxs(num_samples);  // RandomSampler::Initialize
    std::iota(idxs.begin(), idxs.end(), 0);
    double xr[14], yr[14];
    size_t max_num_trials = (size_t)opt.max_num_trials;  // min with sampler.MaxNumSamples() = SIZE_MAX
    size_t dyn_max_num_trials = max_num_trials;
    size_t num_trials = 0;
    for (num_trials = 0; num_trials < max_num_trials; ++num_trials) {
        if (abort) {
            num_trials += 1;
            break;
        }
        {   // RandomSampler::Sample -> Shuffle(7, &sample_idxs_) (util/random.h:115-122)
            const uint32_t last_idx = static_cast<uint32_t>(idxs.size() - 1);
            for (uint32_t i = 0; i < (uint32_t)kMin; ++i) {
                std::uniform_int_distribution<uint32_t> distribution(i, last_idx);
                const uint32_t j = distribution(gen);
                std::swap(idxs[i], idxs[j]);
            }
            for (int i = 0; i < kMin; ++i) {
                xr[2 * i] = p1[2 * idxs[i]], xr[2 * i + 1] = p1[2 * idxs[i] + 1];
                yr[2 * i] = p2[2 * idxs[i]], yr[2 * i + 1] = p2[2 * idxs[i] + 1];
                if (samples_out && (int64_t)num_trials < max_samples_out) samples_out[7 * num_trials + i] = (int32_t)idxs[i];
            }
        }
        Mat3 models[3];
        const int n_models = seven_point(xr, yr, models);
        for (int mi = 0; mi < n_models; ++mi) {
            sampson(p1, p2, n, models[mi].m, residuals.data());
            const Support support = evaluate(residuals, max_residual);
            if (better(support, best)) {
                best = support, best_model = models[mi], best_is_local = false;
                if (support.num_inliers > (size_t)kMin && support.num_inliers >= (size_t)kMinLocal) {
                    xin.clear(), yin.clear();
                    for (size_t i = 0; i < residuals.size(); ++i)
                        if (residuals[i] <= max_residual) {
                            xin.push_back(p1[2 * i]), xin.push_back(p1[2 * i + 1]);
                            yin.push_back(p2[2 * i]), yin.push_back(p2[2 * i + 1]);
                        }
                    Mat3 local;
                    eight_point(xin.data(), yin.data(), (int)(xin.size() / 2), &local);
                    sampson(p1, p2, n, local.m, residuals.data());
                    const Support ls = evaluate(residuals, max_residual);
                    if (better(ls, best)) best = ls, best_model = local, best_is_local = true;
                }
                dyn_max_num_trials = compute_num_trials(best.num_inliers, num_samples, opt.confidence, kMin);
            }
            if (num_trials >= dyn_max_num_trials && num_trials >= (size_t)opt.min_num_trials) {
                abort = true;
                break;
            }
        }
    }
    rep->num_trials = (int64_t)num_trials;
    rep->num_inliers = (int64_t)best.num_inliers;
    rep->residual_sum = best.residual_sum;
    rep->best_is_local = best_is_local;
    std::memcpy(rep->F, best_model.m, sizeof(best_model.m));
    if (best.num_inliers < (size_t)kMin) return 0;  // "No valid model was found"
    rep->success = 1;
    sampson(p1, p2, n, rep->F, residuals.data());
    if (inlier_mask)
        for (size_t i = 0; i < residuals.size(); ++i) inlier_mask[i] = residuals[i] <= max_residual ? 1 : 0;
    return 1;
}

// The caller's acceptance (feature_processing.cc:225-227, 260-296): returns the number of
// matches kept (0 = pair dropped) and compacts the inlier matches in place.
int xro_fm_filter_pair(int n_matches, int num_inliers, const char *inlier_mask, int32_t (*matches)[2]) {
    constexpr int min_num_matches = 15, min_num_inlier = 15;
    constexpr double min_ratio_inlier = 0.25;
    if (n_matches < min_num_matches) return 0;
    const int inlier_threshold = std::max(min_num_inlier, (int)(min_ratio_inlier * n_matches));
    if (num_inliers < inlier_threshold) return 0;
    int j = 0;
    for (int i = 0; i < n_matches; ++i)
        if (inlier_mask[i]) matches[j][0] = matches[i][0], matches[j][1] = matches[i][1], ++j;
    return j;
}

}  // extern "C"
