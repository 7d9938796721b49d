// ba_kernels.cu — per-observation model and the point-major LM kernels of path B.
// See ba_kernels.cuh for the pipeline and DESIGN.md §B for layout and rooflines.
#include <cuda_runtime.h>

#include <cfloat>
#include <climits>

#include "ba_kernels.cuh"
#include "ba_model.cuh"

namespace xrb {

// sum over the LPP lanes that share one point (LPP = 32: the warp, 16: a half warp)
template <int LPP>
__device__ __forceinline__ double group_sum(double v, unsigned gmask) {
#pragma unroll
    for (int d = LPP / 2; d > 0; d >>= 1) v += __shfl_xor_sync(gmask, v, d);
    return v;
}

__device__ __forceinline__ void atomic_max_nonneg(double *addr, double v) {
    // non-negative doubles order like their bit patterns
    atomicMax(reinterpret_cast<unsigned long long *>(addr),
              (unsigned long long)__double_as_longlong(v));
}

// Scaled, robustified blocks of one observation: Jc[12] = [Jd | Jt] (2x6, zero where the
// camera block is constant), JX scaled by the point's Jacobi scale (zero for constant points).
struct LinObs {
    double Jc[12];
    double JX[6];
    double r0, r1;
    int cols[6];
    bool active;
};

__device__ __forceinline__ void lin_obs(const BAProblemDev &P, const BAStateDev &x,
                                        const BAConsts &k, const BALinSys &L, int o, int p,
                                        bool pvar, LinObs &lo) {
    const int c = P.obs_cam[o];
    const int cq = P.colq[c], ct = P.colt[c];
    lo.active = cq >= 0 || ct >= 0 || pvar;
#pragma unroll
    for (int j = 0; j < 3; ++j) lo.cols[j] = cq >= 0 ? cq + j : -1, lo.cols[3 + j] = ct >= 0 ? ct + j : -1;
    if (!lo.active) return;
    const int ii = P.cam_intr[c];
    Obs e;
    eval_obs<true>(x.q + 4 * (size_t)c, x.t + 3 * (size_t)c, x.X + 3 * (size_t)p, P.intr_model[ii],
                   P.intr + 8 * (size_t)ii, P.obs_uv[2 * (size_t)o], P.obs_uv[2 * (size_t)o + 1], k,
                   true, e);
    lo.r0 = e.r0, lo.r1 = e.r1;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            lo.Jc[r * 6 + j] = cq >= 0 ? e.Jd[r * 3 + j] * L.sc[cq + j] : 0.0;
            lo.Jc[r * 6 + 3 + j] = ct >= 0 ? e.Jt[r * 3 + j] * L.sc[ct + j] : 0.0;
            lo.JX[r * 3 + j] = pvar ? e.JX[r * 3 + j] * L.sp[3 * (size_t)p + j] : 0.0;
        }
}

constexpr int kWarpsPerCta = 4;
constexpr int kChunk = 32;  // observations of one point handled per pass (one per lane)

// =====================================================================================
// 1. Jacobi scaling (iteration 0): squared column norms of the robustified Jacobian.
//    Point columns are local to the point; camera columns are accumulated with atomics and
//    finished (after the all-reduce in multi-GPU mode) by k_finish_scaling.
// =====================================================================================
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_colnorm(BAProblemDev P, BAStateDev x, BAConsts k, BALinSys L) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int p = warp; p < P.n_pts_local; p += nwarps) {
        const int k0 = P.pt_ptr[p], kn = P.pt_ptr[p + 1] - k0;
        const bool pvar = P.pt_var[p] != 0;
        double n0 = 0.0, n1 = 0.0, n2 = 0.0;
        for (int base = 0; base < kn; base += kChunk) {
            const int i = base + lane;
            if (i < kn) {
                LinObs lo;
                lin_obs(P, x, k, L, k0 + i, p, pvar, lo);  // sc == sp == 1 at this point
                if (lo.active) {
#pragma unroll
                    for (int j = 0; j < 6; ++j)
                        if (lo.cols[j] >= 0)
                            atomicAdd(&L.n2c[lo.cols[j]], lo.Jc[j] * lo.Jc[j] + lo.Jc[6 + j] * lo.Jc[6 + j]);
                    n0 += lo.JX[0] * lo.JX[0] + lo.JX[3] * lo.JX[3];
                    n1 += lo.JX[1] * lo.JX[1] + lo.JX[4] * lo.JX[4];
                    n2 += lo.JX[2] * lo.JX[2] + lo.JX[5] * lo.JX[5];
                }
            }
        }
        n0 = warp_sum(n0), n1 = warp_sum(n1), n2 = warp_sum(n2);
        if (lane == 0 && pvar) {
            L.sp[3 * (size_t)p + 0] = 1.0 / (1.0 + sqrt(n0));
            L.sp[3 * (size_t)p + 1] = 1.0 / (1.0 + sqrt(n1));
            L.sp[3 * (size_t)p + 2] = 1.0 / (1.0 + sqrt(n2));
        }
    }
}

__global__ void k_finish_scaling(int nc, const double *__restrict__ n2c, double *__restrict__ sc) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < nc) sc[j] = 1.0 / (1.0 + sqrt(n2c[j]));
}

static int point_grid(int n_pts) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int want = (n_pts + kWarpsPerCta - 1) / kWarpsPerCta;
    const int cap = sms * 16;  // persistent-ish: a multiple of the SM count
    return want < cap ? (want > 0 ? want : 1) : cap;
}

int ba_launch_colnorm(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                      const BALinSys &L, cudaStream_t st) {
    if (P.n_pts_local > 0) {
        k_colnorm<<<point_grid(P.n_pts_local), kWarpsPerCta * 32, 0, st>>>(P, x, k, L);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

int ba_launch_finish_scaling(const BAProblemDev &P, const BALinSys &L, cudaStream_t st) {
    if (P.nc > 0) {
        k_finish_scaling<<<(P.nc + 255) / 256, 256, 0, st>>>(P.nc, L.n2c, L.sc);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// =====================================================================================
// 2. Schur complement without atomics.
//   k_lin         point-major, one warp per point: phases A and B as in k_schur, then ONE
//                 18-double record per observation, Tt_i = W_i (V + D^2)^-1/2, plus per
//                 point (V + D^2)^-1 and h = (V + D^2)^-1 g.        (streaming writes only)
//   k_gather      8 lanes per off-diagonal block (a, b): S_ab = - sum_{p in ab} Tt_i Tt_j^T
//                 over the precomputed incidence list (ba_struct.cu), plain stores.
//   k_cam_blocks  one CTA per camera: U_c = sum Jc^T Jc, g_c = sum Jc^T r,
//                 rhs_c = sum Jc^T (r - JX h), Ud_c = sum Tt_i Tt_i^T, block-reduced.  It never
//                 touches S, so it runs next to k_gather on a second stream; k_cam_diag (after the
//                 exchange) folds U - Ud + D^2 into the diagonal blocks.
// The result is deterministic (fixed summation order) and S is written exactly once.
// =====================================================================================
template <int LPP>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_lin(BAProblemDev P, BAStateDev x, BAConsts k, BALinSys L, double inv_radius, double *__restrict__ scalars) {
    // LPP lanes per point: a 32-lane warp serves 32 / LPP points at once (short tracks waste
    // most of a full warp); `lane` below is the lane within the point's group
    constexpr int kGroups = 32 / LPP;
    const int lane = threadIdx.x & (LPP - 1);
    const unsigned gmask = LPP == 32 ? 0xFFFFFFFFu : (((1u << LPP) - 1u) << (((threadIdx.x & 31) / LPP) * LPP));
    const int warp = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kGroups + ((threadIdx.x & 31) / LPP);
    const int nwarps = ((gridDim.x * blockDim.x) >> 5) * kGroups;
    double gmax = 0.0;
    for (int p = warp; p < P.n_pts_local; p += nwarps) {
        const int k0 = P.pt_ptr[p], kn = P.pt_ptr[p + 1] - k0;
        if (kn == 0) continue;
        const bool pvar = P.pt_var[p] != 0;
        const int nchunks = (kn + LPP - 1) / LPP;
        double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
        LinObs lo;
        lo.active = false;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int i = ch * LPP + lane;
            lo.active = false;
            if (i < kn) lin_obs(P, x, k, L, k0 + i, p, pvar, lo);
            if (lo.active) {
                const double *J = lo.JX;
                V[0] += J[0] * J[0] + J[3] * J[3], V[1] += J[0] * J[1] + J[3] * J[4];
                V[2] += J[0] * J[2] + J[3] * J[5], V[3] += J[1] * J[1] + J[4] * J[4];
                V[4] += J[1] * J[2] + J[4] * J[5], V[5] += J[2] * J[2] + J[5] * J[5];
                g[0] += J[0] * lo.r0 + J[3] * lo.r1;
                g[1] += J[1] * lo.r0 + J[4] * lo.r1;
                g[2] += J[2] * lo.r0 + J[5] * lo.r1;
            }
        }
#pragma unroll
        for (int j = 0; j < 6; ++j) V[j] = group_sum<LPP>(V[j], gmask);
#pragma unroll
        for (int j = 0; j < 3; ++j) g[j] = group_sum<LPP>(g[j], gmask);
        // (V + D^2) = Lc Lc^T (3x3 Cholesky); Tt = W Lc^-T so that Tt_i Tt_j^T = W_i (V+D^2)^-1 W_j^T
        double l00 = 1, l10 = 0, l11 = 1, l20 = 0, l21 = 0, l22 = 1;
        if (pvar) {
            const double a = V[0] + fmin(fmax(V[0], 1e-6), 1e32) * inv_radius;
            const double d = V[3] + fmin(fmax(V[3], 1e-6), 1e32) * inv_radius;
            const double f = V[5] + fmin(fmax(V[5], 1e-6), 1e32) * inv_radius;
            const double b = V[1], c = V[2], e = V[4];
            l00 = sqrt(a), l10 = b / l00, l20 = c / l00;
            const double t11 = d - l10 * l10;
            l11 = sqrt(t11), l21 = (e - l20 * l10) / l11;
            const double t22 = f - l20 * l20 - l21 * l21;
            l22 = sqrt(t22);
            const bool bad = !(a > 0.0) || !(t11 > 0.0) || !(t22 > 0.0) || !isfinite(l22);
            if (lane == 0) {
                if (bad) scalars[SC_FAIL] = 1.0;
                // inverse through the factor: Vi = Lc^-T Lc^-1
                const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
                const double m10 = -l10 * i00 * i11, m21 = -l21 * i11 * i22;
                const double m20 = (l10 * l21 - l20 * l11) * i00 * i11 * i22;  // (Lc^-1)[2][0]
                double Vi[6];
                Vi[0] = i00 * i00 + m10 * m10 + m20 * m20, Vi[1] = m10 * i11 + m20 * m21, Vi[2] = m20 * i22;
                Vi[3] = i11 * i11 + m21 * m21, Vi[4] = m21 * i22, Vi[5] = i22 * i22;
#pragma unroll
                for (int j = 0; j < 6; ++j) L.Vinv[6 * (size_t)p + j] = Vi[j];
#pragma unroll
                for (int j = 0; j < 3; ++j) L.gp[3 * (size_t)p + j] = g[j];
                L.h[3 * (size_t)p + 0] = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2];
                L.h[3 * (size_t)p + 1] = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2];
                L.h[3 * (size_t)p + 2] = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
                const double *s = L.sp + 3 * (size_t)p;
                gmax = fmax(gmax, fmax(fabs(g[0] / s[0]), fmax(fabs(g[1] / s[1]), fabs(g[2] / s[2]))));
            }
        } else if (lane == 0) {
            L.h[3 * (size_t)p] = L.h[3 * (size_t)p + 1] = L.h[3 * (size_t)p + 2] = 0.0;
        }
        for (int ch = 0; ch < nchunks; ++ch) {
            const int i = ch * LPP + lane;
            if (nchunks > 1) {
                lo.active = false;
                if (i < kn) lin_obs(P, x, k, L, k0 + i, p, pvar, lo);
            }
            if (i < kn) {
                double2 *rec = reinterpret_cast<double2 *>(L.Tt + 18 * (size_t)(k0 + i));
                if (lo.active && pvar) {
                    double T[18];
#pragma unroll
                    for (int a = 0; a < 6; ++a) {
                        const double w0 = lo.Jc[a] * lo.JX[0] + lo.Jc[6 + a] * lo.JX[3];
                        const double w1 = lo.Jc[a] * lo.JX[1] + lo.Jc[6 + a] * lo.JX[4];
                        const double w2 = lo.Jc[a] * lo.JX[2] + lo.Jc[6 + a] * lo.JX[5];
                        const double t0 = w0 / l00;                          // t Lc^T = w
                        const double t1 = (w1 - l10 * t0) / l11;
                        const double t2 = (w2 - l20 * t0 - l21 * t1) / l22;
                        T[a * 3] = t0, T[a * 3 + 1] = t1, T[a * 3 + 2] = t2;
                    }
#pragma unroll
                    for (int j = 0; j < 9; ++j) rec[j] = make_double2(T[2 * j], T[2 * j + 1]);
                } else {
#pragma unroll
                    for (int j = 0; j < 9; ++j) rec[j] = make_double2(0.0, 0.0);
                }
            }
        }
    }
    if (lane == 0 && gmax > 0.0) atomic_max_nonneg(&scalars[SC_GRAD_MAX_PT], gmax);
}

int ba_launch_lin(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, const BALinSys &L,
                  double inv_radius, double *scalars, cudaStream_t st) {
    if (P.n_pts_local > 0) {
        // short tracks (mean <= 16 observations): two points per warp
        if ((long long)P.n_obs_local <= 16LL * P.n_pts_local)
            k_lin<16><<<point_grid((P.n_pts_local + 1) / 2), kWarpsPerCta * 32, 0, st>>>(P, x, k, L, inv_radius, scalars);
        else
            k_lin<32><<<point_grid(P.n_pts_local), kWarpsPerCta * 32, 0, st>>>(P, x, k, L, inv_radius, scalars);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

__device__ __forceinline__ void load_rec(const double *__restrict__ Tt, int o, double T[18]) {
    const double2 *r = reinterpret_cast<const double2 *>(Tt + 18 * (size_t)o);
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        const double2 v = __ldg(r + j);
        T[2 * j] = v.x, T[2 * j + 1] = v.y;
    }
}

// LPB lanes share one block: 8 for the short lists of an unordered scene (C2: ~70 incidences per block),
// a whole warp for the long lists of a sequential one (C4: ~1900 per block, few blocks).
template <int LPB>
__global__ void __launch_bounds__(256)
k_gather(BAProblemDev P, BALinSys L) {
    const int sub = threadIdx.x & (LPB - 1);
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) / LPB;
    const bool live = b < P.n_blocks;
    double acc[36];
#pragma unroll
    for (int j = 0; j < 36; ++j) acc[j] = 0.0;
    if (live) {
        const int i0 = P.blk_ptr[b], i1 = P.blk_ptr[b + 1];
        // the index pair of the next incidence is requested before the records of this one:
        // index -> record is a dependent load chain, one L2 round trip per link
        int2 nxt = i0 + sub < i1 ? __ldg(P.inc + i0 + sub) : make_int2(0, 0);
        for (int it = i0 + sub; it < i1; it += LPB) {
            const int2 oo = nxt;
            if (it + LPB < i1) nxt = __ldg(P.inc + it + LPB);
            double Ti[18], Tj[18];
            load_rec(L.Tt, oo.x, Ti);
            load_rec(L.Tt, oo.y, Tj);
#pragma unroll
            for (int a = 0; a < 6; ++a)
#pragma unroll
                for (int c = 0; c < 6; ++c)
                    acc[a * 6 + c] += Ti[a * 3] * Tj[c * 3] + Ti[a * 3 + 1] * Tj[c * 3 + 1] + Ti[a * 3 + 2] * Tj[c * 3 + 2];
        }
    }
#pragma unroll
    for (int j = 0; j < 36; ++j) {
        double v = acc[j];
#pragma unroll
        for (int d = 1; d < LPB; d <<= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        acc[j] = v;
    }
    if (!live) return;
    const int2 cams = P.blk_cams[b];
    int ra[6], cb[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        ra[j] = P.colq[cams.x] >= 0 ? P.colq[cams.x] + j : -1, ra[3 + j] = P.colt[cams.x] >= 0 ? P.colt[cams.x] + j : -1;
        cb[j] = P.colq[cams.y] >= 0 ? P.colq[cams.y] + j : -1, cb[3 + j] = P.colt[cams.y] >= 0 ? P.colt[cams.y] + j : -1;
    }
    const bool same = cams.x == cams.y;  // two observations of one camera on one point: M + M^T
    // the lanes of the group share the 36 stores
#pragma unroll
    for (int e = 0; e < 36; ++e) {
        if ((e & (LPB - 1)) != sub) continue;
        const int a = e / 6, c = e % 6;
        if (ra[a] < 0 || cb[c] < 0) continue;
        if (same) {
            if (c > a) continue;
            L.S[L.tm.at(ra[a], cb[c])] = -(acc[a * 6 + c] + acc[c * 6 + a]);
        } else {
            L.S[L.tm.at(ra[a], cb[c])] = -acc[e];
        }
    }
}

int ba_launch_gather(const BAProblemDev &P, const BALinSys &L, cudaStream_t st) {
    if (P.n_blocks > 0) {
        if (P.n_inc > 256LL * P.n_blocks) {
            const long long threads = (long long)P.n_blocks * 32;
            k_gather<32><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, L);
        } else {
            const long long threads = (long long)P.n_blocks * 8;
            k_gather<8><<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(P, L);
        }
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

__global__ void __launch_bounds__(128)
k_cam_blocks(BAProblemDev P, BAStateDev x, BAConsts k, BALinSys L) {
    const int c = blockIdx.x;
    const int cq = P.colq[c], ct = P.colt[c];
    if (cq < 0 && ct < 0) return;
    // 21 lower entries of (Jc^T Jc), 21 of (Tt Tt^T), 6 g_c, 6 rhs
    double acc[54];
#pragma unroll
    for (int j = 0; j < 54; ++j) acc[j] = 0.0;
    const int o0 = P.cam_ptr[c], o1 = P.cam_ptr[c + 1];
    for (int it = o0 + threadIdx.x; it < o1; it += 128) {
        const int o = P.cam_obs[it];
        const int p = P.obs_pt[o];
        const bool pvar = P.pt_var[p] != 0;
        LinObs lo;
        lin_obs(P, x, k, L, o, p, pvar, lo);
        if (!lo.active) continue;
        double T[18];
        load_rec(L.Tt, o, T);
        const double h0 = L.h[3 * (size_t)p], h1 = L.h[3 * (size_t)p + 1], h2 = L.h[3 * (size_t)p + 2];
        // r - JX h
        const double e0 = lo.r0 - (lo.JX[0] * h0 + lo.JX[1] * h1 + lo.JX[2] * h2);
        const double e1 = lo.r1 - (lo.JX[3] * h0 + lo.JX[4] * h1 + lo.JX[5] * h2);
        int w = 0;
#pragma unroll
        for (int a = 0; a < 6; ++a) {
#pragma unroll
            for (int b = 0; b <= a; ++b, ++w) {
                acc[w] += lo.Jc[a] * lo.Jc[b] + lo.Jc[6 + a] * lo.Jc[6 + b];
                acc[21 + w] += T[a * 3] * T[b * 3] + T[a * 3 + 1] * T[b * 3 + 1] + T[a * 3 + 2] * T[b * 3 + 2];
            }
            acc[42 + a] += lo.Jc[a] * lo.r0 + lo.Jc[6 + a] * lo.r1;
            acc[48 + a] += lo.Jc[a] * e0 + lo.Jc[6 + a] * e1;
        }
    }
    __shared__ double red[4][54];
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 54; ++j) {
        const double v = warp_sum(acc[j]);
        if (lane == 0) red[wrp][j] = v;
    }
    __syncthreads();
    if (threadIdx.x < 54) {
        const int j = threadIdx.x;
        const double v = red[0][j] + red[1][j] + red[2][j] + red[3][j];
        int cols[6];
#pragma unroll
        for (int q = 0; q < 3; ++q) cols[q] = cq >= 0 ? cq + q : -1, cols[3 + q] = ct >= 0 ? ct + q : -1;
        if (j < 42) {
            int w = j < 21 ? j : j - 21, a = 0;
            while (w >= a + 1) w -= a + 1, ++a;  // w-th lower entry -> (a, b = w)
            const int b = w;
            if (cols[a] >= 0 && cols[b] >= 0) {
                if (j < 21)
                    L.U[(size_t)cols[a] * 6 + b] = v;
                else
                    L.Ud[(size_t)cols[a] * 6 + b] = v;  // sum Tt Tt^T of the diagonal block: k_cam_diag subtracts it
            }
        } else if (j < 48) {
            if (cols[j - 42] >= 0) L.gc[cols[j - 42]] = v;
        } else {
            if (cols[j - 48] >= 0) L.rhs[cols[j - 48]] = v;
        }
    }
}

int ba_launch_cam_blocks(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, const BALinSys &L,
                         cudaStream_t st) {
    if (P.n_cams > 0) {
        k_cam_blocks<<<P.n_cams, 128, 0, st>>>(P, x, k, L);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// =====================================================================================
// 3a. Camera block-diagonal: S += U + D_c^2 with D_c^2 = clamp(diag U, 1e-6, 1e32)/radius,
//     and the camera part of the gradient max-norm ||x - Plus(x, -g)||_inf.
//     Runs after the exchange (U, gc are global sums).
// =====================================================================================
__global__ void k_cam_diag(BAProblemDev P, BAStateDev x, BALinSys L, double inv_radius,
                           double *__restrict__ scalars) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= P.n_cams) return;
    const int cq = P.colq[c], ct = P.colt[c];
    int cols[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) cols[j] = cq >= 0 ? cq + j : -1, cols[3 + j] = ct >= 0 ? ct + j : -1;
#pragma unroll
    for (int a = 0; a < 6; ++a) {
        if (cols[a] < 0) continue;
#pragma unroll
        for (int b = 0; b <= a; ++b) {
            if (cols[b] < 0) continue;
            double v = L.U[(size_t)cols[a] * 6 + b];
            if (a == b) v += fmin(fmax(v, 1e-6), 1e32) * inv_radius;
            L.S[L.tm.at(cols[a], cols[b])] += v - L.Ud[(size_t)cols[a] * 6 + b];
        }
    }
    double gmax = 0.0;
    if (cq >= 0) {
        double qn[4];
        const double *q = x.q + 4 * (size_t)c;
        quat_plus(q, -L.gc[cq] / L.sc[cq], -L.gc[cq + 1] / L.sc[cq + 1], -L.gc[cq + 2] / L.sc[cq + 2], qn);
#pragma unroll
        for (int j = 0; j < 4; ++j) gmax = fmax(gmax, fabs(q[j] - qn[j]));
    }
    if (ct >= 0)
#pragma unroll
        for (int j = 0; j < 3; ++j) gmax = fmax(gmax, fabs(L.gc[ct + j] / L.sc[ct + j]));
    if (gmax > 0.0) atomic_max_nonneg(&scalars[SC_GRAD_MAX_CAM], gmax);
}

int ba_launch_cam_diag(const BAProblemDev &P, const BAStateDev &x, const BALinSys &L,
                       double inv_radius, double *scalars, cudaStream_t st) {
    k_cam_diag<<<(P.n_cams + 127) / 128, 128, 0, st>>>(P, x, L, inv_radius, scalars);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// ||x||^2 over the variable parameter blocks of the start point (what the parameter-tolerance test of the LM
// loop needs) and the unit Jacobi scales of iteration 0, on the device: one CTA, fixed summation order.
__global__ void __launch_bounds__(1024)
k_start_norm(BAProblemDev P, BAStateDev x, BALinSys L, int with_cams, double *__restrict__ out) {
    double s = 0.0;
    if (with_cams)
        for (int c = threadIdx.x; c < P.n_cams; c += 1024) {
            if (P.colq[c] >= 0)
                for (int j = 0; j < 4; ++j) s += x.q[4 * (size_t)c + j] * x.q[4 * (size_t)c + j];
            if (P.colt[c] >= 0)
                for (int j = 0; j < 3; ++j) s += x.t[3 * (size_t)c + j] * x.t[3 * (size_t)c + j];
        }
    for (int p = threadIdx.x; p < P.n_pts_local; p += 1024)
        if (P.pt_var[p])
            for (int j = 0; j < 3; ++j) s += x.X[3 * (size_t)p + j] * x.X[3 * (size_t)p + j];
    for (int i = threadIdx.x; i < 3 * P.n_pts_local; i += 1024) L.sp[i] = 1.0;
    for (int i = threadIdx.x; i < P.nc; i += 1024) L.sc[i] = 1.0, L.n2c[i] = 0.0;
    __shared__ double red[32];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 32; ++w) t += red[w];
        *out = t;
    }
}

int ba_launch_start_norm(const BAProblemDev &P, const BAStateDev &x, const BALinSys &L, int with_cams, double *out,
                         cudaStream_t st) {
    k_start_norm<<<1, 1024, 0, st>>>(P, x, L, with_cams, out);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

__global__ void k_set_holes(BALinSys L, const int32_t *__restrict__ holes, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) L.S[L.tm.at(holes[i], holes[i])] = 1.0;
}

int ba_launch_set_holes(const BALinSys &L, const int32_t *holes, int n_holes, cudaStream_t st) {
    if (n_holes > 0) {
        k_set_holes<<<(n_holes + 255) / 256, 256, 0, st>>>(L, holes, n_holes);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// =====================================================================================
// 4. Back-substitution y_p = V^-1 (g_p - sum W_i^T y_c), candidate point, model cost change
//    -(J s)^T (r + J s / 2) and step / state norms, one warp per point.
// =====================================================================================
template <int LPP>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_backsub(BAProblemDev P, BAStateDev x, BAStateDev cand, BAConsts k, BALinSys L,
          const double *__restrict__ yc, double *__restrict__ step_p, double *__restrict__ scalars) {
    // LPP lanes per point: a 32-lane warp serves 32 / LPP points at once (short tracks waste
    // most of a full warp); `lane` below is the lane within the point's group
    constexpr int kGroups = 32 / LPP;
    const int lane = threadIdx.x & (LPP - 1);
    const unsigned gmask = LPP == 32 ? 0xFFFFFFFFu : (((1u << LPP) - 1u) << (((threadIdx.x & 31) / LPP) * LPP));
    const int warp = ((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * kGroups + ((threadIdx.x & 31) / LPP);
    const int nwarps = ((gridDim.x * blockDim.x) >> 5) * kGroups;
    double model = 0.0, sn2 = 0.0, xn2 = 0.0;
    for (int p = warp; p < P.n_pts_local; p += nwarps) {
        const int k0 = P.pt_ptr[p], kn = P.pt_ptr[p + 1] - k0;
        const bool pvar = P.pt_var[p] != 0;
        const int nchunks = (kn + LPP - 1) / LPP;
        double acc0 = 0, acc1 = 0, acc2 = 0;
        LinObs lo;
        lo.active = false;
        double jy0 = 0, jy1 = 0;
        for (int ch = 0; ch < nchunks; ++ch) {
            const int i = ch * LPP + lane;
            lo.active = false;
            if (i < kn) lin_obs(P, x, k, L, k0 + i, p, pvar, lo);
            jy0 = jy1 = 0.0;
            if (lo.active) {
#pragma unroll
                for (int a = 0; a < 6; ++a)
                    if (lo.cols[a] >= 0) {
                        const double y = yc[lo.cols[a]];
                        jy0 += lo.Jc[a] * y, jy1 += lo.Jc[6 + a] * y;
                    }
                acc0 += lo.JX[0] * jy0 + lo.JX[3] * jy1;
                acc1 += lo.JX[1] * jy0 + lo.JX[4] * jy1;
                acc2 += lo.JX[2] * jy0 + lo.JX[5] * jy1;
            }
        }
        acc0 = group_sum<LPP>(acc0, gmask), acc1 = group_sum<LPP>(acc1, gmask), acc2 = group_sum<LPP>(acc2, gmask);
        double y0 = 0, y1 = 0, y2 = 0;
        if (pvar) {
            const double *Vi = L.Vinv + 6 * (size_t)p;
            const double b0 = L.gp[3 * (size_t)p] - acc0, b1 = L.gp[3 * (size_t)p + 1] - acc1,
                         b2 = L.gp[3 * (size_t)p + 2] - acc2;
            y0 = Vi[0] * b0 + Vi[1] * b1 + Vi[2] * b2;
            y1 = Vi[1] * b0 + Vi[3] * b1 + Vi[4] * b2;
            y2 = Vi[2] * b0 + Vi[4] * b1 + Vi[5] * b2;
            if (lane == 0) {
                const double *s = L.sp + 3 * (size_t)p;
                const double *X = x.X + 3 * (size_t)p;
                double *Xc = cand.X + 3 * (size_t)p;
                const double d0 = -y0 * s[0], d1 = -y1 * s[1], d2 = -y2 * s[2];
                step_p[3 * (size_t)p] = -y0, step_p[3 * (size_t)p + 1] = -y1, step_p[3 * (size_t)p + 2] = -y2;
                Xc[0] = X[0] + d0, Xc[1] = X[1] + d1, Xc[2] = X[2] + d2;
                const double e0 = X[0] - Xc[0], e1 = X[1] - Xc[1], e2 = X[2] - Xc[2];
                sn2 += e0 * e0 + e1 * e1 + e2 * e2;
                xn2 += Xc[0] * Xc[0] + Xc[1] * Xc[1] + Xc[2] * Xc[2];
                if (!isfinite(y0) || !isfinite(y1) || !isfinite(y2)) scalars[SC_FAIL] = 1.0;
            }
        } else if (lane == 0) {
            const double *X = x.X + 3 * (size_t)p;
            double *Xc = cand.X + 3 * (size_t)p;
            Xc[0] = X[0], Xc[1] = X[1], Xc[2] = X[2];
        }
        // model residual m = J s = -(Jc yc + JX yp); contribution -(m . (r + m / 2))
        for (int ch = 0; ch < nchunks; ++ch) {
            const int i = ch * LPP + lane;
            if (nchunks > 1) {
                lo.active = false;
                if (i < kn) lin_obs(P, x, k, L, k0 + i, p, pvar, lo);
                jy0 = jy1 = 0.0;
                if (lo.active) {
#pragma unroll
                    for (int a = 0; a < 6; ++a)
                        if (lo.cols[a] >= 0) {
                            const double y = yc[lo.cols[a]];
                            jy0 += lo.Jc[a] * y, jy1 += lo.Jc[6 + a] * y;
                        }
                }
            }
            if (lo.active) {
                const double m0 = -(jy0 + lo.JX[0] * y0 + lo.JX[1] * y1 + lo.JX[2] * y2);
                const double m1 = -(jy1 + lo.JX[3] * y0 + lo.JX[4] * y1 + lo.JX[5] * y2);
                model -= m0 * (lo.r0 + 0.5 * m0) + m1 * (lo.r1 + 0.5 * m1);
            }
        }
    }
    model = group_sum<LPP>(model, gmask);
    if (lane == 0) {
        atomicAdd(&scalars[SC_MODEL_CHANGE], model);
        atomicAdd(&scalars[SC_STEP_NORM2], sn2);
        atomicAdd(&scalars[SC_XNORM2], xn2);
    }
}

int ba_launch_backsub(const BAProblemDev &P, const BAStateDev &x, const BAStateDev &cand,
                      const BAConsts &k, const BALinSys &L, const double *yc, double *step_p,
                      double *scalars, cudaStream_t st) {
    if (P.n_pts_local > 0) {
        if ((long long)P.n_obs_local <= 16LL * P.n_pts_local)
            k_backsub<16><<<point_grid((P.n_pts_local + 1) / 2), kWarpsPerCta * 32, 0, st>>>(P, x, cand, k, L, yc, step_p, scalars);
        else
            k_backsub<32><<<point_grid(P.n_pts_local), kWarpsPerCta * 32, 0, st>>>(P, x, cand, k, L, yc, step_p, scalars);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// Camera candidate: q+ = Plus(q, -y_q * scale), t+ = t - y_t * scale; norms of the camera part.
__global__ void k_cam_update(BAProblemDev P, BAStateDev x, BAStateDev cand, BALinSys L,
                             const double *__restrict__ yc, double *__restrict__ scalars,
                             int add_norms) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    double sn2 = 0.0, xn2 = 0.0;
    if (c < P.n_cams) {
        const int cq = P.colq[c], ct = P.colt[c];
        const double *q = x.q + 4 * (size_t)c, *t = x.t + 3 * (size_t)c;
        double *qc = cand.q + 4 * (size_t)c, *tc = cand.t + 3 * (size_t)c;
        if (cq >= 0) {
            double qn[4];
            quat_plus(q, -yc[cq] * L.sc[cq], -yc[cq + 1] * L.sc[cq + 1], -yc[cq + 2] * L.sc[cq + 2], qn);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                qc[j] = qn[j];
                sn2 += (q[j] - qn[j]) * (q[j] - qn[j]);
                xn2 += qn[j] * qn[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) qc[j] = q[j];
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            if (ct >= 0) {
                const double v = t[j] - yc[ct + j] * L.sc[ct + j];
                tc[j] = v;
                sn2 += (t[j] - v) * (t[j] - v);
                xn2 += v * v;
            } else {
                tc[j] = t[j];
            }
        }
    }
    if (add_norms) {
        sn2 = warp_sum(sn2), xn2 = warp_sum(xn2);
        if ((threadIdx.x & 31) == 0 && (sn2 != 0.0 || xn2 != 0.0)) {
            atomicAdd(&scalars[SC_STEP_NORM2], sn2);
            atomicAdd(&scalars[SC_XNORM2], xn2);
        }
    }
}

int ba_launch_cam_update(const BAProblemDev &P, const BAStateDev &x, const BAStateDev &cand,
                         const BALinSys &L, const double *yc, double *scalars, int add_norms,
                         cudaStream_t st) {
    k_cam_update<<<(P.n_cams + 127) / 128, 128, 0, st>>>(P, x, cand, L, yc, scalars, add_norms);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// =====================================================================================
// 5. Cost 1/2 sum rho over observations (thread per observation, point looked up by
//    binary search in pt_ptr so the observation stream stays coalesced).
// =====================================================================================
__device__ __forceinline__ int point_of_obs(const int32_t *__restrict__ pt_ptr, int n_pts, int o) {
    int lo = 0, hi = n_pts;  // largest p with pt_ptr[p] <= o
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (pt_ptr[mid] <= o)
            lo = mid;
        else
            hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(256)
k_cost(BAProblemDev P, BAStateDev x, BAConsts k, int mode, double *__restrict__ out) {
    double s = 0.0;
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < P.n_obs_local; o += gridDim.x * blockDim.x) {
        const int p = point_of_obs(P.pt_ptr, P.n_pts_local, o);
        const int c = P.obs_cam[o];
        const bool active = P.colq[c] >= 0 || P.colt[c] >= 0 || P.pt_var[p] != 0;
        if (active != (mode == 0)) continue;
        const int ii = P.cam_intr[c];
        Obs e;
        eval_obs<false>(x.q + 4 * (size_t)c, x.t + 3 * (size_t)c, x.X + 3 * (size_t)p, P.intr_model[ii],
                        P.intr + 8 * (size_t)ii, P.obs_uv[2 * (size_t)o], P.obs_uv[2 * (size_t)o + 1], k,
                        false, e);
        s += 0.5 * e.rho0;
    }
    __shared__ double wsum[8];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 8) {
        double v = wsum[threadIdx.x];
#pragma unroll
        for (int d = 4; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFu, v, d);
        if (threadIdx.x == 0) atomicAdd(out, v);
    }
}

int ba_launch_cost(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, int mode,
                   double *scalar_out, cudaStream_t st) {
    if (P.n_obs_local > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int want = (P.n_obs_local + 255) / 256;
        const int grid = want < sms * 8 ? want : sms * 8;
        k_cost<<<grid, 256, 0, st>>>(P, x, k, mode, scalar_out);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

__global__ void k_residuals(BAProblemDev P, BAStateDev x, BAConsts k,
                            const int32_t *__restrict__ obs_orig, double *__restrict__ out) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= P.n_obs_local) return;
    const int p = point_of_obs(P.pt_ptr, P.n_pts_local, o);
    const int c = P.obs_cam[o], ii = P.cam_intr[c];
    Obs e;
    eval_obs<false>(x.q + 4 * (size_t)c, x.t + 3 * (size_t)c, x.X + 3 * (size_t)p, P.intr_model[ii],
                    P.intr + 8 * (size_t)ii, P.obs_uv[2 * (size_t)o], P.obs_uv[2 * (size_t)o + 1], k, false, e);
    const size_t dst = obs_orig[o];
    out[2 * dst] = e.r0, out[2 * dst + 1] = e.r1;
}

int ba_launch_residuals(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k,
                        const int32_t *obs_orig, double *out, cudaStream_t st) {
    if (P.n_obs_local > 0) {
        k_residuals<<<(P.n_obs_local + 255) / 256, 256, 0, st>>>(P, x, k, obs_orig, out);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}


// =====================================================================================
// 2b. Fused Schur complement for scenes whose points see a narrow range of cameras (a sequence:
//     KITTI-shaped C4).  One kernel replaces k_lin + k_gather + k_cam_blocks and their 144-byte
//     record per observation: a CTA owns a WINDOW of kWinCams consecutive cameras and the points
//     whose first camera falls into the window's stride; everything those points contribute —
//     the off-diagonal 6 x 6 blocks between window cameras, the cameras' diagonal blocks, gradient
//     and right-hand side — is accumulated on chip (blocks in shared memory, camera sums in
//     registers) and leaves once, as a partial window.  Per batch of <= 32 points:
//       b   one thread per observation: residual, Jacobians, Huber (lin_obs)  -> shared memory
//       c1  one thread per point: V + D^2 = Lc Lc^T, V^-1, h                  -> HBM (k_backsub reads them)
//       c2  one thread per observation: T~ = W Lc^-T, e = r - JX h            -> shared memory
//       d   a warp per window camera: U, sum T~T~^T, g, rhs; the camera's observations in batch order
//       e   a warp per camera pair: S_ab -= T~_a T~_b^T over the points that see both, in batch order
//     Windows overlap by the camera span of a point, and a window may be split over several CTAs:
//     k_window_reduce_* add the partial windows in a fixed order.  No atomics on doubles anywhere: the
//     result is bit-reproducible.
// =====================================================================================
constexpr int kWinThreads = 512, kWinPts = 32, kWinObs = 320, kWinRec = 40;
// record (doubles): [0,18) T~ (after c2; JX sits in [34,40) until then) | [18,30) Jc | 30,31 r | 32,33 e | [34,40) JX
constexpr int kWinSmemBytes = (kWinBlocks * 36 + kWinObs * kWinRec + kWinPts * 12) * 8 + kWinPts * kWinCams * 2 + 2 * kWinObs + 512;

__device__ __forceinline__ int win_block(int li, int lj) { return li * (li - 1) / 2 + lj; }  // li > lj

__global__ void __launch_bounds__(kWinThreads, 1)
k_schur_window(BAProblemDev P, BAStateDev x, BAConsts k, BALinSys L, double inv_radius, double *__restrict__ scalars,
               WindowPlanDev W) {
    extern __shared__ __align__(16) unsigned char win_smem[];
    double *Sacc = reinterpret_cast<double *>(win_smem);
    double *rec = Sacc + kWinBlocks * 36;
    double *ptab = rec + kWinObs * kWinRec;  // per batch point: l00 l10 l11 l20 l21 l22 h0 h1 h2 pvar
    unsigned short *slot = reinterpret_cast<unsigned short *>(ptab + kWinPts * 12);  // [pt][cam] -> batch obs, 0xFFFF = none
    unsigned char *cam_of = reinterpret_cast<unsigned char *>(slot + kWinPts * kWinCams);
    unsigned char *pt_of = cam_of + kWinObs;
    int *meta = reinterpret_cast<int *>(pt_of + kWinObs);  // [0] n_pts, [1] n_obs, [2] present mask, [4..36] pt ids, [40..73] obs offsets
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cta = blockIdx.x, c0 = W.cta_cam0[cta];
    const int p_begin = W.cta_ptr[cta], p_end = W.cta_ptr[cta + 1];
    for (int i = tid; i < kWinBlocks * 36; i += kWinThreads) Sacc[i] = 0.0;
    // camera sums: warp w owns window cameras w and w + 16; lane owns outputs lane and lane + 32 of the 54
    double cacc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    int dx[2][3], dy[2][3];  // record offsets of the three products of outputs lane and lane + 32
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int j = lane + 32 * u;
        int a = 0, w = j < 21 ? j : (j < 42 ? j - 21 : 0);
        while (j < 42 && w >= a + 1) w -= a + 1, ++a;  // w-th lower entry -> (a, b = w)
        if (j < 21) {            // U: Jc^T Jc
            dx[u][0] = 18 + a, dy[u][0] = 18 + w, dx[u][1] = 24 + a, dy[u][1] = 24 + w, dx[u][2] = 34, dy[u][2] = 34;
        } else if (j < 42) {     // sum T~ T~^T
            dx[u][0] = 3 * a, dy[u][0] = 3 * w, dx[u][1] = 3 * a + 1, dy[u][1] = 3 * w + 1, dx[u][2] = 3 * a + 2, dy[u][2] = 3 * w + 2;
        } else if (j < 48) {     // g_c = Jc^T r
            dx[u][0] = 18 + j - 42, dy[u][0] = 30, dx[u][1] = 24 + j - 42, dy[u][1] = 31, dx[u][2] = 34, dy[u][2] = 34;
        } else if (j < 54) {     // rhs = Jc^T (r - JX h)
            dx[u][0] = 18 + j - 48, dy[u][0] = 32, dx[u][1] = 24 + j - 48, dy[u][1] = 33, dx[u][2] = 34, dy[u][2] = 34;
        } else {
            dx[u][0] = dx[u][1] = dx[u][2] = dy[u][0] = dy[u][1] = dy[u][2] = 34;
        }
    }
    double gmax = 0.0;
    bool fail = false;
    for (int pos = p_begin; pos < p_end;) {
        __syncthreads();
        if (warp == 0) {  // batch: as many of the next 32 points as fit kWinObs observations
            const bool in = pos + lane < p_end;
            const int p = in ? W.win_pts[pos + lane] : 0;
            const int kn = in ? P.pt_ptr[p + 1] - P.pt_ptr[p] : 0;
            int incl = kn;
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if (lane >= d) incl += t;
            }
            const bool take = in && incl <= kWinObs;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, take);
            const int n_pts = __ffs(~m) - 1 < 0 ? 32 : __ffs(~m) - 1;  // leading run of takes
            meta[4 + lane] = p, meta[40 + lane] = incl - kn;
            if (lane == 0) meta[0] = n_pts, meta[2] = 0;
            if (lane == n_pts - 1) meta[1] = incl, meta[40 + n_pts] = incl;
        }
        for (int i = tid; i < kWinPts * kWinCams; i += kWinThreads) slot[i] = 0xFFFFu;
        __syncthreads();
        const int n_pts = meta[0], n_obs = meta[1];
        // ---- b: one thread per observation
        if (tid < n_obs) {
            int lp = 0;
            for (int step = 16; step > 0; step >>= 1)
                if (lp + step < n_pts && meta[40 + lp + step] <= tid) lp += step;
            const int p = meta[4 + lp];
            const int o = P.pt_ptr[p] + (tid - meta[40 + lp]);
            const bool pvar = P.pt_var[p] != 0;
            LinObs lo;
            lin_obs(P, x, k, L, o, p, pvar, lo);
            const int lc = P.obs_cam[o] - c0;
            double *r = rec + tid * kWinRec;
            if (lo.active) {
#pragma unroll
                for (int j = 0; j < 12; ++j) r[18 + j] = lo.Jc[j];
                r[30] = lo.r0, r[31] = lo.r1;
#pragma unroll
                for (int j = 0; j < 6; ++j) r[34 + j] = lo.JX[j];
            } else {
#pragma unroll
                for (int j = 18; j < kWinRec; ++j) r[j] = 0.0;
            }
            cam_of[tid] = (unsigned char)lc, pt_of[tid] = (unsigned char)lp;
            slot[lp * kWinCams + lc] = (unsigned short)tid;
            atomicOr(&meta[2], 1 << lc);
        }
        __syncthreads();
        // ---- c1: one thread per point
        if (tid < n_pts) {
            const int p = meta[4 + tid];
            const bool pvar = P.pt_var[p] != 0;
            double V[6] = {0, 0, 0, 0, 0, 0}, g[3] = {0, 0, 0};
            for (int t = meta[40 + tid]; t < meta[40 + tid + 1]; ++t) {
                const double *J = rec + t * kWinRec + 34, r0 = rec[t * kWinRec + 30], r1 = rec[t * kWinRec + 31];
                V[0] += J[0] * J[0] + J[3] * J[3], V[1] += J[0] * J[1] + J[3] * J[4];
                V[2] += J[0] * J[2] + J[3] * J[5], V[3] += J[1] * J[1] + J[4] * J[4];
                V[4] += J[1] * J[2] + J[4] * J[5], V[5] += J[2] * J[2] + J[5] * J[5];
                g[0] += J[0] * r0 + J[3] * r1, g[1] += J[1] * r0 + J[4] * r1, g[2] += J[2] * r0 + J[5] * r1;
            }
            double l00 = 1, l10 = 0, l11 = 1, l20 = 0, l21 = 0, l22 = 1, h0 = 0, h1 = 0, h2 = 0;
            if (pvar) {  // same arithmetic as k_lin
                const double a = V[0] + fmin(fmax(V[0], 1e-6), 1e32) * inv_radius;
                const double d = V[3] + fmin(fmax(V[3], 1e-6), 1e32) * inv_radius;
                const double f = V[5] + fmin(fmax(V[5], 1e-6), 1e32) * inv_radius;
                const double b = V[1], c = V[2], e = V[4];
                l00 = sqrt(a), l10 = b / l00, l20 = c / l00;
                const double t11 = d - l10 * l10;
                l11 = sqrt(t11), l21 = (e - l20 * l10) / l11;
                const double t22 = f - l20 * l20 - l21 * l21;
                l22 = sqrt(t22);
                if (!(a > 0.0) || !(t11 > 0.0) || !(t22 > 0.0) || !isfinite(l22)) fail = true;
                const double i00 = 1.0 / l00, i11 = 1.0 / l11, i22 = 1.0 / l22;
                const double m10 = -l10 * i00 * i11, m21 = -l21 * i11 * i22;
                const double m20 = (l10 * l21 - l20 * l11) * i00 * i11 * i22;
                double Vi[6];
                Vi[0] = i00 * i00 + m10 * m10 + m20 * m20, Vi[1] = m10 * i11 + m20 * m21, Vi[2] = m20 * i22;
                Vi[3] = i11 * i11 + m21 * m21, Vi[4] = m21 * i22, Vi[5] = i22 * i22;
#pragma unroll
                for (int j = 0; j < 6; ++j) L.Vinv[6 * (size_t)p + j] = Vi[j];
#pragma unroll
                for (int j = 0; j < 3; ++j) L.gp[3 * (size_t)p + j] = g[j];
                h0 = Vi[0] * g[0] + Vi[1] * g[1] + Vi[2] * g[2];
                h1 = Vi[1] * g[0] + Vi[3] * g[1] + Vi[4] * g[2];
                h2 = Vi[2] * g[0] + Vi[4] * g[1] + Vi[5] * g[2];
                const double *sp = L.sp + 3 * (size_t)p;
                gmax = fmax(gmax, fmax(fabs(g[0] / sp[0]), fmax(fabs(g[1] / sp[1]), fabs(g[2] / sp[2]))));
            }
            L.h[3 * (size_t)p] = h0, L.h[3 * (size_t)p + 1] = h1, L.h[3 * (size_t)p + 2] = h2;
            double *pt = ptab + tid * 12;
            pt[0] = l00, pt[1] = l10, pt[2] = l11, pt[3] = l20, pt[4] = l21, pt[5] = l22, pt[6] = h0, pt[7] = h1, pt[8] = h2;
            pt[9] = pvar ? 1.0 : 0.0;
        }
        __syncthreads();
        // ---- c2: one thread per observation: T~ = W Lc^-T, e = r - JX h
        if (tid < n_obs) {
            double *r = rec + tid * kWinRec;
            const double *pt = ptab + pt_of[tid] * 12;
            const double *Jc = r + 18, *JX = r + 34;
            r[32] = r[30] - (JX[0] * pt[6] + JX[1] * pt[7] + JX[2] * pt[8]);
            r[33] = r[31] - (JX[3] * pt[6] + JX[4] * pt[7] + JX[5] * pt[8]);
            if (pt[9] != 0.0) {
#pragma unroll
                for (int a = 0; a < 6; ++a) {
                    const double w0 = Jc[a] * JX[0] + Jc[6 + a] * JX[3];
                    const double w1 = Jc[a] * JX[1] + Jc[6 + a] * JX[4];
                    const double w2 = Jc[a] * JX[2] + Jc[6 + a] * JX[5];
                    const double t0 = w0 / pt[0];
                    const double t1 = (w1 - pt[1] * t0) / pt[2];
                    const double t2 = (w2 - pt[3] * t0 - pt[4] * t1) / pt[5];
                    r[a * 3] = t0, r[a * 3 + 1] = t1, r[a * 3 + 2] = t2;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 18; ++j) r[j] = 0.0;
            }
            r[34] = 0.0;  // JX is spent: slot 34 is the zero operand of the camera sums below
        }
        __syncthreads();
        const unsigned present = (unsigned)meta[2];
        // ---- d: camera sums, the camera's observations in batch order; every output is x1 y1 + x2 y2 + x3 y3 of
        // record entries (operand table in registers, slot 34 = 0 where a term is missing): no divergence
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
            const int lc = warp + 16 * cc;
            if (lc >= kWinCams || !((present >> lc) & 1u)) continue;
            for (int base = 0; base < n_obs; base += 32) {
                unsigned m = __ballot_sync(0xFFFFFFFFu, base + lane < n_obs && cam_of[base + lane] == lc);
                while (m) {
                    const int t = base + __ffs(m) - 1;
                    m &= m - 1;
                    const double *r = rec + t * kWinRec;
                    cacc[cc][0] += r[dx[0][0]] * r[dy[0][0]] + r[dx[0][1]] * r[dy[0][1]] + r[dx[0][2]] * r[dy[0][2]];
                    cacc[cc][1] += r[dx[1][0]] * r[dy[1][0]] + r[dx[1][1]] * r[dy[1][1]] + r[dx[1][2]] * r[dy[1][2]];
                }
            }
        }
        // ---- e: camera pairs (li > lj), the points that see both in batch order; the two half-warps take
        // alternate points (outputs l, l + 16, l + 32 of the 36 per lane l of a half), halves added at the end
        for (int bi = warp; bi < kWinBlocks; bi += kWinThreads / 32) {
            int li = 1;
            while ((li + 1) * li / 2 <= bi) ++li;  // bi = li (li - 1) / 2 + lj
            const int lj = bi - li * (li - 1) / 2;
            if (!((present >> li) & 1u) || !((present >> lj) & 1u)) continue;
            const unsigned short si = lane < n_pts ? slot[lane * kWinCams + li] : 0xFFFFu;
            const unsigned short sj = lane < n_pts ? slot[lane * kWinCams + lj] : 0xFFFFu;
            unsigned m = __ballot_sync(0xFFFFFFFFu, si != 0xFFFFu && sj != 0xFFFFu);
            if (!m) continue;
            const int half = lane >> 4, l = lane & 15;
            double acc[3] = {0.0, 0.0, 0.0};
            while (m) {
                const int s0 = __ffs(m) - 1;
                m &= m - 1;
                const int s1 = m ? __ffs(m) - 1 : -1;
                m &= m - 1 < m ? m - 1 : 0u;  // drop the second bit if there is one
                const int src = half == 0 ? s0 : s1;
                const int ti = __shfl_sync(0xFFFFFFFFu, (int)si, src < 0 ? 0 : src);
                const int tj = __shfl_sync(0xFFFFFFFFu, (int)sj, src < 0 ? 0 : src);
                if (src >= 0) {
                    const double *Ti = rec + ti * kWinRec, *Tj = rec + tj * kWinRec;
#pragma unroll
                    for (int u = 0; u < 3; ++u) {
                        const int e = l + 16 * u;
                        if (e < 36) {
                            const int a = e / 6, c = e % 6;
                            acc[u] += Ti[a * 3] * Tj[c * 3] + Ti[a * 3 + 1] * Tj[c * 3 + 1] + Ti[a * 3 + 2] * Tj[c * 3 + 2];
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const double tot = acc[u] + __shfl_down_sync(0xFFFFFFFFu, acc[u], 16);
                if (half == 0 && l + 16 * u < 36) Sacc[bi * 36 + l + 16 * u] += tot;
            }
        }
        pos += n_pts;
    }
    __syncthreads();
    // ---- flush the partial window
    double *pS = W.pS + (size_t)cta * kWinBlocks * 36;
    for (int i = tid; i < kWinBlocks * 36; i += kWinThreads) pS[i] = Sacc[i];
    double *pC = W.pC + (size_t)cta * kWinCams * 54;
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        const int lc = warp + 16 * cc;
        if (lc >= kWinCams) continue;
#pragma unroll
        for (int u = 0; u < 2; ++u)
            if (lane + 32 * u < 54) pC[lc * 54 + lane + 32 * u] = cacc[cc][u];
    }
    if (fail) scalars[SC_FAIL] = 1.0;
    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 16));
    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 8));
    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 4));
    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 2));
    gmax = fmax(gmax, __shfl_xor_sync(0xFFFFFFFFu, gmax, 1));
    if (lane == 0 && gmax > 0.0) atomic_max_nonneg(&scalars[SC_GRAD_MAX_PT], gmax);
}

// windows (and their parts) that hold camera pair (hi, lo) / camera c: w in [ceil((hi - 23) / S), floor(lo / S)]
__device__ __forceinline__ void window_range(const WindowPlanDev &W, int hi, int lo, int &w0, int &w1) {
    const int t = hi - (kWinCams - 1);
    w0 = t <= 0 ? 0 : (t + W.stride - 1) / W.stride;
    w1 = min(lo / W.stride, W.n_win - 1);
}

// off-diagonal blocks of S: 8 lanes per structure block, partial windows added in (window, part) order
__global__ void __launch_bounds__(256)
k_window_reduce_blocks(BAProblemDev P, BALinSys L, WindowPlanDev W) {
    const int sub = threadIdx.x & 7;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
    if (b >= P.n_blocks) return;
    const int2 cams = P.blk_cams[b];  // columns(x) > columns(y)
    if (cams.x == cams.y) return;      // cannot occur on this path (no camera twice on one point)
    const bool x_hi = cams.x > cams.y;
    const int hi = x_hi ? cams.x : cams.y, lo = x_hi ? cams.y : cams.x;
    int w0, w1;
    window_range(W, hi, lo, w0, w1);
    int ra[6], cb[6];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        ra[j] = P.colq[cams.x] >= 0 ? P.colq[cams.x] + j : -1, ra[3 + j] = P.colt[cams.x] >= 0 ? P.colt[cams.x] + j : -1;
        cb[j] = P.colq[cams.y] >= 0 ? P.colq[cams.y] + j : -1, cb[3 + j] = P.colt[cams.y] >= 0 ? P.colt[cams.y] + j : -1;
    }
    for (int e = sub; e < 36; e += 8) {
        const int a = e / 6, c = e % 6;  // entry (a of cams.x, c of cams.y)
        if (ra[a] < 0 || cb[c] < 0) continue;
        const int pe = x_hi ? a * 6 + c : c * 6 + a;  // the partial holds M(hi, lo): rows = hi's parameters
        double v = 0.0;
        for (int w = w0; w <= w1; ++w) {
            const int bi = win_block(hi - w * W.stride, lo - w * W.stride);
            for (int part = 0; part < W.parts; ++part)
                v += W.pS[((size_t)(w * W.parts + part) * kWinBlocks + bi) * 36 + pe];
        }
        L.S[L.tm.at(ra[a], cb[c])] = -v;
    }
}

// camera sums: one thread per (camera, output)
__global__ void __launch_bounds__(256)
k_window_reduce_cams(BAProblemDev P, BALinSys L, WindowPlanDev W) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = idx / 54, j = idx % 54;
    if (c >= P.n_cams) return;
    const int cq = P.colq[c], ct = P.colt[c];
    if (cq < 0 && ct < 0) return;
    int w0, w1;
    window_range(W, c, c, w0, w1);
    double v = 0.0;
    for (int w = w0; w <= w1; ++w)
        for (int part = 0; part < W.parts; ++part)
            v += W.pC[((size_t)(w * W.parts + part) * kWinCams + (c - w * W.stride)) * 54 + j];
    int cols[6];
#pragma unroll
    for (int q = 0; q < 3; ++q) cols[q] = cq >= 0 ? cq + q : -1, cols[3 + q] = ct >= 0 ? ct + q : -1;
    if (j < 42) {
        int w = j < 21 ? j : j - 21, a = 0;
        while (w >= a + 1) w -= a + 1, ++a;
        if (cols[a] >= 0 && cols[w] >= 0) {
            if (j < 21)
                L.U[(size_t)cols[a] * 6 + w] = v;
            else
                L.Ud[(size_t)cols[a] * 6 + w] = v;
        }
    } else if (j < 48) {
        if (cols[j - 42] >= 0) L.gc[cols[j - 42]] = v;
    } else {
        if (cols[j - 48] >= 0) L.rhs[cols[j - 48]] = v;
    }
}

int ba_launch_schur_window(const BAProblemDev &P, const BAStateDev &x, const BAConsts &k, const BALinSys &L, double inv_radius,
                           double *scalars, const WindowPlanDev &W, cudaStream_t st) {
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        XRB_CUDA(cudaFuncSetAttribute(k_schur_window, cudaFuncAttributeMaxDynamicSharedMemorySize, kWinSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    if (W.n_ctas > 0) {
        k_schur_window<<<W.n_ctas, kWinThreads, kWinSmemBytes, st>>>(P, x, k, L, inv_radius, scalars, W);
        XRB_LAUNCHED();
    }
    if (P.n_blocks > 0) {
        k_window_reduce_blocks<<<(unsigned)(((long long)P.n_blocks * 8 + 255) / 256), 256, 0, st>>>(P, L, W);
        XRB_LAUNCHED();
    }
    if (P.n_cams > 0) {
        k_window_reduce_cams<<<(unsigned)(((long long)P.n_cams * 54 + 255) / 256), 256, 0, st>>>(P, L, W);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// per local point: first camera (natural index), camera span, and whether the window path can take it
// (<= kWinPts... observations per point bounded, no camera twice)
__global__ void k_point_anchor(BAProblemDev P, int32_t *__restrict__ anchor, int32_t *__restrict__ stats) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P.n_pts_local) return;
    const int k0 = P.pt_ptr[p], k1 = P.pt_ptr[p + 1];
    int lo = INT_MAX, hi = -1, dup = 0;
    for (int a = k0; a < k1; ++a) {
        const int c = P.obs_cam[a];
        lo = min(lo, c), hi = max(hi, c);
        for (int b = k0; b < a; ++b) dup |= P.obs_cam[b] == c;
    }
    anchor[p] = k1 > k0 ? lo : 0;
    if (k1 > k0) {
        atomicMax(&stats[0], hi - lo);
        atomicMax(&stats[1], k1 - k0);
        if (dup) atomicMax(&stats[2], 1);
    }
}

int ba_launch_point_anchor(const BAProblemDev &P, int32_t *anchor, int32_t *stats, cudaStream_t st) {
    XRB_CUDA(cudaMemsetAsync(stats, 0, 4 * sizeof(int32_t), st));
    if (P.n_pts_local > 0) {
        k_point_anchor<<<(P.n_pts_local + 127) / 128, 128, 0, st>>>(P, anchor, stats);
        XRB_LAUNCHED();
    }
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

}  // namespace xrb
