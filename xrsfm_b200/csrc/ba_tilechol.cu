// ba_tilechol.cu — reduced-camera-system solver of path B, generation 2: ONE persistent kernel
// factors the tile-major system, a second one back-substitutes.
//
// Replaces Ceres' SparseSchurComplementSolver factor + solve (selected by ba_solver.cc:74).
//
// Storage (ba_kernels.cuh, TileGeom): the lower triangle of S as 64 x 64 tiles, tile-column-major,
// only the tiles inside the half bandwidth — every tile is one contiguous 32 KB block, a block
// column is one contiguous range (that is what the multi-GPU exchange sends chunk by chunk).
// The right-hand side is a separate vector; the forward substitution rides along with the
// factorisation (y_k = L_kk^-1 (rhs_k - sum_j L_kj y_j)).
//
// Factorisation = a tile DAG executed by resident CTAs that synchronise through flags in global
// memory (release/acquire), no kernel boundaries:
//   CTA 0, the chain : for k = 0 .. nt-1:  [L_k,k-1 = A_k,k-1 L_k-1,k-1^-T ; A_kk -= L_k,k-1 L_k,k-1^T]
//                      -> POTRF(A_kk) in shared memory (16-wide steps, in-register 16 x 16 factor on
//                      one warp, the rest of the trailing update overlapped with the next factor)
//                      -> y_k -> publish.  The diagonal tile never leaves the SM between steps.
//   CTAs 1.. , workers: pull tile tasks from a counter in dependency order:
//                      P(i,k): L_ik = A_ik L_kk^-T (blocked substitution with the 16 x 16 inverses),
//                              rhs_i -= L_ik y_k, tile rewritten TRANSPOSED (k-major) so that
//                      U(i,j,k): A_ij -= L_ik L_jk^T reads both operands with straight 16-byte
//                              cp.async copies into the layout the FP64 FMA core wants.
// Every tile receives its updates in increasing k (flag-ordered): the result is deterministic.
// Back-substitution: the chain CTA solves x_k = L_kk^-T (y_k - sum_i L_ik^T x_i) top-down for the
// two nearest tiles itself; one worker per block column accumulates the far part as the x_i appear.
//
// Roofline: FP64 FMA pipe for the bulk (U tasks), dependency latency for the chain; DESIGN.md §B.4.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "ba_kernels.cuh"

namespace xrb {

namespace {

constexpr int T = 64;        // tile edge
constexpr int SB = 16;       // inner block of the diagonal factorisation
constexpr int PA = 65;       // pitch (doubles) of row-major working tiles in shared memory
constexpr int PD = 17;       // pitch of the 16 x 16 diagonal-block inverses in shared memory
constexpr int kDinvSm = 4 * SB * PD;
constexpr int kThreads = 256;
constexpr unsigned kSpinLimit = 1u << 22;  // polls before a wait gives up (a poll is >= 0.5 us of L2 round trip)

// flags block (ints): [0] task counter, [1] abort, then per-structure arrays
struct Flags {
    int *base;
    int nt, ntiles;
    __host__ __device__ int *next() const { return base; }
    __host__ __device__ int *abort_flag() const { return base + 1; }
    __host__ __device__ int *diag_done() const { return base + 2; }
    __host__ __device__ int *pdone() const { return base + 2 + nt; }
    __host__ __device__ int *upd() const { return base + 2 + nt + ntiles; }
    __host__ __device__ int *xdone() const { return base + 2 + nt + 2 * ntiles; }
    __host__ __device__ int *wdone() const { return base + 2 + 2 * nt + 2 * ntiles; }
    __host__ __device__ static size_t count(int nt, int ntiles) { return 2 + 3 * (size_t)nt + 2 * (size_t)ntiles; }
};

struct CholArgs {
    TileGeom g;
    double *tiles, *rhs, *dinv, *x, *wpart, *fail;
    Flags f;
    const int4 *tasks;  // {type (0 = P, 1 = U), i, j, k}
    int n_tasks;
    long long *trace;   // optional per-step clock of the chain (debug), may be null
};

extern __shared__ __align__(16) double g_sm[];  // the CTA's dynamic shared memory

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Thread 0 polls until *flag >= want (or the run is aborted); the CTA leaves together.
__device__ __forceinline__ bool cta_wait(const int *flag, int want, int *abort_flag, double *fail,
                                         long long *wait_cycles = nullptr) {
    __shared__ int ok_s;
    if (threadIdx.x == 0) {
        const long long t0 = wait_cycles ? clock64() : 0;
        unsigned spins = 0;
        int ok = 1;
        while (ld_acquire(flag) < want) {
            if ((++spins & 255u) == 0 && (spins > kSpinLimit || ld_acquire(abort_flag) != 0)) {
                st_release(abort_flag, 1);
                *fail = 2.0;  // surfaces as an invalid step on the host (SC_FAIL)
                ok = 0;
                break;
            }
        }
        ok_s = ok;
        if (wait_cycles) *wait_cycles += clock64() - t0;
    }
    __syncthreads();
    const bool ok = ok_s != 0;
    __syncthreads();  // ok_s may be rewritten by the next wait
    return ok;
}

// Same for up to three flags at once (one polling loop instead of three round trips in sequence).
__device__ __forceinline__ bool cta_wait3(const int *f0, int w0, const int *f1, int w1, const int *f2, int w2,
                                          int *abort_flag, double *fail, long long *wait_cycles = nullptr) {
    __shared__ int ok3_s;
    if (threadIdx.x == 0) {
        const long long t0 = wait_cycles ? clock64() : 0;
        unsigned spins = 0;
        int ok = 1;
        bool d0 = false, d1 = f1 == nullptr, d2 = f2 == nullptr;
        for (;;) {
            if (!d0) d0 = ld_acquire(f0) >= w0;
            if (!d1) d1 = ld_acquire(f1) >= w1;
            if (!d2) d2 = ld_acquire(f2) >= w2;
            if (d0 && d1 && d2) break;
            if ((++spins & 255u) == 0 && (spins > kSpinLimit || ld_acquire(abort_flag) != 0)) {
                st_release(abort_flag, 1);
                *fail = 2.0;
                ok = 0;
                break;
            }
        }
        ok3_s = ok;
        if (wait_cycles) *wait_cycles += clock64() - t0;
    }
    __syncthreads();
    const bool ok = ok3_s != 0;
    __syncthreads();
    return ok;
}

// All data stores of the CTA happen-before the flag: barrier, then one fence + release store.
__device__ __forceinline__ void cta_publish(int *flag, int v) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        st_release(flag, v);
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// contiguous 32 KB tile -> shared [64][64] (straight copy, L2 path)
__device__ __forceinline__ void tile_to_smem_async(double *dst, const double *__restrict__ src) {
    for (int idx = threadIdx.x; idx < T * T / 2; idx += kThreads) cp_async16(dst + 2 * idx, src + 2 * idx);
}

// row-major global tile -> shared [64][PA]; all loads in flight before the first store
__device__ __forceinline__ void tile_to_smem_padded(double (*dst)[PA], const double *__restrict__ src) {
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(reinterpret_cast<const double2 *>(src) + threadIdx.x + u * kThreads);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int idx = threadIdx.x + u * kThreads, r = idx >> 5, c = (idx & 31) * 2;
        dst[r][c] = v[u].x, dst[r][c + 1] = v[u].y;
    }
}

// ---- 16 x 16 factor of one warp, entirely in registers (row per lane, values exchanged by
// shuffles, pivots through rsqrt).  Spelled as template recursion so that everything stays in
// registers (with `#pragma unroll` the triangular loops stayed rolled and a[] went to local memory).
template <int J, int C>
__device__ __forceinline__ void fac_upd(double (&a)[SB]) {
    if constexpr (C < SB) {
        const double lcj = __shfl_sync(0xFFFFFFFFu, a[J], C);  // L[C][J]
        a[C] = fma(-a[J], lcj, a[C]);
        fac_upd<J, C + 1>(a);
    }
}
template <int J>
__device__ __forceinline__ void fac_col(double (&a)[SB], const int rl, double &inv_mine, bool &bad) {
    if constexpr (J < SB) {
        double d = __shfl_sync(0xFFFFFFFFu, a[J], J);
        if (!(d > 0.0) || !isfinite(d)) bad = true, d = 1.0;
        const double inv = rsqrt(d);
        if (rl == J) inv_mine = inv;
        a[J] *= inv;
        fac_upd<J, J + 1>(a);
        fac_col<J + 1>(a, rl, inv_mine, bad);
    }
}
// rows below a factored 16-block: one thread per row, right-looking forward substitution
template <int C, int C2>
__device__ __forceinline__ void trsm_upd(double (&a)[SB], const double x, double (*A)[PA], const int b) {
    if constexpr (C2 < SB) {
        a[C2] = fma(-x, A[b + C2][b + C], a[C2]);
        trsm_upd<C, C2 + 1>(a, x, A, b);
    }
}
template <int C>
__device__ __forceinline__ void trsm_col(double (&a)[SB], double (*A)[PA], const double *dinvd, const int b) {
    if constexpr (C < SB) {
        const double x = a[C] * dinvd[b + C];
        a[C] = x;
        trsm_upd<C, C + 1>(a, x, A, b);
        trsm_col<C + 1>(a, A, dinvd, b);
    }
}
// inverse of a factored 16-block out of shared memory: lane rl owns column rl of X = L^-1
template <int R, int P>
__device__ __forceinline__ void inv_dot(const double (&x)[SB], double (*A)[PA], const int b, double &v0, double &v1) {
    if constexpr (P < R) {
        const double l = A[b + R][b + P];
        if constexpr ((P & 1) != 0)
            v1 = fma(-l, x[P], v1);
        else
            v0 = fma(-l, x[P], v0);
        inv_dot<R, P + 1>(x, A, b, v0, v1);
    }
}
template <int R>
__device__ __forceinline__ void inv_row(double (&x)[SB], double (*A)[PA], const double *dinvd, const int b, const int rl) {
    if constexpr (R < SB) {
        double v0 = rl == R ? 1.0 : 0.0, v1 = 0.0;
        inv_dot<R, 0>(x, A, b, v0, v1);
        x[R] = (v0 + v1) * dinvd[b + R];
        inv_row<R + 1>(x, A, dinvd, b, rl);
    }
}

// rank-16 update of the square region [lo, 64)^2 x columns [c0, c1) of D (lower part of the region
// as far as it is lower in D) from panel columns [pb, pb + 16): D[i][j] -= sum_p D[i][pb+p] D[j][pb+p].
// `nthr` threads with ids `t` share the outputs.
__device__ __forceinline__ void rank16_update(double (*D)[PA], int pb, int r_lo, int c_lo, int c_hi, int t, int nthr) {
    const int nr = T - r_lo, ncol = c_hi - c_lo;
    for (int idx = t; idx < nr * ncol; idx += nthr) {
        const int i = r_lo + idx / ncol, j = c_lo + idx % ncol;
        if (j > i) continue;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = 0; p < SB; p += 2) {
            s0 = fma(D[i][pb + p], D[j][pb + p], s0);
            s1 = fma(D[i][pb + p + 1], D[j][pb + p + 1], s1);
        }
        D[i][j] -= s0 + s1;
    }
}

// In-place Cholesky of the 64 x 64 tile in shared memory D (lower part); dinvd[64] receives
// 1 / L[r][r], Dinv[4][16][16] the inverses of the four 16 x 16 diagonal blocks of L.
// Serial chain per 16 columns: in-register factor (warp 0) -> rows below (one thread per row) ->
// update of the NEXT 16 columns; the rest of the trailing update runs on warps 1..7 while warp 0
// already factors the next block, the 16-block inverse on warp 7.
__device__ __forceinline__ void potrf64(double (*D)[PA], double *dinvd, double *Dinv, bool &bad_out) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool bad = false;
    for (int b = 0; b < T; b += SB) {
        if (warp == 0) {
            const int rl = lane & 15;
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = D[b + rl][b + c];
            double inv_mine = 1.0;
            fac_col<0>(a, rl, inv_mine, bad);
            if (lane < SB) {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c <= rl) D[b + rl][b + c] = a[c];
                dinvd[b + rl] = inv_mine;
            }
        } else if (b >= SB && b + SB < T) {
            // rest of the update of the previous step: region [b+16, 64)^2 from panel columns [b-16, b)
            rank16_update(D, b - SB, b + SB, b + SB, T, tid - 32, kThreads - 32);
        }
        __syncthreads();
        const int below = T - b - SB;
        if (warp == 7) {
            const int rl = lane & 15;
            double x[SB];
            inv_row<0>(x, D, dinvd, b, rl);
            if (lane < SB) {
                double *X = Dinv + (b / SB) * SB * PD;
#pragma unroll
                for (int r = 0; r < SB; ++r) X[r * PD + rl] = x[r];  // X[r][rl]; zero above the diagonal by construction
            }
        } else if (tid < below) {
            const int i = b + SB + tid;
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = D[i][b + c];
            trsm_col<0>(a, D, dinvd, b);
#pragma unroll
            for (int c = 0; c < SB; ++c) D[i][b + c] = a[c];
        }
        __syncthreads();
        if (below > 0) {
            // next 16 columns [b+16, b+32), rows [b+16, 64)
            rank16_update(D, b, b + SB, b + SB, b + 2 * SB, tid, kThreads);
            __syncthreads();
        }
    }
    bad_out = bad;
}

// X = A L^-T for a 64 x 64 tile: A (in) / X (out) row-major in As, L row-major lower in Lk, Dinv the
// 16 x 16 diagonal-block inverses of L (explicit zeros above their diagonals, pitch PD); Xt[p][r] =
// X[r][p] is written as well (pitch 64).  Right-looking over the four 16-column blocks:
//   [a] X_cb = A_cb Dinv_cb^T            thread (r = tid / 4, q = tid % 4): row r, columns 4q .. 4q+3
//   [b] A_cb' -= X_cb L[cb'][cb]^T, cb' > cb   4 x 4 register tiles, 64 threads per remaining block,
//                                        X read k-major from Xt (two LDS.128 per 16 FMAs)
// Every accumulation is at most 16 deep and there are 8-16 independent ones per thread: an FP64 FMA
// has ~25 cycles of dependent latency, so chain length, not flop count, is what this routine costs.
__device__ __forceinline__ void trsm64(double (*As)[PA], double (*Lk)[PA], const double *Dinv, double *Xt) {
    const int tid = threadIdx.x, r = tid >> 2, q = tid & 3;
    const int ub = tid >> 6, ur0 = ((tid & 63) >> 2) * 4, uc = (tid & 3) * 4;
#pragma unroll 1
    for (int cb = 0; cb < T / SB; ++cb) {
        const int c0 = cb * SB + 4 * q;
        double t16[SB];
#pragma unroll
        for (int p = 0; p < SB; ++p) t16[p] = As[r][cb * SB + p];
        const double *X = Dinv + cb * SB * PD + 4 * q * PD;
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int p = 0; p < SB; p += 2) {
                s0 = fma(t16[p], X[u * PD + p], s0);
                s1 = fma(t16[p + 1], X[u * PD + p + 1], s1);
            }
            v[u] = s0 + s1;
        }
        __syncwarp();  // the row's four threads have read the old values
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            As[r][c0 + u] = v[u];
            Xt[(c0 + u) * T + r] = v[u];
        }
        __syncthreads();
        const int cbp = cb + 1 + ub;
        if (cbp < T / SB) {
            double acc[4][4] = {};
#pragma unroll 4
            for (int p = 0; p < SB; ++p) {
                const double2 a0 = *reinterpret_cast<const double2 *>(Xt + (cb * SB + p) * T + ur0);
                const double2 a1 = *reinterpret_cast<const double2 *>(Xt + (cb * SB + p) * T + ur0 + 2);
                const double av[4] = {a0.x, a0.y, a1.x, a1.y};
                double bv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = Lk[cbp * SB + uc + j][cb * SB + p];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fma(av[i], bv[j], acc[i][j]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) As[ur0 + i][cbp * SB + uc + j] -= acc[i][j];
        }
        __syncthreads();
    }
}

// v[r] -= sum_p As[r][p] * y[p], r = tid / 4; result by the q == 0 lane
__device__ __forceinline__ double row_dot(double (*As)[PA], const double *y) {
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    double s = 0.0;
#pragma unroll
    for (int p = 0; p < SB; ++p) s = fma(As[r][q * SB + p], y[q * SB + p], s);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
    return s;
}

// FP64 FMA core: acc[i][j] += sum_p Xs[p][r0 + i] * Ys[p][c0 + j] over 64 p; operands k-major with
// pitch 64.  Thread tile 4 x 4: warp (wr, wc) covers rows 32 wr .. +31, columns 16 wc .. +15; a warp
// reads 256 contiguous bytes of Xs and 128 of Ys per p (two LDS.128 each per thread).
struct GemmMap {
    int r0, c0;
    __device__ GemmMap() {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        r0 = 32 * (w >> 2) + 4 * (lane >> 2);
        c0 = 16 * (w & 3) + 4 * (lane & 3);
    }
};
__device__ __forceinline__ void gemm64(const double *__restrict__ Xs, const double *__restrict__ Ys, const GemmMap &m,
                                       double (&acc)[4][4]) {
#pragma unroll 4
    for (int p = 0; p < T; ++p) {
        const double2 a0 = *reinterpret_cast<const double2 *>(Xs + p * T + m.r0);
        const double2 a1 = *reinterpret_cast<const double2 *>(Xs + p * T + m.r0 + 2);
        const double2 b0 = *reinterpret_cast<const double2 *>(Ys + p * T + m.c0);
        const double2 b1 = *reinterpret_cast<const double2 *>(Ys + p * T + m.c0 + 2);
        const double a[4] = {a0.x, a0.y, a1.x, a1.y}, b[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
}

// shared-memory plan (doubles): the chain needs D, Lprev, B (3 padded tiles), Xt, 2 x Dinv, vectors;
// a worker needs As, Lk (padded), Xt / or Xs, Ys
constexpr int kPad = T * PA;
constexpr int kOffD0 = 0, kOffD1 = kPad, kOffB = 2 * kPad, kOffXt = 3 * kPad;
constexpr int kOffDinv0 = kOffXt + T * T, kOffDinv1 = kOffDinv0 + kDinvSm;
constexpr int kOffVec = kOffDinv1 + kDinvSm;  // dinvd[64], y0[64], y1[64], rhs[64], tmp[64]
constexpr int kSmemDoubles = kOffVec + 5 * T;
constexpr int kSmemBytes = kSmemDoubles * (int)sizeof(double);

// ---- the chain --------------------------------------------------------------------------------
__device__ __forceinline__ void run_chain(const CholArgs &A) {
    double *const sm = g_sm;
    const TileGeom g = A.g;
    const int tid = threadIdx.x;
    double(*B)[PA] = reinterpret_cast<double(*)[PA]>(sm + kOffB);
    double *Xt = sm + kOffXt;
    double *dinvd = sm + kOffVec, *rhs = sm + kOffVec + 3 * T;
    const GemmMap gm;
    bool any_bad = false;
    const bool tron = A.trace != nullptr && tid == 0;
#define XRB_CLK(slot) \
    if (tron && k < 256) A.trace[k * 16 + (slot)] = clock64();
    for (int k = 0; k < g.nt; ++k) {
        XRB_CLK(0)
        const int par = k & 1;
        double(*D)[PA] = reinterpret_cast<double(*)[PA]>(sm + (par ? kOffD1 : kOffD0));
        double(*Lprev)[PA] = reinterpret_cast<double(*)[PA]>(sm + (par ? kOffD0 : kOffD1));
        double *Dinv = sm + (par ? kOffDinv1 : kOffDinv0), *Dinvprev = sm + (par ? kOffDinv0 : kOffDinv1);
        double *y = sm + kOffVec + (par ? 2 * T : T), *yprev = sm + kOffVec + (par ? T : 2 * T);
        const int t_kk = g.tile(k, k);
        const bool couple = k > 0 && g.h > 1;
        if (couple) {
            const int need = (k - 1) - g.first_col(k);
            const int t_sub = g.tile(k, k - 1);
            if (!cta_wait3(A.f.upd() + t_sub, need, A.f.upd() + t_kk, need, nullptr, 0, A.f.abort_flag(), A.fail)) return;
            tile_to_smem_padded(B, A.tiles + (size_t)t_sub * T * T);
        }
        tile_to_smem_padded(D, A.tiles + (size_t)t_kk * T * T);
        if (tid < T) rhs[tid] = __ldcg(A.rhs + k * T + tid);
        __syncthreads();
        XRB_CLK(1)
        if (couple) {
            const int t_sub = g.tile(k, k - 1);
            trsm64(B, Lprev, Dinvprev, Xt);  // B = L_k,k-1
            __syncthreads();
            XRB_CLK(7)
            // publish L_k,k-1 (transposed) first: the workers' updates of column k-1 wait for it
            double *dst = A.tiles + (size_t)t_sub * T * T;
            for (int idx = tid; idx < T * T / 2; idx += kThreads)
                reinterpret_cast<double2 *>(dst)[idx] = reinterpret_cast<const double2 *>(Xt)[idx];
            XRB_CLK(2)
            {
                const double s = row_dot(B, yprev);
                if ((tid & 3) == 0) rhs[tid >> 2] -= s;
            }
            XRB_CLK(9)
            double acc[4][4] = {};
            gemm64(Xt, Xt, gm, acc);
            XRB_CLK(10)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) D[gm.r0 + i][gm.c0 + j] -= acc[i][j];
            // the flag for L_k,k-1 goes out now: its stores were issued before the product above, so
            // the fence no longer waits for them (nobody needs the tile before the next pivot block)
            cta_publish(A.f.pdone() + t_sub, 1);
            __syncthreads();
        }
        // rows / columns past the true dimension: identity (they only exist in the last tile)
        {
            const int valid = min(T, g.n - k * T);
            if (valid < T)
                for (int idx = tid; idx < T * T; idx += kThreads) {
                    const int r = idx >> 6, c = idx & 63;
                    if (r >= valid || c >= valid) D[r][c] = r == c ? 1.0 : 0.0;
                }
            if (valid < T) __syncthreads();
        }
        XRB_CLK(3)
        bool bad = false;
        potrf64(D, dinvd, Dinv, bad);
        any_bad |= bad;
        XRB_CLK(4)
        // y_k = L_kk^-1 rhs_k on warp 0 while everybody stores L_kk and the block inverses: 16-blocks
        // top-down, lane (l = lane % 16, half = lane / 16) takes every other product of row l
        if (tid < 32) {
            const int l = tid & 15, half = tid >> 4;
            double *tmp = sm + kOffVec + 4 * T;
            for (int cb = 0; cb < T / SB; ++cb) {
                const int row = cb * SB + l;
                double s0 = 0.0, s1 = 0.0;
                for (int p = 2 * half; p < cb * SB; p += 4) {
                    s0 = fma(D[row][p], y[p], s0);
                    s1 = fma(D[row][p + 1], y[p + 1], s1);
                }
                double s = s0 + s1;
                s += __shfl_xor_sync(0xFFFFFFFFu, s, 16);
                if (half == 0) tmp[l] = rhs[row] - s;
                __syncwarp();
                const double *X = Dinv + cb * SB * PD + l * PD;
                double v0 = 0.0, v1 = 0.0;
#pragma unroll
                for (int p = 0; p < SB / 2; p += 2) {
                    v0 = fma(X[half * 8 + p], tmp[half * 8 + p], v0);
                    v1 = fma(X[half * 8 + p + 1], tmp[half * 8 + p + 1], v1);
                }
                double v = v0 + v1;
                v += __shfl_xor_sync(0xFFFFFFFFu, v, 16);
                if (half == 0) y[row] = v;
                __syncwarp();
            }
            XRB_CLK(8)
        }
        {
            double *dst = A.tiles + (size_t)t_kk * T * T;
            for (int idx = tid; idx < T * T; idx += kThreads) dst[idx] = D[idx >> 6][idx & 63];
            double *dd = A.dinv + (size_t)k * 4 * SB * SB;
            for (int idx = tid; idx < 4 * SB * SB; idx += kThreads) dd[idx] = Dinv[(idx >> 4) * PD + (idx & 15)];
        }
        __syncthreads();
        XRB_CLK(5)
        if (tid < T) A.rhs[k * T + tid] = y[tid];
        cta_publish(A.f.diag_done() + k, 1);
        XRB_CLK(6)
    }
#undef XRB_CLK
    if (any_bad && tid == 0) *A.fail = 1.0;
}

// ---- workers ------------------------------------------------------------------------------------
__device__ __forceinline__ bool run_task_P(const CholArgs &A, int i, int k, long long *wc) {
    double *const sm = g_sm;
    const TileGeom g = A.g;
    const int tid = threadIdx.x;
    double(*As)[PA] = reinterpret_cast<double(*)[PA]>(sm + kOffD0);
    double(*Lk)[PA] = reinterpret_cast<double(*)[PA]>(sm + kOffD1);
    double *Xt = sm + kOffXt, *Dinv = sm + kOffDinv0, *y = sm + kOffVec;
    const int t_ik = g.tile(i, k);
    if (!cta_wait3(A.f.upd() + t_ik, k - g.first_col(i), A.f.diag_done() + k, 1, nullptr, 0, A.f.abort_flag(), A.fail, wc))
        return false;
    tile_to_smem_padded(As, A.tiles + (size_t)t_ik * T * T);
    tile_to_smem_padded(Lk, A.tiles + (size_t)g.tile(k, k) * T * T);
    for (int idx = tid; idx < 4 * SB * SB; idx += kThreads)
        Dinv[(idx >> 4) * PD + (idx & 15)] = __ldcg(A.dinv + (size_t)k * 4 * SB * SB + idx);
    if (tid < T) y[tid] = __ldcg(A.rhs + k * T + tid);
    __syncthreads();
    trsm64(As, Lk, Dinv, Xt);
    __syncthreads();
    double *dst = A.tiles + (size_t)t_ik * T * T;
    for (int idx = tid; idx < T * T / 2; idx += kThreads)
        reinterpret_cast<double2 *>(dst)[idx] = reinterpret_cast<const double2 *>(Xt)[idx];
    {
        const double s = row_dot(As, y);
        if ((tid & 3) == 0) {
            double *p = A.rhs + i * T + (tid >> 2);
            *p = __ldcg(p) - s;
        }
    }
    cta_publish(A.f.pdone() + t_ik, 1);
    __syncthreads();  // shared memory is reused by the next task
    return true;
}

__device__ __forceinline__ bool run_task_U(const CholArgs &A, int i, int j, int k, long long *wc) {
    double *const sm = g_sm;
    const TileGeom g = A.g;
    const int tid = threadIdx.x;
    double *Xs = sm, *Ys = sm + T * T;
    const int t_ik = g.tile(i, k), t_jk = g.tile(j, k), t_ij = g.tile(i, j);
    const GemmMap gm;
    const int seq = k - g.first_col(i);
    if (!cta_wait3(A.f.pdone() + t_ik, 1, A.f.upd() + t_ij, seq, i != j ? A.f.pdone() + t_jk : nullptr, 1, A.f.abort_flag(),
                   A.fail, wc))
        return false;
    if (i != j) tile_to_smem_async(Ys, A.tiles + (size_t)t_jk * T * T);
    tile_to_smem_async(Xs, A.tiles + (size_t)t_ik * T * T);
    cp_async_commit();
    double *C = A.tiles + (size_t)t_ij * T * T;
    double2 c[4][2];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        c[r][0] = __ldcg(reinterpret_cast<const double2 *>(C + (gm.r0 + r) * T + gm.c0));
        c[r][1] = __ldcg(reinterpret_cast<const double2 *>(C + (gm.r0 + r) * T + gm.c0 + 2));
    }
    cp_async_wait_all();
    __syncthreads();
    double acc[4][4] = {};
    gemm64(Xs, i == j ? Xs : Ys, gm, acc);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        *reinterpret_cast<double2 *>(C + (gm.r0 + r) * T + gm.c0) = make_double2(c[r][0].x - acc[r][0], c[r][0].y - acc[r][1]);
        *reinterpret_cast<double2 *>(C + (gm.r0 + r) * T + gm.c0 + 2) = make_double2(c[r][1].x - acc[r][2], c[r][1].y - acc[r][3]);
    }
    cta_publish(A.f.upd() + t_ij, seq + 1);
    __syncthreads();
    return true;
}

__global__ void __launch_bounds__(kThreads, 1) k_tile_cholesky(CholArgs A) {
    if (blockIdx.x == 0) {
        run_chain(A);
        return;
    }
    __shared__ int task_s;
    // debug statistics (thread 0, only when the trace is armed): cycles waiting on flags, in P / U tasks
    long long st_wait = 0, st_p = 0, st_u = 0, n_p = 0, n_u = 0;
    long long *wc = A.trace ? &st_wait : nullptr;
    const long long t_begin = A.trace ? clock64() : 0;
    for (;;) {
        if (threadIdx.x == 0) task_s = atomicAdd(A.f.next(), 1);
        __syncthreads();
        const int t = task_s;
        __syncthreads();
        if (t >= A.n_tasks) break;
        const int4 tk = __ldg(A.tasks + t);
        const long long t0 = A.trace ? clock64() : 0;
        const bool ok = tk.x == 0 ? run_task_P(A, tk.y, tk.w, wc) : run_task_U(A, tk.y, tk.z, tk.w, wc);
        if (A.trace) {
            if (tk.x == 0) st_p += clock64() - t0, ++n_p; else st_u += clock64() - t0, ++n_u;
        }
        if (!ok) return;
    }
    if (A.trace && threadIdx.x == 0) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(A.trace) + 4096;
        atomicAdd(st + 0, (unsigned long long)st_wait), atomicAdd(st + 1, (unsigned long long)st_p);
        atomicAdd(st + 2, (unsigned long long)st_u), atomicAdd(st + 3, (unsigned long long)n_p);
        atomicAdd(st + 4, (unsigned long long)n_u), atomicAdd(st + 5, (unsigned long long)(clock64() - t_begin));
        atomicAdd(st + 6, 1ull);
    }
}

// ---- back-substitution -------------------------------------------------------------------------
// out[p] += sum_r Lt[p][r] * x[r] for one transposed tile in global memory (row p contiguous):
// thread (p = tid / 4, q = tid % 4) takes 16 consecutive r.
__device__ __forceinline__ double tile_tdot(const double *__restrict__ Lt, const double *xs) {
    const int p = threadIdx.x >> 2, q = threadIdx.x & 3;
    const double2 *row = reinterpret_cast<const double2 *>(Lt + p * T + q * SB);
    double2 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) v[u] = __ldcg(row + u);
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        s0 = fma(v[u].x, xs[q * SB + 2 * u], s0);
        s1 = fma(v[u].y, xs[q * SB + 2 * u + 1], s1);
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
    return s;  // valid on every lane of the quad
}

constexpr int kNear = 2;  // tiles (k+1, k) .. (k+kNear, k) are handled by the chain itself

__global__ void __launch_bounds__(kThreads, 1) k_tile_backsolve(CholArgs A) {
    double *const sm = g_sm;
    const TileGeom g = A.g;
    const int tid = threadIdx.x;
    if (blockIdx.x == 0) {
        double(*L)[PA] = reinterpret_cast<double(*)[PA]>(sm);
        double *Dinv = sm + kPad, *s = Dinv + kDinvSm, *xk = s + T, *xn = xk + T;  // xn[kNear][64]
        for (int k = g.nt - 1; k >= 0; --k) {
            const int i_hi = min(g.nt - 1, k + g.h - 1);
            tile_to_smem_padded(L, A.tiles + (size_t)g.tile(k, k) * T * T);
            for (int idx = tid; idx < 4 * SB * SB; idx += kThreads)
                Dinv[(idx >> 4) * PD + (idx & 15)] = __ldcg(A.dinv + (size_t)k * 4 * SB * SB + idx);
            double acc = 0.0;
            for (int i = k + 1; i <= min(i_hi, k + kNear); ++i) {
                // x_i of the previous steps is still in shared memory: slot (i % kNear)
                acc += tile_tdot(A.tiles + (size_t)g.tile(i, k) * T * T, xn + (i % kNear) * T);
            }
            if (i_hi > k + kNear) {
                if (!cta_wait(A.f.wdone() + k, 1, A.f.abort_flag(), A.fail)) return;
                acc += __ldcg(A.wpart + k * T + (tid >> 2));
            }
            if ((tid & 3) == 0) s[tid >> 2] = __ldcg(A.rhs + k * T + (tid >> 2)) - acc;
            __syncthreads();
            // x_k = L_kk^-T s, 16-blocks bottom-up on warp 0: x[cb] = Dinv_cb^T (s[cb] - sum_{c' > cb} L[c'][cb]^T x[c'])
            if (tid < 32) {
                const int l = tid & 15;
                for (int cb = T / SB - 1; cb >= 0; --cb) {
                    const int col = cb * SB + l;
                    double v0 = s[col], v1 = 0.0;
                    for (int r = (cb + 1) * SB; r < T; r += 2) {
                        v0 = fma(-L[r][col], xk[r], v0);
                        v1 = fma(-L[r + 1][col], xk[r + 1], v1);
                    }
                    const double v = v0 + v1;
                    const double *X = Dinv + cb * SB * PD;
                    double o = 0.0;
#pragma unroll
                    for (int p = 0; p < SB; ++p) {
                        const double vp = __shfl_sync(0xFFFFFFFFu, v, p);
                        if (p >= l) o = fma(X[p * PD + l], vp, o);  // (Dinv^T)[l][p] = Dinv[p][l]
                    }
                    if (tid < SB) xk[col] = o;
                    __syncwarp();
                }
            }
            __syncthreads();
            if (tid < T) {
                const double v = xk[tid];
                xn[(k % kNear) * T + tid] = v;
                if (k * T + tid < g.n) A.x[k * T + tid] = v;
                A.wpart[g.nt * T + k * T + tid] = v;  // x in padded form for the workers
            }
            cta_publish(A.f.xdone() + k, 1);
            __syncthreads();
        }
        return;
    }
    // workers: column k accumulates sum_{i > k + kNear} L_ik^T x_i as the x_i appear (top-down)
    double *xs = sm;
    const int nw = gridDim.x - 1;
    for (int k = g.nt - 1 - (blockIdx.x - 1); k >= 0; k -= nw) {
        const int i_hi = min(g.nt - 1, k + g.h - 1);
        if (i_hi <= k + kNear) continue;
        double acc = 0.0;
        for (int i = i_hi; i > k + kNear; --i) {
            if (!cta_wait(A.f.xdone() + i, 1, A.f.abort_flag(), A.fail)) return;
            if (tid < T) xs[tid] = __ldcg(A.wpart + g.nt * T + i * T + tid);
            __syncthreads();
            acc += tile_tdot(A.tiles + (size_t)g.tile(i, k) * T * T, xs);
            __syncthreads();
        }
        if ((tid & 3) == 0) A.wpart[k * T + (tid >> 2)] = acc;
        cta_publish(A.f.wdone() + k, 1);
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct TaskList {
    int4 *d = nullptr;
    int n = 0;
};
struct TileCtx {
    std::map<std::pair<int, int>, TaskList> lists;  // (nt, h) -> device task list
    int *flags = nullptr;
    size_t flags_cap = 0;
    double *wpart = nullptr;
    size_t wpart_cap = 0;
    long long *trace = nullptr;
    int grid = 0;
    bool attr_set = false;
};
constexpr int kMaxDevices = 64;
TileCtx g_tctx[kMaxDevices];
std::mutex g_tmutex;
bool g_trace_on = false;

// Tasks in dependency order, U tasks column by column and row by row inside a column (the rows
// next to the diagonal first: the chain waits for those).  The P task of row i for column k+1 is
// emitted right after row i's U tasks of column k — the last thing it depends on — so that it is
// long finished when the U tasks of column k+1 come up: a U task fetched right behind the P task it
// needs would idle for the whole substitution (that was 60 % of the workers' time).
TaskList build_tasks(const TileGeom &g) {
    std::vector<int4> flat;
    auto ihi = [&](int k) { return std::min(g.nt - 1, k + g.h - 1); };
    if (g.nt > 1)
        for (int i = 2; i <= ihi(0); ++i) flat.push_back(make_int4(0, i, 0, 0));
    for (int k = 0; k + 1 < g.nt; ++k) {
        const bool next = k + 1 <= g.nt - 2;
        for (int i = k + 1; i <= ihi(k); ++i) {
            for (int j = k + 1; j <= i; ++j) {
                if (i == k + 1 && j == k + 1) continue;  // the chain's own update
                flat.push_back(make_int4(1, i, j, k));
            }
            if (next && i >= k + 3 && i <= ihi(k + 1)) flat.push_back(make_int4(0, i, k + 1, k + 1));
        }
        if (next)  // rows that enter the band at column k+1
            for (int i = std::max(k + 3, ihi(k) + 1); i <= ihi(k + 1); ++i) flat.push_back(make_int4(0, i, k + 1, k + 1));
    }
    TaskList tl;
    tl.n = (int)flat.size();
    if (tl.n) {
        if (cudaMalloc(&tl.d, flat.size() * sizeof(int4)) != cudaSuccess) {
            tl.d = nullptr, tl.n = -1;
            return tl;
        }
        cudaMemcpy(tl.d, flat.data(), flat.size() * sizeof(int4), cudaMemcpyHostToDevice);
    }
    return tl;
}

}  // namespace

int ba_launch_tile_cholesky_solve(const TileGeom &g, double *tiles, double *rhs, double *dinv, double *x_out,
                                  double *fail_flag, cudaStream_t st, int64_t *launches) {
    if (g.n <= 0) return XRB_OK;
    int dev = 0;
    XRB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("cholesky: device id %d outside [0, %d)", dev, kMaxDevices);
        return XRB_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_tmutex);
    TileCtx &ctx = g_tctx[dev];
    if (!ctx.attr_set) {
        XRB_CUDA(cudaFuncSetAttribute(k_tile_cholesky, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        XRB_CUDA(cudaFuncSetAttribute(k_tile_backsolve, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        int sms = 0, per_sm = 0, coop = 0;
        XRB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        XRB_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
        XRB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tile_cholesky, kThreads, kSmemBytes));
        if (!coop || per_sm < 1) {
            set_error("cholesky: the device cannot co-schedule the persistent factorisation kernel");
            return XRB_ERR_NO_DEVICE;
        }
        ctx.grid = sms;  // one CTA per SM
        ctx.attr_set = true;
    }
    const int ntiles = g.n_tiles();
    auto key = std::make_pair(g.nt, g.h);
    auto it = ctx.lists.find(key);
    if (it == ctx.lists.end()) {
        TaskList tl = build_tasks(g);
        if (tl.n < 0) {
            set_error("cholesky: task list allocation failed");
            return XRB_ERR_CUDA;
        }
        it = ctx.lists.emplace(key, tl).first;
    }
    const size_t nflags = Flags::count(g.nt, ntiles);
    if (ctx.flags_cap < nflags) {
        if (ctx.flags) cudaFree(ctx.flags);
        ctx.flags = nullptr, ctx.flags_cap = 0;
        XRB_CUDA(cudaMalloc(&ctx.flags, nflags * sizeof(int)));
        ctx.flags_cap = nflags;
    }
    const size_t nw = 2 * (size_t)g.nt * T;
    if (ctx.wpart_cap < nw) {
        if (ctx.wpart) cudaFree(ctx.wpart);
        ctx.wpart = nullptr, ctx.wpart_cap = 0;
        XRB_CUDA(cudaMalloc(&ctx.wpart, nw * sizeof(double)));
        ctx.wpart_cap = nw;
    }
    if (g_trace_on && !ctx.trace) XRB_CUDA(cudaMalloc(&ctx.trace, (4096 + 16) * sizeof(long long)));
    if (g_trace_on) XRB_CUDA(cudaMemsetAsync(ctx.trace + 4096, 0, 16 * sizeof(long long), st));
    XRB_CUDA(cudaMemsetAsync(ctx.flags, 0, nflags * sizeof(int), st));
    CholArgs A;
    A.g = g, A.tiles = tiles, A.rhs = rhs, A.dinv = dinv, A.x = x_out, A.wpart = ctx.wpart, A.fail = fail_flag;
    A.f = Flags{ctx.flags, g.nt, ntiles};
    A.tasks = it->second.d, A.n_tasks = it->second.n;
    A.trace = g_trace_on ? ctx.trace : nullptr;
    // workers beyond the number of tasks would only poll the counter once; the back-substitution wants
    // one worker per block column
    const int grid_f = std::max(1, std::min(ctx.grid, 1 + A.n_tasks));
    const int grid_b = std::max(1, std::min(ctx.grid, 1 + g.nt));
    void *args[] = {&A};
    XRB_CUDA(cudaLaunchCooperativeKernel((void *)k_tile_cholesky, dim3(grid_f), dim3(kThreads), args, kSmemBytes, st));
    XRB_CUDA(cudaLaunchCooperativeKernel((void *)k_tile_backsolve, dim3(grid_b), dim3(kThreads), args, kSmemBytes, st));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    if (launches) *launches += 2;
    return XRB_OK;
}

// after a solve has completed (stream synchronised): did a wait give up?
int ba_tile_cholesky_aborted(int *aborted) {
    int dev = 0;
    XRB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tmutex);
    TileCtx &ctx = g_tctx[dev];
    *aborted = 0;
    if (ctx.flags) XRB_CUDA(cudaMemcpy(aborted, ctx.flags + 1, sizeof(int), cudaMemcpyDeviceToHost));
    return XRB_OK;
}

// debug: per-block-column clock64 of the chain CTA of the last factorisation (n = nt values)
int ba_tile_cholesky_trace(int enable, long long *out, int cap) {
    int dev = 0;
    XRB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tmutex);
    if (enable) {
        g_trace_on = true;
        return 0;
    }
    g_trace_on = false;
    TileCtx &ctx = g_tctx[dev];
    if (!ctx.trace || !out) return 0;
    XRB_CUDA(cudaDeviceSynchronize());
    const int n = std::min(cap, 4096 + 16);
    XRB_CUDA(cudaMemcpy(out, ctx.trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
    return n;
}

}  // namespace xrb
