// ba_tilechol.cu — reduced-camera-system solver of path B, generation 3: a sparse tile Cholesky
// executed as a task DAG by resident CTAs (one persistent kernel), back-substitution in a second one.
//
// Replaces Ceres' SparseSchurComplementSolver factor + solve (selected by ba_solver.cc:74).
//
// Storage (ba_plan.cuh): the lower triangle of S as 64 x 64 row-major tiles, only the tiles of the
// symbolic factor (original non-zeros first — that prefix is what a multi-GPU solve all-reduces —
// fill after).  Tasks and the flag values they wait for come from the plan (build_chol_plan):
//   F(k)   chain CTAs: [L_k,kp = A_k,kp L_kp,kp^-T ; A_kk -= L_k,kp L_k,kp^T] for the LAST column kp that
//          updates the diagonal tile, then POTRF(A_kk) in shared memory, y_k = L_kk^-1 rhs_k, publish.
//          With a natural order kp = k - 1 and one chain CTA walks the diagonal with L_kp,kp still in
//          its shared memory; a dissected band runs one chain per interior at the same time.
//   P(i,k) workers: L_ik = A_ik L_kk^-T, blocked substitution with the 16 x 16 diagonal-block inverses.
//   U(i,j,k) workers: A_ij -= L_ik L_jk^T; the diagonal ones (i = j) also carry the forward substitution
//          rhs_i -= L_ik y_k.  A tile receives its updates in a fixed order (sequence numbers in the
//          flags): the factor is bit-reproducible, every rank of a multi-GPU solve gets the same one.
//   PU(i,k) = P(i,k) fused with U(i,pk,k), pk the next column that meets row i (ba_plan.cu).
// The tile products run on the FP64 tensor pipe (mma.sync.m8n8k4.f64, DMMA): operands in shared memory
// with pitch 68 doubles (conflict-free fragment loads), 2 x 4 accumulator fragments per warp.
// CTAs synchronise through release/acquire flags in global memory; the queues are sorted by longest-path
// level, so a CTA never waits on a task queued behind the one it holds (no deadlock by construction).
//
// Roofline: FP64 tensor/FMA pipe for the bulk (U tasks), dependency latency for the chain; DESIGN.md §B.4.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "ba_kernels.cuh"

namespace xrb {

namespace {

constexpr int T = 64;        // tile edge
constexpr int SB = 16;       // inner block of the diagonal factorisation
constexpr int PA = 65;       // pitch (doubles) of the tile POTRF works on (thread-per-row accesses)
constexpr int PM = 68;       // pitch of DMMA operand tiles: (g * 68 + t) mod 16 distinct for g, t < 4
constexpr int PD = 20;       // pitch of the 16 x 16 diagonal-block inverses (DMMA operands as well)
constexpr int kDinvSm = 4 * SB * PD;
constexpr int kThreads = 256;
constexpr unsigned kSpinLimit = 1u << 22;  // polls before a wait gives up (a poll is >= 0.5 us of L2 round trip)

// flags block (ints): counters, abort, then per-structure arrays
struct Flags {
    int *base;
    int nt, ntiles;
    __host__ __device__ int *next_w() const { return base; }
    __host__ __device__ int *abort_flag() const { return base + 1; }
    __host__ __device__ int *next_f() const { return base + 2; }
    __host__ __device__ int *next_b() const { return base + 3; }
    __host__ __device__ int *next_wb() const { return base + 4; }
    __host__ __device__ int *diag_done() const { return base + 8; }
    __host__ __device__ int *pdone() const { return base + 8 + nt; }
    __host__ __device__ int *upd() const { return base + 8 + nt + ntiles; }
    __host__ __device__ int *xdone() const { return base + 8 + nt + 2 * ntiles; }
    __host__ __device__ int *wdone() const { return base + 8 + 2 * nt + 2 * ntiles; }
    __host__ __device__ static size_t count(int nt, int ntiles) { return 8 + 3 * (size_t)nt + 2 * (size_t)ntiles; }
};

struct CholArgs {
    CholPlanDev p;
    double *tiles, *rhs, *dinv, *x, *wpart, *fail;
    Flags f;
    long long *trace;   // optional clocks of the chain CTAs (debug), may be null
    int pipeline;       // workers claim two tasks ahead and prefetch (many tasks per worker) or take one at a time
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

// Lanes 0..3 of warp 0 each poll one flag until *f >= w (null = nothing to wait for) or the run is
// aborted; the CTA leaves together.
__device__ __forceinline__ bool cta_wait(const int *f0, int w0, const int *f1, int w1, const int *f2, int w2, const int *f3,
                                         int w3, int *abort_flag, double *fail, long long *wait_cycles = nullptr) {
    __shared__ int ok_s;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int *f = lane == 0 ? f0 : lane == 1 ? f1 : lane == 2 ? f2 : lane == 3 ? f3 : nullptr;
        const int w = lane == 0 ? w0 : lane == 1 ? w1 : lane == 2 ? w2 : w3;
        const long long t0 = wait_cycles ? clock64() : 0;
        bool done = f == nullptr;
        unsigned spins = 0;
        int ok = 1;
        for (;;) {
            if (!done) done = ld_acquire(f) >= w;
            if (__all_sync(0xFFFFFFFFu, done)) break;
            if ((++spins & 255u) == 0) {
                const bool ab = spins > kSpinLimit || ld_acquire(abort_flag) != 0;
                if (__any_sync(0xFFFFFFFFu, ab)) {
                    ok = 0;
                    break;
                }
            }
        }
        if (lane == 0) {
            if (!ok) {
                st_release(abort_flag, 1);
                *fail = 2.0;  // surfaces as an invalid step on the host (SC_FAIL)
            }
            ok_s = ok;
            if (wait_cycles) *wait_cycles += clock64() - t0;
        }
    }
    __syncthreads();
    const bool ok = ok_s != 0;
    __syncthreads();  // ok_s may be rewritten by the next wait
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

// Same, but the fence + flag store is left to the LAST thread while the others move on (the next
// CTA-wide barrier absorbs it): the drain of the tile stores overlaps the next phase.
__device__ __forceinline__ void cta_publish_async(int *flag, int v) {
    __syncthreads();
    if (threadIdx.x == kThreads - 1) {
        __threadfence();
        st_release(flag, v);
    }
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// contiguous row-major 32 KB tile -> shared [64][PM] (16-byte copies on the L2 path)
__device__ __forceinline__ void tile_to_smem_async(double *dst, const double *__restrict__ src) {
    for (int idx = threadIdx.x; idx < T * T / 2; idx += kThreads) cp_async16(dst + (idx >> 5) * PM + 2 * (idx & 31), src + 2 * idx);
}
// shared [64][PM] -> contiguous row-major global tile
__device__ __forceinline__ void tile_from_smem(double *__restrict__ dst, const double *src) {
    for (int idx = threadIdx.x; idx < T * T / 2; idx += kThreads)
        reinterpret_cast<double2 *>(dst)[idx] = *reinterpret_cast<const double2 *>(src + (idx >> 5) * PM + 2 * (idx & 31));
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

// dense [4][16][16] diagonal-block inverses in global memory -> shared, pitch PD
__device__ __forceinline__ void dinv_to_smem(double *dst, const double *__restrict__ src) {
    for (int idx = threadIdx.x; idx < 4 * SB * SB; idx += kThreads) dst[(idx >> 4) * PD + (idx & 15)] = __ldcg(src + idx);
}

// ---- FP64 tensor core: D(8x8) += A(8x4) B(4x8).  Lane (g = lane / 4, t = lane % 4) holds A[g][t],
// B[t][g] and the accumulator pair D[g][2t], D[g][2t + 1].
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// acc[fr][fc] += sum_p X[r0 + 8 fr + .][p] * Y[c0 + 8 fc + .][p] over 64 p; X, Y row-major with pitch PM.
// Warp w covers rows 16 (w / 2) .. +15 and columns 32 (w % 2) .. +31: 2 x 4 fragments, 6 loads per 8 DMMAs.
struct WarpMap {
    int g, t, r0, c0;
    __device__ WarpMap() {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
        g = lane >> 2, t = lane & 3;
        r0 = 16 * (w >> 1), c0 = 32 * (w & 1);
    }
};
__device__ __forceinline__ void gemm_nt(const double *__restrict__ Xs, const double *__restrict__ Ys, const WarpMap &m,
                                        double (&acc)[2][4][2]) {
    const double *xa = Xs + (m.r0 + m.g) * PM + m.t;
    const double *yb = Ys + (m.c0 + m.g) * PM + m.t;
#pragma unroll 4
    for (int p = 0; p < T; p += 4) {
        const double a0 = xa[p], a1 = xa[8 * PM + p];
        double b[4];
#pragma unroll
        for (int fc = 0; fc < 4; ++fc) b[fc] = yb[fc * 8 * PM + p];
#pragma unroll
        for (int fc = 0; fc < 4; ++fc) {
            dmma(acc[0][fc], a0, b[fc]);
            dmma(acc[1][fc], a1, b[fc]);
        }
    }
}

// X = A L^-T for a 64 x 64 tile, in place in Bs (pitch PM); L row-major lower in Ls (pitch PM), Dinv the
// 16 x 16 diagonal-block inverses of L (explicit zeros above their diagonals, pitch PD).  Rows are
// independent: warp w owns rows 8w .. 8w+7 through all four column blocks, no CTA barrier inside.
//   [a] X_cb = A_cb Dinv_cb^T          [b] A_cb' -= X_cb L[cb'][cb]^T for cb' > cb
__device__ __forceinline__ void trsm64(double *Bs, const double *__restrict__ Ls, const double *__restrict__ Dinv) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
    double *row = Bs + (8 * w + g) * PM;
#pragma unroll
    for (int cb = 0; cb < T / SB; ++cb) {
        double a4[4];
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) a4[ks] = row[cb * SB + 4 * ks + t];
        double x[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const double *Di = Dinv + cb * SB * PD + g * PD + t;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            if (ks < 2) dmma(x[0], a4[ks], Di[4 * ks]);  // rows 0..7 of a lower-triangular inverse end at p = 7
            dmma(x[1], a4[ks], Di[8 * PD + 4 * ks]);
        }
        __syncwarp();
#pragma unroll
        for (int fc = 0; fc < 2; ++fc) row[cb * SB + 8 * fc + 2 * t] = x[fc][0], row[cb * SB + 8 * fc + 2 * t + 1] = x[fc][1];
        __syncwarp();
        if (cb + 1 < T / SB) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) a4[ks] = row[cb * SB + 4 * ks + t];
#pragma unroll
            for (int cbp = cb + 1; cbp < T / SB; ++cbp) {
                double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
                const double *Lb = Ls + (cbp * SB + g) * PM + cb * SB + t;
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) {
                    dmma(acc[0], a4[ks], Lb[4 * ks]);
                    dmma(acc[1], a4[ks], Lb[8 * PM + 4 * ks]);
                }
#pragma unroll
                for (int fc = 0; fc < 2; ++fc) {
                    row[cbp * SB + 8 * fc + 2 * t] -= acc[fc][0];
                    row[cbp * SB + 8 * fc + 2 * t + 1] -= acc[fc][1];
                }
            }
            __syncwarp();
        }
    }
}

// sum_p Xs[r][p] * y[p], r = tid / 4; thread (r, q) takes p = q, q + 4, ...; valid on the q == 0 lane
__device__ __forceinline__ double row_dot(const double *Xs, const double *y) {
    const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int u = 0; u < SB; u += 2) {
        s0 = fma(Xs[r * PM + q + 4 * u], y[q + 4 * u], s0);
        s1 = fma(Xs[r * PM + q + 4 * u + 4], y[q + 4 * u + 4], s1);
    }
    double s = s0 + s1;
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 1);
    s += __shfl_xor_sync(0xFFFFFFFFu, s, 2);
    return s;
}

// ---- Cholesky of the 64 x 64 diagonal tile in shared memory --------------------------------------
// Per 16 columns (a "panel") the only serial work is the panel itself, and it lives in ONE warp's
// registers: lane l holds rows b + l and b + 32 + l of the panel, column values travel by shuffles.
// Within the panel the column-to-column dependency is the next pivot alone:
//     d' = a[J+1][J+1] - a[J+1][J]^2 / d      (reciprocal by MUFU seed + two Newton steps)
// which lane J+1 forms and broadcasts first; the scaling by rsqrt(d) and the rank-1 update of the other
// columns follow in its shadow.  Everything else runs on the other warps while warp 0 is in the next
// panel: the update of the far columns (DMMA on 8 x 8 fragments), the 16 x 16 block inverse, and the
// right-hand side, which rides along as row 64 (forward substitution + its own updates, one panel behind).
__device__ __forceinline__ double rcp_nr(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    y = fma(y, e, y);
    e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
// 1 / sqrt(d) for a positive, finite, normal d: MUFU seed + one third-order step, no special-case branch —
// the library rsqrt carries one, which ends the basic block and keeps the scheduler from interleaving this chain
// with the reciprocal's (a single warp issues in order: chains it cannot interleave simply add up)
__device__ __forceinline__ double rsqrt_nr(double d) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double t = d * y;
    const double e = fma(-t, y, 1.0);
    const double q = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, q, y);
}
template <int J, int C>
__device__ __forceinline__ void fac2_upd(double (&a1)[SB], double (&a2)[SB], const double *col) {
    if constexpr (C < SB) {
        const double lcj = col[C];  // L[b + C][b + J], one broadcast read (measured: a chain of 64-bit shuffles
                                    // costs 27 cycles each and a single warp cannot overlap 30 of them per column)
        a1[C] = fma(-a1[J], lcj, a1[C]);
        a2[C] = fma(-a2[J], lcj, a2[C]);
        fac2_upd<J, C + 1>(a1, a2, col);
    }
}
template <int J>
__device__ __forceinline__ void fac2_col(double (&a1)[SB], double (&a2)[SB], const int lane, double d, double &inv_mine,
                                         bool &bad, double *sc) {
    if constexpr (J < SB) {
        if (!(d > 0.0) || !isfinite(d)) bad = true, d = 1.0;
        double dnext = 1.0;
        if constexpr (J + 1 < SB) {
            const double rd = rcp_nr(d);
            const double dn = fma(-(a1[J] * a1[J]), rd, a1[J + 1]);  // meaningful on lane J + 1
            dnext = __shfl_sync(0xFFFFFFFFu, dn, J + 1);
        }
        const double inv = rsqrt_nr(d);
        if (lane == J) inv_mine = inv;
        a1[J] *= inv, a2[J] *= inv;
        double *col = sc + (J & 1) * SB;  // the scaled column of the diagonal block, double-buffered
        if (lane < SB) col[lane] = a1[J];
        __syncwarp();
        fac2_upd<J, J + 1>(a1, a2, col);
        if constexpr (J + 1 < SB) {
            if (lane == J + 1) a1[J + 1] = dnext;  // one value for the pivot, the one its rsqrt sees
        }
        fac2_col<J + 1>(a1, a2, lane, dnext, inv_mine, bad, sc);
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
__device__ __forceinline__ void block_inverse(double (*D)[PA], const double *dinvd, double *Dinv, const int b) {
    const int lane = threadIdx.x & 31, rl = lane & 15;
    double x[SB];
    inv_row<0>(x, D, dinvd, b, rl);
    if (lane < SB) {
        double *X = Dinv + (b / SB) * SB * PD;
#pragma unroll
        for (int r = 0; r < SB; ++r) X[r * PD + rl] = x[r];  // X[r][rl]; zero above the diagonal by construction
    }
}

// one 8 x 8 fragment of the trailing update from panel columns [pb, pb + 16):
// D[8 fr + .][8 fc + .] -= D[8 fr + .][pb ..] D[8 fc + .][pb ..]^T  (one warp, four DMMAs)
__device__ __forceinline__ void frag_update(double (*D)[PA], const int pb, const int fr, const int fc) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    double acc[2] = {0.0, 0.0};
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) dmma(acc, D[8 * fr + g][pb + 4 * ks + t], D[8 * fc + g][pb + 4 * ks + t]);
    D[8 * fr + g][8 * fc + 2 * t] -= acc[0];
    D[8 * fr + g][8 * fc + 2 * t + 1] -= acc[1];
}

// row 64 (the right-hand side) against panel pb: forward substitution of its 16 entries (lane c owns
// entry c, the solved ones are broadcast), then its share of the trailing update
__device__ __forceinline__ void rhs_row_step(double (*D)[PA], const double *dinvd, const int pb, const bool last) {
    const int lane = threadIdx.x & 31;
    double r = lane < SB ? D[T][pb + lane] : 0.0, x = 0.0;
#pragma unroll
    for (int p = 0; p < SB; ++p) {
        const double xp = __shfl_sync(0xFFFFFFFFu, r * dinvd[pb + p], p);
        if (lane == p) x = xp;
        if (lane > p && lane < SB) r = fma(-xp, D[pb + lane][pb + p], r);
    }
    if (lane < SB) D[T][pb + lane] = x;
    __syncwarp();
    if (!last)
        for (int c2 = pb + SB + lane; c2 < T; c2 += 32) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int p = 0; p < SB; p += 2) {
                s0 = fma(D[T][pb + p], D[c2][pb + p], s0);
                s1 = fma(D[T][pb + p + 1], D[c2][pb + p + 1], s1);
            }
            D[T][c2] -= s0 + s1;
        }
}

// In-place Cholesky of the 64 x 64 tile in shared memory D (lower part); dinvd[64] receives
// 1 / L[r][r], Dinv[4][16][PD] the inverses of the four 16 x 16 diagonal blocks of L.
// Row 64 of D holds the right-hand side of this block column and leaves as y = L^-1 rhs.
__device__ __forceinline__ void potrf64(double (*D)[PA], double *dinvd, double *Dinv, double *sc, bool &bad_out, long long *tr = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    bool bad = false;
    for (int b = 0; b < T; b += SB) {
        // ---- phase 1: warp 0 factors panel b; the others finish panel b - 16
        if (warp == 0) {
            const int r1 = b + lane, r2 = b + 32 + lane;
            double a1[SB], a2[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a1[c] = r1 < T ? D[r1][b + c] : 0.0, a2[c] = r2 < T ? D[r2][b + c] : 0.0;
            double inv_mine = 1.0;
            fac2_col<0>(a1, a2, lane, __shfl_sync(0xFFFFFFFFu, a1[0], 0), inv_mine, bad, sc);
            if (lane < SB) {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c <= lane) D[r1][b + c] = a1[c];
                dinvd[r1] = inv_mine;
            } else if (r1 < T) {
#pragma unroll
                for (int c = 0; c < SB; ++c) D[r1][b + c] = a1[c];
            }
            if (r2 < T) {
#pragma unroll
                for (int c = 0; c < SB; ++c) D[r2][b + c] = a2[c];
            }
            if (tr && b == 0 && tid == 0) tr[11] = clock64();
        } else if (b > 0) {
            const int pb = b - SB;
            if (warp == 1) {
                rhs_row_step(D, dinvd, pb, false);
            } else if (warp == 7) {
                block_inverse(D, dinvd, Dinv, pb);
            } else {  // warps 2..6: columns [b + 16, 64) from panel pb (the first 16 were done in phase 2)
                const int f0 = (b + SB) / 8;
                int f = 0;
                for (int fr = f0; fr < T / 8; ++fr)
                    for (int fc = f0; fc <= fr; ++fc, ++f)
                        if (f % 5 == warp - 2) frag_update(D, pb, fr, fc);
            }
        }
        __syncthreads();
        if (tr && b == 0 && tid == 0) tr[12] = clock64();
        // ---- phase 2: columns [b + 16, b + 32), rows [b + 16, 64) from panel b — what the next panel needs
        if (b + SB < T) {
            const int f0 = (b + SB) / 8;
            int f = 0;
            for (int fc = f0; fc < f0 + 2; ++fc)
                for (int fr = fc; fr < T / 8; ++fr, ++f)
                    if ((f & 7) == warp) frag_update(D, b, fr, fc);
            __syncthreads();
        }
        if (tr && b == 0 && tid == 0) tr[13] = clock64();
    }
    // tail: the last panel's block inverse and the right-hand side against it
    if (warp == 1)
        rhs_row_step(D, dinvd, T - SB, true);
    else if (warp == 7)
        block_inverse(D, dinvd, Dinv, T - SB);
    __syncthreads();
    bad_out = bad;
}

// shared-memory plan (doubles)
constexpr int kOffD = 0;                      // [65][PA]  chain: the diagonal tile + rhs row; backward: L_kk
constexpr int kOffB = (T + 1) * PA + 1;       // [64][PM]  chain / P: the tile being substituted; U: X operand
static_assert(kOffB % 2 == 0, "cp.async destinations are 16-byte aligned");
constexpr int kOffL = kOffB + T * PM;         // [64][PM]  chain / P: L of the pivot column; U: Y operand
constexpr int kOffDinv0 = kOffL + T * PM, kOffDinv1 = kOffDinv0 + kDinvSm;
constexpr int kOffVec = kOffDinv1 + kDinvSm;  // dinvd[64], y[64], yprev[64], rhs[64], tmp[64], + 8 x 64 scratch
constexpr int kChainDoubles = kOffVec + 13 * T;
constexpr int kWorkerLayoutDoubles = 4 * T * PM + 2 * kDinvSm + 2 * T;  // run_workers: two operand sets
constexpr int kSmemDoubles = kChainDoubles > kWorkerLayoutDoubles ? kChainDoubles : kWorkerLayoutDoubles;
constexpr int kSmemBytes = kSmemDoubles * (int)sizeof(double);

// ---- the chain: F tasks -------------------------------------------------------------------------
// A CTA claims its next task while it still works on the current one (the counter round trip and the
// record load hide behind the factorisation).
struct TaskSlot {
    int idx;
    int4 r[3];
};
__device__ __forceinline__ void claim_task(TaskSlot *slot, int *counter, const int4 *tasks, int n, int n_int4) {
    const int idx = atomicAdd(counter, 1);
    slot->idx = idx;
    if (idx < n)
        for (int u = 0; u < n_int4; ++u) slot->r[u] = __ldg(tasks + (size_t)n_int4 * idx + u);
}

__device__ __forceinline__ void run_chain(const CholArgs &A) {
    double *const sm = g_sm;
    const int tid = threadIdx.x;
    double(*D)[PA] = reinterpret_cast<double(*)[PA]>(sm + kOffD);
    double *Bs = sm + kOffB, *Ls = sm + kOffL;
    double *dinvd = sm + kOffVec, *yprev = sm + kOffVec + 2 * T;
    const WarpMap wm;
    __shared__ TaskSlot slot_s;
    bool any_bad = false;
    int cached_k = -1, par = 0;  // L of column cached_k is in Ls, its block inverses in Dinv[par ^ 1]
    const bool tron = A.trace != nullptr && tid == 0;
    if (tid == 0) claim_task(&slot_s, A.f.next_f(), A.p.ftasks, A.p.n_f, 2);
    __syncthreads();
    for (;;) {
        const int ft = slot_s.idx;
        if (ft >= A.p.n_f) break;
        const int4 r0 = slot_s.r[0], r1 = slot_s.r[1];
        __syncthreads();
        if (tid == 0) claim_task(&slot_s, A.f.next_f(), A.p.ftasks, A.p.n_f, 2);  // read after the last barrier of this task
#define XRB_CLK(slot) \
    if (tron && ft < 256) A.trace[ft * 16 + (slot)] = clock64();
        XRB_CLK(0)
        const int k = r0.x, kp = r0.y, s_kk = r0.z, s_kkp = r0.w, s_kpkp = r1.x, need_kk = r1.y, need_kkp = r1.z;
        double *Dinv = sm + (par ? kOffDinv1 : kOffDinv0), *Dinvprev = sm + (par ? kOffDinv0 : kOffDinv1);
        const bool couple = kp >= 0;
        if (!cta_wait(A.f.upd() + s_kk, need_kk, couple ? A.f.upd() + s_kkp : nullptr, need_kkp,
                      couple ? A.f.diag_done() + kp : nullptr, 1, nullptr, 0, A.f.abort_flag(), A.fail))
            return;
        if (couple) {
            tile_to_smem_async(Bs, A.tiles + (size_t)s_kkp * T * T);
            if (cached_k != kp) {
                tile_to_smem_async(Ls, A.tiles + (size_t)s_kpkp * T * T);
                dinv_to_smem(Dinvprev, A.dinv + (size_t)kp * 4 * SB * SB);
            }
            cp_async_commit();
            if (cached_k != kp && tid < T) yprev[tid] = __ldcg(A.rhs + kp * T + tid);
        }
        tile_to_smem_padded(D, A.tiles + (size_t)s_kk * T * T);
        if (tid < T) D[T][tid] = __ldcg(A.rhs + k * T + tid);  // the right-hand side rides along as row 64
        cp_async_wait_all();
        __syncthreads();
        XRB_CLK(1)
        if (couple) {
            trsm64(Bs, Ls, Dinvprev);  // Bs = L_k,kp
            __syncthreads();
            XRB_CLK(7)
            // publish L_k,kp first: the workers' updates of column kp wait for it
            tile_from_smem(A.tiles + (size_t)s_kkp * T * T, Bs);
            XRB_CLK(2)
            {
                const double s = row_dot(Bs, yprev);
                if ((tid & 3) == 0) D[T][tid >> 2] -= s;
            }
            XRB_CLK(9)
            // diagonal update, lower fragments only: the 36 fragments (R, C <= R) are dealt round-robin, 5 or 4
            // per warp — a warp issues one DMMA per ~32 cycles, so the busiest warp sets the time
            {
                const int wq = tid >> 5;
                int fR[5], fC[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const int f = wq + 8 * i;
                    int R = 0;
                    while ((R + 1) * (R + 2) / 2 <= f) ++R;  // f = R (R + 1) / 2 + C
                    fR[i] = f < 36 ? R : -1, fC[i] = f - R * (R + 1) / 2;
                }
                double acc[5][2];
#pragma unroll
                for (int i = 0; i < 5; ++i) acc[i][0] = acc[i][1] = 0.0;
                const double *base = Bs + wm.g * PM + wm.t;
#pragma unroll 2
                for (int p = 0; p < T; p += 4) {
#pragma unroll
                    for (int i = 0; i < 5; ++i)
                        if (fR[i] >= 0) dmma(acc[i], base[fR[i] * 8 * PM + p], base[fC[i] * 8 * PM + p]);
                }
                XRB_CLK(10)
#pragma unroll
                for (int i = 0; i < 5; ++i)
                    if (fR[i] >= 0) {
                        const int r = 8 * fR[i] + wm.g, c = 8 * fC[i] + 2 * wm.t;
                        D[r][c] -= acc[i][0], D[r][c + 1] -= acc[i][1];
                    }
            }
            // the flag for L_k,kp goes out while the factorisation starts (its first barrier absorbs the fence)
            cta_publish_async(A.f.pdone() + s_kkp, 1);
        }
        XRB_CLK(3)
        bool bad = false;
        potrf64(D, dinvd, Dinv, sm + kOffVec + 4 * T, bad, tron && ft < 256 ? A.trace + ft * 16 : nullptr);
        any_bad |= bad;
        XRB_CLK(4)
        {
            double *dst = A.tiles + (size_t)s_kk * T * T;
            for (int idx = tid; idx < T * T / 2; idx += kThreads) {
                const int r = idx >> 5, c = 2 * (idx & 31);
                const double2 v = make_double2(D[r][c], D[r][c + 1]);
                reinterpret_cast<double2 *>(dst)[idx] = v;
                *reinterpret_cast<double2 *>(Ls + r * PM + c) = v;  // stays for the next diagonal tile of this chain
            }
            double *dd = A.dinv + (size_t)k * 4 * SB * SB;
            for (int idx = tid; idx < 4 * SB * SB; idx += kThreads) dd[idx] = Dinv[(idx >> 4) * PD + (idx & 15)];
            if (tid < T) {  // row 64 left the factorisation as y_k = L_kk^-1 rhs_k
                const double v = D[T][tid];
                A.rhs[k * T + tid] = v;
                yprev[tid] = v;
            }
        }
        XRB_CLK(5)
        cta_publish_async(A.f.diag_done() + k, 1);
        cached_k = k, par ^= 1;
        XRB_CLK(6)
#undef XRB_CLK
    }
    if (any_bad && tid == 0) *A.fail = 1.0;
}

// ---- workers ------------------------------------------------------------------------------------
// A worker keeps two tasks claimed beyond the one it runs.  While it waits for the last flag of the
// current task it samples the operand flags of the next one in the same polling round; when they are
// up, the next task's tiles stream into the other half of shared memory (cp.async) underneath the
// current product.
//   P(i,k)   L_ik = A_ik L_kk^-T (blocked substitution with the 16 x 16 diagonal-block inverses)
//   U(i,j,k) A_ij -= L_ik L_jk^T; the diagonal ones also carry the forward substitution rhs_i -= L_ik y_k
//   PU(i,k)  P(i,k) and, while L_ik is still in shared memory, U(i,pk,k) for the next column pk that meets
//            row i: the row's fill advances one task per column
constexpr int kWOffX = 0, kWOffY = T * PM, kWSet = 2 * T * PM;  // two operand sets: [X | Y] [X | Y]
constexpr int kWOffDinv = 2 * kWSet, kWOffVec = kWOffDinv + 2 * kDinvSm;
constexpr int kWorkerDoubles = kWOffVec + 2 * T;
static_assert(kWorkerDoubles <= kSmemDoubles, "the kernel's shared memory covers the worker layout");

// Lanes 0..2 of warp 0 spin on the blocking flags, lanes 3..4 sample the optional ones in the same
// rounds; ready = every optional flag was up when the blocking ones were.
__device__ __forceinline__ bool cta_wait_peek(const int *b0, int w0, const int *b1, int w1, const int *b2, int w2, const int *o0,
                                              int ow0, const int *o1, int ow1, bool have_opt, bool &ready, int *abort_flag,
                                              double *fail, long long *wait_cycles) {
    __shared__ int ok_s, ready_s;
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const int *f = lane == 0 ? b0 : lane == 1 ? b1 : lane == 2 ? b2 : lane == 3 ? o0 : lane == 4 ? o1 : nullptr;
        const int w = lane == 0 ? w0 : lane == 1 ? w1 : lane == 2 ? w2 : lane == 3 ? ow0 : ow1;
        const bool optional = lane >= 3;
        const long long t0 = wait_cycles ? clock64() : 0;
        bool done = f == nullptr;
        unsigned spins = 0;
        int ok = 1;
        for (;;) {
            if (!done) done = ld_acquire(f) >= w;
            if (__all_sync(0xFFFFFFFFu, done || optional)) break;
            if ((++spins & 255u) == 0) {
                const bool ab = spins > kSpinLimit || ld_acquire(abort_flag) != 0;
                if (__any_sync(0xFFFFFFFFu, ab)) {
                    ok = 0;
                    break;
                }
            }
        }
        const bool all_opt = __all_sync(0xFFFFFFFFu, done || !optional);
        if (lane == 0) {
            if (!ok) {
                st_release(abort_flag, 1);
                *fail = 2.0;
            }
            ok_s = ok, ready_s = have_opt && all_opt && ok;
            if (wait_cycles) *wait_cycles += clock64() - t0;
        }
    }
    __syncthreads();
    const bool ok = ok_s != 0;
    ready = ready_s != 0;
    __syncthreads();
    return ok;
}

struct WTask {
    int type, s_ik, s_jk, s_ij, seq, k, i;
    __device__ WTask(const int4 r0, const int4 r1) : type(r0.x), s_ik(r0.y), s_jk(r0.z), s_ij(r0.w), seq(r1.x), k(r1.y), i(r1.z) {}
    __device__ bool diag() const { return s_ik == s_jk; }
};

// operand flags: P / PU wait for the tile's last update and the factored diagonal tile, U for both panel tiles
__device__ __forceinline__ void operand_flags(const CholArgs &A, const WTask &t, const int *&f0, int &w0, const int *&f1, int &w1) {
    if (t.type == TASK_U) {
        f0 = A.f.pdone() + t.s_ik, w0 = 1;
        f1 = t.diag() ? nullptr : A.f.pdone() + t.s_jk, w1 = 1;
    } else {
        f0 = A.f.upd() + t.s_ik, w0 = t.type == TASK_P ? t.seq : (t.seq & 0xFFFF);
        f1 = A.f.diag_done() + t.k, w1 = 1;
    }
}

__device__ __forceinline__ void issue_operands(const CholArgs &A, const WTask &t, int set) {
    double *X = g_sm + set * kWSet + kWOffX, *Y = g_sm + set * kWSet + kWOffY;
    tile_to_smem_async(X, A.tiles + (size_t)t.s_ik * T * T);
    if (t.type != TASK_U) {
        tile_to_smem_async(Y, A.tiles + (size_t)t.s_jk * T * T);  // s_jk holds the diagonal tile's slot
        double *Dinv = g_sm + kWOffDinv + set * kDinvSm;
        const double *src = A.dinv + (size_t)t.k * 4 * SB * SB;
        for (int idx = threadIdx.x; idx < 4 * SB * SB / 2; idx += kThreads) cp_async16(Dinv + (idx >> 3) * PD + 2 * (idx & 7), src + 2 * idx);
    } else if (!t.diag()) {
        tile_to_smem_async(Y, A.tiles + (size_t)t.s_jk * T * T);
    } else if (threadIdx.x < T / 2) {
        cp_async16(g_sm + kWOffVec + set * T + 2 * threadIdx.x, A.rhs + t.k * T + 2 * threadIdx.x);
    }
    cp_async_commit();
}

// C (global tile) -= X Y^T with X, Y in shared memory; the C fragments were requested into c[][] before
__device__ __forceinline__ void update_tile(double *C, const double2 (&c)[2][4], const double *Xs, const double *Ys, const WarpMap &wm) {
    double acc[2][4][2] = {};
    gemm_nt(Xs, Ys, wm, acc);
#pragma unroll
    for (int fr = 0; fr < 2; ++fr)
#pragma unroll
        for (int fc = 0; fc < 4; ++fc)
            *reinterpret_cast<double2 *>(C + (wm.r0 + 8 * fr + wm.g) * T + wm.c0 + 8 * fc + 2 * wm.t) =
                make_double2(c[fr][fc].x - acc[fr][fc][0], c[fr][fc].y - acc[fr][fc][1]);
}
__device__ __forceinline__ void load_c_frags(const double *C, double2 (&c)[2][4], const WarpMap &wm) {
#pragma unroll
    for (int fr = 0; fr < 2; ++fr)
#pragma unroll
        for (int fc = 0; fc < 4; ++fc)
            c[fr][fc] = __ldcg(reinterpret_cast<const double2 *>(C + (wm.r0 + 8 * fr + wm.g) * T + wm.c0 + 8 * fc + 2 * wm.t));
}

__device__ __forceinline__ void run_workers(const CholArgs &A) {
    const int tid = threadIdx.x;
    __shared__ TaskSlot slots[3];
    // debug statistics (thread 0, only when the trace is armed): cycles waiting on flags, in P / U tasks
    long long st_wait = 0, st_p = 0, st_u = 0, n_p = 0, n_u = 0;
    long long *wc = A.trace ? &st_wait : nullptr;
    const long long t_begin = A.trace ? clock64() : 0;
    const WarpMap wm;
    // with few tasks per worker (a dissected band) a claimed-but-waiting task may be the one another CTA could run:
    // then every worker takes one task at a time
    const bool pipe = A.pipeline != 0;
    if (tid == 0) {
        claim_task(&slots[0], A.f.next_w(), A.p.wtasks, A.p.n_w, 2);
        if (pipe) claim_task(&slots[1], A.f.next_w(), A.p.wtasks, A.p.n_w, 2);
    }
    __syncthreads();
    bool have = false;  // the operands of the current task are already on their way (set `set`)
    int set = 0;
    for (int it = 0;; ++it) {
        const TaskSlot &cs = slots[it % 3], &ns = slots[(it + 1) % 3];
        if (cs.idx >= A.p.n_w) break;
        const WTask cur(cs.r[0], cs.r[1]);
        const bool have_next = pipe && ns.idx < A.p.n_w;
        const WTask nxt(ns.r[0], ns.r[1]);
        // the claim of the task after next is spread over this one so that nobody waits for it: counter
        // now, record after the flags, shared-memory slot after the product (read from the next iteration on)
        int claim_idx = 0;
        int4 cr0 = make_int4(0, 0, 0, 0), cr1 = cr0;
        if (pipe && tid == kThreads - 1) claim_idx = atomicAdd(A.f.next_w(), 1);
        TaskSlot &claim_slot = slots[(it + 2) % 3];
        const long long t0 = A.trace ? clock64() : 0;
        const int *f0 = nullptr, *f1 = nullptr, *o0 = nullptr, *o1 = nullptr;
        int w0 = 0, w1 = 0, ow0 = 0, ow1 = 0;
        if (!have) operand_flags(A, cur, f0, w0, f1, w1);
        if (have_next) operand_flags(A, nxt, o0, ow0, o1, ow1);
        bool ready = false;
        if (!cta_wait_peek(f0, w0, f1, w1, cur.type == TASK_U ? A.f.upd() + cur.s_ij : nullptr, cur.seq, o0, ow0, o1, ow1,
                           have_next, ready, A.f.abort_flag(), A.fail, wc))
            return;
        if (!have) issue_operands(A, cur, set);
        if (pipe && tid == kThreads - 1 && claim_idx < A.p.n_w) cr0 = __ldg(A.p.wtasks + 2 * claim_idx), cr1 = __ldg(A.p.wtasks + 2 * claim_idx + 1);
        double *Xs = g_sm + set * kWSet + kWOffX, *Ys = g_sm + set * kWSet + kWOffY;
        if (cur.type == TASK_U) {
            double *C = A.tiles + (size_t)cur.s_ij * T * T;
            double2 c[2][4];
            load_c_frags(C, c, wm);
            if (ready) issue_operands(A, nxt, set ^ 1);
            if (ready) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
            __syncthreads();
            update_tile(C, c, Xs, cur.diag() ? Xs : Ys, wm);
            if (pipe && tid == kThreads - 1) claim_slot.idx = claim_idx, claim_slot.r[0] = cr0, claim_slot.r[1] = cr1;
            if (cur.diag()) {  // forward substitution rides on the diagonal update: rhs_i -= L_ik y_k, in the same order
                const double s = row_dot(Xs, g_sm + kWOffVec + set * T);
                if ((tid & 3) == 0) {
                    double *p = A.rhs + cur.i * T + (tid >> 2);
                    *p = __ldcg(p) - s;
                }
            }
            cta_publish_async(A.f.upd() + cur.s_ij, cur.seq + 1);
            if (A.trace) st_u += clock64() - t0, ++n_u;
        } else {
            if (ready) issue_operands(A, nxt, set ^ 1);
            if (ready) cp_async_wait_group<1>(); else cp_async_wait_group<0>();
            __syncthreads();
            trsm64(Xs, Ys, g_sm + kWOffDinv + set * kDinvSm);
            if (pipe && tid == kThreads - 1) claim_slot.idx = claim_idx, claim_slot.r[0] = cr0, claim_slot.r[1] = cr1;
            __syncthreads();
            tile_from_smem(A.tiles + (size_t)cur.s_ik * T * T, Xs);
            cta_publish_async(A.f.pdone() + cur.s_ik, 1);
            if (cur.type == TASK_PU) {
                // the update of tile (i, pk): Y = L_pk,k (published by the chain or a worker about now)
                const int useq = cur.seq >> 16, s_pk_k = cur.i;
                if (!cta_wait(A.f.pdone() + s_pk_k, 1, A.f.upd() + cur.s_ij, useq, nullptr, 0, nullptr, 0, A.f.abort_flag(), A.fail, wc))
                    return;
                tile_to_smem_async(Ys, A.tiles + (size_t)s_pk_k * T * T);
                cp_async_commit();
                double *C = A.tiles + (size_t)cur.s_ij * T * T;
                double2 c[2][4];
                load_c_frags(C, c, wm);
                cp_async_wait_all();
                __syncthreads();
                update_tile(C, c, Xs, Ys, wm);
                cta_publish_async(A.f.upd() + cur.s_ij, useq + 1);
            }
            if (A.trace) st_p += clock64() - t0, ++n_p;
        }
        have = ready, set ^= 1;
        if (!pipe) {
            __syncthreads();
            if (tid == 0) claim_task(&slots[(it + 1) % 3], A.f.next_w(), A.p.wtasks, A.p.n_w, 2);
            __syncthreads();
        }
    }
    if (A.trace && tid == 0) {
        unsigned long long *st = reinterpret_cast<unsigned long long *>(A.trace) + 4096;
        atomicAdd(st + 0, (unsigned long long)st_wait), atomicAdd(st + 1, (unsigned long long)st_p);
        atomicAdd(st + 2, (unsigned long long)st_u), atomicAdd(st + 3, (unsigned long long)n_p);
        atomicAdd(st + 4, (unsigned long long)n_u), atomicAdd(st + 5, (unsigned long long)(clock64() - t_begin));
        atomicAdd(st + 6, 1ull);
    }
}

__global__ void __launch_bounds__(kThreads, 1) k_tile_cholesky(CholArgs A) {
    if ((int)blockIdx.x < A.p.n_chain_f)
        run_chain(A);
    else
        run_workers(A);
}

// ---- back-substitution -------------------------------------------------------------------------
// x_k = L_kk^-T (y_k - sum_{i in struct(k)} L_ik^T x_i).  Chain CTAs take the columns in reverse-level
// order and handle the (up to) three nearest tiles themselves; a worker per column accumulates the far
// tiles as their x_i appear.
//
// Partial column sums of one row-major tile in global memory: thread (rq = tid / 32, cp = tid % 32)
// adds rows 8 rq .. 8 rq + 7 of columns 2 cp, 2 cp + 1 (a warp reads whole 512-byte rows).
struct TilePart {
    double2 v[8];
};
__device__ __forceinline__ void tile_rows_load(TilePart &tp, const double *__restrict__ Lt) {
    const int rq = threadIdx.x >> 5, cp = threadIdx.x & 31;
#pragma unroll
    for (int u = 0; u < 8; ++u) tp.v[u] = __ldcg(reinterpret_cast<const double2 *>(Lt + (8 * rq + u) * T) + cp);
}
__device__ __forceinline__ void tile_rows_fma(const TilePart &tp, const double *xs, double &s0, double &s1) {
    const int rq = threadIdx.x >> 5;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const double xv = xs[8 * rq + u];
        s0 = fma(tp.v[u].x, xv, s0);
        s1 = fma(tp.v[u].y, xv, s1);
    }
}
// red[8][64] <- the partial sums; after a barrier out[p] = sum_rq red[rq][p] in fixed order
__device__ __forceinline__ void part_store(double *red, double s0, double s1) {
    const int rq = threadIdx.x >> 5, cp = threadIdx.x & 31;
    red[rq * T + 2 * cp] = s0, red[rq * T + 2 * cp + 1] = s1;
}
__device__ __forceinline__ double part_sum(const double *red, int p) {
    double s = 0.0;
#pragma unroll
    for (int rq = 0; rq < 8; ++rq) s += red[rq * T + p];
    return s;
}

__global__ void __launch_bounds__(kThreads, 1) k_tile_backsolve(CholArgs A) {
    double *const sm = g_sm;
    const int tid = threadIdx.x;
    __shared__ int task_s;
    double *red = sm + kOffVec + 5 * T;  // [8][64]  (workers; the chain places its own)
    if ((int)blockIdx.x < A.p.n_chain_b) {
        // the chain: L_kk and its block inverses of the NEXT column stream into the other buffer (cp.async)
        // while this column is solved; the solve itself uses the whole CTA (matrix-vector steps per 16-block)
        constexpr int PB = 66;  // pitch of L_kk here: rows stay 16-byte aligned for cp.async
        double *Lbuf[2] = {sm, sm + T * PB};
        double *Dbuf[2] = {sm + 2 * T * PB, sm + 2 * T * PB + kDinvSm};
        double *vec = sm + 2 * T * PB + 2 * kDinvSm;  // s[64], xk[64], xs[3][64], red[8][64]
        double *s = vec, *xk = vec + T, *xs = vec + 2 * T;
        red = vec + 5 * T;
        __shared__ TaskSlot slots[3];  // this column, the next one (complete), the one after (being claimed)
        auto prefetch = [&](const TaskSlot &ts, int buf) {
            const int k = ts.r[0].x, s_kk = ts.r[0].y;
            const double *src = A.tiles + (size_t)s_kk * T * T;
            for (int idx = tid; idx < T * T / 2; idx += kThreads) cp_async16(Lbuf[buf] + (idx >> 5) * PB + 2 * (idx & 31), src + 2 * idx);
            const double *dsrc = A.dinv + (size_t)k * 4 * SB * SB;
            for (int idx = tid; idx < 4 * SB * SB / 2; idx += kThreads) cp_async16(Dbuf[buf] + (idx >> 3) * PD + 2 * (idx & 7), dsrc + 2 * idx);
            cp_async_commit();
        };
        if (tid == 0) {
            claim_task(&slots[0], A.f.next_b(), A.p.btasks, A.p.n_b, 3);
            claim_task(&slots[1], A.f.next_b(), A.p.btasks, A.p.n_b, 3);
        }
        __syncthreads();
        if (slots[0].idx < A.p.n_b) prefetch(slots[0], 0);
        for (int it = 0;; ++it) {
            const int buf = it & 1;
            const TaskSlot &cs = slots[it % 3], &ns = slots[(it + 1) % 3];
            if (cs.idx >= A.p.n_b) break;
            const int4 r0 = cs.r[0], r1 = cs.r[1], r2 = cs.r[2];
            const int k = r0.x, n_near = r0.z, has_far = r0.w;
            const int row[3] = {r1.x, r1.y, r1.z}, slot[3] = {r2.x, r2.y, r2.z};
            // the claim of the column after next is spread over this step so that nobody waits for it
            int claim_idx = 0;
            int4 cr0 = make_int4(0, 0, 0, 0), cr1 = cr0, cr2 = cr0;
            if (tid == kThreads - 1) claim_idx = atomicAdd(A.f.next_b(), 1);
            // the near tiles do not depend on the flags either: request them before waiting
            TilePart tp[kNearTiles];
#pragma unroll
            for (int a = 0; a < kNearTiles; ++a)
                if (a < n_near) tile_rows_load(tp[a], A.tiles + (size_t)slot[a] * T * T);
            if (!cta_wait(n_near > 0 ? A.f.xdone() + row[0] : nullptr, 1, n_near > 1 ? A.f.xdone() + row[1] : nullptr, 1,
                          n_near > 2 ? A.f.xdone() + row[2] : nullptr, 1, has_far ? A.f.wdone() + k : nullptr, 1,
                          A.f.abort_flag(), A.fail))
                return;
            if (tid == kThreads - 1 && claim_idx < A.p.n_b)
                cr0 = __ldg(A.p.btasks + 3 * claim_idx), cr1 = __ldg(A.p.btasks + 3 * claim_idx + 1), cr2 = __ldg(A.p.btasks + 3 * claim_idx + 2);
            const bool have_next = ns.idx < A.p.n_b;
            if (have_next) prefetch(ns, buf ^ 1);
            if (tid < n_near * T) xs[tid] = __ldcg(A.wpart + A.p.nt * T + row[tid >> 6] * T + (tid & 63));
            __syncthreads();
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int a = 0; a < kNearTiles; ++a)
                if (a < n_near) tile_rows_fma(tp[a], xs + a * T, s0, s1);
            part_store(red, s0, s1);
            if (have_next) cp_async_wait_group<1>(); else cp_async_wait_group<0>();  // this column's L_kk has landed
            __syncthreads();
            if (tid < T) {
                double v = __ldcg(A.rhs + k * T + tid) - part_sum(red, tid);
                if (has_far) v -= __ldcg(A.wpart + k * T + tid);
                s[tid] = v;
            }
            __syncthreads();
            // x_k = L_kk^-T s, 16-blocks bottom-up: x[cb] = Dinv_cb^T s[cb], then s[c] -= L[cb rows][c] x[cb] for c < 16 cb
            const double *L = Lbuf[buf], *Dinv = Dbuf[buf];
            for (int cb = T / SB - 1; cb >= 0; --cb) {
                {
                    const int l = tid >> 4, pp = tid & 15;  // output l, term pp
                    double v = pp >= l ? Dinv[cb * SB * PD + pp * PD + l] * s[cb * SB + pp] : 0.0;
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 8);
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 4);
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
                    if (pp == 0) xk[cb * SB + l] = v;
                }
                __syncthreads();
                if (cb > 0) {
                    const int c = tid >> 2, q = tid & 3;  // column c, rows 4q .. 4q+3 of the block
                    double v = 0.0;
                    if (c < cb * SB) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) v = fma(L[(cb * SB + 4 * q + j) * PB + c], xk[cb * SB + 4 * q + j], v);
                    }
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 2);
                    v += __shfl_xor_sync(0xFFFFFFFFu, v, 1);
                    if (q == 0 && c < cb * SB) s[c] -= v;
                    __syncthreads();
                }
            }
            if (tid < T) {
                const double v = xk[tid];
                A.x[k * T + tid] = v;
                A.wpart[A.p.nt * T + k * T + tid] = v;  // x for the other CTAs
            }
            if (tid == kThreads - 1) {
                TaskSlot &c2 = slots[(it + 2) % 3];
                c2.idx = claim_idx, c2.r[0] = cr0, c2.r[1] = cr1, c2.r[2] = cr2;
            }
            cta_publish_async(A.f.xdone() + k, 1);
        }
        return;
    }
    // workers: column k accumulates sum over its far tiles L_ik^T x_i as the x_i appear
    double *xs = sm + kOffVec;
    for (;;) {
        if (tid == 0) task_s = atomicAdd(A.f.next_wb(), 1);
        __syncthreads();
        const int wt = task_s;
        __syncthreads();
        if (wt >= A.p.n_wb) break;
        const int4 r = __ldg(A.p.wbtasks + wt);
        const int k = r.x;
        double s0 = 0.0, s1 = 0.0;
        for (int e = r.y; e < r.z; ++e) {
            const int i = __ldg(A.p.far_rows + e), sl = __ldg(A.p.far_slots + e);
            TilePart tp;
            tile_rows_load(tp, A.tiles + (size_t)sl * T * T);
            if (!cta_wait(A.f.xdone() + i, 1, nullptr, 0, nullptr, 0, nullptr, 0, A.f.abort_flag(), A.fail)) return;
            if (tid < T) xs[tid] = __ldcg(A.wpart + A.p.nt * T + i * T + tid);
            __syncthreads();
            tile_rows_fma(tp, xs, s0, s1);
            __syncthreads();
        }
        part_store(red, s0, s1);
        __syncthreads();
        if (tid < T) A.wpart[k * T + tid] = part_sum(red, tid);
        cta_publish(A.f.wdone() + k, 1);
        __syncthreads();
    }
}

// ---- host side -----------------------------------------------------------------------------------
struct DeviceInfo {
    int grid = 0;
    bool attr_set = false;
    long long *trace = nullptr;
};
constexpr int kMaxDevices = 64;
DeviceInfo g_dev[kMaxDevices];
std::mutex g_tmutex;
bool g_trace_on = false;

}  // namespace

void CholWorkspace::release() { flags.release(), wpart.release(); }

// `ws` belongs to the calling solver: several solvers may factor on one device at the same time
// (batched local BA), each with its own flags and partial sums.
int ba_launch_tile_cholesky_solve(const CholPlanDev &plan, CholWorkspace &ws, double *tiles, double *rhs, double *dinv,
                                  double *x_out, double *fail_flag, cudaStream_t st, int64_t *launches) {
    if (plan.nt <= 0) return XRB_OK;
    int dev = 0;
    XRB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDevices) {
        set_error("cholesky: device id %d outside [0, %d)", dev, kMaxDevices);
        return XRB_ERR_INVALID;
    }
    int grid = 0;
    long long *trace = nullptr;
    {
        std::lock_guard<std::mutex> lock(g_tmutex);
        DeviceInfo &di = g_dev[dev];
        if (!di.attr_set) {
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
            di.grid = sms;  // one CTA per SM
            di.attr_set = true;
        }
        if (g_trace_on && !di.trace) XRB_CUDA(cudaMalloc(&di.trace, (4096 + 16) * sizeof(long long)));
        grid = di.grid;
        trace = g_trace_on ? di.trace : nullptr;
    }
    const size_t nflags = Flags::count(plan.nt, plan.n_tiles);
    int rc;
    if ((rc = ws.flags.reserve(nflags * sizeof(int)))) return rc;
    if ((rc = ws.wpart.reserve(2 * (size_t)plan.nt * T * sizeof(double)))) return rc;
    if (trace) XRB_CUDA(cudaMemsetAsync(trace, 0, (4096 + 16) * sizeof(long long), st));
    XRB_CUDA(cudaMemsetAsync(ws.flags.p, 0, nflags * sizeof(int), st));
    CholArgs A;
    A.p = plan, A.tiles = tiles, A.rhs = rhs, A.dinv = dinv, A.x = x_out, A.wpart = ws.wpart.as<double>(), A.fail = fail_flag;
    A.f = Flags{ws.flags.as<int>(), plan.nt, plan.n_tiles};
    A.trace = trace;
    // CTAs beyond the number of tasks would only poll a counter once
    const int grid_f = std::max(1, std::min(grid, plan.n_chain_f + plan.n_w));
    A.pipeline = plan.n_w >= 32 * std::max(1, grid_f - plan.n_chain_f);
    const int grid_b = std::max(1, std::min(grid, plan.n_chain_b + plan.n_wb));
    void *args[] = {&A};
    XRB_CUDA(cudaLaunchCooperativeKernel((void *)k_tile_cholesky, dim3(grid_f), dim3(kThreads), args, kSmemBytes, st));
    XRB_CUDA(cudaLaunchCooperativeKernel((void *)k_tile_backsolve, dim3(grid_b), dim3(kThreads), args, kSmemBytes, st));
    g_launches.fetch_add(2, std::memory_order_relaxed);
    if (launches) *launches += 2;
    return XRB_OK;
}

// after a solve has completed (stream synchronised): did a wait give up?
int ba_tile_cholesky_aborted(const CholWorkspace &ws, int *aborted) {
    *aborted = 0;
    if (ws.flags.p) XRB_CUDA(cudaMemcpy(aborted, ws.flags.as<int>() + 1, sizeof(int), cudaMemcpyDeviceToHost));
    return XRB_OK;
}

// debug: clock64 stamps of the chain CTAs, 16 per F task (first 256 tasks), of the last factorisation
int ba_tile_cholesky_trace(int enable, long long *out, int cap) {
    int dev = 0;
    XRB_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lock(g_tmutex);
    if (enable) {
        g_trace_on = true;
        return 0;
    }
    g_trace_on = false;
    DeviceInfo &di = g_dev[dev];
    if (!di.trace || !out) return 0;
    XRB_CUDA(cudaDeviceSynchronize());
    const int n = std::min(cap, 4096 + 16);
    XRB_CUDA(cudaMemcpy(out, di.trace, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost));
    return n;
}

}  // namespace xrb
