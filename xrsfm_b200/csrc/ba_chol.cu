// ba_chol.cu — reduced-camera-system solver of path B: blocked FP64 Cholesky (lower, in
// place) limited to a half-bandwidth, forward substitution folded into the factorisation
// (the right-hand side rides along as row n of the matrix), blocked backward substitution.
//
// Replaces Ceres' SparseSchurComplementSolver factor+solve (selected by ba_solver.cc:74).
// Right-looking, block size 64:
//   chol_diag    one CTA: L_kk = chol(A_kk) and its inverse, both in shared memory
//   chol_panel   L_ik = A_ik Linv_kk^T   — FP64 GEMM tiles (64 x 64 x 64)
//   chol_update  A_ij -= L_ik L_jk^T     — FP64 GEMM tiles (128 x 128 x 64 or 64 x 64 x 64),
//                lower-triangular tile pairs inside the band; this is where the n^3/3 flops are
//   chol_backsolve  one launch per block column, bottom-up
// The whole launch sequence is fixed for a given (n, bandwidth) and is replayed as a CUDA
// graph.  Roofline: FP64 FMA pipe (64 FMA/clk/SM); see DESIGN.md §B.4.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "ba_kernels.cuh"

namespace xrb {

constexpr int NB = 64;  // block-column width == GEMM depth

// ---- optional timeline (xrb_debug_chol_trace): when switched on, CTA 0 of every kernel of the
// factorisation records (kernel id, step, %globaltimer at entry/exit) and the fused diagonal
// CTA also its phase cycle counts.  Off by default: one predicated load per kernel.
constexpr int kTraceCap = 4096, kTraceWords = 12;
__device__ int g_trace_on = 0;
__device__ unsigned int g_trace_n = 0;
__device__ long long g_trace[kTraceCap * kTraceWords];

__device__ __forceinline__ long long gtimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct Trace {  // lives in thread 0 of the recording CTA
    long long t0 = 0, ph[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    bool on = false;
    __device__ void begin(bool recorder) {
        on = recorder && g_trace_on != 0;
        if (on) t0 = gtimer();
    }
    __device__ void end(int id, int step) {
        if (!on) return;
        const unsigned slot = atomicAdd(&g_trace_n, 1u);
        if (slot >= (unsigned)kTraceCap) return;
        long long *r = g_trace + (size_t)slot * kTraceWords;
        r[0] = id, r[1] = step, r[2] = t0, r[3] = gtimer();
        for (int i = 0; i < 8; ++i) r[4 + i] = ph[i];
    }
};

// ---- 1. diagonal block ---------------------------------------------------------------------
// 64 x 64 Cholesky + inverse in one CTA (256 threads), blocked by 16:
//   for each 16-block: warp 0 factors the 16 x 16 diagonal sub-block and inverts it entirely
//   in registers (row per lane, values exchanged by shuffles); all warps then form the panel
//   below it (multiply by the 16 x 16 inverse) and apply the rank-16 trailing update.
//   Finally the off-diagonal 16-blocks of L^-1 are filled block-diagonal by block-diagonal:
//   X_ij = -X_ii (sum_k L_ik X_kj).
// ~13 barriers in total instead of 128 (one per column) in the unblocked form.
constexpr int SB = 16;
constexpr int kDiagSmem = (2 * NB * (NB + 1) + NB * (SB + 1) + NB) * (int)sizeof(double);  // A, X, T, 1/diag
__device__ int g_diag_variant = 2;  // 1: inverse of each 16-block on the chain; 2: on warp 7 (XRB_CHOL_DIAG)

// In-place factorisation of the 64 x 64 block held in shared memory A (lower part, rows >= kb
// made identity by the caller); on return A holds L and X holds L^-1.  256 threads.
// 16 x 16 factor and inverse of one warp, entirely in registers.  The loops are spelled as
// template recursion: with `#pragma unroll` the compiler kept the triangular inner loops rolled,
// which put a[] and x[] in local memory and made this serial section 5x slower (r01 trace:
// 14 k cycles per 16-block, 57 % of the fused diagonal CTA).
template <int J, int C>
__device__ __forceinline__ void fac_upd(double (&a)[SB], const int rl) {
    if constexpr (C < SB) {
        // applied on every lane: rows above C only collect values nobody reads (the entries
        // right of a row's diagonal), which saves the two FSELs a predicated update costs
        const double lcj = __shfl_sync(0xFFFFFFFFu, a[J], C);  // L[C][J]
        a[C] = fma(-a[J], lcj, a[C]);
        fac_upd<J, C + 1>(a, rl);
    }
}
template <int J>
__device__ __forceinline__ void fac_col(double (&a)[SB], const int rl, double &inv_mine, bool &bad) {
    if constexpr (J < SB) {
        double d = __shfl_sync(0xFFFFFFFFu, a[J], J);
        if (!(d > 0.0) || !isfinite(d)) bad = true, d = 1.0;
        // the 64 pivots are a serial chain: rsqrt + multiplies instead of the much longer
        // software sqrt and divide sequences
        const double inv = rsqrt(d);
        if (rl == J) inv_mine = inv;
        a[J] *= inv;  // L[r][J] for r >= J (lane J holds d itself: d * inv = sqrt(d))
        fac_upd<J, J + 1>(a, rl);
        fac_col<J + 1>(a, rl, inv_mine, bad);
    }
}
template <int R, int P>
__device__ __forceinline__ void inv_dot(const double (&a)[SB], const double (&x)[SB], double &v0, double &v1) {
    if constexpr (P < R) {
        const double l = __shfl_sync(0xFFFFFFFFu, a[P], R);  // L[R][P]
        if constexpr ((P & 1) != 0)
            v1 = fma(-l, x[P], v1);
        else
            v0 = fma(-l, x[P], v0);
        inv_dot<R, P + 1>(a, x, v0, v1);
    }
}
template <int R>
__device__ __forceinline__ void inv_row(const double (&a)[SB], double (&x)[SB], const int rl, const double inv_mine) {
    if constexpr (R < SB) {  // lane rl owns column rl of the inverse: x[R] = X[R][rl]
        double v0 = rl == R ? 1.0 : 0.0, v1 = 0.0;
        inv_dot<R, 0>(a, x, v0, v1);
        x[R] = (v0 + v1) * __shfl_sync(0xFFFFFFFFu, inv_mine, R);
        inv_row<R + 1>(a, x, rl, inv_mine);
    }
}

// Off-diagonal 16-blocks of X = L^-1 once the four diagonal 16-block inverses are in X.  256 threads;
// ends with a barrier.
__device__ void diag_inverse_levels(double (*A)[NB + 1], double (*X)[NB + 1], double (*T)[SB + 1]) {
    const int tid = threadIdx.x, ti = tid >> 4, tj = tid & 15;
    // ---- off-diagonal 16-blocks of X = L^-1 by recursive doubling:
    //   level 1: X10 = -X11 (L10 X00) and X32 = -X33 (L32 X22), two independent 16-block merges
    //   level 2: X[32:64, 0:32] = -X_BB (L_BA X_AA) with the completed 32 x 32 triangles
    {
        // level 1, T1[s][r][c] = sum_p L[o1 + r][o0 + p] X[o0 + p][o0 + c], o0 = 32 s, o1 = o0 + 16
        const int s = tid >> 7, r = (tid & 127) >> 4, c = tj, o0 = 32 * s, o1 = o0 + 16;
        double t0 = 0.0, t1 = 0.0;
#pragma unroll
        for (int p = 0; p < SB; ++p) {
            const double xv = X[o0 + p][o0 + c];
            t0 = fma(A[o1 + r][o0 + p], xv, t0);
            t1 = fma(A[o1 + r + 8][o0 + p], xv, t1);
        }
        T[16 * s + r][c] = t0, T[16 * s + r + 8][c] = t1;
        __syncthreads();
        double x0 = 0.0, x1 = 0.0;
#pragma unroll
        for (int p = 0; p < SB; ++p) {
            const double tv = T[16 * s + p][c];
            x0 = fma(X[o1 + r][o1 + p], tv, x0);
            x1 = fma(X[o1 + r + 8][o1 + p], tv, x1);
        }
        X[o1 + r][o0 + c] = -x0, X[o1 + r + 8][o0 + c] = -x1;  // block (1, 0): nobody reads it in this phase
    }
    __syncthreads();
    {
        // level 2: 2 x 2 register block per thread, rows ti and ti + 16, columns tj and tj + 16
        double *T2 = &T[0][0];  // viewed as [32][33]
        constexpr int L2 = 33, H = 32;
        double t[2][2] = {};
#pragma unroll 4
        for (int p = 0; p < H; ++p) {
            const double l0 = A[H + ti][p], l1 = A[H + ti + 16][p];
            const double x0 = X[p][tj], x1 = X[p][tj + 16];
            t[0][0] = fma(l0, x0, t[0][0]), t[0][1] = fma(l0, x1, t[0][1]);
            t[1][0] = fma(l1, x0, t[1][0]), t[1][1] = fma(l1, x1, t[1][1]);
        }
        T2[ti * L2 + tj] = t[0][0], T2[ti * L2 + tj + 16] = t[0][1];
        T2[(ti + 16) * L2 + tj] = t[1][0], T2[(ti + 16) * L2 + tj + 16] = t[1][1];
        __syncthreads();
        double y[2][2] = {};
#pragma unroll 4
        for (int p = 0; p < H; ++p) {
            const double b0 = X[H + ti][H + p], b1 = X[H + ti + 16][H + p];
            const double t0 = T2[p * L2 + tj], t1 = T2[p * L2 + tj + 16];
            y[0][0] = fma(b0, t0, y[0][0]), y[0][1] = fma(b0, t1, y[0][1]);
            y[1][0] = fma(b1, t0, y[1][0]), y[1][1] = fma(b1, t1, y[1][1]);
        }
        X[H + ti][tj] = -y[0][0], X[H + ti][tj + 16] = -y[0][1];
        X[H + ti + 16][tj] = -y[1][0], X[H + ti + 16][tj + 16] = -y[1][1];
    }
    __syncthreads();
}

__device__ void diag_factor(double (*A)[NB + 1], double (*X)[NB + 1], double (*T)[SB + 1], double *__restrict__ fail,
                            Trace *tr = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ti = tid >> 4, tj = tid & 15;
    bool bad = false;
    const bool tron = tr && tr->on;  // phase cycles: ph[1] warp factor, ph[2] panel, ph[3] trailing, ph[4] inverse
    long long tc = tron ? clock64() : 0;
#define XRB_TRACE_PHASE(k)                    \
    if (tron) {                               \
        const long long now = clock64();      \
        tr->ph[k] += now - tc, tc = now;      \
    }
    for (int b = 0; b < NB; b += SB) {
        if (warp == 0) {
            // ---- 16 x 16 factor: lane l (< 16) owns row b + l (lanes 16..31 mirror them)
            const int rl = lane & 15;
            double a[SB], x[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = A[b + rl][b + c];
            double inv_mine = 1.0;  // 1 / L[rl][rl]
            fac_col<0>(a, rl, inv_mine, bad);
            // ---- its inverse: lane l owns column l; L[r][p] broadcast from lane r
            inv_row<0>(a, x, rl, inv_mine);
            if (lane < SB) {
#pragma unroll
                for (int c = 0; c < SB; ++c) {
                    if (c <= rl) A[b + rl][b + c] = a[c];
                    X[b + c][b + rl] = c >= rl ? x[c] : 0.0;  // x[c] = X[c][column rl]
                }
            }
        }
        __syncthreads();
        XRB_TRACE_PHASE(1)
        const int below = NB - b - SB;  // rows under the sub-block
        if (below > 0) {
            // ---- panel: P[i][c] = sum_{p <= c} A[i][b+p] * Xbb[c][p]   (= A_panel * Lbb^-T)
            for (int idx = tid; idx < below * SB; idx += 256) {
                const int i = b + SB + (idx >> 4), c = idx & 15;
                double v = 0.0;
#pragma unroll
                for (int p = 0; p < SB; ++p)
                    if (p <= c) v = fma(A[i][b + p], X[b + c][b + p], v);
                T[i][c] = v;
            }
            __syncthreads();
            for (int idx = tid; idx < below * SB; idx += 256) {
                const int i = b + SB + (idx >> 4), c = idx & 15;
                A[i][b + c] = T[i][c];
            }
            __syncthreads();
            XRB_TRACE_PHASE(2)
            // ---- trailing update A[i][j] -= sum_p A[i][b+p] A[j][b+p], j <= i: thread (ti, tj) owns
            // the lattice rows ti + 16 u, columns tj + 16 v (v <= u), a 3 x 3 register block at most
            const int base = b + SB, nb16 = below / SB;
            double acc[3][3] = {};
#pragma unroll 4
            for (int p = 0; p < SB; ++p) {
                double ai[3], aj[3];
#pragma unroll
                for (int u = 0; u < 3; ++u) {
                    ai[u] = u < nb16 ? A[base + ti + 16 * u][b + p] : 0.0;
                    aj[u] = u < nb16 ? A[base + tj + 16 * u][b + p] : 0.0;
                }
#pragma unroll
                for (int u = 0; u < 3; ++u)
#pragma unroll
                    for (int v = 0; v <= u; ++v) acc[u][v] = fma(ai[u], aj[v], acc[u][v]);
            }
            // (reads touch columns b .. b+15 only, the writes columns >= b+16: no barrier between)
#pragma unroll
            for (int u = 0; u < 3; ++u)
#pragma unroll
                for (int v = 0; v <= u; ++v) {
                    const int i = ti + 16 * u, jj = tj + 16 * v;
                    if (u < nb16 && jj <= i) A[base + i][base + jj] -= acc[u][v];
                }
            __syncthreads();
            XRB_TRACE_PHASE(3)
        }
    }
    if (bad && tid == 0) *fail = 1.0;
    diag_inverse_levels(A, X, T);
    XRB_TRACE_PHASE(4)
#undef XRB_TRACE_PHASE
}

// ---- variant 2 of the diagonal block: the serial chain per 16-block is factor -> triangular solve
// of the rows below -> trailing update; the inverse of the 16-block (needed only for the final
// L^-1) runs on warp 7 next to it, off the chain.
template <int C, int C2>
__device__ __forceinline__ void trsm_upd(double (&a)[SB], const double x, double (*A)[NB + 1], const int b) {
    if constexpr (C2 < SB) {
        a[C2] = fma(-x, A[b + C2][b + C], a[C2]);  // L[C2][C]: same address on every lane (broadcast)
        trsm_upd<C, C2 + 1>(a, x, A, b);
    }
}
template <int C>
__device__ __forceinline__ void trsm_col(double (&a)[SB], double (*A)[NB + 1], const double *dinv, const int b) {
    if constexpr (C < SB) {
        const double x = a[C] * dinv[b + C];
        a[C] = x;
        trsm_upd<C, C + 1>(a, x, A, b);
        trsm_col<C + 1>(a, A, dinv, b);
    }
}
template <int R, int P>
__device__ __forceinline__ void inv2_dot(const double (&x)[SB], double (*A)[NB + 1], const int b, double &v0, double &v1) {
    if constexpr (P < R) {
        const double l = A[b + R][b + P];
        if constexpr ((P & 1) != 0)
            v1 = fma(-l, x[P], v1);
        else
            v0 = fma(-l, x[P], v0);
        inv2_dot<R, P + 1>(x, A, b, v0, v1);
    }
}
template <int R>
__device__ __forceinline__ void inv2_row(double (&x)[SB], double (*A)[NB + 1], const double *dinv, const int b, const int rl) {
    if constexpr (R < SB) {
        double v0 = rl == R ? 1.0 : 0.0, v1 = 0.0;
        inv2_dot<R, 0>(x, A, b, v0, v1);
        x[R] = (v0 + v1) * dinv[b + R];
        inv2_row<R + 1>(x, A, dinv, b, rl);
    }
}

__device__ void diag_factor2(double (*A)[NB + 1], double (*X)[NB + 1], double (*T)[SB + 1], double *dinv,
                             double *__restrict__ fail, Trace *tr = nullptr) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tj = tid & 15;
    bool bad = false;
    const bool tron = tr && tr->on;  // ph[1] factor (+ barrier), ph[2] solve, ph[3] trailing, ph[4] inverse levels
    long long tc = tron ? clock64() : 0;
#define XRB_TRACE_PHASE(k)                    \
    if (tron) {                               \
        const long long now = clock64();      \
        tr->ph[k] += now - tc, tc = now;      \
    }
    for (int b = 0; b < NB; b += SB) {
        if (warp == 0) {
            const int rl = lane & 15;
            double a[SB];
#pragma unroll
            for (int c = 0; c < SB; ++c) a[c] = A[b + rl][b + c];
            double inv_mine = 1.0;
            fac_col<0>(a, rl, inv_mine, bad);
            if (lane < SB) {
#pragma unroll
                for (int c = 0; c < SB; ++c)
                    if (c <= rl) A[b + rl][b + c] = a[c];
                dinv[b + rl] = inv_mine;
            }
        }
        __syncthreads();  // warp 7 arrives here once the previous 16-block's inverse is done
        XRB_TRACE_PHASE(1)
        const int below = NB - b - SB;
        if (warp == 7) {
            // inverse of the 16-block from shared memory: lane l owns column l of X_bb
            const int rl = lane & 15;
            double x[SB];
            inv2_row<0>(x, A, dinv, b, rl);
            if (lane < SB) {
#pragma unroll
                for (int r = 0; r < SB; ++r) X[b + r][b + rl] = x[r];  // zero above the diagonal by construction
            }
        } else if (below > 0) {
            // rows below: P = A_panel L_bb^-T by forward substitution, one thread per row, right-looking
            if (tid < below) {
                const int i = b + SB + tid;
                double a[SB];
#pragma unroll
                for (int c = 0; c < SB; ++c) a[c] = A[i][b + c];
                trsm_col<0>(a, A, dinv, b);
#pragma unroll
                for (int c = 0; c < SB; ++c) A[i][b + c] = a[c];
            }
            asm volatile("bar.sync 1, 224;" ::: "memory");
            XRB_TRACE_PHASE(2)
            // trailing update by the 224 threads of warps 0..6: rows ti + 14 u, columns tj + 16 v
            const int base = b + SB, nb16 = below / SB, ti = tid >> 4;
            double acc[4][3] = {};
#pragma unroll 4
            for (int p = 0; p < SB; ++p) {
                double ai[4], aj[3];
#pragma unroll
                for (int u = 0; u < 4; ++u) ai[u] = (ti + 14 * u) < below ? A[base + ti + 14 * u][b + p] : 0.0;
#pragma unroll
                for (int v = 0; v < 3; ++v) aj[v] = v < nb16 ? A[base + tj + 16 * v][b + p] : 0.0;
#pragma unroll
                for (int u = 0; u < 4; ++u)
#pragma unroll
                    for (int v = 0; v < 3; ++v) acc[u][v] = fma(ai[u], aj[v], acc[u][v]);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int v = 0; v < 3; ++v) {
                    const int i = ti + 14 * u, jj = tj + 16 * v;
                    if (i < below && v < nb16 && jj <= i) A[base + i][base + jj] -= acc[u][v];
                }
            asm volatile("bar.sync 1, 224;" ::: "memory");
            XRB_TRACE_PHASE(3)
        }
    }
    if (bad && tid == 0) *fail = 1.0;
    __syncthreads();  // the last 16-block's inverse (warp 7) is in X
    diag_inverse_levels(A, X, T);
    XRB_TRACE_PHASE(4)
#undef XRB_TRACE_PHASE
}

__device__ void diag_store(double (*A)[NB + 1], double (*X)[NB + 1], double *__restrict__ S, int ld, int k0, int kb,
                           double *__restrict__ linv_out) {
    for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        if (r < kb && c <= r) S[(size_t)(k0 + r) * ld + k0 + c] = A[r][c];
        linv_out[idx] = (r < kb && c < kb) ? X[r][c] : 0.0;
    }
}

__global__ void __launch_bounds__(256)
chol_diag(double *__restrict__ S, int ld, int k0, int kb, double *__restrict__ linv_out,
          double *__restrict__ fail) {
    extern __shared__ __align__(16) double smem_d[];
    double(*A)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(smem_d);                      // L when done
    double(*X)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(smem_d + NB * (NB + 1));      // L^-1
    double(*T)[SB + 1] = reinterpret_cast<double(*)[SB + 1]>(smem_d + 2 * NB * (NB + 1));  // scratch
    for (int idx = threadIdx.x; idx < NB * NB; idx += 256) {
        const int r = idx >> 6, c = idx & 63;
        A[r][c] = (r < kb && c <= r) ? S[(size_t)(k0 + r) * ld + k0 + c] : (r == c ? 1.0 : 0.0);
        X[r][c] = 0.0;
    }
    __syncthreads();
    Trace tr;
    tr.begin(threadIdx.x == 0);
    if (g_diag_variant == 2)
        diag_factor2(A, X, T, smem_d + 2 * NB * (NB + 1) + NB * (SB + 1), fail, &tr);
    else
        diag_factor(A, X, T, fail, &tr);
    diag_store(A, X, S, ld, k0, kb, linv_out);
    tr.end(0, k0);
}

// ---- FP64 GEMM tile: acc[i][j] = sum_p X[row(ty,i)][p] * Y[col(tx,j)][p], p < 64 -------------
// 256 threads as 16 x 16; thread (ty, tx) owns rows ty + 16 i and columns tx + 16 j, so that
// for a fixed p the 16 tx lanes read 16 consecutive doubles (one shared-memory wavefront) and
// the two ty values of a warp read two adjacent doubles (broadcast).  Operands are staged
// p-major: Xs[p][row], Ys[p][col].
template <int TM, int TN>
struct Tile {
    static constexpr int RM = TM / 16, RN = TN / 16;
    static constexpr int LDX = TM + 1, LDY = TN + 1;
    static constexpr int kSmemBytes = NB * (LDX + LDY) * (int)sizeof(double);

    // global row-major [row][p] -> shared [p][row]; rows >= nrows and p >= kb are zero
    template <int T>
    __device__ static void stage(double *dst, int ldd, const double *__restrict__ src, size_t ld, int nrows,
                                 int kb) {
        // 8 independent 16-byte loads in flight per thread before any shared-memory store:
        // the tile staging is latency-bound otherwise (one L2 round trip per element pair)
        constexpr int kBatch = 8;
        static_assert((T * (NB / 2)) % (256 * kBatch) == 0, "tile must be a multiple of the batch");
        for (int base = threadIdx.x; base < T * (NB / 2); base += 256 * kBatch) {
            double2 v[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int idx = base + u * 256, row = idx >> 5, p = (idx & 31) * 2;
                v[u] = make_double2(0.0, 0.0);
                if (row < nrows) {
                    const double *g = src + (size_t)row * ld + p;
                    if (p + 1 < kb)
                        v[u] = __ldg(reinterpret_cast<const double2 *>(g));
                    else if (p < kb)
                        v[u].x = __ldg(g);
                }
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
                const int idx = base + u * 256, row = idx >> 5, p = (idx & 31) * 2;
                dst[p * ldd + row] = v[u].x;
                dst[(p + 1) * ldd + row] = v[u].y;
            }
        }
    }

    __device__ static void mma(const double *Xs, const double *Ys, double (&acc)[RM][RN]) {
        const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll 2
        for (int p = 0; p < NB; ++p) {
            double a[RM], b[RN];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = Xs[p * LDX + ty + 16 * i];
#pragma unroll
            for (int j = 0; j < RN; ++j) b[j] = Ys[p * LDY + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
    }
};

// ---- 2. panel: L_ik = A_ik * Linv^T ----------------------------------------------------------
// The source block column must be 16-byte aligned for the double2 loads: ld % 2 == 0 and
// k0 % 2 == 0 hold by construction (ld is a multiple of 8, k0 a multiple of 64).
__global__ void __launch_bounds__(256)
chol_panel(double *__restrict__ S, int ld, int k0, int kb, int r0, int r1, int rhs_row,
           const double *__restrict__ linv) {
    using T = Tile<64, 64>;  // 64-row tiles: twice the CTAs of a 128-row split, half the latency
    extern __shared__ __align__(16) double smem_d[];
    Trace tr;
    tr.begin(threadIdx.x == 0 && blockIdx.x == 0);
    double *Xs = smem_d, *Ys = smem_d + NB * T::LDX;
    int row0 = r0 + blockIdx.x * 64;
    int nrows = min(64, r1 - row0);
    if (row0 >= r1) row0 = rhs_row, nrows = 1;  // extra CTA: the right-hand-side row alone
    T::stage<64>(Xs, T::LDX, S + (size_t)row0 * ld + k0, (size_t)ld, nrows, kb);
    T::stage<64>(Ys, T::LDY, linv, (size_t)NB, NB, NB);  // Ys[p][j] = Linv[j][p]
    __syncthreads();
    double acc[T::RM][T::RN] = {};
    T::mma(Xs, Ys, acc);
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
#pragma unroll
    for (int i = 0; i < T::RM; ++i)
#pragma unroll
        for (int j = 0; j < T::RN; ++j) {
            const int r = ty + 16 * i, c = tx + 16 * j;
            if (r < nrows && c < kb) S[(size_t)(row0 + r) * ld + k0 + c] = acc[i][j];
        }
    tr.end(1, k0);
}

// ---- 3. trailing update: A_ij -= L_ik L_jk^T for tiles j <= i inside the band ----------------
// Software-pipelined over the depth: the 64-deep product is cut into 4 chunks of 16; while
// chunk c is multiplied out of one shared-memory buffer, chunk c+1 travels global -> registers
// -> the other buffer, so the L2 latency of the operand fetch hides behind the FP64 FMAs.
constexpr int KC = 16;  // depth of one pipeline chunk

template <int TT>
struct UpdateCfg {
    static constexpr int LDT = TT + 1;
    static constexpr int kPerThread = TT * KC / 2 / 256;  // double2 loads per operand and thread
    static constexpr int kSmemBytes = 2 * 2 * KC * LDT * (int)sizeof(double);
};

template <int TT>
__device__ __forceinline__ void upd_fetch(const double *__restrict__ src, size_t ld, int nrows, int kb, int pc,
                                          double2 (&v)[UpdateCfg<TT>::kPerThread]) {
#pragma unroll
    for (int u = 0; u < UpdateCfg<TT>::kPerThread; ++u) {
        const int idx = threadIdx.x + u * 256, row = idx >> 3, p = pc + (idx & 7) * 2;  // 8 double2 per row
        v[u] = make_double2(0.0, 0.0);
        if (row < nrows) {
            const double *g = src + (size_t)row * ld + p;
            if (p + 1 < kb)
                v[u] = __ldg(reinterpret_cast<const double2 *>(g));
            else if (p < kb)
                v[u].x = __ldg(g);
        }
    }
}

template <int TT>
__device__ __forceinline__ void upd_store(double *dst, const double2 (&v)[UpdateCfg<TT>::kPerThread]) {
#pragma unroll
    for (int u = 0; u < UpdateCfg<TT>::kPerThread; ++u) {
        const int idx = threadIdx.x + u * 256, row = idx >> 3, p = (idx & 7) * 2;
        dst[p * UpdateCfg<TT>::LDT + row] = v[u].x;
        dst[(p + 1) * UpdateCfg<TT>::LDT + row] = v[u].y;
    }
}

// kDeep (TT == 64, the tiles of the critical first tile column): every operand chunk and the C tile
// are requested from L2 before anything is consumed — one exposed round trip instead of one per
// pipeline chunk plus one for C.  The bulk tiles keep the double-buffered pipeline (less shared
// memory and registers per CTA, more CTAs per SM).
template <int TT, bool kDeep = false>
__global__ void __launch_bounds__(256, (TT == 64 && !kDeep) ? 2 : 1)
chol_update(double *__restrict__ S, int ld, int k0, int kb, int r0, int r1, int rhs_row, int tj_lo,
            double *__restrict__ fuse_linv, double *__restrict__ fail) {
    // fuse_linv != nullptr (TT == 64 only): the CTA of tile (0, 0) goes on to factor the block it
    // has just updated — the next diagonal block — and writes its inverse to fuse_linv, which
    // takes the stand-alone chol_diag launch off the critical path.
    using C = UpdateCfg<TT>;
    constexpr int RM = TT / 16;
    extern __shared__ __align__(16) double smem_d[];
    double *Xs = smem_d;                      // [2][KC][LDT]
    double *Ys = smem_d + 2 * KC * C::LDT;    // [2][KC][LDT]
    const int ntile = (r1 - r0 + TT - 1) / TT;
    const int ti = blockIdx.y, tj = blockIdx.x + tj_lo;  // tile-column range [tj_lo, tj_lo + gridDim.x)
    Trace tr;
    tr.begin(threadIdx.x == 0 && blockIdx.x == 0 && (blockIdx.y == 0 || blockIdx.y == gridDim.y - 1));
    long long tc = tr.on ? clock64() : 0;
    int row0, nrows;
    if (ti < ntile) {
        if (tj > ti) return;
        row0 = r0 + ti * TT, nrows = min(TT, r1 - row0);
    } else {
        row0 = rhs_row, nrows = 1;  // right-hand-side row against every column tile
    }
    if (tj >= ntile) return;
    const int col0 = r0 + tj * TT, ncols = min(TT, r1 - col0);
    const double *xsrc = S + (size_t)row0 * ld + k0, *ysrc = S + (size_t)col0 * ld + k0;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

    double acc[RM][RM] = {};
    const bool diag_tile = (ti < ntile) && (ti == tj);
    double cpre[kDeep ? RM : 1][kDeep ? RM : 1];
    if constexpr (kDeep) {
        static_assert(TT == 64, "deep prefetch is sized for 64 x 64 tiles");
        constexpr int NC = NB / KC;
        double2 ax[NC][C::kPerThread], ay[NC][C::kPerThread];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            upd_fetch<TT>(xsrc, (size_t)ld, nrows, kb, c * KC, ax[c]);
            upd_fetch<TT>(ysrc, (size_t)ld, ncols, kb, c * KC, ay[c]);
        }
#pragma unroll
        for (int i = 0; i < RM; ++i)
#pragma unroll
            for (int j = 0; j < RM; ++j) {
                const int r = ty + 16 * i, cc = tx + 16 * j;
                const bool ok = r < nrows && cc < ncols && (!diag_tile || cc <= r);
                cpre[i][j] = ok ? __ldcg(S + (size_t)(row0 + r) * ld + col0 + cc) : 0.0;
            }
        double *Xd = smem_d, *Yd = smem_d + NB * C::LDT;  // [64][LDT] each, p-major
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            upd_store<TT>(Xd + c * KC * C::LDT, ax[c]);
            upd_store<TT>(Yd + c * KC * C::LDT, ay[c]);
        }
        __syncthreads();
#pragma unroll 4
        for (int p = 0; p < NB; ++p) {
            double a[RM], b[RM];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = Xd[p * C::LDT + ty + 16 * i];
#pragma unroll
            for (int j = 0; j < RM; ++j) b[j] = Yd[p * C::LDT + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RM; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        __syncthreads();  // the stage is dead: the fused path reuses it for the diagonal block
    } else {
    double2 vx[C::kPerThread], vy[C::kPerThread];
    upd_fetch<TT>(xsrc, (size_t)ld, nrows, kb, 0, vx);
    upd_fetch<TT>(ysrc, (size_t)ld, ncols, kb, 0, vy);
    upd_store<TT>(Xs, vx);
    upd_store<TT>(Ys, vy);
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < NB / KC; ++c) {
        const int cur = c & 1;
        if (c + 1 < NB / KC) {
            upd_fetch<TT>(xsrc, (size_t)ld, nrows, kb, (c + 1) * KC, vx);
            upd_fetch<TT>(ysrc, (size_t)ld, ncols, kb, (c + 1) * KC, vy);
        }
        const double *xb = Xs + cur * KC * C::LDT, *yb = Ys + cur * KC * C::LDT;
#pragma unroll 4
        for (int p = 0; p < KC; ++p) {
            double a[RM], b[RM];
#pragma unroll
            for (int i = 0; i < RM; ++i) a[i] = xb[p * C::LDT + ty + 16 * i];
#pragma unroll
            for (int j = 0; j < RM; ++j) b[j] = yb[p * C::LDT + tx + 16 * j];
#pragma unroll
            for (int i = 0; i < RM; ++i)
#pragma unroll
                for (int j = 0; j < RM; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
        }
        if (c + 1 < NB / KC) {
            upd_store<TT>(Xs + (cur ^ 1) * KC * C::LDT, vx);
            upd_store<TT>(Ys + (cur ^ 1) * KC * C::LDT, vy);
        }
        __syncthreads();
    }
    }
    // C -= acc: a whole row of the register block is loaded before anything is stored, so the
    // RM loads of a row are in flight together (a load/subtract/store chain per element
    // exposes one L2 round trip per element: that was 60 % of this kernel's stall samples).
    if (tr.on) tr.ph[0] = clock64() - tc, tc = clock64();
    const bool fuse = kDeep && fuse_linv != nullptr && ti == 0 && tj == 0;  // only the critical column fuses
    double(*FA)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(smem_d);  // aliases the operand stage
#pragma unroll
    for (int i = 0; i < RM; ++i) {
        const int r = ty + 16 * i;
        double cv[RM];
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const int cc = tx + 16 * j;
            const bool ok = r < nrows && cc < ncols && (!diag_tile || cc <= r);
            if constexpr (kDeep)
                cv[j] = cpre[i][j];
            else
                cv[j] = ok ? __ldcg(S + (size_t)(row0 + r) * ld + col0 + cc) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const int cc = tx + 16 * j;
            const bool ok = r < nrows && cc < ncols && (!diag_tile || cc <= r);
            if (kDeep && fuse) {
                if constexpr (kDeep) FA[r][cc] = ok ? cv[j] - acc[i][j] : (r == cc ? 1.0 : 0.0);
            } else if (ok) {
                S[(size_t)(row0 + r) * ld + col0 + cc] = cv[j] - acc[i][j];
            }
        }
    }
    if constexpr (kDeep) if (fuse) {  // block-uniform branch
        double(*FX)[NB + 1] = reinterpret_cast<double(*)[NB + 1]>(smem_d + NB * (NB + 1));
        double(*FT)[SB + 1] = reinterpret_cast<double(*)[SB + 1]>(smem_d + 2 * NB * (NB + 1));
        for (int idx = threadIdx.x; idx < NB * NB; idx += 256) FX[idx >> 6][idx & 63] = 0.0;
        __syncthreads();
        if (tr.on) tr.ph[5] = clock64() - tc, tc = clock64();
        if (g_diag_variant == 2)
            diag_factor2(FA, FX, FT, smem_d + 2 * NB * (NB + 1) + NB * (SB + 1), fail, &tr);
        else
            diag_factor(FA, FX, FT, fail, &tr);
        if (tr.on) tc = clock64();
        diag_store(FA, FX, S, ld, r0, nrows, fuse_linv);
        if (tr.on) tr.ph[6] = clock64() - tc;
        tr.end(4, k0);
        return;
    }
    tr.end(TT == 64 ? 3 : 2, k0);
}

// ---- 4. backward substitution L^T x = y, kBackGroup block columns per launch ------------------
// Thread layout (c, q): column c of a 64-column chunk, quarter q = one 64-row block of the
// group; every dot product over 64 rows runs as 16 independent chains so 16 loads are in
// flight per thread (the row stride makes each load a separate L2 sector stream).
//   phase 1 (every CTA, redundantly, ~16 K FMAs): solve the group's kBackGroup x 64 unknowns
//            bottom-up: y_k -= sum_{earlier blocks of the group} L_pk^T x_p ; x_k = Linv_k^T y_k
//   phase 2: remove the group's contribution from this CTA's 64 earlier unknowns:
//            y_j -= sum_g sum_r L[k0_g + r][j] x_g[r]
constexpr int kBackGroup = 4;

__device__ __forceinline__ double dot64_col(const double *__restrict__ base, size_t ld, int nrow,
                                            const double *__restrict__ x) {
    // sum_{r < nrow} base[r * ld] * x[r], 16 independent chains
    double part[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) part[u] = 0.0;
#pragma unroll
    for (int r0 = 0; r0 < NB; r0 += 16) {
        double v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) v[u] = (r0 + u < nrow) ? __ldcg(base + (size_t)(r0 + u) * ld) : 0.0;
#pragma unroll
        for (int u = 0; u < 16; ++u) part[u] = fma(v[u], x[r0 + u], part[u]);
    }
#pragma unroll
    for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int u = 0; u < w; ++u) part[u] += part[u + w];
    return part[0];
}

__global__ void __launch_bounds__(256)
chol_backsolve(const double *__restrict__ S, int ld, int n, int blk_hi, int nblk_group, int j0,
               const double *__restrict__ linv, double *__restrict__ y, double *__restrict__ x_out) {
    __shared__ double xk[kBackGroup][NB];
    __shared__ double red[4][NB];
    __shared__ double yt[NB];
    const int tid = threadIdx.x, c = tid & 63, q = tid >> 6;
    Trace tr;
    tr.begin(tid == 0 && blockIdx.x == 0);
    for (int g = 0; g < nblk_group; ++g) {
        const int k0 = (blk_hi - g) * NB, kb = min(NB, n - k0);
        // correction from the blocks of the group already solved: quarter q handles block gp = q
        double corr = 0.0;
        if (q < g && c < kb) {
            const int kp0 = (blk_hi - q) * NB, kpb = min(NB, n - kp0);
            corr = dot64_col(S + (size_t)kp0 * ld + k0 + c, (size_t)ld, kpb, xk[q]);
        }
        red[q][c] = corr;
        __syncthreads();
        if (tid < NB) yt[tid] = tid < kb ? y[k0 + tid] - (red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid]) : 0.0;
        __syncthreads();
        {
            // x_k[c] = sum_{r >= c} Linv[r][c] yt[r]; quarter q takes rows r = q, q+4, ...
            const double *li = linv + (size_t)(blk_hi - g) * NB * NB;
            double v = 0.0;
            if (c < kb)
                for (int r = c + q; r < kb; r += 4) v = fma(__ldg(li + r * NB + c), yt[r], v);
            red[q][c] = v;
        }
        __syncthreads();
        if (tid < NB) {
            const double v = tid < kb ? red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid] : 0.0;
            xk[g][tid] = v;
            if (blockIdx.x == 0 && tid < kb) x_out[k0 + tid] = v;
        }
        __syncthreads();
    }
    const int k_low = (blk_hi - nblk_group + 1) * NB;
    const int j = j0 + blockIdx.x * NB + c;
    double part = 0.0;
    if (j < k_low && q < nblk_group) {
        const int k0 = (blk_hi - q) * NB, kb = min(NB, n - k0);
        part = dot64_col(S + (size_t)k0 * ld + j, (size_t)ld, kb, xk[q]);
    }
    red[q][c] = part;
    __syncthreads();
    if (tid < NB && j < k_low) y[j] -= red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
    tr.end(5, blk_hi);
}

namespace {

// Side stream + events for the look-ahead: the bulk of the trailing update of step k runs on
// `side` while the main stream already factors the next diagonal block and forms the next
// panel, which only need the FIRST tile column of that update.
struct GraphEntry {
    cudaGraphExec_t exec = nullptr;
    int64_t launches = 0;
};
using GraphKey = std::tuple<double *, int, int, int, double *, double *, double *>;

// Streams, events, function attributes and instantiated graphs belong to one device: the state is
// kept per device id so that a process driving several GPUs (one solver each) stays correct.
struct DevCtx {
    cudaStream_t side = nullptr, side2 = nullptr;
    std::vector<cudaEvent_t> events;
    std::map<GraphKey, GraphEntry> graphs;
    bool attr_set = false;
    cudaEvent_t get_event(size_t i) {
        while (events.size() <= i) {
            cudaEvent_t e;
            cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
            events.push_back(e);
        }
        return events[i];
    }
};
constexpr int kMaxDevices = 64;
DevCtx g_ctx[kMaxDevices];
std::mutex g_ctx_mutex;  // enqueueing is host work; one solve is enqueued at a time

template <int TT, bool kDeep = false>
constexpr int update_smem(bool fused) {
    const int stage = kDeep ? 2 * NB * UpdateCfg<TT>::LDT * (int)sizeof(double) : UpdateCfg<TT>::kSmemBytes;
    return fused && kDiagSmem > stage ? kDiagSmem : stage;
}

template <int TT, bool kDeep = false>
void launch_update(double *S, int ld, int k0, int kb, int r0, int r1, int n, int tj_lo, int tj_hi,
                   double *fuse_linv, double *fail, cudaStream_t st) {
    const int nt = (r1 - r0 + TT - 1) / TT;
    if (tj_hi <= tj_lo) return;
    dim3 grid(tj_hi - tj_lo, nt + 1);
    chol_update<TT, kDeep><<<grid, 256, update_smem<TT, kDeep>(fuse_linv != nullptr), st>>>(S, ld, k0, kb, r0, r1, n, tj_lo,
                                                                                         fuse_linv, fail);
}

int enqueue_all(DevCtx &ctx, double *S, int n, int ld, int bw, double *linv, double *x_out, double *fail_flag,
                cudaStream_t st, int64_t *count) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (!ctx.side) {
        int least = 0, greatest = 0;
        cudaDeviceGetStreamPriorityRange(&least, &greatest);
        cudaStreamCreateWithPriority(&ctx.side, cudaStreamNonBlocking, least);
        cudaStreamCreateWithPriority(&ctx.side2, cudaStreamNonBlocking, least);
    }
    cudaStream_t g_side = ctx.side, g_side2 = ctx.side2;
    auto get_event = [&](size_t i) { return ctx.get_event(i); };
    const int nblk = (n + NB - 1) / NB;
    int64_t nl = 0;
    size_t ev = 0;
    // Look-ahead of depth 2.  Step k (block column k factored) updates the trailing matrix in three
    // pieces: `crit` = block column k+1 (main stream, its (0, 0) CTA factors the next diagonal
    // block), `second` = block column k+2 (side stream 2) and `rest` = everything right of it
    // (side stream 1, the bulk of the flops).  crit(k) only needs second(k-1); second(k) needs
    // rest(k-1): the big update has two chain steps to finish before anybody waits for it.
    cudaEvent_t second_done = nullptr, rest_done = nullptr;
    chol_diag<<<1, 256, kDiagSmem, st>>>(S, ld, 0, min(NB, n), linv, fail_flag);  // block 0 only
    ++nl;
    for (int kblk = 0; kblk < nblk; ++kblk) {
        const int k0 = kblk * NB, kb = min(NB, n - k0);
        double *li = linv + (size_t)kblk * NB * NB;
        const int r0 = k0 + kb;
        const int r1 = min(n, r0 + bw);  // rows that can be non-zero in this block column
        const int np = (r1 - r0 + 63) / 64;
        chol_panel<<<np + 1, 256, Tile<64, 64>::kSmemBytes, st>>>(S, ld, k0, kb, r0, r1, n, li);
        ++nl;
        cudaEvent_t panel_done = get_event(ev++);
        cudaEventRecord(panel_done, st);
        if (second_done) cudaStreamWaitEvent(st, second_done, 0);
        // a band narrower than two blocks: `rest` of the previous step is empty or tiny, but the
        // main stream must still see it before the back substitution; order it here
        if (r1 > r0) {
            // fusing needs the whole next diagonal block inside this update's row range (a band
            // narrower than one block leaves rows the update never loads)
            const int kb_next = min(NB, n - r0);
            const bool fuse_ok = kblk + 1 < nblk && (r1 - r0) >= kb_next;
            double *li_next = linv + (size_t)(kblk + 1) * NB * NB;
            launch_update<64, true>(S, ld, k0, kb, r0, r1, n, 0, 1, fuse_ok ? li_next : nullptr, fail_flag, st);
            ++nl;
            if (!fuse_ok && kblk + 1 < nblk) {
                chol_diag<<<1, 256, kDiagSmem, st>>>(S, ld, r0, kb_next, li_next, fail_flag);
                ++nl;
            }
        }
        // second: block column k+2 = tile column 1 of the 64-wide tiling anchored at r0
        cudaEvent_t prev_rest = rest_done;
        second_done = nullptr;
        if (r1 > r0 + NB) {
            cudaStreamWaitEvent(g_side2, panel_done, 0);
            if (prev_rest) cudaStreamWaitEvent(g_side2, prev_rest, 0);
            launch_update<64>(S, ld, k0, kb, r0, r1, n, 1, 2, nullptr, fail_flag, g_side2);
            ++nl;
            second_done = get_event(ev++);
            cudaEventRecord(second_done, g_side2);
        } else if (prev_rest) {
            cudaStreamWaitEvent(st, prev_rest, 0);  // nothing else will order it before the main stream
        }
        // rest: everything from column r0 + 128 on
        rest_done = nullptr;
        const int b0 = r0 + 2 * NB;
        if (r1 > b0) {
            cudaStreamWaitEvent(g_side, panel_done, 0);
            const int nt128 = (r1 - b0 + 127) / 128;
            if (nt128 * (nt128 + 1) / 2 >= sms)
                launch_update<128>(S, ld, k0, kb, b0, r1, n, 0, nt128, nullptr, fail_flag, g_side);
            else
                launch_update<64>(S, ld, k0, kb, b0, r1, n, 0, (r1 - b0 + 63) / 64, nullptr, fail_flag, g_side);
            ++nl;
            rest_done = get_event(ev++);
            cudaEventRecord(rest_done, g_side);
        }
    }
    if (second_done) cudaStreamWaitEvent(st, second_done, 0);
    if (rest_done) cudaStreamWaitEvent(st, rest_done, 0);
    double *y = S + (size_t)n * ld;  // y = L^-1 rhs now sits in row n
    for (int hi = nblk - 1; hi >= 0; hi -= kBackGroup) {
        const int ng = min(kBackGroup, hi + 1);
        const int k_low = (hi - ng + 1) * NB;
        const int j0 = max(0, k_low - bw - NB);
        const int ncols = k_low - j0;
        const int grid = ncols > 0 ? (ncols + NB - 1) / NB : 1;
        chol_backsolve<<<grid, 256, 0, st>>>(S, ld, n, hi, ng, j0, linv, y, x_out);
        ++nl;
    }
    *count = nl;
    return cudaGetLastError() == cudaSuccess ? XRB_OK : XRB_ERR_CUDA;
}

}  // namespace

int ba_launch_cholesky_solve(double *S, int n, int ld, int bw, double *linv, double *x_out,
                             double *fail_flag, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return XRB_OK;
    int dev_id = 0;
    XRB_CUDA(cudaGetDevice(&dev_id));
    if (dev_id < 0 || dev_id >= kMaxDevices) {
        set_error("cholesky: device id %d outside [0, %d)", dev_id, kMaxDevices);
        return XRB_ERR_INVALID;
    }
    std::lock_guard<std::mutex> lock(g_ctx_mutex);
    DevCtx &ctx = g_ctx[dev_id];
    auto &g_graphs = ctx.graphs;
    if (!ctx.attr_set) {
        XRB_CUDA(cudaFuncSetAttribute(chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kDiagSmem));
        XRB_CUDA(cudaFuncSetAttribute(chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Tile<64, 64>::kSmemBytes));
        XRB_CUDA(cudaFuncSetAttribute(chol_update<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      update_smem<128>(false)));
        XRB_CUDA(cudaFuncSetAttribute(chol_update<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      update_smem<64>(true)));
        XRB_CUDA(cudaFuncSetAttribute(chol_update<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      update_smem<64, true>(true)));
        const char *dv = getenv("XRB_CHOL_DIAG");
        const int variant = (dv && atoi(dv) == 1) ? 1 : 2;
        XRB_CUDA(cudaMemcpyToSymbol(g_diag_variant, &variant, sizeof(variant)));
        ctx.attr_set = true;
    }
    // The launch sequence depends only on (n, ld, bw) and the buffers: replay it as a graph.
    const GraphKey key{S, n, ld, bw, linv, x_out, fail_flag};
    auto it = g_graphs.find(key);
    if (it == g_graphs.end()) {
        GraphEntry e;
        cudaGraph_t graph = nullptr;
        cudaStream_t cs = nullptr;
        int pr_least = 0, pr_greatest = 0;  // the chain's kernels go first whenever an SM frees up
        cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest);
        bool ok = cudaStreamCreateWithPriority(&cs, cudaStreamNonBlocking, pr_greatest) == cudaSuccess &&
                  cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            const int rc = enqueue_all(ctx, S, n, ld, bw, linv, x_out, fail_flag, cs, &e.launches);
            ok = cudaStreamEndCapture(cs, &graph) == cudaSuccess && rc == XRB_OK && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&e.exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (cs) cudaStreamDestroy(cs);
        if (!ok) {
            cudaGetLastError();
            e.exec = nullptr;
        }
        if (g_graphs.size() > 64) {  // bounded cache
            for (auto &kv : g_graphs)
                if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
            g_graphs.clear();
        }
        it = g_graphs.emplace(key, e).first;
    }
    int64_t nl = 0;
    if (it->second.exec) {
        XRB_CUDA(cudaGraphLaunch(it->second.exec, st));
        nl = it->second.launches;
    } else {
        const int rc = enqueue_all(ctx, S, n, ld, bw, linv, x_out, fail_flag, st, &nl);
        if (rc) {
            set_error("cholesky launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return rc;
        }
    }
    g_launches.fetch_add((uint64_t)nl, std::memory_order_relaxed);
    if (launches) *launches += nl;
    return XRB_OK;
}

int ba_chol_trace(int enable, long long *out, int cap_records) {
    int n = 0;
    if (enable) {
        const int one = 1;
        const unsigned zero = 0;
        XRB_CUDA(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(zero)));
        XRB_CUDA(cudaMemcpyToSymbol(g_trace_on, &one, sizeof(one)));
        return 0;
    }
    const int off = 0;
    unsigned cnt = 0;
    XRB_CUDA(cudaDeviceSynchronize());
    XRB_CUDA(cudaMemcpyToSymbol(g_trace_on, &off, sizeof(off)));
    XRB_CUDA(cudaMemcpyFromSymbol(&cnt, g_trace_n, sizeof(cnt)));
    n = (int)std::min<unsigned>(cnt, (unsigned)std::min(cap_records, kTraceCap));
    if (n && out) XRB_CUDA(cudaMemcpyFromSymbol(out, g_trace, (size_t)n * kTraceWords * sizeof(long long)));
    return n;
}

}  // namespace xrb
