// ba_chol.cu — reduced-camera-system solver of path B: blocked FP64 Cholesky (lower, in
// place) limited to a half-bandwidth, forward substitution folded into the factorisation
// (the right-hand side rides along as row n of the matrix), blocked backward substitution.
//
// Replaces Ceres' SparseSchurComplementSolver factor+solve (selected by ba_solver.cc:74).
// Generation 1: 64x64 FP64 FMA tiles staged in shared memory; see DESIGN.md §B.4 for the
// roofline (FP64 pipe) and the planned DMMA variant.
#include <cuda_runtime.h>

#include "ba_kernels.cuh"

namespace xrb {

constexpr int NB = 64;        // block size
constexpr int TLD = NB + 2;   // shared tile leading dimension (doubles)

// ---- 1. diagonal block: L_kk = chol(A_kk) in place, Linv = L_kk^-1 -------------------------
__global__ void __launch_bounds__(256)
chol_diag(double *__restrict__ S, int ld, int k0, int kb, double *__restrict__ linv_out,
          double *__restrict__ fail) {
    extern __shared__ __align__(16) double smem_d[];
    double(*A)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d);
    double(*Li)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d + NB * TLD);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx / NB, c = idx % NB;
        A[r][c] = (r < kb && c <= r) ? S[(size_t)(k0 + r) * ld + k0 + c] : (r == c ? 1.0 : 0.0);
        Li[r][c] = 0.0;
    }
    __syncthreads();
    for (int j = 0; j < kb; ++j) {
        if (tid == 0) {
            double d = A[j][j];
            if (!(d > 0.0) || !isfinite(d)) {
                *fail = 1.0;
                d = 1.0;
            }
            A[j][j] = sqrt(d);
        }
        __syncthreads();
        const double inv = 1.0 / A[j][j];
        for (int i = j + 1 + tid; i < kb; i += 256) A[i][j] *= inv;
        __syncthreads();
        // rank-1 update of the trailing lower triangle
        const int m = kb - j - 1;
        for (int idx = tid; idx < m * m; idx += 256) {
            const int r = j + 1 + idx / m, c = j + 1 + idx % m;
            if (c <= r) A[r][c] -= A[r][j] * A[c][j];
        }
        __syncthreads();
    }
    // inverse of the lower-triangular block: column c of Linv by forward substitution
    if (tid < kb) {
        const int c = tid;
        for (int r = c; r < kb; ++r) {
            double v = (r == c) ? 1.0 : 0.0;
            for (int p = c; p < r; ++p) v -= A[r][p] * Li[p][c];
            Li[r][c] = v / A[r][r];
        }
    }
    __syncthreads();
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int r = idx / NB, c = idx % NB;
        if (r < kb && c <= r) S[(size_t)(k0 + r) * ld + k0 + c] = A[r][c];
        linv_out[idx] = (r < kb && c < kb) ? Li[r][c] : 0.0;
    }
}

// 64x64x64 FP64 tile product helper: acc[4][4] += sum_p At[p][row] * Bt[p][col]
// (both operands stored p-major so a thread's 4 rows / 4 cols are contiguous).
__device__ __forceinline__ void tile_mma(double (*At)[TLD], double (*Bt)[TLD], int ty,
                                         int tx, double acc[4][4], int kb) {
#pragma unroll 4
    for (int p = 0; p < kb; ++p) {
        double a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = At[p][ty * 4 + i], b[i] = Bt[p][tx * 4 + i];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
}

// ---- 2. panel: L_ik = A_ik * Linv^T for the rows below the diagonal block -------------------
// row_list semantics: tile t covers rows r0 + 64 t ... ; the last tile of the launch is the
// single right-hand-side row n when it lies outside the band.
__global__ void __launch_bounds__(256)
chol_panel(double *__restrict__ S, int ld, int k0, int kb, int r0, int r1, int rhs_row,
           const double *__restrict__ linv) {
    extern __shared__ __align__(16) double smem_d[];
    double(*At)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d);             // At[p][i] = A[row i][k0+p]
    double(*Bt)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d + NB * TLD);  // Bt[p][j] = Linv[j][p]
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    int row0 = r0 + blockIdx.x * NB;
    int nrows = min(NB, r1 - row0);
    if (row0 >= r1) {  // extra block: the rhs row alone
        row0 = rhs_row;
        nrows = 1;
    }
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx / NB, p = idx % NB;
        At[p][i] = (i < nrows && p < kb) ? S[(size_t)(row0 + i) * ld + k0 + p] : 0.0;
        Bt[p][i] = linv[i * NB + p];  // Linv[j=i][p]
    }
    __syncthreads();
    double acc[4][4] = {};
    tile_mma(At, Bt, ty, tx, acc, kb);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ty * 4 + i, c = tx * 4 + j;
            if (r < nrows && c < kb) S[(size_t)(row0 + r) * ld + k0 + c] = acc[i][j];
        }
}

// ---- 3. trailing update: A_ij -= L_ik L_jk^T for tiles j <= i inside the band ---------------
__global__ void __launch_bounds__(256)
chol_update(double *__restrict__ S, int ld, int k0, int kb, int r0, int r1, int rhs_row) {
    extern __shared__ __align__(16) double smem_d[];
    double(*At)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d);
    double(*Bt)[TLD] = reinterpret_cast<double(*)[TLD]>(smem_d + NB * TLD);
    const int ntile = (r1 - r0 + NB - 1) / NB;
    int ti = blockIdx.y, tj = blockIdx.x;
    int row0, nrows;
    if (ti < ntile) {
        if (tj > ti) return;
        row0 = r0 + ti * NB;
        nrows = min(NB, r1 - row0);
    } else {  // rhs row against every column tile
        row0 = rhs_row;
        nrows = 1;
    }
    if (tj >= ntile) return;
    const int col0 = r0 + tj * NB, ncols = min(NB, r1 - col0);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    for (int idx = tid; idx < NB * NB; idx += 256) {
        const int i = idx / NB, p = idx % NB;
        At[p][i] = (i < nrows && p < kb) ? S[(size_t)(row0 + i) * ld + k0 + p] : 0.0;
        Bt[p][i] = (i < ncols && p < kb) ? S[(size_t)(col0 + i) * ld + k0 + p] : 0.0;
    }
    __syncthreads();
    double acc[4][4] = {};
    tile_mma(At, Bt, ty, tx, acc, kb);
    const bool diag_tile = (ti < ntile) && (ti == tj);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = ty * 4 + i, c = tx * 4 + j;
            if (r < nrows && c < ncols && (!diag_tile || c <= r))
                S[(size_t)(row0 + r) * ld + col0 + c] -= acc[i][j];
        }
}

// ---- 4. backward substitution L^T x = y, one block column per launch ------------------------
// Every CTA recomputes x_k = Linv_k^T y_k (64x64, trivial) and then removes x_k's
// contribution from its slice of the earlier unknowns: y_j -= sum_r L[k0+r][j] x_k[r].
__global__ void __launch_bounds__(256)
chol_backsolve(const double *__restrict__ S, int ld, int k0, int kb, int j0,
               const double *__restrict__ linv, double *__restrict__ y, double *__restrict__ x_out) {
    __shared__ double xk[NB];
    const int tid = threadIdx.x;
    if (tid < NB) {
        double v = 0.0;
        if (tid < kb)
            for (int r = tid; r < kb; ++r) v += linv[r * NB + tid] * y[k0 + r];  // Linv^T
        xk[tid] = v;
    }
    __syncthreads();
    if (blockIdx.x == 0 && tid < kb) x_out[k0 + tid] = xk[tid];
    const int j = j0 + blockIdx.x * 256 + tid;
    if (j < k0) {
        double acc = 0.0;
        for (int r = 0; r < kb; ++r) acc += S[(size_t)(k0 + r) * ld + j] * xk[r];
        y[j] -= acc;
    }
}

int ba_launch_cholesky_solve(double *S, int n, int ld, int bw, double *linv, double *x_out,
                             double *fail_flag, cudaStream_t st, int64_t *launches) {
    if (n <= 0) return XRB_OK;
    constexpr int kSmem = 2 * NB * TLD * (int)sizeof(double);
    static bool attr_set = false;
    if (!attr_set) {
        XRB_CUDA(cudaFuncSetAttribute(chol_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        XRB_CUDA(cudaFuncSetAttribute(chol_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        XRB_CUDA(cudaFuncSetAttribute(chol_update, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
        attr_set = true;
    }
    const int nblk = (n + NB - 1) / NB;
    int64_t nl = 0;
    for (int kblk = 0; kblk < nblk; ++kblk) {
        const int k0 = kblk * NB, kb = min(NB, n - k0);
        double *li = linv + (size_t)kblk * NB * NB;
        chol_diag<<<1, 256, kSmem, st>>>(S, ld, k0, kb, li, fail_flag);
        ++nl;
        const int r0 = k0 + kb;
        const int r1 = min(n, r0 + bw);  // rows that can be non-zero in this block column
        const int ntile = (r1 - r0 + NB - 1) / NB;
        // the rhs row n always participates (it is dense)
        chol_panel<<<ntile + 1, 256, kSmem, st>>>(S, ld, k0, kb, r0, r1, n, li);
        ++nl;
        dim3 grid(ntile > 0 ? ntile : 1, ntile + 1);
        chol_update<<<grid, 256, kSmem, st>>>(S, ld, k0, kb, r0, r1, n);
        ++nl;
    }
    // y = row n of S; solve L^T x = y block by block from the bottom
    double *y = S + (size_t)n * ld;
    for (int kblk = nblk - 1; kblk >= 0; --kblk) {
        const int k0 = kblk * NB, kb = min(NB, n - k0);
        const int j0 = max(0, k0 - bw - NB);
        const int ncols = k0 - j0;
        const int grid = ncols > 0 ? (ncols + 255) / 256 : 1;
        chol_backsolve<<<grid, 256, 0, st>>>(S, ld, k0, kb, j0, linv + (size_t)kblk * NB * NB, y, x_out);
        ++nl;
    }
    g_launches.fetch_add((uint64_t)nl, std::memory_order_relaxed);
    if (launches) *launches += nl;
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

}  // namespace xrb
