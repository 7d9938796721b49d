// match_kernels.cu — generation-1 score kernel (dp4a tiles) and the shared finalize /
// pack / helper kernels of the fused SIFT matcher.  See match_kernels.cuh for the
// contract and the reference lines each piece restates.
#include <cuda_runtime.h>

#include "match_kernels.cuh"

namespace xrb {

// ---------------------------------------------------------------------------
// Threshold arithmetic — identical expressions to ProgramCU.cu:1830-1835,1865-1870:
// int -> float multiply by 2^-18, min(float,double) promotes, DOUBLE acos, float compare.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float dist_of_dot(int dot) {
    return (float)acos(min(dot * 0.000003814697265625f, 1.0));
}
__device__ __forceinline__ bool accept(int best, int second, float distmax, float ratiomax) {
    float dist = dist_of_dot(best);
    float distn = dist_of_dot(second);
    return (dist < distmax) && (dist < distn * ratiomax);
}

// v_low: the largest dot that can influence no decision, i.e. for every v <= v_low
//   (a) v fails the distance test as a best match, and
//   (b) v as a runner-up lets every best that passes the distance test pass the ratio
//       test too (dist(v)*ratiomax >= distmax > dist(best)).
// Entries <= v_low are skipped by the score kernels; the result is unchanged.
__global__ void vlow_kernel(float distmax, float ratiomax, int *vlow) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int result = -1;
    if (distmax > 0.f && ratiomax > 0.f) {  // otherwise: no filtering (NaN lands here too)
        int lo = 0, hi = 262144;             // dist(v) == 0 for v >= 2^18
        // predicate is monotone non-increasing in v
        auto pred = [&](int v) {
            float d = dist_of_dot(v);
            return (d >= distmax) && (d * ratiomax >= distmax);
        };
        if (pred(0)) {
            while (lo < hi) {  // largest v with pred(v)
                int mid = (lo + hi + 1) >> 1;
                if (pred(mid))
                    lo = mid;
                else
                    hi = mid - 1;
            }
            result = lo;
        }
    }
    *vlow = result;
}

int launch_vlow(float distmax, float ratiomax, int *vlow_dev, cudaStream_t st) {
    vlow_kernel<<<1, 32, 0, st>>>(distmax, ratiomax, vlow_dev);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

__global__ void dist_table_kernel(float *out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = dist_of_dot(i);
}

int launch_dist_table(float *out_dev, int n, cudaStream_t st) {
    dist_table_kernel<<<(n + 255) / 256, 256, 0, st>>>(out_dev, n);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// ---------------------------------------------------------------------------
// Candidate push: lock-free top-2 per row and per column.
// atomicMax on the packed key returns the previous best; whichever of (old, new) lost is
// offered to `second`.  Every key except the final maximum loses exactly once against a
// larger key, so `second` ends as the largest non-maximal dot (ties count) — the same
// value the reference's "t_dotnxt = test ? t_dotmax : max(t_dotnxt, v)" produces.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void push_top2(unsigned long long *best, unsigned int *second,
                                          unsigned long long key) {
    unsigned long long old = atomicMax(best, key);
    unsigned long long loser = old < key ? old : key;
    unsigned int lv = (unsigned int)(loser >> 32);
    if (lv) atomicMax(second, lv);
}

__device__ __forceinline__ void push_candidate(const Top2State &rows, const Top2State &cols,
                                               size_t base, int i, int j, int v) {
    unsigned long long hv = (unsigned long long)(unsigned int)v << 32;
    push_top2(rows.best + base + i, rows.second + base + i,
              hv | (0xFFFFFFFFu - row_tie_rank((uint32_t)j)));
    push_top2(cols.best + base + j, cols.second + base + j, hv | (0xFFFFFFFFu - (uint32_t)i));
}

// ---------------------------------------------------------------------------
// Generation 1 score kernel: one CTA = one 128x128 tile of dot(A_i, B_j), K = 128 bytes.
// 256 threads, each an 8x8 register block; operands staged once in shared memory as
// 32-bit words (row stride 36 words -> conflict-free 128-bit reads), u8 x u8 -> u32 dp4a.
// ---------------------------------------------------------------------------
constexpr int kTile = 128;
constexpr int kWordStride = 36;  // 32 data words + 4 pad, keeps uint4 alignment

__global__ void __launch_bounds__(256)
score_dp4a_kernel(const PairDesc *__restrict__ pairs, int state_stride, Top2State rows,
                  Top2State cols, const int *__restrict__ vlow_ptr) {
    __shared__ __align__(16) uint32_t As[kTile * kWordStride];
    __shared__ __align__(16) uint32_t Bs[kTile * kWordStride];

    const PairDesc pd = pairs[blockIdx.z];
    const int row0 = blockIdx.y * kTile, col0 = blockIdx.x * kTile;
    if (row0 >= pd.n1 || col0 >= pd.n2) return;
    const int tid = threadIdx.x;

    // Stage tiles: 128 rows x 8 uint4 per operand; rows past the end are zero.
    for (int idx = tid; idx < kTile * 8; idx += 256) {
        int r = idx >> 3, q = idx & 7;
        uint4 va = make_uint4(0, 0, 0, 0), vb = make_uint4(0, 0, 0, 0);
        if (row0 + r < pd.n1)
            va = __ldg(reinterpret_cast<const uint4 *>(pd.a + (size_t)(row0 + r) * kDim) + q);
        if (col0 + r < pd.n2)
            vb = __ldg(reinterpret_cast<const uint4 *>(pd.b + (size_t)(col0 + r) * kDim) + q);
        *reinterpret_cast<uint4 *>(&As[r * kWordStride + q * 4]) = va;
        *reinterpret_cast<uint4 *>(&Bs[r * kWordStride + q * 4]) = vb;
    }
    __syncthreads();

    const int tx = tid & 15, ty = tid >> 4;
    unsigned int acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0;

#pragma unroll 2
    for (int q = 0; q < 8; ++q) {
        uint4 a[8], b[8];
#pragma unroll
        for (int r = 0; r < 8; ++r)
            a[r] = *reinterpret_cast<const uint4 *>(&As[(ty + 16 * r) * kWordStride + q * 4]);
#pragma unroll
        for (int c = 0; c < 8; ++c)
            b[c] = *reinterpret_cast<const uint4 *>(&Bs[(tx + 16 * c) * kWordStride + q * 4]);
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                unsigned int s = acc[r][c];
                s = __dp4a(a[r].x, b[c].x, s);
                s = __dp4a(a[r].y, b[c].y, s);
                s = __dp4a(a[r].z, b[c].z, s);
                s = __dp4a(a[r].w, b[c].w, s);
                acc[r][c] = s;
            }
    }

    const int vlow = *vlow_ptr;
    const size_t base = (size_t)blockIdx.z * state_stride;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int i = row0 + ty + 16 * r;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int j = col0 + tx + 16 * c;
            const int v = (int)acc[r][c];
            if (v > vlow && i < pd.n1 && j < pd.n2) push_candidate(rows, cols, base, i, j, v);
        }
    }
}

int launch_score_dp4a(const PairDesc *pairs_dev, int n_pairs, int max_n1, int max_n2,
                      int state_stride, Top2State rows, Top2State cols, const int *vlow_dev,
                      cudaStream_t st) {
    if (n_pairs <= 0 || max_n1 <= 0 || max_n2 <= 0) return XRB_OK;
    dim3 grid((max_n2 + kTile - 1) / kTile, (max_n1 + kTile - 1) / kTile, n_pairs);
    score_dp4a_kernel<<<grid, 256, 0, st>>>(pairs_dev, state_stride, rows, cols, vlow_dev);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// ---------------------------------------------------------------------------
// Finalize: thresholds + mutual test + ordered compaction, one CTA per pair
// (RowMatch/ColMatch threshold lines + SiftMatchCU.cpp:199-207), then wipe the state so
// the next batch starts clean.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
finalize_kernel(const PairDesc *__restrict__ pairs, int state_stride, Top2State rows,
                Top2State cols, float distmax, float ratiomax, int mbm, int max_match,
                int32_t *__restrict__ counts, uint32_t (*__restrict__ out)[2], int out_stride) {
    __shared__ int warp_sum[32];
    __shared__ int running;
    const PairDesc pd = pairs[blockIdx.x];
    const size_t base = (size_t)blockIdx.x * state_stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t(*dst)[2] = out + (size_t)blockIdx.x * out_stride;
    if (tid == 0) running = 0;
    __syncthreads();

    for (int chunk = 0; chunk < pd.n1; chunk += 1024) {
        const int i = chunk + tid;
        int flag = 0, j = -1;
        if (i < pd.n1) {
            unsigned long long key = rows.best[base + i];
            int v = (int)(key >> 32);
            // v == 0 can only be stored when nothing is filtered; the reference never
            // promotes a zero dot to "best" (strict '>' against the initial 0).
            if (v > 0 && accept(v, (int)rows.second[base + i], distmax, ratiomax)) {
                j = (int)row_tie_unrank(0xFFFFFFFFu - (uint32_t)key);
                if (mbm) {
                    unsigned long long ck = cols.best[base + j];
                    int cv = (int)(ck >> 32);
                    flag = cv > 0 && (int)(0xFFFFFFFFu - (uint32_t)ck) == i &&
                           accept(cv, (int)cols.second[base + j], distmax, ratiomax);
                } else {
                    flag = 1;
                }
            }
        }
        // block-wide exclusive scan of flag
        unsigned ballot = __ballot_sync(0xFFFFFFFFu, flag);
        int excl = __popc(ballot & ((1u << lane) - 1));
        if (lane == 0) warp_sum[warp] = __popc(ballot);
        __syncthreads();
        if (warp == 0) {
            int s = warp_sum[lane];
            int inc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
                if (lane >= d) inc += t;
            }
            warp_sum[lane] = inc - s;  // exclusive
        }
        __syncthreads();
        const int pos = running + warp_sum[warp] + excl;
        if (flag && pos < max_match) {
            dst[pos][0] = (uint32_t)i;
            dst[pos][1] = (uint32_t)j;
        }
        __syncthreads();
        if (tid == 1023) running = pos + flag;  // last thread holds the inclusive total
        __syncthreads();
    }
    if (tid == 0) counts[blockIdx.x] = running < max_match ? running : max_match;
    __syncthreads();
    for (int i = tid; i < pd.n1; i += 1024) {
        rows.best[base + i] = 0ull;
        rows.second[base + i] = 0u;
    }
    for (int j = tid; j < pd.n2; j += 1024) {
        cols.best[base + j] = 0ull;
        cols.second[base + j] = 0u;
    }
}

int launch_finalize(const PairDesc *pairs_dev, int n_pairs, int state_stride, Top2State rows,
                    Top2State cols, float distmax, float ratiomax, int mbm, int max_match,
                    int32_t *counts_dev, uint32_t (*out_dev)[2], int out_stride,
                    cudaStream_t st) {
    if (n_pairs <= 0) return XRB_OK;
    finalize_kernel<<<n_pairs, 1024, 0, st>>>(pairs_dev, state_stride, rows, cols, distmax,
                                              ratiomax, mbm, max_match, counts_dev, out_dev,
                                              out_stride);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// ---------------------------------------------------------------------------
// PairDesc table from an index list (device-resident API).
// ---------------------------------------------------------------------------
__global__ void build_pairs_kernel(const int32_t (*__restrict__ idx)[2], int n_pairs, int n_images,
                                   const int64_t *__restrict__ row_offsets,
                                   const uint8_t *__restrict__ block, int max_features,
                                   PairDesc *__restrict__ out) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    int a = idx[p][0], b = idx[p][1];
    PairDesc pd;
    if ((unsigned)a >= (unsigned)n_images || (unsigned)b >= (unsigned)n_images) {
        // an index outside the resident set: an empty pair (0 matches), never an out-of-bounds read
        pd.a = pd.b = block, pd.n1 = pd.n2 = 0;
        out[p] = pd;
        return;
    }
    pd.a = block + row_offsets[a] * kDim;
    pd.b = block + row_offsets[b] * kDim;
    int n1 = (int)(row_offsets[a + 1] - row_offsets[a]);
    int n2 = (int)(row_offsets[b + 1] - row_offsets[b]);
    pd.n1 = n1 < max_features ? n1 : max_features;  // SiftMatchCU.cpp:113-114
    pd.n2 = n2 < max_features ? n2 : max_features;
    out[p] = pd;
}

int launch_build_pairs(const int32_t (*pairs_idx_dev)[2], int n_pairs, int n_images,
                       const int64_t *row_offsets_dev, const uint8_t *block_dev,
                       int max_features, PairDesc *out_dev, cudaStream_t st) {
    if (n_pairs <= 0) return XRB_OK;
    build_pairs_kernel<<<(n_pairs + 255) / 256, 256, 0, st>>>(
        pairs_idx_dev, n_pairs, n_images, row_offsets_dev, block_dev, max_features, out_dev);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

// ---------------------------------------------------------------------------
// Pack strided per-pair lists into one contiguous list + prefix offsets.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scan_counts_kernel(const int32_t *__restrict__ counts, int n, int64_t base,
                   int64_t *__restrict__ offsets) {
    __shared__ long long warp_sum[32];
    __shared__ long long running;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) running = base;
    __syncthreads();
    for (int chunk = 0; chunk < n; chunk += 1024) {
        int i = chunk + tid;
        long long v = i < n ? counts[i] : 0, inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            long long t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sum[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            long long s = warp_sum[lane], winc = s;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                long long t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
                if (lane >= d) winc += t;
            }
            warp_sum[lane] = winc - s;
        }
        __syncthreads();
        long long excl = running + warp_sum[warp] + inc - v;
        if (i < n) offsets[i] = excl;
        __syncthreads();
        if (tid == 1023) running = excl + v;
        __syncthreads();
    }
    if (tid == 0) offsets[n] = running;
}

__global__ void pack_kernel(const int32_t *__restrict__ counts,
                            const uint32_t (*__restrict__ strided)[2], int stride,
                            const int64_t *__restrict__ offsets,
                            uint32_t (*__restrict__ packed)[2], int64_t capacity) {
    const int p = blockIdx.x;
    const int n = counts[p];
    const int64_t off = offsets[p];
    if (off + n > capacity) return;
    const uint2 *src = reinterpret_cast<const uint2 *>(strided + (size_t)p * stride);
    uint2 *dst = reinterpret_cast<uint2 *>(packed + off);
    for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[k];
}

int launch_pack(const int32_t *counts_dev, int n_pairs, const uint32_t (*strided_dev)[2],
                int stride, int64_t *offsets_dev, uint32_t (*packed_dev)[2],
                int64_t base_offset, int64_t capacity, cudaStream_t st) {
    if (n_pairs <= 0) return XRB_OK;
    scan_counts_kernel<<<1, 1024, 0, st>>>(counts_dev, n_pairs, base_offset, offsets_dev);
    XRB_LAUNCHED();
    pack_kernel<<<n_pairs, 256, 0, st>>>(counts_dev, strided_dev, stride, offsets_dev,
                                         packed_dev, capacity);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

}  // namespace xrb
