// ba_struct.cu — one-off structure of the reduced camera system for path B, generation 2.
//
// The sparsity of S = U - W V^-1 W^T is fixed for a problem: block (a, b) receives one 6x6
// contribution per point that cameras a and b both observe.  This file builds, on the device
// (thrust radix sort — plumbing, not a hot-path kernel), the per-block incidence lists the
// gather kernel (ba_kernels.cu: k_gather) walks every LM iteration:
//     blocks   : (cam_a, cam_b), first reduced column of a > first reduced column of b
//     blk_ptr  : CSR offsets into `inc`
//     inc      : (obs_i, obs_j) point-major observation indices, obs_i belongs to cam_a
// sorted by (cam_a, cam_b) so the a-side operand of consecutive blocks is shared in L2.
#include <cuda_runtime.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/binary_search.h>
#include <thrust/scan.h>
#include <thrust/sort.h>

#include <algorithm>
#include <vector>

#include "ba_structure.cuh"

namespace xrb {

__device__ __forceinline__ int first_col(const int32_t *colq, const int32_t *colt, int c) {
    return colq[c] >= 0 ? colq[c] : colt[c];
}

// one thread per local point: enumerate its unordered observation pairs
__global__ void k_enum_pairs(int n_pts, const int32_t *__restrict__ pt_ptr, const int64_t *__restrict__ pair_ptr,
                             const int32_t *__restrict__ obs_cam, const uint8_t *__restrict__ pt_var,
                             const int32_t *__restrict__ colq, const int32_t *__restrict__ colt,
                             unsigned long long *__restrict__ keys, int2 *__restrict__ vals) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pts) return;
    const int k0 = pt_ptr[p], kn = pt_ptr[p + 1] - k0;
    int64_t w = pair_ptr[p];
    const bool pv = pt_var[p] != 0;
    for (int i = 1; i < kn; ++i)
        for (int j = 0; j < i; ++j, ++w) {
            const int ci = obs_cam[k0 + i], cj = obs_cam[k0 + j];
            const int fi = first_col(colq, colt, ci), fj = first_col(colq, colt, cj);
            if (!pv || fi < 0 || fj < 0) {  // constant point or constant camera: no Schur term
                keys[w] = ~0ull;
                vals[w] = make_int2(-1, -1);
                continue;
            }
            const bool i_is_a = fi >= fj;
            const int ca = i_is_a ? ci : cj, cb = i_is_a ? cj : ci;
            keys[w] = ((unsigned long long)(unsigned)ca << 32) | (unsigned)cb;
            vals[w] = i_is_a ? make_int2(k0 + i, k0 + j) : make_int2(k0 + j, k0 + i);
        }
}

// head flags of the sorted key runs -> block table
__global__ void k_mark_heads(int64_t n, const unsigned long long *__restrict__ keys, int32_t *__restrict__ head) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    head[i] = (keys[i] != ~0ull) && (i == 0 || keys[i] != keys[i - 1]);
}

__global__ void k_fill_blocks(int64_t n, const unsigned long long *__restrict__ keys,
                              const int32_t *__restrict__ head_scan, const int32_t *__restrict__ head,
                              int2 *__restrict__ blk_cams, int32_t *__restrict__ blk_ptr) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !head[i]) return;
    const int b = head_scan[i] - 1;  // inclusive scan
    blk_cams[b] = make_int2((int)(keys[i] >> 32), (int)(keys[i] & 0xFFFFFFFFu));
    blk_ptr[b] = (int32_t)i;
}

// unordered observation pairs of each local point: k (k - 1) / 2, and a trailing 0 for the scan
__global__ void k_pair_count(int n_pts, const int32_t *__restrict__ pt_ptr, int64_t *__restrict__ cnt) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p > n_pts) return;
    const int64_t k = p < n_pts ? pt_ptr[p + 1] - pt_ptr[p] : 0;
    cnt[p] = k * (k - 1) / 2;
}

int ba_build_block_lists(const BAProblemDev &P, BAStructScratch &W, DevBuf &d_inc, DevBuf &d_blk_ptr,
                         DevBuf &d_blk_cams, int *n_blocks, int64_t *n_inc, cudaStream_t st) {
    *n_blocks = 0, *n_inc = 0;
    int rc;
    if ((rc = d_blk_ptr.reserve(16))) return rc;
    if (P.n_pts_local == 0) return XRB_OK;
    if ((rc = W.pair_ptr.reserve(((size_t)P.n_pts_local + 1) * 8))) return rc;
    auto pol = thrust::cuda::par_nosync(W.pool).on(st);
    int64_t total = 0;
    k_pair_count<<<(P.n_pts_local + 1 + 255) / 256, 256, 0, st>>>(P.n_pts_local, P.pt_ptr, W.pair_ptr.as<int64_t>());
    XRB_LAUNCHED();
    try {
        thrust::device_ptr<int64_t> c(W.pair_ptr.as<int64_t>());
        thrust::exclusive_scan(pol, c, c + P.n_pts_local + 1, c);
    } catch (const std::exception &e) {
        set_error("ba: building the block lists failed: %s", e.what());
        return XRB_ERR_CUDA;
    }
    XRB_CUDA(cudaMemcpyAsync(&total, W.pair_ptr.as<int64_t>() + P.n_pts_local, 8, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    if (total == 0) return XRB_OK;
    if (total >= (int64_t)INT32_MAX) {
        set_error("ba: %lld observation pairs exceed the 2^31 limit of the block lists", (long long)total);
        return XRB_ERR_INVALID;
    }
    if ((rc = W.keys64.reserve((size_t)total * 8))) return rc;
    if ((rc = d_inc.reserve((size_t)total * 8))) return rc;
    if ((rc = W.head.reserve((size_t)total * 4))) return rc;
    if ((rc = W.scan.reserve((size_t)total * 4))) return rc;
    unsigned long long *keys = W.keys64.as<unsigned long long>();
    k_enum_pairs<<<(P.n_pts_local + 127) / 128, 128, 0, st>>>(P.n_pts_local, P.pt_ptr, W.pair_ptr.as<int64_t>(), P.obs_cam,
                                                           P.pt_var, P.colq, P.colt, keys, d_inc.as<int2>());
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    int64_t valid = total;
    try {
        thrust::device_ptr<unsigned long long> k(keys);
        thrust::device_ptr<int2> v(d_inc.as<int2>());
        thrust::sort_by_key(pol, k, k + total, v);
        k_mark_heads<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, keys, W.head.as<int32_t>());
        XRB_LAUNCHED();
        thrust::device_ptr<int32_t> h(W.head.as<int32_t>()), s(W.scan.as<int32_t>());
        thrust::inclusive_scan(pol, h, h + total, s);
        // number of valid incidences = index of the first invalid key (they sort last)
        valid = thrust::lower_bound(thrust::cuda::par(W.pool).on(st), k, k + total, ~0ull) - k;
    } catch (const std::exception &e) {
        set_error("ba: building the block lists failed: %s", e.what());
        return XRB_ERR_CUDA;
    }
    int32_t nb = 0;
    XRB_CUDA(cudaMemcpyAsync(&nb, W.scan.as<int32_t>() + (total - 1), 4, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    if ((rc = d_blk_ptr.reserve((size_t)(nb + 1) * 4))) return rc;
    if ((rc = d_blk_cams.reserve(std::max<size_t>(1, (size_t)nb) * 8))) return rc;
    k_fill_blocks<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(total, keys, W.scan.as<int32_t>(), W.head.as<int32_t>(),
                                                                    d_blk_cams.as<int2>(), d_blk_ptr.as<int32_t>());
    XRB_LAUNCHED();
    const int32_t v32 = (int32_t)valid;
    XRB_CUDA(cudaMemcpyAsync(d_blk_ptr.as<int32_t>() + nb, &v32, 4, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaStreamSynchronize(st));
    *n_blocks = nb, *n_inc = valid;
    return XRB_OK;
}

}  // namespace xrb
