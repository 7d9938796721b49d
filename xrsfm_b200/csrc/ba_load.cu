// ba_load.cu — problem structure of path B, built on the device (see ba_structure.cuh).
//
// Host work left in xrb_ba_load: argument checks, the O(cameras) assignment of reduced columns
// (which cameras ceres::Problem would hold: ba_solver.cc:334-356, constant blocks :611-614) and,
// with several ranks, the prefix walk of xrb_ba_shard_range.  The 2M-observation passes (counting
// sorts by point and by camera, bandwidth, variable flags) are radix sorts and flat kernels here.
#include <cuda_runtime.h>
#include <thrust/binary_search.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/sequence.h>
#include <thrust/sort.h>

#include <algorithm>
#include <climits>
#include <vector>

#include "ba_structure.cuh"

namespace xrb {

namespace {

enum { ST_BAD = 0, ST_VAR_PTS = 1, ST_RES_BLOCKS = 2, ST_BW = 3, ST_COUNT = 4 };

__global__ void k_init_stats(int32_t *stats) {
    stats[ST_BAD] = INT_MAX, stats[ST_VAR_PTS] = 0, stats[ST_RES_BLOCKS] = 0, stats[ST_BW] = 0;
}

// range check of every observation (first offender wins) + which cameras are observed at all
__global__ void k_validate(int n_obs, const int32_t *__restrict__ cam, const int32_t *__restrict__ pt, int C,
                           int NP, uint8_t *__restrict__ cam_seen, int32_t *__restrict__ stats) {
    for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < n_obs; o += gridDim.x * blockDim.x) {
        const int c = cam[o], p = pt[o];
        if ((unsigned)c >= (unsigned)C || (unsigned)p >= (unsigned)NP) {
            atomicMin(&stats[ST_BAD], o);
            continue;
        }
        if (!cam_seen[c]) cam_seen[c] = 1;  // benign race: every writer stores 1
    }
}

// one thread per point: variable flag, residual blocks with a variable parameter block, and the
// span of reduced columns its cameras touch (half bandwidth of S)
__global__ void k_point_struct(int NP, const int32_t *__restrict__ ptr, const int32_t *__restrict__ pt_obs,
                               const int32_t *__restrict__ raw_cam, const uint8_t *__restrict__ fixed,
                               const int32_t *__restrict__ colq, const int32_t *__restrict__ colt,
                               uint8_t *__restrict__ pt_var, int32_t *__restrict__ stats) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    int var = 0, nres = 0, span = 0;
    if (p < NP) {
        const int k0 = ptr[p], k1 = ptr[p + 1];
        var = (k1 > k0) && !(fixed && fixed[p]);
        int lo = INT_MAX, hi = -1;
        for (int k = k0; k < k1; ++k) {
            const int c = raw_cam[pt_obs[k]];
            const int cq = colq[c], ct = colt[c];
            nres += (cq >= 0 || ct >= 0 || var) ? 1 : 0;
            if (cq >= 0) lo = min(lo, cq), hi = max(hi, cq + 2);
            if (ct >= 0) lo = min(lo, ct), hi = max(hi, ct + 2);
        }
        if (var && hi >= 0) span = hi - lo;
        pt_var[p] = (uint8_t)var;
    }
    var = __reduce_add_sync(0xffffffffu, var);
    nres = __reduce_add_sync(0xffffffffu, nres);
    span = __reduce_max_sync(0xffffffffu, span);
    if ((threadIdx.x & 31) == 0) {
        if (var) atomicAdd(&stats[ST_VAR_PTS], var);
        if (nres) atomicAdd(&stats[ST_RES_BLOCKS], nres);
        if (span) atomicMax(&stats[ST_BW], span);
    }
}

// point-major observation arrays of this rank's shard
__global__ void k_gather_local(int n_local, int o_lo, int p_lo, const int32_t *__restrict__ pt_obs,
                               const int32_t *__restrict__ sorted_pt, const int32_t *__restrict__ raw_cam,
                               const double2 *__restrict__ raw_uv, int32_t *__restrict__ obs_cam,
                               int32_t *__restrict__ obs_orig, int32_t *__restrict__ obs_pt,
                               double2 *__restrict__ obs_uv) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_local) return;
    const int o = pt_obs[o_lo + k];
    obs_cam[k] = raw_cam[o];
    obs_orig[k] = o;
    obs_pt[k] = sorted_pt[o_lo + k] - p_lo;
    obs_uv[k] = raw_uv[o];
}

__global__ void k_local_ptr(int n, int p_lo, int o_lo, const int32_t *__restrict__ ptr_g, int32_t *__restrict__ ptr_l) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) ptr_l[p] = ptr_g[p_lo + p] - o_lo;
}

inline unsigned blocks_for(int64_t n, int threads) { return (unsigned)std::max<int64_t>(1, (n + threads - 1) / threads); }

}  // namespace

void BAStructScratch::release() {
    DevBuf *b[] = {&raw_cam, &raw_pt, &raw_uv, &keys, &vals, &pt_ptr_g, &pt_fixed, &pt_var_g,
                   &cam_seen, &stats, &pair_ptr, &keys64, &head, &scan};
    for (DevBuf *x : b) x->release();
    pool.release();
    if (h_stats) cudaFreeHost(h_stats);
    if (h_seen) cudaFreeHost(h_seen);
    h_stats = nullptr, h_seen = nullptr, h_seen_cap = 0;
}

int ba_build_structure(const xrb_ba_problem *P, int rank, int world, BAStructScratch &W,
                       const BAStructBufs &out, BAStructInfo *info, cudaStream_t st) {
    const int C = P->n_cams, NP = P->n_pts, NO = P->n_obs;
    int rc;
    // ---- raw uploads (the only pass over the caller's observation arrays)
    if ((rc = W.raw_cam.reserve(std::max<size_t>(1, NO) * 4))) return rc;
    if ((rc = W.raw_pt.reserve(std::max<size_t>(1, NO) * 4))) return rc;
    if ((rc = W.raw_uv.reserve(std::max<size_t>(1, NO) * 16))) return rc;
    if ((rc = W.keys.reserve(std::max<size_t>(1, NO) * 4))) return rc;
    if ((rc = W.vals.reserve(std::max<size_t>(1, NO) * 4))) return rc;
    if ((rc = W.pt_ptr_g.reserve(((size_t)NP + 1) * 4))) return rc;
    if ((rc = W.pt_fixed.reserve(std::max<size_t>(1, NP)))) return rc;
    if ((rc = W.pt_var_g.reserve(std::max<size_t>(1, NP)))) return rc;
    if ((rc = W.cam_seen.reserve(std::max<size_t>(1, C)))) return rc;
    if ((rc = W.stats.reserve(ST_COUNT * 4))) return rc;
    if (!W.h_stats) XRB_CUDA(cudaMallocHost(&W.h_stats, ST_COUNT * sizeof(int32_t)));
    if (W.h_seen_cap < (size_t)C) {
        if (W.h_seen) cudaFreeHost(W.h_seen);
        W.h_seen = nullptr, W.h_seen_cap = 0;
        XRB_CUDA(cudaMallocHost(&W.h_seen, std::max<size_t>(64, C)));
        W.h_seen_cap = std::max<size_t>(64, C);
    }
    if (NO) {
        XRB_CUDA(cudaMemcpyAsync(W.raw_cam.p, P->obs_cam, (size_t)NO * 4, cudaMemcpyHostToDevice, st));
        XRB_CUDA(cudaMemcpyAsync(W.raw_pt.p, P->obs_pt, (size_t)NO * 4, cudaMemcpyHostToDevice, st));
        XRB_CUDA(cudaMemcpyAsync(W.raw_uv.p, P->obs_uv, (size_t)NO * 16, cudaMemcpyHostToDevice, st));
    }
    const bool has_fixed = P->pt_fixed != nullptr;
    if (has_fixed && NP) XRB_CUDA(cudaMemcpyAsync(W.pt_fixed.p, P->pt_fixed, (size_t)NP, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaMemsetAsync(W.cam_seen.p, 0, std::max<size_t>(1, C), st));
    k_init_stats<<<1, 1, 0, st>>>(W.stats.as<int32_t>());
    XRB_LAUNCHED();
    k_validate<<<std::min(blocks_for(NO, 256), 148u * 8u), 256, 0, st>>>(NO, W.raw_cam.as<int32_t>(), W.raw_pt.as<int32_t>(), C,
                                                                    NP, W.cam_seen.as<uint8_t>(), W.stats.as<int32_t>());
    XRB_LAUNCHED();
    XRB_CUDA(cudaMemcpyAsync(W.h_stats, W.stats.p, 4, cudaMemcpyDeviceToHost, st));
    XRB_CUDA(cudaMemcpyAsync(W.h_seen, W.cam_seen.p, std::max<size_t>(1, C), cudaMemcpyDeviceToHost, st));

    // ---- stable sort of the observation indices by point -> point-major order (original order
    // kept inside a point, like the counting sort it replaces), then the CSR offsets
    try {
        auto pol = thrust::cuda::par_nosync(W.pool).on(st);
        thrust::device_ptr<int32_t> k(W.keys.as<int32_t>()), v(W.vals.as<int32_t>());
        if (NO) XRB_CUDA(cudaMemcpyAsync(W.keys.p, W.raw_pt.p, (size_t)NO * 4, cudaMemcpyDeviceToDevice, st));
        thrust::sequence(pol, v, v + NO);
        thrust::stable_sort_by_key(pol, k, k + NO, v);
        thrust::device_ptr<int32_t> pp(W.pt_ptr_g.as<int32_t>());
        thrust::lower_bound(pol, k, k + NO, thrust::counting_iterator<int32_t>(0),
                            thrust::counting_iterator<int32_t>(NP + 1), pp);
    } catch (const std::exception &e) {
        set_error("ba_load: sorting the observations by point failed: %s", e.what());
        return XRB_ERR_CUDA;
    }
    XRB_CUDA(cudaStreamSynchronize(st));
    if (W.h_stats[ST_BAD] != INT_MAX) {
        const int o = W.h_stats[ST_BAD];
        set_error("ba_load: observation %d references camera %d / point %d out of range", o, P->obs_cam[o], P->obs_pt[o]);
        return XRB_ERR_INVALID;
    }

    // ---- reduced columns (host, O(cameras)): a camera without observations is not in the problem
    std::vector<int32_t> colq(std::max(1, C), -1), colt(std::max(1, C), -1);
    info->nc = 0, info->n_var_q = info->n_var_t = 0;
    for (int c = 0; c < C; ++c) {
        if (!W.h_seen[c]) continue;
        if (!(P->cam_q_fixed && P->cam_q_fixed[c])) colq[c] = info->nc, info->nc += 3, info->n_var_q++;
        if (!(P->cam_t_fixed && P->cam_t_fixed[c])) colt[c] = info->nc, info->nc += 3, info->n_var_t++;
    }
    if ((rc = out.colq->reserve(colq.size() * 4))) return rc;
    if ((rc = out.colt->reserve(colt.size() * 4))) return rc;
    XRB_CUDA(cudaMemcpyAsync(out.colq->p, colq.data(), colq.size() * 4, cudaMemcpyHostToDevice, st));
    XRB_CUDA(cudaMemcpyAsync(out.colt->p, colt.data(), colt.size() * 4, cudaMemcpyHostToDevice, st));
    info->h_colq = colq, info->h_colt = colt;
    k_point_struct<<<blocks_for(NP, 128), 128, 0, st>>>(NP, W.pt_ptr_g.as<int32_t>(), W.vals.as<int32_t>(),
                                                        W.raw_cam.as<int32_t>(), has_fixed ? W.pt_fixed.as<uint8_t>() : nullptr,
                                                        out.colq->as<int32_t>(), out.colt->as<int32_t>(),
                                                        W.pt_var_g.as<uint8_t>(), W.stats.as<int32_t>());
    XRB_LAUNCHED();
    XRB_CUDA(cudaMemcpyAsync(W.h_stats, W.stats.p, ST_COUNT * 4, cudaMemcpyDeviceToHost, st));

    // ---- shard of this rank: contiguous point range balanced by the Schur work
    int p_lo = 0, p_hi = NP, o_lo = 0, o_hi = NO;
    if (world > 1) {
        std::vector<int32_t> ptr((size_t)NP + 1), kp(std::max(1, NP));
        XRB_CUDA(cudaMemcpyAsync(ptr.data(), W.pt_ptr_g.p, ptr.size() * 4, cudaMemcpyDeviceToHost, st));
        XRB_CUDA(cudaStreamSynchronize(st));
        for (int p = 0; p < NP; ++p) kp[p] = ptr[p + 1] - ptr[p];
        info->shard_lo.assign((size_t)world + 1, NP);
        for (int r = 0; r < world; ++r) {
            int32_t lo32 = 0, hi32 = NP;
            if ((rc = xrb_ba_shard_range(NP, kp.data(), r, world, &lo32, &hi32))) return rc;
            info->shard_lo[r] = lo32;
            if (r == rank) p_lo = lo32, p_hi = hi32;
        }
        o_lo = ptr[p_lo], o_hi = ptr[p_hi];
    } else {
        info->shard_lo = {0, NP};
    }
    const int PL = p_hi - p_lo, OL = o_hi - o_lo;
    info->p_lo = p_lo, info->P_local = PL, info->O_local = OL;
    if ((rc = out.pt_ptr->reserve(((size_t)PL + 1) * 4))) return rc;
    if ((rc = out.obs_cam->reserve(std::max<size_t>(1, OL) * 4))) return rc;
    if ((rc = out.obs_orig->reserve(std::max<size_t>(1, OL) * 4))) return rc;
    if ((rc = out.obs_pt->reserve(std::max<size_t>(1, OL) * 4))) return rc;
    if ((rc = out.obs_uv->reserve(std::max<size_t>(1, OL) * 16))) return rc;
    if ((rc = out.pt_var->reserve(std::max<size_t>(1, PL)))) return rc;
    if ((rc = out.cam_ptr->reserve(((size_t)C + 1) * 4))) return rc;
    if ((rc = out.cam_obs->reserve(std::max<size_t>(1, OL) * 4))) return rc;
    k_local_ptr<<<blocks_for(PL + 1, 256), 256, 0, st>>>(PL + 1, p_lo, o_lo, W.pt_ptr_g.as<int32_t>(), out.pt_ptr->as<int32_t>());
    XRB_LAUNCHED();
    k_gather_local<<<blocks_for(OL, 256), 256, 0, st>>>(OL, o_lo, p_lo, W.vals.as<int32_t>(), W.keys.as<int32_t>(),
                                                        W.raw_cam.as<int32_t>(), W.raw_uv.as<double2>(),
                                                        out.obs_cam->as<int32_t>(), out.obs_orig->as<int32_t>(),
                                                        out.obs_pt->as<int32_t>(), out.obs_uv->as<double2>());
    XRB_LAUNCHED();
    if (PL) XRB_CUDA(cudaMemcpyAsync(out.pt_var->p, W.pt_var_g.as<uint8_t>() + p_lo, (size_t)PL, cudaMemcpyDeviceToDevice, st));

    // ---- camera-major CSR of the local observations (k_cam_blocks walks it)
    try {
        auto pol = thrust::cuda::par_nosync(W.pool).on(st);
        thrust::device_ptr<int32_t> k(W.keys.as<int32_t>()), v(out.cam_obs->as<int32_t>());
        if (OL) XRB_CUDA(cudaMemcpyAsync(W.keys.p, out.obs_cam->p, (size_t)OL * 4, cudaMemcpyDeviceToDevice, st));
        thrust::sequence(pol, v, v + OL);
        thrust::stable_sort_by_key(pol, k, k + OL, v);
        thrust::device_ptr<int32_t> cp(out.cam_ptr->as<int32_t>());
        thrust::lower_bound(pol, k, k + OL, thrust::counting_iterator<int32_t>(0),
                            thrust::counting_iterator<int32_t>(C + 1), cp);
    } catch (const std::exception &e) {
        set_error("ba_load: sorting the observations by camera failed: %s", e.what());
        return XRB_ERR_CUDA;
    }
    XRB_CUDA(cudaStreamSynchronize(st));
    info->n_var_pts = W.h_stats[ST_VAR_PTS];
    info->n_res_blocks = W.h_stats[ST_RES_BLOCKS];
    info->bw = std::max(W.h_stats[ST_BW], 5);
    return XRB_OK;
}

}  // namespace xrb
