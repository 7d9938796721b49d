// ba_structure.cuh — device-side construction of the bundle-adjustment problem structure.
//
// xrb_ba_load receives the flat problem the reference's BASolver::SetUp walks
// (ba_solver.cc:330-356: one ReProjectionCost per (frame, track) observation).  Everything
// derived from it — the point-major CSR, the camera-major CSR, the half bandwidth of S, the
// variable flags, the incidence lists of the 6x6 blocks — is built on the device from ONE upload of
// the raw observation arrays (ba_load.cu, ba_struct.cu): the host keeps only O(cameras) work.
#pragma once
#include <cstddef>
#include <new>
#include <vector>

#include "ba_kernels.cuh"

namespace xrb {

// Temporary-storage pool for thrust (radix sort double buffers): blocks are kept between loads,
// so a steady-state xrb_ba_load performs no cudaMalloc/cudaFree.
struct PoolAlloc {
    typedef char value_type;
    struct Blk {
        char *p;
        size_t n;
        bool used;
    };
    std::vector<Blk> blks;
    char *allocate(std::ptrdiff_t n) {
        Blk *best = nullptr;
        for (auto &b : blks)
            if (!b.used && b.n >= (size_t)n && (!best || b.n < best->n)) best = &b;
        if (best) {
            best->used = true;
            return best->p;
        }
        char *p = nullptr;
        if (cudaMalloc(&p, (size_t)n) != cudaSuccess) throw std::bad_alloc();
        blks.push_back({p, (size_t)n, true});
        return p;
    }
    void deallocate(char *p, size_t) {
        for (auto &b : blks)
            if (b.p == p) b.used = false;
    }
    void release() {
        for (auto &b : blks) cudaFree(b.p);
        blks.clear();
    }
};

struct BAStructScratch {
    DevBuf raw_cam, raw_pt, raw_uv;  // the caller's observation arrays, uploaded as they are
    DevBuf keys, vals;               // sort keys / original observation indices
    DevBuf pt_ptr_g, pt_fixed, pt_var_g, cam_seen, stats;
    DevBuf pair_ptr, keys64, head, scan;  // block lists (ba_struct.cu)
    PoolAlloc pool;
    int32_t *h_stats = nullptr;  // pinned: [bad obs, n_var_pts, n_res_blocks, bandwidth]
    uint8_t *h_seen = nullptr;   // pinned, grow-only
    size_t h_seen_cap = 0;
    void release();
};

struct BAStructBufs {  // outputs (owned by the solver)
    DevBuf *colq, *colt, *pt_ptr, *obs_cam, *obs_uv, *pt_var, *obs_orig, *obs_pt, *cam_ptr, *cam_obs;
};

struct BAStructInfo {
    int nc, n_var_q, n_var_t, n_var_pts, n_res_blocks, bw;
    int p_lo, P_local, O_local;
    std::vector<int32_t> shard_lo;  // [world + 1]: rank r owns points [shard_lo[r], shard_lo[r+1])
    std::vector<int32_t> h_colq, h_colt;  // natural-order reduced columns (host copy; xrb_ba_load reorders them)
};

// Upload obs_cam/obs_pt/obs_uv and derive the structure for points [p_lo, p_hi) of this rank.
int ba_build_structure(const xrb_ba_problem *P, int rank, int world, BAStructScratch &W,
                       const BAStructBufs &out, BAStructInfo *info, cudaStream_t st);

// Incidence lists of the off-diagonal blocks of S for the local points (generation-2 Schur).
int ba_build_block_lists(const BAProblemDev &P, BAStructScratch &W, DevBuf &d_inc, DevBuf &d_blk_ptr,
                         DevBuf &d_blk_cams, int *n_blocks, int64_t *n_inc, cudaStream_t st);

}  // namespace xrb
