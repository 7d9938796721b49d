// ba_plan.cuh — symbolic side of the reduced-camera-system solver (path B, generation 3).
//
// Replaces what Ceres' SparseSchurComplementSolver asks of SuiteSparse (ordering + symbolic
// factorisation, selected by ba_solver.cc:74 SPARSE_SCHUR): the reduced camera system S is stored
// as 64 x 64 tiles, only the structurally non-zero ones plus the fill the factorisation creates.
// A plan is built once per loaded problem:
//   1. column order   natural, or — when S is narrow-banded (a sequential scene: KITTI-shaped C4) —
//                     a one-level nested dissection: contiguous camera ranges separated by camera
//                     ranges at least one bandwidth wide, interiors first, separators last, every
//                     part starting on a tile boundary (padding columns are identity rows).
//                     The interiors factor independently: depth nt/P + P instead of nt.
//   2. tile pattern   from the actual point-camera incidences (k_tile_pattern, all ranks see all
//                     observations at load time, so every rank derives the same plan).
//   3. symbolic fill  elimination tree on tile columns, struct(L_k) by child merging.
//   4. task lists     F (diagonal tile: last substitution + last update + factor, the "chain"),
//                     P (substitution of one tile), U (update of one tile), each with the flag
//                     values it waits for; sorted by longest-path level so that CTAs pulling them
//                     in order never wait on a task that is queued behind them.  Backward
//                     substitution: B (one column) / W (far part of a column) lists by reverse level.
// Updates reach a tile in increasing column order (sequence numbers), so the factorisation is
// bit-reproducible — every rank of a multi-GPU solve computes the identical factor.
#pragma once
#include <cstdint>
#include <vector>

#include "common.cuh"

namespace xrb {

// Device view of the tile map: element (r, c), r >= c, of the lower triangle.
struct TileMap {
    int nt = 0;
    const int32_t *tab = nullptr;  // [nt * nt]: slot of tile (i, j), i >= j, or -1
#ifdef __CUDACC__
    __device__ __forceinline__ size_t at(int r, int c) const {
        const int s = __ldg(tab + (size_t)(r >> 6) * nt + (c >> 6));
        return (size_t)s * 4096 + (size_t)((r & 63) * 64 + (c & 63));
    }
#endif
};

constexpr int kNearTiles = 3;   // tiles of a column the backward chain handles itself
constexpr int kMaxChainCtas = 32;

enum { TASK_P = 0, TASK_U = 1, TASK_PU = 2 };

// records are int32[8] (two int4 loads)
//   F : k, kp, s_kk, s_kkp, s_kpkp, need_kk, need_kkp, level
//   W : type, s_ik, s_jk (P: s_kk), s_ij (P: unused), seq (P: need), k, i, level
//       PU (fused P(i,k) + U(i,pk,k)): 2, s_ik, s_kk, s_ij = (i,pk), need | seq << 16, k, s_jk = (pk,k), level
//   B : k, s_kk, n_near, has_far, row[3], level | slot[3], 0 ...   (int32[12])
//   WB: k, far_begin, far_end, level
struct CholPlanHost {
    int nt = 0, n_tiles = 0, n_tiles_orig = 0;
    std::vector<int32_t> tab;
    std::vector<int32_t> colptr, rowidx, slot;  // strictly-lower structure of L by tile column (with fill)
    std::vector<int32_t> ftasks, wtasks, btasks, wbtasks;
    std::vector<int32_t> far_rows, far_slots;
    int n_f = 0, n_w = 0, n_b = 0, n_wb = 0;
    int n_chain_f = 1, n_chain_b = 1;
    int depth_f = 0, depth_b = 0;  // longest dependency paths (tasks)
    double flops = 0.0;            // factor + both substitutions, as executed (tile granularity)
};

// pat[i * nt + j] != 0 for i >= j: tile (i, j) holds an original non-zero.  Diagonal tiles are always kept.
int build_chol_plan(int nt, const uint8_t *pat, CholPlanHost &H);

// Column order.  widths[v] = reduced columns (3 or 6) of the v-th variable camera in natural order, bw =
// largest column span (last - first column) of the cameras of one point in natural numbering.
// Output: start[v] first column of camera v, n_pad = padded dimension (multiple of 64), parts = number of
// independent interiors (1 = natural order).
void plan_column_order(const std::vector<int> &widths, int bw, bool allow_nd, std::vector<int32_t> &start, int &n_pad,
                       int &parts);

struct CholPlanDev {
    int nt = 0, n_tiles = 0;
    const int32_t *tab = nullptr;
    const int4 *ftasks = nullptr, *wtasks = nullptr, *btasks = nullptr, *wbtasks = nullptr;
    const int32_t *far_rows = nullptr, *far_slots = nullptr;
    int n_f = 0, n_w = 0, n_b = 0, n_wb = 0, n_chain_f = 1, n_chain_b = 1;
};

struct CholPlan {
    CholPlanHost h;
    CholPlanDev d;
    DevBuf buf;  // one allocation behind every device array of `d`
    int upload(cudaStream_t st);
    void release() { buf.release(); }
};

// Marks pat (device, [nt * nt] bytes, zeroed by the caller) from the point-major observation lists of ALL
// points: tiles (i, j) touched by camera pairs that share a variable point, and each camera's own tiles.
int launch_tile_pattern(int n_pts, const int32_t *pt_ptr, const int32_t *pt_obs, const int32_t *raw_cam,
                        const uint8_t *pt_var, int n_cams, const int32_t *colq, const int32_t *colt, int nt, uint8_t *pat,
                        cudaStream_t st);

}  // namespace xrb
