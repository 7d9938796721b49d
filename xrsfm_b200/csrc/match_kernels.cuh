// match_kernels.cuh — device-side contract of the fused SIFT matcher (path M).
//
// Replaces the reference's three kernels + n1 x n2 int32 matrix in HBM
//   MultiplyDescriptor_Kernel / RowMatch_Kernel / ColMatch_Kernel
//   (3rdparty/SiftGPU/ProgramCU.cu:1491-1578, 1780-1837, 1852-1872)
// by: a score kernel that forms dot tiles on chip and pushes only the entries that can
// influence a decision ("candidates", dot > v_low) into per-row / per-column top-2 state,
// and a finalize kernel that applies the acos thresholds, the mutual test and the ordered
// compaction of SiftMatchCU::GetBestMatch (SiftMatchCU.cpp:186-215).
#pragma once
#include <cstdint>

#include "common.cuh"

namespace xrb {

constexpr int kDim = 128;  // SIFT descriptor bytes (src/base/types.h:9-10)

// One image pair as the kernels see it.
struct PairDesc {
    const uint8_t *a;  // n1 x 128, row-major
    const uint8_t *b;  // n2 x 128
    int32_t n1, n2;
};

// Top-2 state of one row (or column): written only through atomics.
//   best  : (dot << 32) | (0xFFFFFFFF - tie_rank(index))   0 == empty
//   second: largest dot that lost against `best` (ties included)   0 == none
struct Top2State {
    unsigned long long *best;
    unsigned int *second;
};

// tie_rank orders equal dots the way the reference's scans do.
//  rows   (RowMatch_Kernel, ProgramCU.cu:1798-1826): each of 32 lanes keeps its first
//         maximum (lowest j in the lane); the 16/8/4/2/1 tree then keeps the LOWER tree
//         position on ties, which ranks lanes in bit-reversed order (0,16,8,24,4,20,...):
//         rank = (bitrev5(j % 32), j / 32).
//  columns(MultiplyDescriptor partials + ColMatch, :1556-1570,1858-1864): lowest i
__host__ __device__ inline uint32_t bitrev5(uint32_t x) {
    return ((x & 1u) << 4) | ((x & 2u) << 2) | (x & 4u) | ((x & 8u) >> 2) | ((x & 16u) >> 4);
}
__host__ __device__ inline uint32_t row_tie_rank(uint32_t j) {
    return (bitrev5(j & 31u) << 20) | (j >> 5);
}
__host__ __device__ inline uint32_t row_tie_unrank(uint32_t r) {
    return ((r & 0xFFFFFu) << 5) | bitrev5(r >> 20);
}

int launch_vlow(float distmax, float ratiomax, int *vlow_dev, cudaStream_t st);

// Generation 1: 128x128 dp4a tiles on the CUDA cores.
int launch_score_dp4a(const PairDesc *pairs_dev, int n_pairs, int max_n1, int max_n2,
                      int state_stride, Top2State rows, Top2State cols, const int *vlow_dev,
                      cudaStream_t st);

// Generation 2: tcgen05 int8 MMA, accumulators in TMEM (match_tc.cu).
// baseA/baseB: device blocks every PairDesc::a / ::b points into (rowsA/rowsB descriptors);
// the TMA tensor maps are built over them.
int launch_score_tc(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA,
                    const uint8_t *baseB, uint64_t rowsB, int state_stride, Top2State rows,
                    Top2State cols, const int *vlow_dev, cudaStream_t st);
bool score_tc_available();

// Generation 3: the same tensor-core kernel with the pair's top-2 state in shared memory and
// thresholds + mutual test + ordered compaction fused in (images up to fused_max_features()
// descriptors).  dist_tab = launch_dist_table output for v = 0 .. 2^18.
int fused_max_features();
int launch_match_fused(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA,
                       const uint8_t *baseB, uint64_t rowsB, const int *vlow_dev, const float *dist_tab,
                       float distmax, float ratiomax, int mbm, int max_match, int32_t *counts_dev,
                       uint32_t (*out_dev)[2], int out_stride, cudaStream_t st);

int launch_finalize(const PairDesc *pairs_dev, int n_pairs, int state_stride, Top2State rows,
                    Top2State cols, float distmax, float ratiomax, int mbm, int max_match,
                    int32_t *counts_dev, uint32_t (*out_dev)[2], int out_stride,
                    cudaStream_t st);

int launch_build_pairs(const int32_t (*pairs_idx_dev)[2], int n_pairs, int n_images,
                       const int64_t *row_offsets_dev, const uint8_t *block_dev,
                       int max_features, PairDesc *out_dev, cudaStream_t st);

int launch_pack(const int32_t *counts_dev, int n_pairs, const uint32_t (*strided_dev)[2],
                int stride, int64_t *offsets_dev, uint32_t (*packed_dev)[2],
                int64_t base_offset, int64_t capacity, cudaStream_t st);

// float(acos(double(min(float(dot)*2^-18,1.0)))) for dot = 0..n-1 (test hook: lets the
// GPU test compare CUDA's double acos against the oracle's libm over the whole domain).
int launch_dist_table(float *out_dev, int n, cudaStream_t st);

}  // namespace xrb
