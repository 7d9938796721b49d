// match_tc.cu — generation-2 score kernel of the fused SIFT matcher: the u8 x u8 -> s32
// descriptor products run on the 5th-generation tensor cores (tcgen05.mma kind::i8,
// accumulators in TMEM), operands arrive by TMA into 128B-swizzled shared memory, and the
// epilogue reads the accumulators back with tcgen05.ld and keeps only entries above v_low.
//
// Replaces MultiplyDescriptor_Kernel (3rdparty/SiftGPU/ProgramCU.cu:1491-1578) — the exact
// integer dot products — without ever materialising the n1 x n2 matrix in HBM.
//
// One persistent CTA per SM; each CTA owns whole image pairs:
//   warp 0   TMA producer: B group (up to 1024 descriptors = 128 KB, resident) and a 3-stage
//            ring of A tiles (128 descriptors = 16 KB each)
//   warp 1   TMEM allocator + MMA issuer: per (A tile, 256-column slice of the B group)
//            4 x tcgen05.mma.cta_group::1.kind::i8 (M128 N256 K32), double-buffered in TMEM
//   warps 2-17 epilogue (four per TMEM lane quarter, each owns 64 of a tile's 256 columns): TMEM -> registers
//            (2 x 32x32b.x32 per wait), VIMNMX3 max-filter against v_low, rare candidates pushed into the
//            per-row / per-column top-2 state.  Sixteen warps instead of eight put four on every scheduler:
//            one warp's tcgen05.ld round trip and candidate path overlap with three others (135 k -> 197 k
//            pairs/s; profiles/r02_match_ncu_summary.md)
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "match_kernels.cuh"

namespace xrb {

namespace {

constexpr int kTileM = 128;          // descriptors of image 1 per MMA tile (TMEM lanes)
constexpr int kTileN = 256;          // descriptors of image 2 per MMA tile (TMEM columns)
constexpr int kAStages = 3;
constexpr int kATileBytes = kTileM * kDim;    // 16 KB
constexpr int kBoxRows = 128;                 // TMA box: 128 rows x 128 bytes
constexpr int kEpiWarps = 16;        // epilogue warps: 4 per TMEM lane quarter, each owns kTileN / 4 columns of a tile
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kColsPerWarp = kTileN / (kEpiWarps / 4);
constexpr int kThreads = 32 * (2 + kEpiWarps);
// Two flavours of the kernel (template parameter kFused):
//  false: top-2 state in global memory (any image size), resident B group of 1024 descriptors,
//         thresholds/compaction in finalize_kernel afterwards;
//  true : images up to kFusedMax descriptors: the pair's whole top-2 state lives in shared
//         memory as one packed 64-bit word per row/column, candidates go through per-warp
//         queues, and thresholds + mutual test + ordered compaction run in the same kernel.
constexpr int kFusedMax = 4096;
constexpr int kLaneQueue = 4;  // candidates a lane can park before the warp drains
template <bool kFused> struct Cfg {
    static constexpr int kGroupN = kFused ? 512 : 1024;  // resident B descriptors per group
    static constexpr int kBGroupBytes = kGroupN * kDim;
    static constexpr int kStateBytes = kFused ? 2 * kFusedMax * 8 : 0;
    static constexpr int kQueueBytes = kFused ? kEpiWarps * kLaneQueue * 32 * 8 : 0;
    static constexpr int kScratchBytes = kFused ? kEpiWarps * 16 * 32 * 4 : 0;  // per lane: one group of 16 accumulators
    static constexpr int kSmemBytes = kBGroupBytes + kAStages * kATileBytes + kStateBytes + kQueueBytes +
                                      kScratchBytes + 1024 /*align*/ + 256 /*barriers*/;
};

// ---- PTX helpers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(bar),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }

// u8 x u8 -> s32, A and B both K-major in shared memory (descriptors), D in TMEM
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                       uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
        : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version [46,48) = 1 |
// layout_type [61,64) = 2 (SWIZZLE_128B).  Rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    const uint32_t lo = ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    return ((uint64_t)hi << 32) | lo;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format S32 = 2 at [4,6),
// a/b format UINT8 = 0, K-major both, N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (2u << 4) | ((uint32_t)(kTileN >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);

__device__ __forceinline__ void push_top2(unsigned long long *best, unsigned int *second,
                                          unsigned long long key) {
    unsigned long long old = atomicMax(best, key);
    unsigned long long loser = old < key ? old : key;
    unsigned int lv = (unsigned int)(loser >> 32);
    if (lv) atomicMax(second, lv);
}

// 32 lanes x 32 consecutive TMEM columns -> 32 registers per thread
__device__ __forceinline__ void ldtm32(uint32_t taddr, int *v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}

struct Bars {
    unsigned long long b_full, b_empty;
    unsigned long long a_full[kAStages], a_empty[kAStages];
    unsigned long long t_full[2], t_empty[2];
    uint32_t tmem_base;
};

// ---- fused mode: packed top-2 word ----------------------------------------------------------
//   [63:41] best dot (23 bits)   [40:28] 8191 - tie_rank (13 bits)   [27:5] runner-up dot (23 bits)
// 0 == empty.  One 64-bit shared-memory CAS loop per update keeps best, argmax and runner-up
// consistent; the update rule is the reference's (ProgramCU.cu:1802-1807): a strictly larger
// (dot, tie order) takes over and demotes the old best to runner-up, anything else only feeds
// the runner-up.
__device__ __forceinline__ uint32_t rank13_row(uint32_t j) { return (bitrev5(j & 31u) << 8) | (j >> 5); }
__device__ __forceinline__ uint32_t unrank13_row(uint32_t r) { return ((r & 0xFFu) << 5) | bitrev5(r >> 8); }

__device__ __forceinline__ void top2_update(unsigned long long *addr, uint32_t v, uint32_t rk) {
    unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(addr), assumed;
    const unsigned long long keyc = ((unsigned long long)v << 13) | rk;
    do {
        assumed = old;
        unsigned long long nw;
        if (keyc > (assumed >> 28)) {
            nw = (keyc << 28) | ((assumed >> 41) << 5);
        } else {
            const uint32_t sec = (uint32_t)(assumed >> 5) & 0x7FFFFFu;
            if (sec >= v) return;
            nw = (assumed & ~(0x7FFFFFull << 5)) | ((unsigned long long)v << 5);
        }
        old = atomicCAS(addr, assumed, nw);
    } while (old != assumed);
}

// Apply the candidates parked in one lane's private queue to the pair's top-2 state.  Kept out
// of line on purpose: it is the only place with the 64-bit shared-memory CAS loops, it runs
// rarely (a lane's queue fills up, or the warp drains after a tile), and the epilogue's hot
// loop must stay small.
__device__ __noinline__ void lane_flush(const unsigned long long *q, int qn, unsigned long long *st_rows,
                                        unsigned long long *st_cols) {
    for (int k = 0; k < qn; ++k) {
        const unsigned long long e = q[k * 32];
        const uint32_t v = (uint32_t)(e >> 26), i = (uint32_t)(e >> 13) & 0x1FFFu, j = (uint32_t)e & 0x1FFFu;
        top2_update(st_rows + i, v, 8191u - rank13_row(j));
        top2_update(st_cols + j, v, 8191u - i);
    }
}

// Same for the global-state flavour: lock-free pushes into the per-row / per-column state.
__device__ __noinline__ void global_slow16(Top2State rows, Top2State cols, size_t sbase, int vlow, int i, int j0,
                                           int n2, int4 a, int4 b, int4 c, int4 d) {
    const int v[16] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
#pragma unroll
    for (int e = 0; e < 16; ++e) {
        const int j = j0 + e;
        if (v[e] > vlow && j < n2) {
            const unsigned long long hv = (unsigned long long)(unsigned int)v[e] << 32;
            push_top2(rows.best + sbase + i, rows.second + sbase + i, hv | (0xFFFFFFFFu - row_tie_rank((uint32_t)j)));
            push_top2(cols.best + sbase + j, cols.second + sbase + j, hv | (0xFFFFFFFFu - (uint32_t)i));
        }
    }
}

__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

struct FusedArgs {
    const float *dist_tab;  // float(acos(double(min(v * 2^-18, 1)))) for v = 0 .. 2^18
    float distmax, ratiomax;
    int mbm, max_match;
    int32_t *counts;
    uint32_t (*out)[2];
    int out_stride;
};

__device__ __forceinline__ bool accept_tab(const float *__restrict__ tab, uint32_t best, uint32_t second,
                                           float distmax, float ratiomax) {
    const float dist = __ldg(tab + min(best, 262144u));
    const float distn = __ldg(tab + min(second, 262144u));
    return (dist < distmax) && (dist < distn * ratiomax);
}

template <bool kFused>
__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const uint8_t *baseA, const uint8_t *baseB, const PairDesc *__restrict__ pairs,
                int n_pairs, int state_stride, Top2State rows, Top2State cols,
                const int *__restrict__ vlow_ptr, FusedArgs fa) {
    using C = Cfg<kFused>;
    constexpr int kGroupN = C::kGroupN;
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024 B alignment
    const uint32_t sB = base;
    const uint32_t sA = base + C::kBGroupBytes;
    unsigned char *after = smem_raw + (base - raw) + C::kBGroupBytes + kAStages * kATileBytes;
    unsigned long long *st_rows = reinterpret_cast<unsigned long long *>(after);  // [kFusedMax]
    unsigned long long *st_cols = st_rows + (kFused ? kFusedMax : 0);             // [kFusedMax]
    unsigned long long *queues = st_cols + (kFused ? kFusedMax : 0);              // [8][64]
    int *scratch_all = reinterpret_cast<int *>(after + C::kStateBytes + C::kQueueBytes);        // [8][16][32]
    Bars *bars = reinterpret_cast<Bars *>(after + C::kStateBytes + C::kQueueBytes + C::kScratchBytes);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars->b_full), 1);
        mbar_init(smem_u32(&bars->b_empty), 1);
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(smem_u32(&bars->a_full[s]), 1);
            mbar_init(smem_u32(&bars->a_empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->t_full[b]), 1);
            mbar_init(smem_u32(&bars->t_empty[b]), kEpiWarps);  // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // 512 TMEM columns = two 128 x 256 s32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            smem_u32(&bars->tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t a_it = 0, b_it = 0;
            for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
                const PairDesc pd = pairs[p];
                if (pd.n1 <= 0 || pd.n2 <= 0) continue;
                const int rowA = (int)((pd.a - baseA) / kDim), rowB = (int)((pd.b - baseB) / kDim);
                const int mt = (pd.n1 + kTileM - 1) / kTileM;
                for (int g0 = 0; g0 < pd.n2; g0 += kGroupN, ++b_it) {
                    const int gcols = min(kGroupN, pd.n2 - g0);
                    const int nbox = (gcols + kBoxRows - 1) / kBoxRows;
                    mbar_wait(smem_u32(&bars->b_empty), (b_it & 1) ^ 1);
                    mbar_expect_tx(smem_u32(&bars->b_full), nbox * kBoxRows * kDim);
                    for (int bx = 0; bx < nbox; ++bx)
                        tma_load_2d(sB + bx * kBoxRows * kDim, &mapB, 0, rowB + g0 + bx * kBoxRows,
                                    smem_u32(&bars->b_full));
                    for (int m = 0; m < mt; ++m, ++a_it) {
                        const int s = a_it % kAStages;
                        mbar_wait(smem_u32(&bars->a_empty[s]), ((a_it / kAStages) & 1) ^ 1);
                        mbar_expect_tx(smem_u32(&bars->a_full[s]), kATileBytes);
                        tma_load_2d(sA + s * kATileBytes, &mapA, 0, rowA + m * kTileM, smem_u32(&bars->a_full[s]));
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        uint32_t a_it = 0, b_it = 0, t_it = 0;
        for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
            const PairDesc pd = pairs[p];
            if (pd.n1 <= 0 || pd.n2 <= 0) continue;
            const int mt = (pd.n1 + kTileM - 1) / kTileM;
            for (int g0 = 0; g0 < pd.n2; g0 += kGroupN, ++b_it) {
                const int gcols = min(kGroupN, pd.n2 - g0);
                const int nsub = (gcols + kTileN - 1) / kTileN;
                mbar_wait(smem_u32(&bars->b_full), b_it & 1);
                for (int m = 0; m < mt; ++m, ++a_it) {
                    const int s = a_it % kAStages;
                    mbar_wait(smem_u32(&bars->a_full[s]), (a_it / kAStages) & 1);
                    for (int ns = 0; ns < nsub; ++ns, ++t_it) {
                        const int tb = t_it & 1;
                        mbar_wait(smem_u32(&bars->t_empty[tb]), ((t_it >> 1) & 1) ^ 1);
                        tc_fence_after();
                        if (lane == 0) {
                            const uint64_t da = make_desc(sA + s * kATileBytes);
                            const uint64_t db = make_desc(sB + ns * kTileN * kDim);
#pragma unroll
                            for (int kk = 0; kk < kDim / 32; ++kk)  // K = 32 bytes per instruction
                                mma_i8(tmem + tb * kTileN, da + (uint64_t)(kk * 2), db + (uint64_t)(kk * 2), kIdesc,
                                       kk > 0);
                            mma_commit(smem_u32(&bars->t_full[tb]));
                            if (ns == nsub - 1) mma_commit(smem_u32(&bars->a_empty[s]));
                            if (ns == nsub - 1 && m == mt - 1) mma_commit(smem_u32(&bars->b_empty));
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        const int ew = warp - 2;            // 0..kEpiWarps-1
        const int et = threadIdx.x - 64;    // 0..kEpiThreads-1 among the epilogue threads
        const int quarter = warp & 3;       // TMEM lane quarter this warp may touch (warp id % 4)
        const int part = ew >> 2;           // which kColsPerWarp-column part of the 256-column tile
        const int vlow = *vlow_ptr;
        // per-lane private candidate queues: q[slot][lane]; no coordination needed to append
        unsigned long long *q = queues + ew * (kLaneQueue * 32) + lane;
        int *scratch = scratch_all + ew * (16 * 32) + lane;  // [e][lane]: conflict-free
        int qn = 0;  // this lane's fill
        uint32_t t_it = 0;
        if (kFused) {
            for (int z = et; z < 2 * kFusedMax; z += kEpiThreads) st_rows[z] = 0ull;
            epi_bar();
        }
        auto drain = [&]() {
            lane_flush(q, qn, st_rows, st_cols);
            qn = 0;
            __syncwarp();
        };
        for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
            const PairDesc pd = pairs[p];
            if (pd.n1 <= 0 || pd.n2 <= 0) {
                if (kFused && et == 0) fa.counts[p] = 0;
                continue;
            }
            const size_t sbase = (size_t)p * state_stride;
            const int mt = (pd.n1 + kTileM - 1) / kTileM;
            for (int g0 = 0; g0 < pd.n2; g0 += kGroupN) {
                const int gcols = min(kGroupN, pd.n2 - g0);
                const int nsub = (gcols + kTileN - 1) / kTileN;
                for (int m = 0; m < mt; ++m) {
                    const int i = m * kTileM + quarter * 32 + lane;
                    for (int ns = 0; ns < nsub; ++ns, ++t_it) {
                        const int tb = t_it & 1;
                        mbar_wait(smem_u32(&bars->t_full[tb]), (t_it >> 1) & 1);
                        tc_fence_after();
                        const int col_tile = g0 + ns * kTileN + part * kColsPerWarp;
                        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + tb * kTileN + part * kColsPerWarp;
                        // 64 columns (2 x 32x32b.x32) in flight per wait: 18 warps share the SM's 64 K
                        // registers, which caps a thread at 112 (ptxas settles on 96) — more accumulators
                        // in flight would spill
#pragma unroll 1
                        for (int hh = 0; hh < kColsPerWarp / 64; ++hh) {
                        int v[64];
                        ldtm32(taddr + hh * 64, v);
                        ldtm32(taddr + hh * 64 + 32, v + 32);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (hh == kColsPerWarp / 64 - 1) {  // accumulator fully read: hand the TMEM buffer back
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(smem_u32(&bars->t_empty[tb]));
                        }
#pragma unroll
                        for (int g = 0; g < 64; g += 16) {
                            const int j0 = col_tile + hh * 64 + g;
                            int mx = v[g];
#pragma unroll
                            for (int e = 1; e < 16; ++e) mx = max(mx, v[g + e]);
                            if (kFused) {
                                // lane-divergent slow path (a few % of the lanes): descend by quarters
                                // of the group, park the hits in the lane's own queue
                                // lane-divergent slow path with a small code footprint: the lane dumps the
                                // group's 16 accumulators to its private scratch column (static register
                                // indices end here) and scans them in a ROLLED loop; hits are parked in the
                                // lane's private queue.  (Unrolled compares or an ABI call per group cost
                                // more in instruction-cache misses / register save-restore than the work.)
                                if (mx > vlow && i < pd.n1) {
                                    // hit mask as a shallow tree (16 independent compares, 3 levels of
                                    // 3-input ORs) rather than a 16-long dependent chain
                                    unsigned hb[16];
#pragma unroll
                                    for (int e = 0; e < 16; ++e) hb[e] = v[g + e] > vlow ? (1u << e) : 0u;
                                    const unsigned h0 = hb[0] | hb[1] | hb[2], h1 = hb[3] | hb[4] | hb[5];
                                    const unsigned h2 = hb[6] | hb[7] | hb[8], h3 = hb[9] | hb[10] | hb[11];
                                    const unsigned h4 = hb[12] | hb[13] | hb[14];
                                    unsigned hits = (h0 | h1 | h2) | (h3 | h4 | hb[15]);
                                    if ((hits & (hits - 1)) == 0) {
                                        // the usual case, a single hit: it is the group's maximum, no re-read
                                        const int e = __ffs(hits) - 1;
                                        if (j0 + e < pd.n2) {
                                            if (qn == kLaneQueue) {
                                                lane_flush(q, qn, st_rows, st_cols);
                                                qn = 0;
                                            }
                                            q[qn * 32] = ((unsigned long long)(unsigned)mx << 26) |
                                                         ((unsigned long long)(unsigned)i << 13) | (unsigned)(j0 + e);
                                            ++qn;
                                        }
                                    } else {
#pragma unroll
                                        for (int e = 0; e < 16; ++e) scratch[e * 32] = v[g + e];
                                        while (hits) {
                                            const int e = __ffs(hits) - 1;
                                            hits &= hits - 1;
                                            if (j0 + e < pd.n2) {
                                                if (qn == kLaneQueue) {  // rare: this lane's queue is full
                                                    lane_flush(q, qn, st_rows, st_cols);
                                                    qn = 0;
                                                }
                                                q[qn * 32] = ((unsigned long long)(unsigned)scratch[e * 32] << 26) |
                                                             ((unsigned long long)(unsigned)i << 13) | (unsigned)(j0 + e);
                                                ++qn;
                                            }
                                        }
                                    }
                                }
                            } else if (mx > vlow && i < pd.n1) {
                                global_slow16(rows, cols, sbase, vlow, i, j0, pd.n2,
                                              make_int4(v[g], v[g + 1], v[g + 2], v[g + 3]),
                                              make_int4(v[g + 4], v[g + 5], v[g + 6], v[g + 7]),
                                              make_int4(v[g + 8], v[g + 9], v[g + 10], v[g + 11]),
                                              make_int4(v[g + 12], v[g + 13], v[g + 14], v[g + 15]));
                            }
                        }
                        __syncwarp();  // reconverge before the next .aligned tcgen05.ld
                        }
                        if (kFused && __any_sync(0xFFFFFFFFu, qn >= kLaneQueue - 1)) drain();
                    }
                }
            }
            if (kFused) {
                // ---- the pair is complete: thresholds, mutual test, ordered compaction -------
                // (RowMatch/ColMatch threshold lines + SiftMatchCU.cpp:199-207), then wipe the state
                drain();
                epi_bar();
                int *wsum = reinterpret_cast<int *>(queues);  // queues are idle now
                int running = 0;
                uint32_t(*dst)[2] = fa.out + (size_t)p * fa.out_stride;
                for (int chunk = 0; chunk < pd.n1; chunk += kEpiThreads) {
                    const int r = chunk + et;
                    int flag = 0, j = -1;
                    if (r < pd.n1) {
                        const unsigned long long sr = st_rows[r];
                        const uint32_t bv = (uint32_t)(sr >> 41);
                        if (bv > 0 && accept_tab(fa.dist_tab, bv, (uint32_t)(sr >> 5) & 0x7FFFFFu, fa.distmax, fa.ratiomax)) {
                            j = (int)unrank13_row(8191u - ((uint32_t)(sr >> 28) & 0x1FFFu));
                            if (fa.mbm) {
                                const unsigned long long sc = st_cols[j];
                                const uint32_t cv = (uint32_t)(sc >> 41);
                                flag = cv > 0 && (int)(8191u - ((uint32_t)(sc >> 28) & 0x1FFFu)) == r &&
                                       accept_tab(fa.dist_tab, cv, (uint32_t)(sc >> 5) & 0x7FFFFFu, fa.distmax, fa.ratiomax);
                            } else {
                                flag = 1;
                            }
                        }
                    }
                    const unsigned ballot = __ballot_sync(0xFFFFFFFFu, flag);
                    if (lane == 0) wsum[ew] = __popc(ballot);
                    epi_bar();
                    int before = 0, total = 0;
#pragma unroll
                    for (int w = 0; w < kEpiWarps; ++w) {
                        const int c = wsum[w];
                        before += w < ew ? c : 0;
                        total += c;
                    }
                    const int pos = running + before + __popc(ballot & ((1u << lane) - 1u));
                    if (flag && pos < fa.max_match) dst[pos][0] = (uint32_t)r, dst[pos][1] = (uint32_t)j;
                    running += total;
                    epi_bar();
                }
                if (et == 0) fa.counts[p] = running < fa.max_match ? running : fa.max_match;
                for (int z = et; z < pd.n1; z += kEpiThreads) st_rows[z] = 0ull;
                for (int z = et; z < pd.n2; z += kEpiThreads) st_cols[z] = 0ull;
                epi_bar();
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

// ---- host: tensor maps ---------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// rows x 128 bytes, box 128 rows x 128 bytes, 128B swizzle, out-of-range rows read as zero
int make_map(CUtensorMap *map, const uint8_t *base, uint64_t rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return XRB_ERR_CUDA;
    }
    cuuint64_t gdim[2] = {(cuuint64_t)kDim, rows ? rows : 1};
    cuuint64_t gstride[1] = {(cuuint64_t)kDim};
    cuuint32_t box[2] = {(cuuint32_t)kDim, (cuuint32_t)kBoxRows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, (void *)base, gdim, gstride, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d) for base %p rows %llu", (int)r, (const void *)base,
                  (unsigned long long)rows);
        return XRB_ERR_CUDA;
    }
    return XRB_OK;
}

}  // namespace

bool score_tc_available() {
    const char *env = getenv("XRB_MATCH_VARIANT");
    if (env && atoi(env) == 1) return false;
    return get_encode() != nullptr;
}

namespace {
template <bool kFused>
int launch_impl(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA, const uint8_t *baseB,
                uint64_t rowsB, int state_stride, Top2State rows, Top2State cols, const int *vlow_dev,
                const FusedArgs &fa, cudaStream_t st) {
    if (n_pairs <= 0) return XRB_OK;
    if (((uintptr_t)baseA & 15) || ((uintptr_t)baseB & 15)) {
        set_error("tcgen05 matcher: descriptor blocks must be 16-byte aligned");
        return XRB_ERR_INVALID;
    }
    CUtensorMap mapA, mapB;
    int rc;
    if ((rc = make_map(&mapA, baseA, rowsA))) return rc;
    if ((rc = make_map(&mapB, baseB, rowsB))) return rc;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};  // function attributes are per device
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        XRB_CUDA(cudaFuncSetAttribute(score_tc_kernel<kFused>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      Cfg<kFused>::kSmemBytes));
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_pairs < sms ? n_pairs : sms;
    score_tc_kernel<kFused><<<grid, kThreads, Cfg<kFused>::kSmemBytes, st>>>(
        mapA, mapB, baseA, baseB, pairs_dev, n_pairs, state_stride, rows, cols, vlow_dev, fa);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}
}  // namespace

int launch_score_tc(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA,
                    const uint8_t *baseB, uint64_t rowsB, int state_stride, Top2State rows, Top2State cols,
                    const int *vlow_dev, cudaStream_t st) {
    FusedArgs fa{};
    return launch_impl<false>(pairs_dev, n_pairs, baseA, rowsA, baseB, rowsB, state_stride, rows, cols, vlow_dev,
                              fa, st);
}

int fused_max_features() { return kFusedMax; }

int launch_match_fused(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA,
                       const uint8_t *baseB, uint64_t rowsB, const int *vlow_dev, const float *dist_tab,
                       float distmax, float ratiomax, int mbm, int max_match, int32_t *counts_dev,
                       uint32_t (*out_dev)[2], int out_stride, cudaStream_t st) {
    FusedArgs fa{dist_tab, distmax, ratiomax, mbm, max_match, counts_dev, out_dev, out_stride};
    Top2State none{nullptr, nullptr};
    return launch_impl<true>(pairs_dev, n_pairs, baseA, rowsA, baseB, rowsB, 0, none, none, vlow_dev, fa, st);
}

}  // namespace xrb
