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
//   warps 2-9 epilogue: TMEM -> registers (32x32b.x32), VIMNMX3 max-filter against v_low,
//            rare candidates pushed into the per-row / per-column top-2 state
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdlib>
#include <mutex>

#include "match_kernels.cuh"

namespace xrb {

namespace {

constexpr int kTileM = 128;          // descriptors of image 1 per MMA tile (TMEM lanes)
constexpr int kTileN = 256;          // descriptors of image 2 per MMA tile (TMEM columns)
constexpr int kGroupN = 1024;        // resident B descriptors per group
constexpr int kAStages = 3;
constexpr int kATileBytes = kTileM * kDim;    // 16 KB
constexpr int kBGroupBytes = kGroupN * kDim;  // 128 KB
constexpr int kBoxRows = 128;                 // TMA box: 128 rows x 128 bytes
constexpr int kThreads = 32 * 10;
constexpr int kSmemBytes = kBGroupBytes + kAStages * kATileBytes + 1024 /*align*/ + 256 /*barriers*/;

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

struct Bars {
    unsigned long long b_full, b_empty;
    unsigned long long a_full[kAStages], a_empty[kAStages];
    unsigned long long t_full[2], t_empty[2];
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1)
score_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const uint8_t *baseA, const uint8_t *baseB, const PairDesc *__restrict__ pairs,
                int n_pairs, int state_stride, Top2State rows, Top2State cols,
                const int *__restrict__ vlow_ptr) {
    extern __shared__ unsigned char smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024 B alignment
    const uint32_t sB = base;
    const uint32_t sA = base + kBGroupBytes;
    Bars *bars = reinterpret_cast<Bars *>(smem_raw + (base - raw) + kBGroupBytes + kAStages * kATileBytes);
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
            mbar_init(smem_u32(&bars->t_empty[b]), 8);  // one arrival per epilogue warp
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
        const int ew = warp - 2;            // 0..7
        const int quarter = warp & 3;       // TMEM lane quarter this warp may touch (warp id % 4)
        const int half = ew >> 2;           // which 128-column half of the 256-column tile
        const int vlow = *vlow_ptr;
        uint32_t t_it = 0;
        for (int p = blockIdx.x; p < n_pairs; p += gridDim.x) {
            const PairDesc pd = pairs[p];
            if (pd.n1 <= 0 || pd.n2 <= 0) continue;
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
                        const int col_tile = g0 + ns * kTileN + half * 128;
                        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16) + tb * kTileN + half * 128;
#pragma unroll 1
                        for (int ch = 0; ch < 4; ++ch) {
                            int v[32];
                            asm volatile(
                                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                                  "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                                  "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
                                  "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
                                  "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                : "r"(taddr + ch * 32));
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                            if (ch == 3) {  // accumulator fully read: hand the TMEM buffer back
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(smem_u32(&bars->t_empty[tb]));
                            }
                            const int j0 = col_tile + ch * 32;
#pragma unroll
                            for (int g = 0; g < 32; g += 16) {
                                int mx = v[g];
#pragma unroll
                                for (int e = 1; e < 16; ++e) mx = max(mx, v[g + e]);
                                if (mx > vlow && i < pd.n1) {
#pragma unroll
                                    for (int e = 0; e < 16; ++e) {
                                        const int val = v[g + e], j = j0 + g + e;
                                        if (val > vlow && j < pd.n2) {
                                            const unsigned long long hv = (unsigned long long)(unsigned int)val << 32;
                                            push_top2(rows.best + sbase + i, rows.second + sbase + i,
                                                      hv | (0xFFFFFFFFu - row_tie_rank((uint32_t)j)));
                                            push_top2(cols.best + sbase + j, cols.second + sbase + j,
                                                      hv | (0xFFFFFFFFu - (uint32_t)i));
                                        }
                                    }
                                }
                            }
                            __syncwarp();  // reconverge before the next .aligned tcgen05.ld
                        }
                    }
                }
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

int launch_score_tc(const PairDesc *pairs_dev, int n_pairs, const uint8_t *baseA, uint64_t rowsA,
                    const uint8_t *baseB, uint64_t rowsB, int state_stride, Top2State rows, Top2State cols,
                    const int *vlow_dev, cudaStream_t st) {
    if (n_pairs <= 0) return XRB_OK;
    if (((uintptr_t)baseA & 15) || ((uintptr_t)baseB & 15)) {
        set_error("tcgen05 matcher: descriptor blocks must be 16-byte aligned");
        return XRB_ERR_INVALID;
    }
    CUtensorMap mapA, mapB;
    int rc;
    if ((rc = make_map(&mapA, baseA, rowsA))) return rc;
    if ((rc = make_map(&mapB, baseB, rowsB))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        XRB_CUDA(cudaFuncSetAttribute(score_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
        attr_set = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = n_pairs < sms ? n_pairs : sms;
    score_tc_kernel<<<grid, kThreads, kSmemBytes, st>>>(mapA, mapB, baseA, baseB, pairs_dev, n_pairs, state_stride,
                                                        rows, cols, vlow_dev);
    XRB_LAUNCHED();
    XRB_CUDA(cudaGetLastError());
    return XRB_OK;
}

}  // namespace xrb
