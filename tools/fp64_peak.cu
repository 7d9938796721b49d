// fp64_peak.cu — micro-benchmark of the two FP64 pipes of one B200: DFMA (vector) and
// DMMA (mma.sync.m8n8k4.f64, the only FP64 tensor shape sm_100a has), plus the larger
// m16n8k16-style shapes where ptxas accepts them.  Prints one JSON object; bench.py reads the copy
// committed under profiles/ as the FP64 roofline denominator (MEASURED_PEAKS.json carries none).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/fp64_peak tools/fp64_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <algorithm>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) acc[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) acc[c] = fma(acc[c], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += acc[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// m8n8k4: A 8x4 (1 double per lane), B 4x8 (1 per lane), C/D 8x8 (2 per lane): 2*8*8*4 = 512 flop
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) c0[c] = threadIdx.x * 1e-3 + c, c1[c] = c;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c0[c]), "+d"(c1[c]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

#ifdef XRB_DMMA_BIG
// m16n8k8 f64 (sm_90+ PTX): A 16x8 (4 per lane), B 8x8 (2 per lane), C 16x8 (4 per lane): 2048 flop
template <int CHAINS>
__global__ void __launch_bounds__(256) k_dmma1688(double *out, int iters, double a, double b) {
    double c0[CHAINS], c1[CHAINS], c2[CHAINS], c3[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) c0[c] = threadIdx.x * 1e-3 + c, c1[c] = c, c2[c] = 1, c3[c] = 2;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c0[c]), "+d"(c1[c]), "+d"(c2[c]), "+d"(c3[c])
                             : "d"(a), "d"(b), "d"(a), "d"(b), "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += c0[c] + c1[c] + c2[c] + c3[c];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
#endif

template <class F>
double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0));
        launch();
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    double *out; CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 8));
    const int iters = 4096;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, sms);
    // DFMA: blocks/SM x chains sweep, keep the best
    double best_dfma = 0; int bd_b = 0, bd_c = 0;
    for (int bps : {1, 2, 4, 8}) {
        auto run = [&](auto kern, int chains, double flop_per_thread_iter, double &best, int &bb, int &bc) {
            const double ms = time_ms([&] { kern<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            const double tf = flop_per_thread_iter * iters * 256.0 * sms * bps / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf, bb = bps, bc = chains;
        };
        run(k_dfma<4>, 4, 2.0 * 4 * 4, best_dfma, bd_b, bd_c);
        run(k_dfma<8>, 8, 2.0 * 4 * 8, best_dfma, bd_b, bd_c);
        run(k_dfma<16>, 16, 2.0 * 4 * 16, best_dfma, bd_b, bd_c);
    }
    printf(", \"dfma_tflops\": %.2f, \"dfma_cfg\": [%d, %d]", best_dfma, bd_b, bd_c);
    double best_mma = 0; int bm_b = 0, bm_c = 0;
    for (int bps : {1, 2, 4, 8}) {
        auto run = [&](auto kern, int chains, double flop_per_warp_iter, double &best, int &bb, int &bc) {
            const double ms = time_ms([&] { kern<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            const double tf = flop_per_warp_iter * iters * 8.0 * sms * bps / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf, bb = bps, bc = chains;
        };
        run(k_dmma884<2>, 2, 512.0 * 4 * 2, best_mma, bm_b, bm_c);
        run(k_dmma884<4>, 4, 512.0 * 4 * 4, best_mma, bm_b, bm_c);
        run(k_dmma884<8>, 8, 512.0 * 4 * 8, best_mma, bm_b, bm_c);
    }
    printf(", \"dmma_m8n8k4_tflops\": %.2f, \"dmma_m8n8k4_cfg\": [%d, %d]", best_mma, bm_b, bm_c);
#ifdef XRB_DMMA_BIG
    double best_big = 0; int bb_b = 0, bb_c = 0;
    for (int bps : {1, 2, 4, 8}) {
        auto run = [&](auto kern, int chains, double flop_per_warp_iter, double &best, int &bb, int &bc) {
            const double ms = time_ms([&] { kern<<<sms * bps, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
            const double tf = flop_per_warp_iter * iters * 8.0 * sms * bps / (ms * 1e-3) / 1e12;
            if (tf > best) best = tf, bb = bps, bc = chains;
        };
        run(k_dmma1688<2>, 2, 2048.0 * 4 * 2, best_big, bb_b, bb_c);
        run(k_dmma1688<4>, 4, 2048.0 * 4 * 4, best_big, bb_b, bb_c);
    }
    printf(", \"dmma_m16n8k8_tflops\": %.2f, \"dmma_m16n8k8_cfg\": [%d, %d]", best_big, bb_b, bb_c);
#endif
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf(", \"sm_clock_khz_attr\": %d, \"how\": \"best of 5 launches, %d x 4 dependent ops per chain, CUDA events\"}\n", clk, iters);
    return 0;
}
