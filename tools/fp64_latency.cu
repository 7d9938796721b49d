// fp64_latency.cu — dependent-issue latencies of the operations on the factorisation's critical chain,
// one warp, clock64 around N dependent operations.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_bin/fp64_latency tools/fp64_latency.cu
#include <cuda_runtime.h>
#include <cstdio>
constexpr int N = 2048;
__global__ void k(double *out, long long *cyc, double a, double b) {
    double x = a + threadIdx.x * 1e-9;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, b, a);
    long long t1 = clock64();
    double y = x;
#pragma unroll 16
    for (int i = 0; i < N; ++i) y = rsqrt(y) + a;
    long long t2 = clock64();
    double z = y;
#pragma unroll 16
    for (int i = 0; i < N; ++i) z = __shfl_sync(0xFFFFFFFFu, z, (threadIdx.x + 1) & 31);
    long long t3 = clock64();
    double w = z;
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        double r;
        asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(w));
        w = r + a;
    }
    long long t4 = clock64();
    double c0 = w, c1 = 0.0;
#pragma unroll 16
    for (int i = 0; i < N; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
    long long t5 = clock64();
    double v = c0 + c1;
#pragma unroll 16
    for (int i = 0; i < N; ++i) v = sqrt(v) + a;
    long long t6 = clock64();
    double q = v;
#pragma unroll 16
    for (int i = 0; i < N; ++i) q = a / q + b;
    long long t7 = clock64();
    if (threadIdx.x == 0) {
        cyc[0] = t1 - t0, cyc[1] = t2 - t1, cyc[2] = t3 - t2, cyc[3] = t4 - t3, cyc[4] = t5 - t4, cyc[5] = t6 - t5, cyc[6] = t7 - t6;
    }
    out[threadIdx.x] = q;
}
int main() {
    double *out; long long *cyc, h[7];
    cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 7 * 8);
    for (int rep = 0; rep < 2; ++rep) { k<<<1, 32>>>(out, cyc, 1.000001, 0.999999); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 7 * 8, cudaMemcpyDeviceToHost);
    printf("{\"dependent_cycles\": {\"dfma\": %.1f, \"rsqrt_f64_plus_dadd\": %.1f, \"shfl_f64\": %.1f, \"rcp_approx_f64_plus_dadd\": %.1f, \"dmma_m8n8k4\": %.1f, \"sqrt_f64_plus_dadd\": %.1f, \"ddiv_plus_dadd\": %.1f}}\n",
           h[0] / (double)N, h[1] / (double)N, h[2] / (double)N, h[3] / (double)N, h[4] / (double)N, h[5] / (double)N, h[6] / (double)N);
    return 0;
}
