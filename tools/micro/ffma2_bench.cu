// FFMA vs FFMA2 (packed FP32, sm_100): dependent-issue latency and per-SM throughput.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS, bool PACKED>
__global__ void k(int iters, float* out, long long* cyc) {
    float2 a[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) a[i] = make_float2(threadIdx.x * 1e-3f + i, threadIdx.x * 2e-3f + i);
    const float2 m = make_float2(0.999999f, 0.999998f), c = make_float2(1e-6f, 2e-6f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < CHAINS; ++i) {
                if (PACKED) a[i] = __ffma2_rn(a[i], m, c);
                else { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a[i].x + a[i].y;
    if (s == -1.f) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int CHAINS, bool PACKED>
void run(const char* name, int blocks, int threads) {
    float* out; long long* cyc;
    cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<CHAINS, PACKED><<<blocks, threads>>>(iters, out, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CHAINS, PACKED><<<blocks, threads>>>(iters, out, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double fma_per_thread = (double)iters * 8 * CHAINS * 2;   // scalar FMAs (2 per float2)
    double instr_per_warp = PACKED ? fma_per_thread / 2 : fma_per_thread;
    printf("%-34s blocks %4d thr %4d: %.3f ms, %.1f TFLOP/s, %.2f cycles per warp-instruction (warp 0)\n", name, blocks,
           threads, ms, fma_per_thread * 2 * blocks * threads / (ms * 1e-3) / 1e12, (double)c / instr_per_warp);
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    // latency: one warp per SM sub-partition (4 warps per SM), one chain
    run<1, false>("latency FFMA  (1 chain, 1 warp/SMSP)", sms, 128);
    run<1, true >("latency FFMA2 (1 chain, 1 warp/SMSP)", sms, 128);
    run<2, false>("FFMA  2 chains, 1 warp/SMSP", sms, 128);
    run<2, true >("FFMA2 2 chains, 1 warp/SMSP", sms, 128);
    run<4, true >("FFMA2 4 chains, 1 warp/SMSP", sms, 128);
    // throughput: 8 chains, many warps
    run<8, false>("throughput FFMA  (8 chains)", sms * 8, 256);
    run<8, true >("throughput FFMA2 (8 chains)", sms * 8, 256);
    return 0;
}
