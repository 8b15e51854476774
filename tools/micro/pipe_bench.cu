// pipe_bench.cu — per-SM throughput of the SIMT instructions the tensor-core epilogues are made of (sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>

template <int OP> __device__ __forceinline__ void step(float& a, float& b, uint32_t& u) {
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a));
    if (OP == 1) { asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u) : "f"(a), "f"(b)); a = __uint_as_float(u); }
    if (OP == 2) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(b));
    if (OP == 3) asm volatile("max.f32 %0, %0, %1;" : "+f"(a) : "f"(b));
    if (OP == 4) asm volatile("lop3.b32 %0, %0, %1, 0x55555555, 0x96;" : "+r"(u) : "r"(__float_as_uint(b)));
    if (OP == 5) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(u) : "r"(__float_as_uint(b)));
    if (OP == 6) asm volatile("add.u32 %0, %0, 0x8000;" : "+r"(u));
    if (OP == 7) asm volatile("{ .reg .pred p; setp.gt.f32 p, %0, 0f00000000; selp.f32 %0, %0, %1, p; }" : "+f"(a) : "f"(b));
    if (OP == 8) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a));
    if (OP == 9) asm volatile("shl.b32 %0, %0, 23;" : "+r"(u));
}
template <int OP> __global__ void k(float* out, int iters) {
    float a[8], b = 0.999f;
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { a[i] = 0.5f + threadIdx.x * 1e-3f + i; u[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) step<OP>(a[i], b, u[i]);
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float(u[i]);
    if (s == 12345.678f) out[0] = s;
}
template <int OP> void run(const char* name, int warps_per_sm) {
    float* d; cudaMalloc(&d, 4);
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<sms, warps_per_sm * 32>>>(d, 16);
    cudaEventRecord(e0);
    k<OP><<<sms, warps_per_sm * 32>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)sms * warps_per_sm * 32 * iters * 64;
    printf("%-28s warps/SM %2d: %8.2f Gop/s  = %6.2f thread-ops/clk/SM at %d MHz nominal\n", name, warps_per_sm, ops / ms * 1e-6,
           ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
    cudaFree(d);
}
int main() {
    for (int w : {8, 16}) {
        run<0>("ex2.approx.ftz.f32 (MUFU)", w);
        run<8>("rcp.approx.ftz.f32 (MUFU)", w);
        run<1>("cvt.rn.bf16x2.f32 (F2FP)", w);
        run<2>("fma.rn.f32", w);
        run<3>("max.f32 (FMNMX)", w);
        run<4>("lop3", w);
        run<5>("prmt", w);
        run<6>("add.u32", w);
        run<7>("setp+selp", w);
        run<9>("shl", w);
    }
    return 0;
}
