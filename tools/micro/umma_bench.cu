// umma_bench.cu — cycles per tcgen05.mma (kind::f16, M = 128, K = 16) as a function of N, operands in shared memory (SS) or
// A in TMEM (TS).  One CTA; `reps` back-to-back MMAs on one accumulator, then one commit.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -I../../knode-cosserat_b200/csrc -o umma_bench umma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "kc_umma.cuh"

template <int NACC, int TS> __global__ void __launch_bounds__(128) bench(int N, int reps, long long* out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t slot;
    __shared__ __align__(8) uint64_t bar;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < (128 + 256) * 64 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (warp == 0) umma::tmem_alloc(&slot, 512);
    if (tid == 0) umma::mbar_init(&bar, 1);
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tb = slot;
    if (warp == 0) {
        const uint32_t idesc = umma::make_idesc_bf16(128, N);
        const uint64_t da = umma::make_desc(umma::smem_u32(smem), 128, 512);
        const uint64_t db = umma::make_desc(umma::smem_u32(smem + 16384), 128, 512);
        long long t0 = clock64();
        for (int r = 0; r < reps; r += 16) {
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                if (TS) umma::mma_bf16_ts_w(tb + (u % NACC) * 64, tb + 256 + (u & 7) * 8, db + (u & 1) * 16, idesc, 1u);
                else umma::mma_bf16_w(tb + (u % NACC) * 64, da + (u & 1) * 16, db + (u & 1) * 16, idesc, 1u);
            }
        }
        umma::commit_w(&bar);
        umma::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (tid == 0) out[0] = t1 - t0;
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tb, 512);
}
int main() {
    long long* d; cudaMalloc(&d, 8);
    auto go = [&](auto kern, const char* name, int nacc) {
        for (int N : {32, 64, 128, 256}) {
            if (N * nacc > 256 && nacc > 1) continue;
            const int reps = 1024;
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            kern<<<1, 128, 64 * 1024>>>(N, reps, d);
            long long c; cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
            printf("%s N=%3d nacc=%d: %8lld cycles = %6.1f cycles/MMA (%s)\n", name, N, nacc, c, (double)c / reps, cudaGetErrorString(cudaGetLastError()));
        }
    };
    go(bench<1, 0>, "SS", 1); go(bench<2, 0>, "SS", 2); go(bench<4, 0>, "SS", 4);
    go(bench<1, 1>, "TS", 1); go(bench<2, 1>, "TS", 2); go(bench<4, 1>, "TS", 4);
    return 0;
}
