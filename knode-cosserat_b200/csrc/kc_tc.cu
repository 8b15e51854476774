// kc_tc.cu — tensor-core (tcgen05 / TMEM) kernels.  Step 1: a self-test GEMM that pins down the operand layout,
// descriptors and TMEM read-back used by the fused KNODE MLP kernels: D[128][N] = A[128][K] * B[N][K]^T (tf32, fp32 acc).
#include <cuda_runtime.h>
#include "kc_common.cuh"
#include "kc_umma.cuh"
#include <cuda_bf16.h>

__global__ void __launch_bounds__(128)
kc_umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int N, int K,
                        int mn_major) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    float* sA = reinterpret_cast<float*>(smem);                       // 128 x K, K-major interleaved
    float* sB = reinterpret_cast<float*>(smem + (size_t)128 * K * 4);  // N x K
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t ncols = 32;
    while ((int)ncols < N + (mn_major >= 3 ? K / 2 : 0)) ncols <<= 1;
    if (warp == 0) umma::tmem_alloc(&tmem_slot, ncols);
    if (tid == 0) umma::mbar_init(&bar, 1);
    unsigned char* bA = reinterpret_cast<unsigned char*>(sA);
    unsigned char* bB = bA + (mn_major == 2 ? (size_t)128 * K * 2 : (size_t)128 * K * 4);
    if (mn_major >= 3) bB = bA;   // modes 3, 4: A goes to TMEM, only B is staged in shared memory (bf16)
    for (int e = tid; e < 128 * K && mn_major < 3; e += 128) {
        const int r = e / K, k = e - r * K;
        if (mn_major == 2) *reinterpret_cast<__nv_bfloat16*>(bA + umma::mnmajor_off_b16(r, k, K)) = __float2bfloat16(A[e]);
        else *reinterpret_cast<float*>(bA + (mn_major ? umma::mnmajor_off(r, k, K) : umma::kmajor_off(r, k, K))) = A[e];
    }
    for (int e = tid; e < N * K; e += 128) {
        const int r = e / K, k = e - r * K;
        if (mn_major == 4) *reinterpret_cast<__nv_bfloat16*>(bB + umma::kmajor_off_b16(k, r, N)) = __float2bfloat16(B[e]);   // B^T [K][N], K-major image
        else if (mn_major == 3) *reinterpret_cast<__nv_bfloat16*>(bB + umma::kmajor_off_b16(r, k, K)) = __float2bfloat16(B[e]);
        else if (mn_major == 2) *reinterpret_cast<__nv_bfloat16*>(bB + umma::mnmajor_off_b16(r, k, K)) = __float2bfloat16(B[e]);
        else *reinterpret_cast<float*>(bB + (mn_major ? umma::mnmajor_off(r, k, K) : umma::kmajor_off(r, k, K))) = B[e];
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = tmem_slot;
    if (mn_major >= 3) {
        // A[row][k] as packed bf16 pairs in TMEM columns N .. N + K/2 of this thread's lane (row = warp*32 + lane)
        const int row = warp * 32 + lane;
        for (int c0 = 0; c0 < K / 2; c0 += 16) {
            uint32_t v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const __nv_bfloat162 pr = __floats2bfloat162_rn(A[(size_t)row * K + 2 * (c0 + c)], A[(size_t)row * K + 2 * (c0 + c) + 1]);
                v[c] = *reinterpret_cast<const uint32_t*>(&pr);
            }
            umma::st16(tbase + ((uint32_t)(warp * 32) << 16) + N + c0, v);
        }
        umma::wait_st();
        umma::fence_before();
        __syncthreads();
        umma::fence_after();
    }
    if (tid == 0) {
        if (mn_major == 4) {  // as mode 3, but B is the K-major image of B^T ([K rows][N]): read as an MN-major B operand whose
                              // k-groups are (N/8)*128 bytes apart (LBO) and whose n-groups are 128 bytes apart (SBO)
            const uint32_t idesc = umma::make_idesc_bf16(128, N, 0, 1);
            const uint32_t lbo = (uint32_t)(N / 8) * 128;
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t db = umma::make_desc(umma::smem_u32(bB) + kk * 2 * lbo, lbo, 128);
                umma::mma_bf16_ts(tbase, tbase + N + kk * 8, db, idesc, kk > 0 ? 1u : 0u);
            }
        } else if (mn_major == 3) {  // bf16, A from TMEM (8 packed columns per K = 16), B K-major in shared memory
            const uint32_t idesc = umma::make_idesc_bf16(128, N, 0, 0);
            const uint32_t sbo = (uint32_t)(K / 8) * 128;
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t db = umma::make_desc(umma::smem_u32(bB) + kk * 256, 128, sbo);
                umma::mma_bf16_ts(tbase, tbase + N + kk * 8, db, idesc, kk > 0 ? 1u : 0u);
            }
        } else if (mn_major == 2) {  // bf16, both operands MN-major, K = 16 per instruction
            const uint32_t idesc = umma::make_idesc_bf16(128, N, 1, 1);
            const uint32_t sbo = (uint32_t)(K / 8) * 128;
            for (int kk = 0; kk < K / 16; ++kk) {
                const uint64_t da = umma::make_desc(umma::smem_u32(bA) + kk * 256, 128, sbo);
                const uint64_t db = umma::make_desc(umma::smem_u32(bB) + kk * 256, 128, sbo);
                umma::mma_bf16(tbase, da, db, idesc, kk > 0 ? 1u : 0u);
            }
        } else {
            const uint32_t idesc = umma::make_idesc_tf32(128, N, mn_major, mn_major);
            const uint32_t sbo = mn_major ? (uint32_t)(K / 8) * 128 : (uint32_t)(K / 4) * 128;
            const uint32_t kstep = mn_major ? 128 : 256;  // bytes per K = 8
            for (int kk = 0; kk < K / 8; ++kk) {
                const uint64_t da = umma::make_desc(umma::smem_u32(bA) + kk * kstep, 128, sbo);
                const uint64_t db = umma::make_desc(umma::smem_u32(bB) + kk * kstep, 128, sbo);
                umma::mma_tf32(tbase, da, db, idesc, kk > 0 ? 1u : 0u);
            }
        }
        umma::commit(&bar);
    }
    umma::mbar_wait(&bar, 0);
    umma::fence_after();
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        umma::ld32(tbase + ((uint32_t)(warp * 32) << 16) + c0, v);
        umma::wait_ld();
        const int row = warp * 32 + lane;
        for (int c = 0; c < 32 && c0 + c < N; ++c) D[(size_t)row * N + c0 + c] = __uint_as_float(v[c]);
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, ncols);
}

// mode 3: bf16, A operand written to TMEM with tcgen05.st (".ts" MMA), B K-major in shared memory, K % 32 == 0.
// D[128][N] = A[128][K] B[N][K]^T on the tensor cores; N % 16 == 0, 16 <= N <= 256, K % 8 == 0, (128+N)*K*4 <= 200 KB.
extern "C" int kc_umma_selftest(const void* A, const void* B, void* D, int32_t N, int32_t K, int32_t mn_major, void* stream) {
    KC_CHECK_ARG(A && B && D, "NULL pointer");
    KC_CHECK_ARG(N % 16 == 0 && N >= 16 && N <= 256 && K % 8 == 0 && K >= 8 && (mn_major < 2 || K % 32 == 0 || (mn_major == 2 && K % 16 == 0)) && mn_major <= 4, "need N %% 16 == 0 in [16,256], K %% 8 == 0 (16 for bf16)");
    const size_t smem = (size_t)(128 + N) * K * 4;
    KC_CHECK_ARG(smem <= 200 * 1024, "tile too large for shared memory");
    cudaFuncSetAttribute(kc_umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kc_umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const float*)A, (const float*)B, (float*)D, N, K, mn_major);
    KC_CHECK_LAUNCH("kc_umma_selftest_kernel");
    return KC_OK;
}
