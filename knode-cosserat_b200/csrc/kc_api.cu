// kc_api.cu — library-level entry points: version, thread-local error text, FMA-pipe micro-benchmark.
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>
#include "kc_common.cuh"

static thread_local char g_err[512] = "";

void kc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int kc_version(void) { return 100; }
extern "C" const char* kc_last_error(void) { return g_err; }

// Eight independent FMA chains per thread, `iters` rounds: 16*iters FLOP per thread.  Used by bench.py to measure the
// FP32 / FP64 pipe peak that the rollout's roofline fraction is quoted against.
template <typename T>
__global__ void __launch_bounds__(256) kc_fma_peak_kernel(int64_t iters, T* out) {
    T a0 = T(threadIdx.x) * T(1e-3), a1 = a0 + T(1), a2 = a0 + T(2), a3 = a0 + T(3);
    T a4 = a0 + T(4), a5 = a0 + T(5), a6 = a0 + T(6), a7 = a0 + T(7);
    const T m = T(0.999999), c = T(1e-6);
#pragma unroll 1
    for (int64_t i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
            a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
        }
    }
    const T s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == T(-1)) out[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true; keeps the chains alive
}

extern "C" int kc_fma_peak(int dtype, int64_t iters, double* flops_host, void* scratch, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(iters > 0 && flops_host, "iters must be > 0 and flops_host non-NULL");
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int blocks = sms * 8, threads = 256;
    if (dtype == KC_F32) kc_fma_peak_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, (float*)scratch);
    else kc_fma_peak_kernel<double><<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, (double*)scratch);
    KC_CHECK_LAUNCH("kc_fma_peak_kernel");
    *flops_host = (double)blocks * threads * (double)iters * 8.0 * 8.0 * 2.0;
    return KC_OK;
}
