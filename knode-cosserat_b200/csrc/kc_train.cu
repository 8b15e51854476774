// kc_train.cu — teacher-forced KNODE training step, ODE reverse mode, Adam + clamp.
#include <cuda_runtime.h>
#include "kc_rod.cuh"

// torch.optim.Adam semantics (L2 weight decay added to the gradient, bias-corrected moments,
// denom = sqrt(v)/sqrt(1-beta2^t) + eps) followed by the reference's clamp(min=0) of Linear weights
// (physics_train.py:199,296-304).
template <typename T>
__global__ void kc_adam_clamp_kernel(int64_t n, T* __restrict__ p, const T* __restrict__ g, T* __restrict__ m,
                                     T* __restrict__ v, T lr, T b1, T b2, T eps, T wd, T bc1, T bc2s, int clamp) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    T gi = g[i] + wd * p[i];
    const T mi = b1 * m[i] + (T(1) - b1) * gi;
    const T vi = b2 * v[i] + (T(1) - b2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const T denom = sqrt(vi) / bc2s + eps;
    T pi = p[i] - (lr / bc1) * (mi / denom);
    if (clamp && pi < T(0)) pi = T(0);
    p[i] = pi;
}

extern "C" int kc_adam_clamp(int dtype, int64_t n, void* param, const void* grad, void* exp_avg, void* exp_avg_sq,
                             int32_t step, double lr, double beta1, double beta2, double eps, double weight_decay,
                             int32_t clamp_min_zero, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(n >= 0 && step >= 1, "n must be >= 0 and step >= 1");
    KC_CHECK_ARG(n == 0 || (param && grad && exp_avg && exp_avg_sq), "NULL data pointer");
    if (n == 0) return KC_OK;
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2s = sqrt(1.0 - pow(beta2, (double)step));
    const int threads = 256;
    const unsigned grid = (unsigned)((n + threads - 1) / threads);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        kc_adam_clamp_kernel<float><<<grid, threads, 0, st>>>(n, (float*)param, (const float*)grad, (float*)exp_avg,
                                                             (float*)exp_avg_sq, (float)lr, (float)beta1, (float)beta2,
                                                             (float)eps, (float)weight_decay, (float)bc1, (float)bc2s,
                                                             clamp_min_zero);
    else
        kc_adam_clamp_kernel<double><<<grid, threads, 0, st>>>(n, (double*)param, (const double*)grad, (double*)exp_avg,
                                                              (double*)exp_avg_sq, lr, beta1, beta2, eps, weight_decay,
                                                              bc1, bc2s, clamp_min_zero);
    KC_CHECK_LAUNCH("kc_adam_clamp_kernel");
    return KC_OK;
}

// ---- placeholders until the fused kernels land (next commit) ----
extern "C" int64_t kc_ode_bwd_workspace_bytes(int, const kc_mlp*, int64_t) { return 0; }
extern "C" int kc_ode_bwd(int, const kc_rod_params*, const kc_mlp*, int64_t, const void*, const void*, const void*,
                          const void*, const void*, const void*, void*, void*, void*, void*, void*, void*, void*, void*,
                          void*, int64_t, void*) {
    kc_set_error("kc_ode_bwd: not built yet");
    return KC_EINVAL;
}
extern "C" int64_t kc_train_step_workspace_bytes(int, const kc_mlp*, int64_t, int64_t, int32_t) { return 0; }
extern "C" int kc_train_step(int, const kc_rod_params*, const kc_mlp*, int64_t, int64_t, int32_t, const int32_t*,
                             const void*, const void*, double*, void*, void*, void*, void*, void*, void*, int64_t, void*) {
    kc_set_error("kc_train_step: not built yet");
    return KC_EINVAL;
}
