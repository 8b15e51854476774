// kc_ode.cu — batched node evaluations, shooting marches and teacher-forced segment steps in the reference's own
// tensor layouts (these entry points are what the drop-in CosseratRodTorch / CosseratRod methods call).
//   kc_ode_fwd     : ODE_parallel / ODE            (cosserat_ode_torch.py:217-322, :137-214; cosserat_ode.py:114-186)
//   kc_march       : getResidualEuler / RK4        (cosserat_ode.py:188-255; cosserat_ode_torch.py:325-367)
//   kc_segment_fwd : (parallel)GetNextSegmentEuler (cosserat_ode_torch.py:370-437)
// One sample (or one rod) per thread, everything in registers.  These calls are HBM/latency bound (≈300 B and
// ≈450 FLOP per sample without the MLP): inputs are staged through shared memory so that global accesses are
// coalesced even though the reference layouts are array-of-structures.
#include <cuda_runtime.h>
#include "kc_rod.cuh"

template <typename T> int kc_pack_mlp(const kc_mlp* mlp, T* Wp, MlpC<T>& M, cudaStream_t st);
int kc_check_mlp(const kc_mlp* mlp);

// ---------------------------------------------------------------------------------------------------------------
// kc_ode_fwd
// ---------------------------------------------------------------------------------------------------------------
constexpr int ODE_THREADS = 128;

// Cooperative, coalesced copy of `rows` consecutive records of `width` values between global AoS and a shared tile
// [width][ODE_THREADS+1]-style transposed layout is overkill here: records are short, so we copy the contiguous span
// linearly (coalesced) into shared memory and let each thread read its own record (stride `width`, odd => no conflicts
// for 19; 6 and 3 give 2-way/…-way conflicts on a few loads, negligible next to the arithmetic).
template <typename T>
__device__ __forceinline__ void stage_in(T* sm, const T* __restrict__ g, int64_t first, int64_t count, int width) {
    const int64_t n = count * width;
    const T* src = g + first * width;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) sm[i] = src[i];
}
template <typename T>
__device__ __forceinline__ void stage_out(const T* sm, T* __restrict__ g, int64_t first, int64_t count, int width) {
    const int64_t n = count * width;
    T* dst = g + first * width;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) dst[i] = sm[i];
}

template <typename T, bool DIAG, int IN>
__global__ void __launch_bounds__(ODE_THREADS)
kc_ode_fwd_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t Q, const T* __restrict__ y,
                  const T* __restrict__ yh, const T* __restrict__ zh, const T* __restrict__ tf, T* __restrict__ ys,
                  T* __restrict__ z) {
    __shared__ T s_y[ODE_THREADS * 19], s_yh[ODE_THREADS * 19], s_zh[ODE_THREADS * 6], s_tf[ODE_THREADS * 3];
    const int64_t first = (int64_t)blockIdx.x * ODE_THREADS;
    const int64_t count = min((int64_t)ODE_THREADS, Q - first);
    stage_in(s_y, y, first, count, 19);
    stage_in(s_yh, yh, first, count, 19);
    stage_in(s_zh, zh, first, count, 6);
    stage_in(s_tf, tf, first, count, 3);
    __syncthreads();
    T ry[19], rh[25], rt[3], rys[19], rz[6];
    const int i = threadIdx.x;
    if (i < count) {
#pragma unroll
        for (int k = 0; k < 19; ++k) { ry[k] = s_y[i * 19 + k]; rh[k] = s_yh[i * 19 + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) rh[19 + k] = s_zh[i * 6 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) rt[k] = s_tf[i * 3 + k];
        node_eval<T, DIAG, IN, 25>(P, M, ry, rh, rt, rys, rz);
    }
    __syncthreads();
    if (i < count) {
#pragma unroll
        for (int k = 0; k < 19; ++k) s_y[i * 19 + k] = rys[k];
#pragma unroll
        for (int k = 0; k < 6; ++k) s_zh[i * 6 + k] = rz[k];
    }
    __syncthreads();
    stage_out(s_y, ys, first, count, 19);
    stage_out(s_zh, z, first, count, 6);
}

// Scratch for the packed MLP of the layout-API calls: these entry points take no workspace (the reference methods they
// replace are called with nothing but tensors), so the packed copy lives in a small per-process device buffer that is
// grown on demand (never shrunk); the pack kernel runs on the caller's stream before the consumer.
static void* g_wp_buf = nullptr;
static size_t g_wp_bytes = 0;
static void* wp_scratch(size_t bytes) {
    if (bytes > g_wp_bytes) {
        if (g_wp_buf) cudaFree(g_wp_buf);  // synchronises: safe w.r.t. earlier consumers
        if (cudaMalloc(&g_wp_buf, bytes) != cudaSuccess) { g_wp_buf = nullptr; g_wp_bytes = 0; return nullptr; }
        g_wp_bytes = bytes;
    }
    return g_wp_buf;
}
template <typename T>
static int prep_mlp(const kc_mlp* mlp, MlpC<T>& M, cudaStream_t st) {
    M = MlpC<T>{};
    if (!mlp) return KC_OK;
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    const size_t bytes = (size_t)mlp->hidden * (((mlp->in_dim + 3) & ~3) + 32) * sizeof(T);
    void* buf = wp_scratch(bytes);
    if (!buf) { kc_set_error("cudaMalloc of %zu B for the packed MLP failed", bytes); return KC_ECUDA; }
    return kc_pack_mlp<T>(mlp, (T*)buf, M, st);
}

#define KC_DISPATCH_IN(D, in_dim, CALL)               \
    do {                                              \
        if (in_dim == 0) { CALL(D, 0); }              \
        else if (in_dim == 28) { CALL(D, 28); }       \
        else { CALL(D, 53); }                         \
    } while (0)
#define KC_DISPATCH(diag, in_dim, CALL)                                   \
    do {                                                                  \
        if (diag) KC_DISPATCH_IN(true, in_dim, CALL);                     \
        else KC_DISPATCH_IN(false, in_dim, CALL);                         \
    } while (0)

// Tensor-core KNODE forward (fp32, 28 inputs, hidden <= 512, >= 4096 samples): the physics of every sample on the SIMT
// pipes (kc_ode_prep_kernel -> X[Q][32] = [y; z; tf], PHYS[Q][25] = [ys; z]), then the MLP contraction and the residual add
// on tcgen05 (kc_train_tc_kernel<1>) writing ys[Q][19] and z[Q][6] directly.
void* kc_tc_scratch(size_t bytes);
bool kc_tc_forward_ok(const kc_mlp* mlp, int64_t Q);
int kc_train_tc_grid(int64_t Q);
int kc_tc_launch_mode(int mode, const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                      const float* TGT, float* W1hl, float* W2c_, float* partial, int64_t NP, double* loss_part,
                      float* pred_out, const float* dOin, int grid, cudaStream_t st, int xpitch = 32, int dopitch = 32);

template <typename T, bool DIAG>
__global__ void __launch_bounds__(ODE_THREADS)
kc_ode_prep_kernel(const __grid_constant__ RodC<T> P, int64_t Q, const T* __restrict__ y, const T* __restrict__ yh,
                   const T* __restrict__ zh, const T* __restrict__ tf, T* __restrict__ X, T* __restrict__ PHYS) {
    __shared__ T s_a[ODE_THREADS * 32], s_b[ODE_THREADS * 25], s_zh[ODE_THREADS * 6], s_tf[ODE_THREADS * 3];
    const int64_t first = (int64_t)blockIdx.x * ODE_THREADS;
    const int64_t count = min((int64_t)ODE_THREADS, Q - first);
    stage_in(s_a, y, first, count, 19);
    stage_in(s_b, yh, first, count, 19);
    stage_in(s_zh, zh, first, count, 6);
    stage_in(s_tf, tf, first, count, 3);
    __syncthreads();
    T ry[19], rh[25], rt[3], rys[19], rz[6];
    const int i = threadIdx.x;
    if (i < count) {
#pragma unroll
        for (int k = 0; k < 19; ++k) { ry[k] = s_a[i * 19 + k]; rh[k] = s_b[i * 19 + k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) rh[19 + k] = s_zh[i * 6 + k];
#pragma unroll
        for (int k = 0; k < 3; ++k) rt[k] = s_tf[i * 3 + k];
        rod_ode<T, DIAG>(P, ry, rh + 13, rh + 16, rh + 19, rh + 22, rt, rys, rz);
    }
    __syncthreads();
    if (i < count) {
#pragma unroll
        for (int k = 0; k < 19; ++k) { s_a[i * 32 + k] = ry[k]; s_b[i * 25 + k] = rys[k]; }
#pragma unroll
        for (int k = 0; k < 6; ++k) { s_a[i * 32 + 19 + k] = rz[k]; s_b[i * 25 + 19 + k] = rz[k]; }
#pragma unroll
        for (int k = 0; k < 3; ++k) s_a[i * 32 + 25 + k] = rt[k];
#pragma unroll
        for (int k = 28; k < 32; ++k) s_a[i * 32 + k] = T(0);
    }
    __syncthreads();
    stage_out(s_a, X, first, count, 32);
    stage_out(s_b, PHYS, first, count, 25);
}

template <typename T>
static int ode_fwd_tc(const RodC<T>&, const kc_mlp*, int64_t, const void*, const void*, const void*, const void*, void*, void*,
                      cudaStream_t) { return 1; }
template <>
int ode_fwd_tc<float>(const RodC<float>& P, const kc_mlp* mlp, int64_t Q, const void* y, const void* yh, const void* zh,
                      const void* tf, void* ys, void* z, cudaStream_t st) {
    if (!kc_tc_forward_ok(mlp, Q)) return 1;   // 1: not taken
    if (int rc = kc_check_mlp(mlp)) return rc;
    const size_t tcw_b = (size_t)4 * (2 * 128 * 32 * 4 + 32768 + 16384);
    const size_t xb = ((size_t)Q * 32 * 4 + 255) & ~(size_t)255, pb = ((size_t)Q * 25 * 4 + 255) & ~(size_t)255;
    unsigned char* buf = (unsigned char*)kc_tc_scratch(tcw_b + 2048 + xb + pb);
    if (!buf) { kc_set_error("cudaMalloc of the tensor-core scratch failed"); return KC_ECUDA; }
    float* tcw = (float*)buf;
    double* lossp = (double*)(buf + tcw_b);
    float* X = (float*)(buf + tcw_b + 2048);
    float* PHYS = (float*)(buf + tcw_b + 2048 + xb);
    const unsigned grid = (unsigned)((Q + ODE_THREADS - 1) / ODE_THREADS);
    if (P.diag)
        kc_ode_prep_kernel<float, true><<<grid, ODE_THREADS, 0, st>>>(P, Q, (const float*)y, (const float*)yh, (const float*)zh,
                                                                      (const float*)tf, X, PHYS);
    else
        kc_ode_prep_kernel<float, false><<<grid, ODE_THREADS, 0, st>>>(P, Q, (const float*)y, (const float*)yh, (const float*)zh,
                                                                       (const float*)tf, X, PHYS);
    KC_CHECK_LAUNCH("kc_ode_prep_kernel");
    // ys += o[0:19], z += o[19:25] (cosserat_ode_torch.py:199-212): scale 1 on every row; outputs split (z via `partial`)
    return kc_tc_launch_mode(1, mlp, 1.f, Q, 2, 1, X, PHYS, nullptr, tcw, tcw + 4 * 2 * 128 * 32, (float*)z, 0, lossp,
                             (float*)ys, nullptr, kc_train_tc_grid(Q), st, 32);
}

template <typename T>
static int ode_fwd_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t Q, const void* y, const void* yh,
                         const void* zh, const void* tf, void* ys, void* z, cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    {
        const int rc = ode_fwd_tc<T>(P, mlp, Q, y, yh, zh, tf, ys, z, st);
        if (rc <= 0) return rc;
    }
    MlpC<T> M;
    int rc = prep_mlp<T>(mlp, M, st);
    if (rc) return rc;
    if (Q == 0) return KC_OK;
    const unsigned grid = (unsigned)((Q + ODE_THREADS - 1) / ODE_THREADS);
    const int in_dim = mlp ? mlp->in_dim : 0;
#define CALL(D, I) kc_ode_fwd_kernel<T, D, I><<<grid, ODE_THREADS, 0, st>>>(P, M, Q, (const T*)y, (const T*)yh, (const T*)zh, (const T*)tf, (T*)ys, (T*)z)
    KC_DISPATCH(P.diag, in_dim, CALL);
#undef CALL
    KC_CHECK_LAUNCH("kc_ode_fwd_kernel");
    return KC_OK;
}

extern "C" int kc_ode_fwd(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t Q, const void* y,
                          const void* yh, const void* zh, const void* tf, void* ys, void* z, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(Q >= 0, "Q must be >= 0");
    KC_CHECK_ARG(Q == 0 || (y && yh && zh && tf && ys && z), "NULL data pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32) return ode_fwd_typed<float>(P, mlp, Q, y, yh, zh, tf, ys, z, st);
    return ode_fwd_typed<double>(P, mlp, Q, y, yh, zh, tf, ys, z, st);
}

// ---------------------------------------------------------------------------------------------------------------
// kc_march — one rod per thread, reference layout [19][N] / [6][N], in place
// ---------------------------------------------------------------------------------------------------------------
template <typename T, int NH> struct RefHist {
    const T* yh; const T* zh; int N;
    KC_HD void load(int j, T hist[NH]) const {
        if (NH == 12) {
#pragma unroll
            for (int s = 0; s < 6; ++s) { hist[s] = yh[(13 + s) * N + j]; hist[6 + s] = zh[s * N + j]; }
        } else {
#pragma unroll
            for (int r = 0; r < 19; ++r) hist[r % NH] = yh[r * N + j];
#pragma unroll
            for (int c = 0; c < 6; ++c) hist[(19 + c) % NH] = zh[c * N + j];
        }
    }
};
template <typename T> struct RefSink {
    T* y; T* z; int N;
    KC_HD void put(int j, const T v[19]) {
#pragma unroll
        for (int r = 0; r < 19; ++r) y[r * N + j] = v[r];
    }
    KC_HD void putz(int j, const T v[6]) {
#pragma unroll
        for (int c = 0; c < 6; ++c) z[c * N + j] = v[c];
    }
};

template <typename T, bool DIAG, int IN, int NH, int METHOD>
__global__ void kc_march_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t B,
                                const T* __restrict__ G, T* y, T* z, const T* __restrict__ yh,
                                const T* __restrict__ zh, const T* __restrict__ tensions, T* __restrict__ res) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int N = P.N;
    T g[6], tn[4], tf[3], r[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) g[i] = G[b * 6 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) tn[i] = tensions[b * 4 + i];
    tendon_force(P, tn, tf);
    RefHist<T, NH> H{yh + (size_t)b * 19 * N, zh + (size_t)b * 6 * N, N};
    RefSink<T> S{y + (size_t)b * 19 * N, z + (size_t)b * 6 * N, N};
    if (METHOD == KC_MARCH_RK4) rod_march_rk4<T, DIAG, IN, NH>(P, M, g, tf, H, S, r);
    else rod_march<T, DIAG, IN, NH>(P, M, g, tf, H, S, r);
#pragma unroll
    for (int i = 0; i < 6; ++i) res[b * 6 + i] = r[i];
}

template <typename T>
static int march_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int method, int64_t B, const void* G, void* y,
                       void* z, const void* yh, const void* zh, const void* tensions, void* res, cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    MlpC<T> M;
    int rc = prep_mlp<T>(mlp, M, st);
    if (rc) return rc;
    if (B == 0) return KC_OK;
    const int threads = 32;
    const unsigned grid = (unsigned)((B + threads - 1) / threads);
    const int in_dim = mlp ? mlp->in_dim : 0;
#define ARGS P, M, B, (const T*)G, (T*)y, (T*)z, (const T*)yh, (const T*)zh, (const T*)tensions, (T*)res
#define MARCH(D, I, H)                                                                               \
    do {                                                                                             \
        if (method == KC_MARCH_RK4) kc_march_kernel<T, D, I, H, KC_MARCH_RK4><<<grid, threads, 0, st>>>(ARGS);  \
        else kc_march_kernel<T, D, I, H, KC_MARCH_EULER><<<grid, threads, 0, st>>>(ARGS);            \
    } while (0)
    if (P.diag) {
        if (in_dim == 0) MARCH(true, 0, 12);
        else if (in_dim == 28) MARCH(true, 28, 12);
        else MARCH(true, 53, 25);
    } else {
        if (in_dim == 0) MARCH(false, 0, 12);
        else if (in_dim == 28) MARCH(false, 28, 12);
        else MARCH(false, 53, 25);
    }
#undef MARCH
#undef ARGS
    KC_CHECK_LAUNCH("kc_march_kernel");
    return KC_OK;
}

extern "C" int kc_march(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int method, int64_t B, const void* G,
                        void* y, void* z, const void* yh, const void* zh, const void* tensions, void* res, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(method == KC_MARCH_EULER || method == KC_MARCH_RK4, "method must be KC_MARCH_EULER or KC_MARCH_RK4");
    KC_CHECK_ARG(B >= 0, "B must be >= 0");
    KC_CHECK_ARG(B == 0 || (G && y && z && yh && zh && tensions && res), "NULL data pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32) return march_typed<float>(P, mlp, method, B, G, y, z, yh, zh, tensions, res, st);
    return march_typed<double>(P, mlp, method, B, G, y, z, yh, zh, tensions, res, st);
}

// ---------------------------------------------------------------------------------------------------------------
// kc_segment_fwd — one (step, key node) sample per thread
// ---------------------------------------------------------------------------------------------------------------
struct KeyIdx { int32_t k[64]; };

template <typename T, bool DIAG, int IN>
__global__ void kc_segment_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t S, int K,
                                  const __grid_constant__ KeyIdx key, const T* __restrict__ Gs,
                                  const T* __restrict__ yh, const T* __restrict__ zh,
                                  const T* __restrict__ tensions, T* __restrict__ out) {
    const int N = P.N;
    const int per = K > 0 ? K : N - 1;   // samples per step
    const int ocol = K > 0 ? K : N;      // output columns per step
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S * per) return;
    const int64_t s = idx / per;
    const int kk = (int)(idx % per);
    const int j = K > 0 ? key.k[kk] - 1 : kk;   // node whose ODE is evaluated
    const int oc = K > 0 ? kk : kk + 1;          // output column
    const T* g = Gs + (size_t)s * 25 * N;
    T y[19], hist[25], tn[4], tf[3], ys[19], z[6];
#pragma unroll
    for (int r = 0; r < 19; ++r) { y[r] = g[r * N + j]; hist[r] = yh[((size_t)s * 19 + r) * N + j]; }
#pragma unroll
    for (int c = 0; c < 6; ++c) hist[19 + c] = zh[((size_t)s * 6 + c) * N + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) tn[i] = tensions[s * 4 + i];
    tendon_force(P, tn, tf);
    node_eval<T, DIAG, IN, 25>(P, M, y, hist, tf, ys, z);
    T* o = out + (size_t)s * 25 * ocol;
#pragma unroll
    for (int r = 0; r < 19; ++r) o[r * ocol + oc] = y[r] + P.ds * ys[r];
#pragma unroll
    for (int c = 0; c < 6; ++c) o[(19 + c) * ocol + oc] = z[c];
    if (K == 0 && kk == 0) {  // full_rod[:, 0] = G[:, 0] (cosserat_ode_torch.py:379)
#pragma unroll
        for (int r = 0; r < 25; ++r) o[r * ocol] = g[r * N];
    }
}

template <typename T>
static int segment_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t S, int K, const int32_t* key_host,
                         const void* Gs, const void* yh, const void* zh, const void* tensions, void* out, cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    MlpC<T> M;
    int rc = prep_mlp<T>(mlp, M, st);
    if (rc) return rc;
    KeyIdx key{};
    for (int i = 0; i < K; ++i) {
        KC_CHECK_ARG(key_host[i] >= 1 && key_host[i] <= P.N - 1, "key index %d out of range [1, N-1]", key_host[i]);
        key.k[i] = key_host[i];
    }
    const int per = K > 0 ? K : P.N - 1;
    const int64_t total = S * per;
    if (total == 0) return KC_OK;
    const int threads = 128;
    const unsigned grid = (unsigned)((total + threads - 1) / threads);
    const int in_dim = mlp ? mlp->in_dim : 0;
#define CALL(D, I) kc_segment_kernel<T, D, I><<<grid, threads, 0, st>>>(P, M, S, K, key, (const T*)Gs, (const T*)yh, (const T*)zh, (const T*)tensions, (T*)out)
    KC_DISPATCH(P.diag, in_dim, CALL);
#undef CALL
    KC_CHECK_LAUNCH("kc_segment_kernel");
    return KC_OK;
}

extern "C" int kc_segment_fwd(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t S, int32_t K,
                              const int32_t* key_idx_host, const void* Gs, const void* yh, const void* zh,
                              const void* tensions, void* out, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(S >= 0 && K >= 0 && K <= 64, "S must be >= 0 and 0 <= K <= 64");
    KC_CHECK_ARG(K == 0 || key_idx_host, "key_idx_host is NULL");
    KC_CHECK_ARG(S == 0 || (Gs && yh && zh && tensions && out), "NULL data pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32) return segment_typed<float>(P, mlp, S, K, key_idx_host, Gs, yh, zh, tensions, out, st);
    return segment_typed<double>(P, mlp, S, K, key_idx_host, Gs, yh, zh, tensions, out, st);
}
