// kc_bptt.cu — kc_rollout_bwd: reverse mode through the batched rollout (north-star subsystem 3, SURVEY kernel K5).
// One rod per thread (kc_bptt_core.cuh); the four history-cotangent arrays of a rod live in shared memory; every node
// evaluation of the two adjoint marches emits an MLP sample (x, dL/do) and kc_mlp_bwd's tiled reduction kernels turn the
// B*(T-1)*(N-1)*2 samples into gW1, gb1, gW2, gb2.
#include <cuda_runtime.h>
#include <cstdlib>
#include "kc_bptt_core.cuh"
#include "kc_mlp_coop.cuh"

#include <type_traits>
template <typename T> int kc_pack_mlp(const kc_mlp* mlp, T* Wp, MlpC<T>& M, cudaStream_t st);
// kc_knode_tc.cu: joint adjoint march with the MLP input-VJP on tcgen05
bool kc_knode_tc_eligible(int dtype, const kc_mlp* mlp, int N, int method);
size_t kc_knode_tc_img_bytes();
size_t kc_knode_tc_bwd_scratch_bytes(int N, int64_t B);
int kc_knode_tc_bwd(const RodC<float>& P, const kc_mlp* mlp, int64_t B, int T_, const float* tensions, const float* traj,
                    const float* gtraj, float* gten, float* xs, float* gos, unsigned char* img, unsigned char* scratch,
                    cudaStream_t st);
int kc_check_mlp(const kc_mlp* mlp);

template <typename T, bool DIAG, int IN, int NH>
__global__ void __launch_bounds__(32)
kc_rollout_bwd_kernel(const __grid_constant__ RodC<T> P, const MlpC<T> M, int64_t B, int T_, int rpw,
                      const T* __restrict__ tensions, const T* __restrict__ traj, const T* __restrict__ gtraj,
                      T* __restrict__ gten, T* __restrict__ xs, T* __restrict__ gos, T fd_eps) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = P.N;
    const int64_t b = (int64_t)blockIdx.x * rpw + threadIdx.x;
    if ((int)threadIdx.x >= rpw || b >= B) return;
    T* Hs = reinterpret_cast<T*>(kc_smem) + threadIdx.x;
    const size_t per_rod = (size_t)(T_ - 1) * (N - 1) * 2;
    bptt_rod<T, DIAG, IN, NH, 32>(P, M, traj + (size_t)b * T_ * 25 * N, gtraj + (size_t)b * T_ * 25 * N,
                                  tensions + (size_t)b * T_ * 4, gten ? gten + (size_t)b * T_ * 4 : nullptr, T_, Hs,
                                  IN > 0 ? xs + b * per_rod * (IN > 0 ? IN : 1) : nullptr,
                                  IN > 0 ? gos + b * per_rod * 25 : nullptr, fd_eps);
}

// warp-cooperative variant: one rod per warp, the MLP (and its input VJP) split over the lanes (kc_mlp_coop.cuh)
constexpr int KC_BCOOP_WARPS = 8;   // rods (warps) per CTA sharing one shared-memory copy of the MLP weights
template <typename T, bool DIAG, int IN, int NH>
__global__ void __launch_bounds__(32 * KC_BCOOP_WARPS, 1)
kc_rollout_bwd_coop_kernel(const __grid_constant__ RodC<T> P, MlpCoop<T> M, int64_t B, int T_,
                           const T* __restrict__ tensions, const T* __restrict__ traj, const T* __restrict__ gtraj,
                           T* __restrict__ gten, T* __restrict__ xs, T* __restrict__ gos, T fd_eps, int wc_elems) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    const int N = P.N;
    T* Wsm = reinterpret_cast<T*>(kc_smem);
    if (wc_elems) {   // stage the packed weights once per CTA (see kc_rollout_coop_kernel)
        for (int e = threadIdx.x; e < wc_elems; e += blockDim.x) Wsm[e] = M.Wc[e];
        __syncthreads();
        M.Wc = Wsm;
    }
    const int warp = threadIdx.x >> 5;
    const int64_t b = (int64_t)blockIdx.x * KC_BCOOP_WARPS + warp;
    if (b >= B) return;
    // lane stride 1: the warp holds ONE rod (every lane addresses the same element), so the per-rod arrays take 1/32 of
    // the one-rod-per-lane footprint
    T* Hs = Wsm + wc_elems + (size_t)warp * 4 * NH * (N - 1);
    const size_t per_rod = (size_t)(T_ - 1) * (N - 1) * 2;
    bptt_rod<T, DIAG, IN, NH, 1>(P, M, traj + (size_t)b * T_ * 25 * N, gtraj + (size_t)b * T_ * 25 * N,
                                 tensions + (size_t)b * T_ * 4, gten ? gten + (size_t)b * T_ * 4 : nullptr, T_, Hs,
                                 xs + b * per_rod * IN, gos + b * per_rod * 25, fd_eps);
}

static inline size_t a256(size_t x) { return (x + 255) & ~(size_t)255; }
struct BpttWs { size_t xs, gos, wp, wc, mlpws, tcimg, tcscr, total; int64_t Qs; int64_t mlp_bytes; };
static BpttWs bptt_ws(int dtype, int N, const kc_mlp* mlp, int64_t B, int64_t T_) {
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    BpttWs w{};
    w.Qs = mlp ? B * (T_ - 1) * (N - 1) * 2 : 0;
    size_t off = 0;
    w.xs = off; off += a256((size_t)w.Qs * (mlp ? mlp->in_dim : 0) * sz);
    w.gos = off; off += a256((size_t)w.Qs * 25 * sz);
    w.wp = off; off += mlp ? a256((size_t)mlp->hidden * (((mlp->in_dim + 3) & ~3) + 32) * sz) : 0;
    w.wc = off; off += mlp ? a256((size_t)kc_coop_row(mlp->in_dim) * (size_t)((mlp->hidden + 31) & ~31) * sz) : 0;
    w.mlp_bytes = mlp ? kc_ode_bwd_workspace_bytes(dtype, mlp, w.Qs) : 0;
    w.mlpws = off; off += a256((size_t)w.mlp_bytes);
    w.tcimg = off; off += mlp ? a256(kc_knode_tc_img_bytes()) : 0;
    w.tcscr = off; off += (mlp && dtype == KC_F32) ? a256(kc_knode_tc_bwd_scratch_bytes(N, B)) : 0;
    w.total = off + 256;
    return w;
}

extern "C" int64_t kc_rollout_bwd_workspace_bytes(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_) {
    if (!P || P->N < 2 || B < 0 || T_ < 1 || (dtype != KC_F32 && dtype != KC_F64)) return KC_EINVAL;
    if (mlp && kc_check_mlp(mlp)) return KC_EINVAL;
    return (int64_t)bptt_ws(dtype, P->N, mlp, B, T_).total;
}

template <typename T>
static int bwd_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t B, int64_t T_, const void* tensions,
                     const void* traj, const void* gtraj, void* gten, void* gW1, void* gb1, void* gW2, void* gb2,
                     void* workspace, cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    const int N = P.N, dtype = sizeof(T) == 4 ? KC_F32 : KC_F64;
    const BpttWs w = bptt_ws(dtype, N, mlp, B, T_);
    unsigned char* ws = (unsigned char*)workspace;
    MlpC<T> M{};
    const int in_dim = mlp ? mlp->in_dim : 0;
    if (mlp) {
        int rc = kc_pack_mlp<T>(mlp, (T*)(ws + w.wp), M, st);
        if (rc) return rc;
    }
    const int NH = in_dim == 53 ? 25 : 12;
    const size_t smem = (size_t)4 * NH * (N - 1) * 32 * sizeof(T);
    KC_CHECK_ARG(smem <= 227 * 1024, "N=%d too large for the shared-memory history cotangents (%zu B)", N, smem);
    int rpw = 32;
    {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int64_t slots = (int64_t)sms * 4;
        while (rpw > 8 && (B + rpw / 2 - 1) / (rpw / 2) <= slots) rpw >>= 1;
    }
    const T fd_eps = sizeof(T) == 4 ? T(1e-2) : T(1e-6);
    T* xs = (T*)(ws + w.xs);
    T* gos = (T*)(ws + w.gos);
    bool coop = in_dim != 0 && B <= 8192;
    {
        const char* e = getenv("KC_ROLLOUT_COOP");
        if (e && e[0] == '0') coop = false;
        if (e && e[0] == '1' && in_dim != 0) coop = true;
    }
    bool tc = kc_knode_tc_eligible(dtype, mlp, N, KC_MARCH_EULER) && !getenv("KC_ROLLOUT_COOP");
    {
        const char* e = getenv("KC_ROLLOUT_TC");
        if (e && e[0] == '1' && kc_knode_tc_eligible(dtype, mlp, N, KC_MARCH_EULER)) tc = true;
    }
    int64_t Qred = w.Qs;   // samples handed to the weight-gradient reduction
    if (B > 0 && T_ > 1 && tc) {
        if constexpr (std::is_same<T, float>::value) {
            int rc = kc_knode_tc_bwd(P, mlp, B, (int)T_, (const float*)tensions, (const float*)traj, (const float*)gtraj,
                                     (float*)gten, xs, gos, ws + w.tcimg, ws + w.tcscr, st);
            if (rc) return rc;
        }
        Qred = B * (T_ - 1) * (N - 1);   // the joint adjoint march emits ONE (x, dL/do) sample per node
    } else if (B > 0 && T_ > 1 && coop) {
        MlpCoop<T> MC;
        static_cast<MlpC<T>&>(MC) = M;
        MC.Hp = (mlp->hidden + 31) & ~31;
        T* Wc = (T*)(ws + w.wc);
        MC.Wc = Wc;
        kc_pack_mlp_coop_kernel<T><<<64, 256, 0, st>>>((const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Wc, in_dim,
                                                      (in_dim + 3) & ~3, mlp->hidden, MC.Hp);
        KC_CHECK_LAUNCH("kc_pack_mlp_coop_kernel");
#define KC_LAUNCH_BWDC(D, I, H)                                                                                        \
    do {                                                                                                               \
        auto kern = kc_rollout_bwd_coop_kernel<T, D, I, H>;                                                            \
        const size_t state_b = (size_t)KC_BCOOP_WARPS * 4 * H * (N - 1) * sizeof(T);                                   \
        const size_t wc_b = (size_t)kc_coop_row(I) * MC.Hp * sizeof(T);                                                \
        const int wc_elems = wc_b + state_b <= 200 * 1024 ? (int)(wc_b / sizeof(T)) : 0;                               \
        const size_t csmem = state_b + (size_t)wc_elems * sizeof(T);                                                   \
        if (csmem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)csmem);    \
        kern<<<(unsigned)((B + KC_BCOOP_WARPS - 1) / KC_BCOOP_WARPS), 32 * KC_BCOOP_WARPS, csmem, st>>>(               \
            P, MC, B, (int)T_, (const T*)tensions, (const T*)traj, (const T*)gtraj, (T*)gten, xs, gos, fd_eps,         \
            wc_elems);                                                                                                 \
    } while (0)
        if (P.diag) { if (in_dim == 28) KC_LAUNCH_BWDC(true, 28, 12); else KC_LAUNCH_BWDC(true, 53, 25); }
        else { if (in_dim == 28) KC_LAUNCH_BWDC(false, 28, 12); else KC_LAUNCH_BWDC(false, 53, 25); }
#undef KC_LAUNCH_BWDC
        KC_CHECK_LAUNCH("kc_rollout_bwd_coop_kernel");
    } else if (B > 0 && T_ > 1) {
        const unsigned grid = (unsigned)((B + rpw - 1) / rpw);
#define KC_LAUNCH_BWD(D, I, H)                                                                                         \
    do {                                                                                                               \
        auto kern = kc_rollout_bwd_kernel<T, D, I, H>;                                                                 \
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        kern<<<grid, 32, smem, st>>>(P, M, B, (int)T_, rpw, (const T*)tensions, (const T*)traj, (const T*)gtraj,       \
                                     (T*)gten, xs, gos, fd_eps);                                                       \
    } while (0)
        if (P.diag) {
            if (in_dim == 0) KC_LAUNCH_BWD(true, 0, 12);
            else if (in_dim == 28) KC_LAUNCH_BWD(true, 28, 12);
            else KC_LAUNCH_BWD(true, 53, 25);
        } else {
            if (in_dim == 0) KC_LAUNCH_BWD(false, 0, 12);
            else if (in_dim == 28) KC_LAUNCH_BWD(false, 28, 12);
            else KC_LAUNCH_BWD(false, 53, 25);
        }
#undef KC_LAUNCH_BWD
        KC_CHECK_LAUNCH("kc_rollout_bwd_kernel");
    }
    if (mlp && (gW1 || gb1 || gW2 || gb2))
        return kc_mlp_bwd(dtype, mlp, (B > 0 && T_ > 1) ? Qred : 0, xs, gos, nullptr, gW1, gb1, gW2, gb2, ws + w.mlpws,
                          w.mlp_bytes, (void*)st);
    return KC_OK;
}

extern "C" int kc_rollout_bwd(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_,
                              const void* tensions, const void* traj, const void* g_traj, void* g_tensions, void* gW1,
                              void* gb1, void* gW2, void* gb2, void* workspace, int64_t workspace_bytes, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(B >= 0 && T_ >= 1, "B must be >= 0 and T >= 1");
    KC_CHECK_ARG(B == 0 || (tensions && traj && g_traj && workspace), "NULL tensions/traj/g_traj/workspace");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    const int64_t need = kc_rollout_bwd_workspace_bytes(dtype, P, mlp, B, T_);
    if (workspace_bytes < need) {
        kc_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return KC_ENOSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        return bwd_typed<float>(P, mlp, B, T_, tensions, traj, g_traj, g_tensions, gW1, gb1, gW2, gb2, workspace, st);
    return bwd_typed<double>(P, mlp, B, T_, tensions, traj, g_traj, g_tensions, gW1, gb1, gW2, gb2, workspace, st);
}

// ---------------------------------------------------------------------------------------------------------------------
// Loss of a rollout against a target trajectory and its cotangent, fused (north-star C3(ii): the training loss of
// physics_train.py:345-352 applied to a ROLLOUT instead of a teacher-forced step): for every trajectory b, time index
// t = 1..T-1 and key node k:  MSE(p) + MSE(n,m,q,w) + MSE(euler(h)) at node k and MSE(z) at node k-1, every block
// averaged over its own entries, the key nodes and the T-1 steps, times `scale`.
// -> *loss (double) and g_traj[B][T][25][N] = d loss / d traj (dense; zero where the loss does not look).
struct KeyIdx { int32_t k[64]; };
constexpr int KC_LOSS_BLOCKS = 1024;
__device__ double kc_loss_partials[KC_LOSS_BLOCKS];

template <typename T>
__global__ void __launch_bounds__(256)
kc_rollout_loss_kernel(int64_t B, int T_, int N, int K, KeyIdx key, const T* __restrict__ traj, const T* __restrict__ target,
                       double scale, T* __restrict__ g_traj) {
    __shared__ double red[8];
    const int64_t total = B * (T_ - 1) * K;
    const T S = T(T_ - 1);
    const T wp = T(scale) / (T(3 * K) * S), wf = T(scale) / (T(12 * K) * S), wz = T(scale) / (T(6 * K) * S);
    double acc = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int kk = (int)(i % K);
        const int64_t bt = i / K;
        const int t = 1 + (int)(bt % (T_ - 1));
        const int64_t b = bt / (T_ - 1);
        const int node = key.k[kk];
        const size_t base = ((size_t)b * T_ + t) * 25 * N;
        const T* pr = traj + base;
        const T* tg = target + base;
        T* g = g_traj + base;
        T a = T(0);
#pragma unroll
        for (int r = 0; r < 3; ++r) { const T e = pr[r * N + node] - tg[r * N + node]; a += wp * e * e; g[r * N + node] = T(2) * wp * e; }
#pragma unroll
        for (int r = 7; r < 19; ++r) { const T e = pr[r * N + node] - tg[r * N + node]; a += wf * e * e; g[r * N + node] = T(2) * wf * e; }
#pragma unroll
        for (int c = 19; c < 25; ++c) { const T e = pr[c * N + node - 1] - tg[c * N + node - 1]; a += wz * e * e; g[c * N + node - 1] = T(2) * wz * e; }
        T qp[4], qt[4], ep[3], et[3], ge[3], gq[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) { qp[r] = pr[(3 + r) * N + node]; qt[r] = tg[(3 + r) * N + node]; }
        quat_to_euler(qp, ep);
        quat_to_euler(qt, et);
#pragma unroll
        for (int r = 0; r < 3; ++r) { const T e = ep[r] - et[r]; a += wp * e * e; ge[r] = T(2) * wp * e; }
        quat_to_euler_vjp(qp, ge, gq);
#pragma unroll
        for (int r = 0; r < 4; ++r) g[(3 + r) * N + node] = gq[r];
        acc += (double)a;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w];
        kc_loss_partials[blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) kc_loss_sum_kernel(int n, double* __restrict__ loss) {   // fixed tree: bitwise reproducible
    __shared__ double s[256];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) a += kc_loss_partials[i];
    s[threadIdx.x] = a;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if ((int)threadIdx.x < w) s[threadIdx.x] += s[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *loss = s[0];
}

extern "C" int kc_rollout_loss(int dtype, int64_t B, int64_t T_, int32_t N, int32_t K, const int32_t* key_idx_host,
                               const void* traj, const void* target, double scale, double* loss, void* g_traj, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(B >= 0 && T_ >= 2 && N >= 2, "need B >= 0, T >= 2, N >= 2");
    KC_CHECK_ARG(K >= 1 && K <= 64 && key_idx_host, "1 <= K <= 64 and key_idx_host non-NULL");
    KC_CHECK_ARG(loss && (B == 0 || (traj && target && g_traj)), "NULL pointer");
    KeyIdx key{};
    for (int i = 0; i < K; ++i) {
        KC_CHECK_ARG(key_idx_host[i] >= 1 && key_idx_host[i] < N, "key node indices must lie in [1, N-1]");
        for (int j = 0; j < i; ++j) KC_CHECK_ARG(key_idx_host[j] != key_idx_host[i], "key node indices must be distinct");
        key.k[i] = key_idx_host[i];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    const int64_t total = B * (T_ - 1) * K;
    int grid = (int)((total + 255) / 256);
    if (grid > KC_LOSS_BLOCKS) grid = KC_LOSS_BLOCKS;
    if (grid < 1) grid = 1;
    if (B > 0) {
        cudaError_t e = cudaMemsetAsync(g_traj, 0, (size_t)B * T_ * 25 * N * sz, st);
        if (e != cudaSuccess) { kc_set_error("cudaMemsetAsync: %s", cudaGetErrorString(e)); return KC_ECUDA; }
    }
    if (dtype == KC_F32)
        kc_rollout_loss_kernel<float><<<grid, 256, 0, st>>>(B, (int)T_, N, K, key, (const float*)traj, (const float*)target, scale, (float*)g_traj);
    else
        kc_rollout_loss_kernel<double><<<grid, 256, 0, st>>>(B, (int)T_, N, K, key, (const double*)traj, (const double*)target, scale, (double*)g_traj);
    KC_CHECK_LAUNCH("kc_rollout_loss_kernel");
    kc_loss_sum_kernel<<<1, 256, 0, st>>>(grid, loss);
    KC_CHECK_LAUNCH("kc_loss_sum_kernel");
    return KC_OK;
}
