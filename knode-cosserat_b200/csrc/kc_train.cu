// kc_train.cu — teacher-forced KNODE training step (forward + 4-term loss + reverse mode to the MLP weights), the
// reverse mode of the batched node ODE, and Adam + clamp.
//
// Reference: physics_train.py:313-368 (fast path) == :215-267 (slow path at its key nodes) == train_segment.py:140-185;
// loss terms physics_train.py:345-352 with Utils/transformations.py:3-31.  The physics is a constant offset of the
// prediction, so the weight gradient needs no physics adjoint:
//     pred[0:19] = y + ds*(ys_phys + o[0:19]),  pred[19:25] = z_phys + o[19:25],  o = W2 ELU(W1 x + b1) + b2.
//
// Kernel plan of one step (SIMT FP32/FP64; Q = B*(T-1)*K samples):
//   1. kc_train_prep_kernel   : gather (b,t,key) from traj, BDF2 history, physics ODE -> X[Q][XP], PHYS[Q][25], TGT[Q][25]
//   2. kc_train_fwd_kernel    : one sample per thread, MLP forward with broadcast weights, loss terms, dL/do -> dO[Q][32]
//   3. kc_train_bwd_kernel    : CTA = (32 hidden units) x (a slice of the samples); recomputes z1/ELU for its units from X,
//                               forms dz1 = (W2^T dO) * ELU'(z1), accumulates gW1, gb1, gW2 (and gb2) in registers as
//                               shared-memory-tiled outer products; writes per-slice partials
//   4. kc_train_reduce_kernel : deterministic sum of the partials into gW1, gb1, gW2, gb2 and of the loss partials
#include <cuda_runtime.h>
#include <algorithm>
#include <cstdlib>
#include "kc_rod.cuh"
#include "kc_adjoint.cuh"
#include "kc_train_prep.cuh"

template <typename T> int kc_pack_mlp(const kc_mlp* mlp, T* Wp, MlpC<T>& M, cudaStream_t st);
int kc_train_tc_grid(int64_t Q);
int kc_train_tc2_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS, const float* TGT,
                        unsigned char* img, float* partial, int64_t NP, double* loss_part, float* pred_out, int grid,
                        cudaStream_t st);
int kc_train_tc3_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, float* X, float* PHYS, float* TGT,
                        unsigned char* img, float* partial, int64_t NP, double* loss_part, float* pred_out, int grid,
                        cudaStream_t st, const RodC<float>* RP, const KeyIdx64* key, const float* traj, const float* controls,
                        int64_t B);
int kc_train_tc_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                       const float* TGT, float* W1hl, float* W2c, float* partial, int64_t NP, double* loss_part,
                       float* pred_out, int grid, cudaStream_t st);
template <typename T> struct kc_is_float { static constexpr bool value = false; };
template <> struct kc_is_float<float> { static constexpr bool value = true; };
int kc_check_mlp(const kc_mlp* mlp);

static inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// ---------------------------------------------------------------------------------------------------------------
// 1. prep
// ---------------------------------------------------------------------------------------------------------------
// thread per sample, straight from global memory (any T): the fallback when a trajectory does not fit shared memory
template <typename T, bool DIAG, int IN>
__global__ void __launch_bounds__(128)
kc_train_prep_kernel(const __grid_constant__ RodC<T> P, const __grid_constant__ KeyIdx64 key, int64_t B, int T_, int K,
                     const T* __restrict__ traj, const T* __restrict__ controls, T* __restrict__ X, int XP,
                     T* __restrict__ PHYS, T* __restrict__ TGT) {
    const int64_t Q = B * (int64_t)(T_ - 1) * K;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    const int N = P.N;
    const int kk = (int)(q % K);
    const int64_t bt = q / K;
    const int t = (int)(bt % (T_ - 1));
    const int64_t b = bt / (T_ - 1);
    prep_sample<T, DIAG, IN>(P, key.k[kk], traj + ((size_t)(b * T_ + t + 1) * 25) * N, traj + ((size_t)(b * T_ + t) * 25) * N,
                             traj + ((size_t)(b * T_ + (t > 0 ? t - 1 : 0)) * 25) * N, controls + (size_t)(b * T_ + t) * 4,
                             X + (size_t)q * XP, PHYS + (size_t)q * 25, TGT + (size_t)q * 25);
}

// One CTA per trajectory: the whole [T][25][N] block is loaded once, coalesced, into shared memory (every state is used
// by three consecutive steps and the per-sample accesses are column gathers), the samples are computed from there,
// staged in shared memory and written out as contiguous runs.  (56 -> ~25 us at C3; the thread-per-sample kernel moves
// 4-byte elements at a 40-byte stride in both directions.)
// x goes straight from registers to its 128 / 224-byte row as 16-byte stores; phys | tgt (100-byte rows) are staged.  Shared
// memory per CTA: the trajectory + 128 x 50 values (4 CTAs per SM at C3 instead of 2 with x staged as well: 31 -> ~20 us).
constexpr int PREP_ROW = 50;   // staging row: phys 25 | tgt 25
template <typename T, bool DIAG, int IN>
__global__ void __launch_bounds__(128)
kc_train_prep_traj_kernel(const __grid_constant__ RodC<T> P, const __grid_constant__ KeyIdx64 key, int64_t B, int T_, int K,
                          const T* __restrict__ traj, const T* __restrict__ controls, T* __restrict__ X,
                          T* __restrict__ PHYS, T* __restrict__ TGT) {
    constexpr int XPG = IN == 28 ? 32 : 56;
    extern __shared__ __align__(16) unsigned char kc_smem[];
    T* s_traj = reinterpret_cast<T*>(kc_smem);
    const int N = P.N, slab = 25 * N, total = T_ * slab;
    T* s_out = s_traj + total;                      // [128][PREP_ROW]
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    const T* src = traj + (size_t)b * total;
    if (((size_t)total * sizeof(T)) % 16 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {     // 128-bit loads
        const int nv = (int)((size_t)total * sizeof(T) / 16);
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(s_traj);
        for (int e = tid; e < nv; e += 128) d4[e] = s4[e];
    } else {
        for (int e = tid; e < total; e += 128) s_traj[e] = src[e];
    }
    __syncthreads();
    const int S = (T_ - 1) * K;
    for (int s0 = 0; s0 < S; s0 += 128) {
        const int s = s0 + tid, n = S - s0 < 128 ? S - s0 : 128;
        const size_t q0 = (size_t)b * S + s0;
        if (s < S) {
            const int t = s / K, kk = s - t * K;
            T* row = s_out + (size_t)tid * PREP_ROW;
            T x[XPG];
            prep_sample<T, DIAG, IN>(P, key.k[kk], s_traj + (size_t)(t + 1) * slab, s_traj + (size_t)t * slab,
                                     s_traj + (size_t)(t > 0 ? t - 1 : 0) * slab, controls + (size_t)(b * T_ + t) * 4,
                                     x, row, row + 25);
            T* xr = X + (q0 + tid) * XPG;            // rows are 16-byte aligned (XPG * sizeof(T) is a multiple of 16)
            constexpr int V = 16 / sizeof(T);
#pragma unroll
            for (int i = 0; i < XPG; i += V) {
                if (sizeof(T) == 4) *reinterpret_cast<float4*>(xr + i) = make_float4((float)x[i], (float)x[i + 1], (float)x[i + 2], (float)x[i + 3]);
                else *reinterpret_cast<double2*>(xr + i) = make_double2((double)x[i], (double)x[i + 1]);
            }
        }
        __syncthreads();
        for (int e = tid; e < n * 25; e += 128) {
            const int r = e / 25, c = e - r * 25;
            PHYS[q0 * 25 + e] = s_out[(size_t)r * PREP_ROW + c];
            TGT[q0 * 25 + e] = s_out[(size_t)r * PREP_ROW + 25 + c];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 2. forward + loss + dL/do
// ---------------------------------------------------------------------------------------------------------------
constexpr int FWD_THREADS = 512;
// One sample per thread; the packed weights are staged in shared memory once per CTA (when they fit: 123 KB in fp32 at
// H = 512) and read as 16-byte broadcasts; CTAs are persistent over tiles of FWD_THREADS samples.
template <typename T, int IN>
__global__ void __launch_bounds__(FWD_THREADS, 1)
kc_train_fwd_kernel(MlpC<T> M, int wp_in_smem, T ds, int64_t Q, int T_, int K, const T* __restrict__ X, int XP,
                    const T* __restrict__ PHYS, const T* __restrict__ TGT, T* __restrict__ dO,
                    double* __restrict__ loss_part, T* __restrict__ pred_out) {
    extern __shared__ __align__(16) unsigned char kc_smem[];
    if (wp_in_smem) {
        T* sw = reinterpret_cast<T*>(kc_smem);
        const int n = M.hidden * M.stride;
        for (int e = threadIdx.x; e < n; e += FWD_THREADS) sw[e] = M.Wp[e];
        __syncthreads();
        M.Wp = sw;
    }
    double my = 0.0;
    for (int64_t q = (int64_t)blockIdx.x * FWD_THREADS + threadIdx.x; q < Q; q += (int64_t)gridDim.x * FWD_THREADS) {
        T x[IN], o[25];
        const T* xr = X + (size_t)q * XP;
#pragma unroll
        for (int i = 0; i < IN; ++i) x[i] = xr[i];
        mlp_eval<T, IN>(M, x, o);
        T pred[25], tg[25];
#pragma unroll
        for (int r = 0; r < 19; ++r) pred[r] = PHYS[(size_t)q * 25 + r] + ds * o[r];
#pragma unroll
        for (int c = 19; c < 25; ++c) pred[c] = PHYS[(size_t)q * 25 + c] + o[c];
#pragma unroll
        for (int r = 0; r < 25; ++r) tg[r] = TGT[(size_t)q * 25 + r];
        const T S = T(T_ - 1);
        const T wp = T(1) / (T(3 * K) * S), wf = T(1) / (T(12 * K) * S), wz = T(1) / (T(6 * K) * S);
        T g[25];
        T acc = T(0);
#pragma unroll
        for (int r = 0; r < 3; ++r) { const T e = pred[r] - tg[r]; acc += wp * e * e; g[r] = T(2) * wp * e * ds; }
#pragma unroll
        for (int r = 7; r < 19; ++r) { const T e = pred[r] - tg[r]; acc += wf * e * e; g[r] = T(2) * wf * e * ds; }
#pragma unroll
        for (int r = 19; r < 25; ++r) { const T e = pred[r] - tg[r]; acc += wz * e * e; g[r] = T(2) * wz * e; }
        T ep[3], et[3], ge[3], gq[4];
        quat_to_euler(pred + 3, ep);
        quat_to_euler(tg + 3, et);
#pragma unroll
        for (int i = 0; i < 3; ++i) { const T e = ep[i] - et[i]; acc += wp * e * e; ge[i] = T(2) * wp * e; }
        quat_to_euler_vjp(pred + 3, ge, gq);
#pragma unroll
        for (int i = 0; i < 4; ++i) g[3 + i] = gq[i] * ds;
        T* go = dO + (size_t)q * 32;
#pragma unroll
        for (int r = 0; r < 25; ++r) go[r] = g[r];
#pragma unroll
        for (int r = 25; r < 32; ++r) go[r] = T(0);
        my += (double)acc;
        if (pred_out) {  // pred[B][T-1][25][K]
            const int kk = (int)(q % K);
            const int64_t bt = q / K;
            T* po = pred_out + (size_t)bt * 25 * K + kk;
#pragma unroll
            for (int r = 0; r < 25; ++r) po[r * K] = pred[r];
        }
    }
    // block reduction of the loss in double
    __shared__ double red[FWD_THREADS / 32];
#pragma unroll
    for (int o2 = 16; o2 > 0; o2 >>= 1) my += __shfl_down_sync(0xffffffffu, my, o2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = my;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int i = 0; i < FWD_THREADS / 32; ++i) s += red[i];
        loss_part[blockIdx.x] = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// 3. backward: weight gradients
// ---------------------------------------------------------------------------------------------------------------
constexpr int BWD_THREADS = 256;
constexpr int BWD_TS = 128;  // samples per tile
constexpr int BWD_HC = 32;   // hidden units per CTA

template <typename T, int IN>
__global__ void __launch_bounds__(BWD_THREADS)
kc_train_bwd_kernel(int hidden, const T* __restrict__ W1, const T* __restrict__ b1, const T* __restrict__ W2,
                    int64_t Q, const T* __restrict__ X, const T* __restrict__ dO, T* __restrict__ partial,
                    int64_t NP, int64_t tiles_per_split) {
    constexpr int XP = (IN + 3) & ~3;      // 28 -> 32? no: 28 stays 28; rows of X are padded to XPG in global memory
    constexpr int XPG = IN == 28 ? 32 : 56;  // global row pitch of X
    constexpr int NG = XP / 4;             // float4 groups per x row (7 or 14)
    extern __shared__ __align__(16) unsigned char kc_smem[];
    T* Xs = reinterpret_cast<T*>(kc_smem);              // [TS][XPG]
    T* dOs = Xs + BWD_TS * XPG;                          // [TS][32]
    T* As = dOs + BWD_TS * 32;                           // [TS][HC+1]
    T* dZs = As + BWD_TS * (BWD_HC + 1);                 // [TS][HC+1]
    T* W1c = dZs + BWD_TS * (BWD_HC + 1);                // [HC][XP]
    T* W2c = W1c + BWD_HC * XP;                          // [HC][28] (transposed: unit-major, 25 used)
    T* b1c = W2c + BWD_HC * 28;                          // [HC]
    const int tid = threadIdx.x;
    const int u0 = blockIdx.x * BWD_HC;
    // stage this CTA's slice of the weights (zero beyond `hidden`)
    for (int e = tid; e < BWD_HC * XP; e += BWD_THREADS) {
        const int i = e / XP, k = e - i * XP;
        W1c[e] = (u0 + i < hidden && k < IN) ? W1[(size_t)(u0 + i) * IN + k] : T(0);
    }
    for (int e = tid; e < BWD_HC * 28; e += BWD_THREADS) {
        const int i = e / 28, c = e - i * 28;
        W2c[e] = (u0 + i < hidden && c < 25) ? W2[(size_t)c * hidden + u0 + i] : T(0);
    }
    for (int e = tid; e < BWD_HC; e += BWD_THREADS) b1c[e] = (u0 + e < hidden) ? b1[u0 + e] : T(0);
    // accumulator ownership
    const int wi = tid >> 3, wg = tid & 7;          // gW1: unit wi, float4 groups wg (and wg+8 when NG > 8)
    const int vi = tid & 31, vc = (tid >> 5) * 4;   // gW2: unit vi, outputs vc..vc+3
    T aW1a[4] = {0, 0, 0, 0}, aW1b[4] = {0, 0, 0, 0}, aW2[4] = {0, 0, 0, 0}, ab1 = T(0), ab2 = T(0);
    const int64_t tile0 = (int64_t)blockIdx.y * tiles_per_split;
    const int64_t ntiles = (Q + BWD_TS - 1) / BWD_TS;
    for (int64_t tile = tile0; tile < tile0 + tiles_per_split && tile < ntiles; ++tile) {
        const int64_t q0 = tile * BWD_TS;
        const int cnt = (int)min((int64_t)BWD_TS, Q - q0);
        __syncthreads();
        for (int e = tid; e < BWD_TS * XPG; e += BWD_THREADS) Xs[e] = (e < cnt * XPG) ? X[(size_t)q0 * XPG + e] : T(0);
        for (int e = tid; e < BWD_TS * 32; e += BWD_THREADS) dOs[e] = (e < cnt * 32) ? dO[(size_t)q0 * 32 + e] : T(0);
        __syncthreads();
        {   // phase A: a = ELU(z1), dz = (W2^T dO) ELU'(z1) for (sample s, 16 of the 32 units)
            const int s = tid & (BWD_TS - 1), ib = (tid >> 7) * 16;
            T x[IN], g[25];
#pragma unroll
            for (int k = 0; k < IN; ++k) x[k] = Xs[s * XPG + k];
#pragma unroll
            for (int c = 0; c < 25; ++c) g[c] = dOs[s * 32 + c];
#pragma unroll 2
            for (int i = ib; i < ib + 16; ++i) {
                const T z1 = mlp_unit_dot<T, IN>(W1c + i * XP, x, b1c[i]);
                const T da = mlp_unit_dot<T, 25>(W2c + i * 28, g, T(0));
                As[s * (BWD_HC + 1) + i] = kc_elu(z1);
                dZs[s * (BWD_HC + 1) + i] = da * kc_elu_grad(z1);
            }
        }
        __syncthreads();
        // phase B: outer-product accumulation over the tile's samples
        for (int s = 0; s < BWD_TS; ++s) {
            const T dz = dZs[s * (BWD_HC + 1) + wi];
            if (wg < NG) {
                T xv[4];
                kc_ld4(Xs + s * XPG + wg * 4, xv);
#pragma unroll
                for (int k = 0; k < 4; ++k) aW1a[k] += dz * xv[k];
            }
            if (NG > 8 && wg + 8 < NG) {
                T xv[4];
                kc_ld4(Xs + s * XPG + (wg + 8) * 4, xv);
#pragma unroll
                for (int k = 0; k < 4; ++k) aW1b[k] += dz * xv[k];
            }
            const T a = As[s * (BWD_HC + 1) + vi];
            T gv[4];
            kc_ld4(dOs + s * 32 + vc, gv);
#pragma unroll
            for (int c = 0; c < 4; ++c) aW2[c] += gv[c] * a;
            if (tid < BWD_HC) ab1 += dZs[s * (BWD_HC + 1) + tid];
            if (blockIdx.x == 0 && tid >= 32 && tid < 32 + 25) ab2 += dOs[s * 32 + tid - 32];
        }
    }
    // write this CTA's partials: flat parameter order [W1 (hidden x IN) | b1 | W2 (25 x hidden) | b2]
    T* out = partial + (size_t)blockIdx.y * NP;
    const int64_t oW1 = 0, ob1 = (int64_t)hidden * IN, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
    if (u0 + wi < hidden) {
        if (wg < NG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if (wg * 4 + k < IN) out[oW1 + (int64_t)(u0 + wi) * IN + wg * 4 + k] = aW1a[k];
        }
        if (NG > 8 && wg + 8 < NG) {
#pragma unroll
            for (int k = 0; k < 4; ++k) if ((wg + 8) * 4 + k < IN) out[oW1 + (int64_t)(u0 + wi) * IN + (wg + 8) * 4 + k] = aW1b[k];
        }
    }
    if (u0 + vi < hidden) {
#pragma unroll
        for (int c = 0; c < 4; ++c) if (vc + c < 25) out[oW2 + (int64_t)(vc + c) * hidden + u0 + vi] = aW2[c];
    }
    if (tid < BWD_HC && u0 + tid < hidden) out[ob1 + u0 + tid] = ab1;
    if (blockIdx.x == 0 && tid >= 32 && tid < 32 + 25) out[ob2 + tid - 32] = ab2;
}

// 256 threads = 4 slice groups x 64 outputs: group g sums slices g, g+4, ... with four independent accumulators (the
// loads overlap instead of forming one 148-long dependent chain), the four group sums are combined in a fixed order.
template <typename T>
__global__ void __launch_bounds__(256)
kc_train_reduce_kernel(const T* __restrict__ partial, int splits, int64_t NP, int hidden, int in_dim,
                       T* gW1, T* gb1, T* gW2, T* gb2, const double* __restrict__ loss_part,
                       int nloss, double* loss) {
    __shared__ T red[4][64];
    const int g = threadIdx.x >> 6, pl = threadIdx.x & 63;
    const int64_t p = (int64_t)blockIdx.x * 64 + pl;
    T a0 = T(0), a1 = T(0), a2 = T(0), a3 = T(0);
    if (p < NP) {
        int k = g;
        for (; k + 12 < splits; k += 16) {
            a0 += partial[(size_t)k * NP + p];
            a1 += partial[(size_t)(k + 4) * NP + p];
            a2 += partial[(size_t)(k + 8) * NP + p];
            a3 += partial[(size_t)(k + 12) * NP + p];
        }
        for (; k < splits; k += 4) a0 += partial[(size_t)k * NP + p];
    }
    red[g][pl] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (g == 0 && p < NP) {
        const T s = (red[0][pl] + red[1][pl]) + (red[2][pl] + red[3][pl]);
        const int64_t ob1 = (int64_t)hidden * in_dim, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
        if (p < ob1) { if (gW1) gW1[p] = s; }
        else if (p < oW2) { if (gb1) gb1[p - ob1] = s; }
        else if (p < ob2) { if (gW2) gW2[p - oW2] = s; }
        else { if (gb2) gb2[p - ob2] = s; }
    }
    if (loss && blockIdx.x == 0 && threadIdx.x < 32) {  // deterministic: fixed lane partition, fixed shuffle tree
        double s = 0.0;
        for (int i = threadIdx.x; i < nloss; i += 32) s += loss_part[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *loss = s;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// host side of the train step
// ---------------------------------------------------------------------------------------------------------------
struct TrainWs { size_t X, PHYS, TGT, dO, lossp, part, wp, tcw, total; int XPG; int64_t Q, NP; int nfwd; int splits; int64_t tps; int chunks; };

static size_t bwd_smem_bytes(int in_dim, size_t sz) {
    const int XP = (in_dim + 3) & ~3, XPG = in_dim == 28 ? 32 : 56;
    return (size_t)(BWD_TS * XPG + BWD_TS * 32 + 2 * BWD_TS * (BWD_HC + 1) + BWD_HC * XP + BWD_HC * 28 + BWD_HC) * sz;
}

static TrainWs train_ws(int dtype, const kc_mlp* mlp, int64_t Q) {
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    TrainWs w;
    w.Q = Q;
    w.XPG = mlp->in_dim == 28 ? 32 : 56;
    w.NP = (int64_t)mlp->hidden * mlp->in_dim + mlp->hidden + 25 * (int64_t)mlp->hidden + 25;
    w.nfwd = (int)std::min<int64_t>((Q + FWD_THREADS - 1) / FWD_THREADS, 148);
    if (w.nfwd < 1) w.nfwd = 1;
    w.chunks = (mlp->hidden + BWD_HC - 1) / BWD_HC;
    const int64_t ntiles = (Q + BWD_TS - 1) / BWD_TS;
    // one resident wave of bwd CTAs: ~3 (fp32) / 1 (fp64) CTAs per SM on 148 SMs
    const int resident = 148 * (int)((227 * 1024) / bwd_smem_bytes(mlp->in_dim, sz));
    int splits = resident / w.chunks;
    if (splits < 1) splits = 1;
    if (splits > ntiles) splits = (int)(ntiles > 0 ? ntiles : 1);
    w.splits = splits;
    w.tps = (ntiles + splits - 1) / splits;
    size_t off = 0;
    w.X = off; off += al256((size_t)Q * w.XPG * sz);
    w.PHYS = off; off += al256((size_t)Q * 25 * sz);
    w.TGT = off; off += al256((size_t)Q * 25 * sz);
    w.dO = off; off += al256((size_t)Q * 32 * sz);
    w.lossp = off; off += al256((size_t)std::max(w.nfwd, 160) * 8);
    w.part = off; off += al256((size_t)std::max(splits, 160) * w.NP * sz);   // >= one slice per SM for the tensor-core path
    w.wp = off; off += al256((size_t)mlp->hidden * (((mlp->in_dim + 3) & ~3) + 32) * sz);
    w.tcw = off; off += al256((size_t)4 * (2 * 128 * 32 * 4 + 32768 + 16384));   // tensor-core path: split / transposed weights
    w.total = off;
    return w;
}

extern "C" int64_t kc_train_step_workspace_bytes(int dtype, const kc_mlp* mlp, int64_t B, int64_t T_, int32_t K) {
    if (!mlp || (dtype != KC_F32 && dtype != KC_F64) || B < 0 || T_ < 2 || K < 1) return KC_EINVAL;
    if (kc_check_mlp(mlp)) return KC_EINVAL;
    return (int64_t)train_ws(dtype, mlp, B * (T_ - 1) * K).total;
}

template <typename T>
static int train_typed(const kc_rod_params* Pp, const kc_mlp* mlp, int64_t B, int64_t T_, int K, const int32_t* key_host,
                       const void* traj, const void* controls, double* loss, void* gW1, void* gb1, void* gW2, void* gb2,
                       void* pred, void* workspace, cudaStream_t st) {
    const RodC<T> P = make_rodc<T>(*Pp);
    KeyIdx64 key{};
    for (int i = 0; i < K; ++i) {
        KC_CHECK_ARG(key_host[i] >= 1 && key_host[i] <= P.N - 1, "key index %d out of range [1, N-1]", key_host[i]);
        key.k[i] = key_host[i];
    }
    const int64_t Q = B * (T_ - 1) * K;
    const TrainWs w = train_ws(sizeof(T) == 4 ? KC_F32 : KC_F64, mlp, Q);
    unsigned char* ws = (unsigned char*)workspace;
    T* X = (T*)(ws + w.X); T* PHYS = (T*)(ws + w.PHYS); T* TGT = (T*)(ws + w.TGT); T* dO = (T*)(ws + w.dO);
    double* lossp = (double*)(ws + w.lossp); T* part = (T*)(ws + w.part);
    MlpC<T> M{};
    const int in_dim = mlp->in_dim;
    int tc_slices = 0;
    if (Q > 0) {
        const unsigned g1 = (unsigned)((Q + 127) / 128);
        const size_t psmem = ((size_t)T_ * 25 * P.N + (size_t)128 * PREP_ROW) * sizeof(T);
        const bool per_traj = psmem <= 100 * 1024;   // >= 2 CTAs per SM
#define PREP(D, I)                                                                                                     \
    do {                                                                                                               \
        if (per_traj) {                                                                                                \
            auto pk = kc_train_prep_traj_kernel<T, D, I>;                                                              \
            if (psmem > 48 * 1024) cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);  \
            pk<<<(unsigned)B, 128, psmem, st>>>(P, key, B, (int)T_, K, (const T*)traj, (const T*)controls, X, PHYS, TGT); \
        } else {                                                                                                       \
            kc_train_prep_kernel<T, D, I><<<g1, 128, 0, st>>>(P, key, B, (int)T_, K, (const T*)traj,                   \
                                                              (const T*)controls, X, w.XPG, PHYS, TGT);                \
        }                                                                                                              \
    } while (0)
        // tensor-core path (tcgen05): fp32 model, 28 inputs, hidden <= 512; KC_TRAIN_MODE=simt forces the SIMT kernels
        bool use_tc = kc_is_float<T>::value && in_dim == 28 && mlp->hidden <= 512;
        {
            const char* e = getenv("KC_TRAIN_MODE");
            if (e && e[0] == 's') use_tc = false;
        }
        // Small batches (at most one tile per CTA: the strong-scaling regime): the third-generation kernel forms its own samples
        // in its prologue and the prep launch is skipped (0.073 -> 0.069 ms at 128 trajectories).  With several tiles per CTA the
        // in-kernel gathers (4-byte, 40-byte stride) lose against the coalescing prep kernel (0.192 vs 0.179 ms at 1024), so the
        // stand-alone kernel stays.  KC_TRAIN_FUSE_PREP=0|1 forces either.
        bool fuse_prep = false;
        if (use_tc) {
            const char* gen = getenv("KC_TRAIN_TC");
            const char* fe = getenv("KC_TRAIN_FUSE_PREP");
            fuse_prep = !(gen && (gen[0] == '1' || gen[0] == '2')) && (Q + 127) / 128 <= kc_train_tc_grid(Q);
            if (fe && (fe[0] == '0' || fe[0] == '1')) fuse_prep = fe[0] == '1' && !(gen && (gen[0] == '1' || gen[0] == '2'));
        }
        if (!fuse_prep) {
            if (P.diag) { if (in_dim == 28) PREP(true, 28); else PREP(true, 53); }
            else { if (in_dim == 28) PREP(false, 28); else PREP(false, 53); }
            KC_CHECK_LAUNCH("kc_train_prep_kernel");
        }
#undef PREP
        if (use_tc) {
            float* tcw = (float*)(ws + w.tcw);
            tc_slices = kc_train_tc_grid(Q);
            // third-generation kernel (kc_train_tc3.cu: transposed backward, activations stay in TMEM) unless KC_TRAIN_TC=2
            // (warp-specialised pipeline with shared-memory activation tiles, kc_train_tc2.cu) or =1 (kc_train_tc.cu)
            const char* gen = getenv("KC_TRAIN_TC");
            int rc2;
            if (gen && gen[0] == '1')
                rc2 = kc_train_tc_launch(mlp, (float)P.ds, Q, (int)T_, K, (const float*)X, (const float*)PHYS, (const float*)TGT,
                                         tcw, tcw + 4 * 2 * 128 * 32, (float*)part, w.NP, lossp, (float*)pred, tc_slices, st);
            else if (gen && gen[0] == '2')
                rc2 = kc_train_tc2_launch(mlp, (float)P.ds, Q, (int)T_, K, (const float*)X, (const float*)PHYS, (const float*)TGT,
                                          (unsigned char*)tcw, (float*)part, w.NP, lossp, (float*)pred, tc_slices, st);
            else
{
                if constexpr (kc_is_float<T>::value)
                    rc2 = kc_train_tc3_launch(mlp, (float)P.ds, Q, (int)T_, K, (float*)X, (float*)PHYS, (float*)TGT, (unsigned char*)tcw,
                                              (float*)part, w.NP, lossp, (float*)pred, tc_slices, st, fuse_prep ? &P : nullptr, &key,
                                              (const float*)traj, (const float*)controls, B);
                else
                    rc2 = KC_EINVAL;
            }
            if (rc2) return rc2;
        } else {
        {
            int rc = kc_pack_mlp<T>(mlp, (T*)(ws + w.wp), M, st);   // SIMT weight image (the tensor-core path has its own)
            if (rc) return rc;
            const size_t wbytes = (size_t)M.hidden * M.stride * sizeof(T);
            const int in_smem = wbytes <= 200 * 1024 ? 1 : 0;
            const size_t fsmem = in_smem ? wbytes : 0;
            if (in_dim == 28) {
                auto k = kc_train_fwd_kernel<T, 28>;
                if (fsmem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
                k<<<w.nfwd, FWD_THREADS, fsmem, st>>>(M, in_smem, P.ds, Q, (int)T_, K, X, w.XPG, PHYS, TGT, dO, lossp, (T*)pred);
            } else {
                auto k = kc_train_fwd_kernel<T, 53>;
                if (fsmem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem);
                k<<<w.nfwd, FWD_THREADS, fsmem, st>>>(M, in_smem, P.ds, Q, (int)T_, K, X, w.XPG, PHYS, TGT, dO, lossp, (T*)pred);
            }
        }
        KC_CHECK_LAUNCH("kc_train_fwd_kernel");
        const size_t smem = bwd_smem_bytes(in_dim, sizeof(T));
        dim3 grid((unsigned)w.chunks, (unsigned)w.splits);
        if (in_dim == 28) {
            auto k = kc_train_bwd_kernel<T, 28>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.NP, w.tps);
        } else {
            auto k = kc_train_bwd_kernel<T, 53>;
            cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            k<<<grid, BWD_THREADS, smem, st>>>(mlp->hidden, (const T*)mlp->W1, (const T*)mlp->b1, (const T*)mlp->W2, Q, X, dO, part, w.NP, w.tps);
        }
        KC_CHECK_LAUNCH("kc_train_bwd_kernel");
        }
    }
    const int splits = Q > 0 ? (tc_slices > 0 ? tc_slices : w.splits) : 0;
    kc_train_reduce_kernel<T><<<(unsigned)((w.NP + 63) / 64), 256, 0, st>>>(part, splits, w.NP, mlp->hidden, in_dim, (T*)gW1, (T*)gb1,
                                                                            (T*)gW2, (T*)gb2, lossp, Q > 0 ? (tc_slices > 0 ? tc_slices : w.nfwd) : 0, loss);
    KC_CHECK_LAUNCH("kc_train_reduce_kernel");
    return KC_OK;
}

extern "C" int kc_train_step(int dtype, const kc_rod_params* P, const kc_mlp* mlp, int64_t B, int64_t T_, int32_t K,
                             const int32_t* key_idx_host, const void* traj, const void* controls, double* loss, void* gW1,
                             void* gb1, void* gW2, void* gb2, void* pred, void* workspace, int64_t workspace_bytes,
                             void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(P && P->N >= 2, "rod params missing or N < 2");
    KC_CHECK_ARG(mlp, "kc_train_step needs the MLP (use_nn)");
    int rc = kc_check_mlp(mlp);
    if (rc) return rc;
    KC_CHECK_ARG(B >= 0 && T_ >= 2, "B must be >= 0 and T >= 2");
    KC_CHECK_ARG(K >= 1 && K <= 64 && key_idx_host, "1 <= K <= 64 and key_idx_host non-NULL");
    KC_CHECK_ARG(loss && gW1 && gb1 && gW2 && gb2 && workspace, "NULL output/workspace pointer");
    KC_CHECK_ARG(B == 0 || (traj && controls), "NULL traj/controls");
    const int64_t need = kc_train_step_workspace_bytes(dtype, mlp, B, T_, K);
    if (workspace_bytes < need) {
        kc_set_error("workspace too small: %lld < %lld", (long long)workspace_bytes, (long long)need);
        return KC_ENOSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        return train_typed<float>(P, mlp, B, T_, K, key_idx_host, traj, controls, loss, gW1, gb1, gW2, gb2, pred, workspace, st);
    return train_typed<double>(P, mlp, B, T_, K, key_idx_host, traj, controls, loss, gW1, gb1, gW2, gb2, pred, workspace, st);
}

// ---------------------------------------------------------------------------------------------------------------
// Adam + clamp
// ---------------------------------------------------------------------------------------------------------------
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

// ---- all parameter tensors in ONE launch, hyper-parameters resident on the device --------------------------------
// For a training step captured in a CUDA graph: nothing of the launch depends on the host (the step count lives in
// *step_dev and is advanced by the kernel itself, the learning rate is read from *lr_dev, which the host rewrites when
// the plateau scheduler fires).  Up to 8 tensors; a block handles 256 consecutive elements of one tensor.
struct AdamTensors {
    void* p[8]; const void* g[8]; void* m[8]; void* v[8];
    int64_t n[8]; int32_t clamp[8]; int32_t first_block[9];
    int32_t count;
};
template <typename T>
__global__ void __launch_bounds__(256)
kc_adam_clamp_multi_kernel(const __grid_constant__ AdamTensors A, int64_t* step_dev, const double* lr_dev, double b1d,
                           double b2d, T eps, T wd, int32_t* ticket) {
    __shared__ T s_lr1, s_bc2s;
    if (threadIdx.x == 0) {
        const double step = (double)(*step_dev + 1);
        s_lr1 = (T)(*lr_dev / (1.0 - pow(b1d, step)));
        s_bc2s = (T)sqrt(1.0 - pow(b2d, step));
    }
    __syncthreads();
    int k = 0;
#pragma unroll
    for (int i = 1; i < 8; ++i) if (i < A.count && (int)blockIdx.x >= A.first_block[i]) k = i;
    const int64_t i = (int64_t)((int)blockIdx.x - A.first_block[k]) * 256 + threadIdx.x;
    if (i < A.n[k]) {
        T* p = (T*)A.p[k]; const T* g = (const T*)A.g[k]; T* m = (T*)A.m[k]; T* v = (T*)A.v[k];
        const T b1 = (T)b1d, b2 = (T)b2d;
        const T gi = g[i] + wd * p[i];
        const T mi = b1 * m[i] + (T(1) - b1) * gi;
        const T vi = b2 * v[i] + (T(1) - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const T denom = sqrt(vi) / s_bc2s + eps;
        T pi = p[i] - s_lr1 * (mi / denom);
        if (A.clamp[k] && pi < T(0)) pi = T(0);
        p[i] = pi;
    }
    // the block that finishes last advances the step count (every block has read it by then) and re-arms the ticket
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            *step_dev += 1;
            *ticket = 0;
        }
    }
}

extern "C" int kc_adam_clamp_multi(int dtype, int32_t n_tensors, const kc_adam_tensor* t, int64_t* step_dev,
                                   const double* lr_dev, double beta1, double beta2, double eps, double weight_decay,
                                   int32_t* ticket_dev, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(n_tensors >= 1 && n_tensors <= 8 && t, "1..8 tensors");
    KC_CHECK_ARG(step_dev && lr_dev && ticket_dev, "NULL step/lr/ticket pointer");
    AdamTensors A{};
    A.count = n_tensors;
    int blocks = 0;
    for (int i = 0; i < n_tensors; ++i) {
        KC_CHECK_ARG(t[i].n >= 0 && (t[i].n == 0 || (t[i].param && t[i].grad && t[i].exp_avg && t[i].exp_avg_sq)),
                     "tensor %d: NULL data pointer or negative size", i);
        A.p[i] = t[i].param; A.g[i] = t[i].grad; A.m[i] = t[i].exp_avg; A.v[i] = t[i].exp_avg_sq;
        A.n[i] = t[i].n; A.clamp[i] = t[i].clamp_min_zero;
        A.first_block[i] = blocks;
        blocks += (int)((t[i].n + 255) / 256);
    }
    A.first_block[n_tensors] = blocks;
    if (blocks == 0) blocks = 1;   // still advances the step count
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        kc_adam_clamp_multi_kernel<float><<<blocks, 256, 0, st>>>(A, step_dev, lr_dev, beta1, beta2, (float)eps,
                                                                 (float)weight_decay, ticket_dev);
    else
        kc_adam_clamp_multi_kernel<double><<<blocks, 256, 0, st>>>(A, step_dev, lr_dev, beta1, beta2, eps, weight_decay,
                                                                  ticket_dev);
    KC_CHECK_LAUNCH("kc_adam_clamp_multi_kernel");
    return KC_OK;
}


// ---------------------------------------------------------------------------------------------------------------------
// Gradient all-reduce + Adam as TWO kernels over NVLink peer memory (replaces kc_train_reduce's output -> ncclAllReduce ->
// kc_adam_clamp_multi for the 110 KB [gradients | loss] buffer of the training step, SURVEY §8e: the only collective of the
// path is latency bound).  Every rank owns a region of symmetric (peer-mapped) memory:
//     float data[2][n_pad]   two slots, used alternately by consecutive optimiser steps
//     uint32 flags[8]        flags[p] = number of the last step whose data rank p has published HERE
//   kc_peer_publish     : copy the local flat buffer into the own region's slot (step & 1), fence, then store the step
//                         number into flags[rank] of EVERY rank's region (st.release.sys over NVLink);
//   kc_peer_gather_adam : wait until all flags of the own region have reached this step (ld.acquire.sys), then every
//                         thread sums its element over the ranks' slots IN RANK ORDER — all ranks add the same numbers in
//                         the same order, so the result (and therefore the weights) are bitwise identical everywhere —
//                         writes the sum back to the local flat buffer and applies Adam + clamp to its parameter.
// A slot is rewritten two steps later; by then every peer has passed the gather of the step in between, which waited for
// this rank's publish of that step, i.e. for the end of this rank's previous gather: no reader can still be in the slot.
struct PeerRegions { void* r[8]; int32_t world, rank; };
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
static inline int64_t peer_npad(int64_t n) { return (n + 31) / 32 * 32; }

template <typename T>
__global__ void __launch_bounds__(256)
kc_peer_publish_kernel(const __grid_constant__ PeerRegions R, int64_t n, int64_t n_pad, const T* __restrict__ flat,
                       const int64_t* __restrict__ step_dev, int32_t* ticket) {
    const uint32_t e = (uint32_t)(*step_dev + 1);
    T* slot = reinterpret_cast<T*>(R.r[R.rank]) + (size_t)(e & 1u) * n_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) slot[i] = flat[i];
    __threadfence_system();
    __syncthreads();
    __shared__ int last;
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last) {
        if (threadIdx.x == 0) *ticket = 0;
        __threadfence_system();
        if ((int)threadIdx.x < R.world) {
            uint32_t* flags = reinterpret_cast<uint32_t*>(reinterpret_cast<T*>(R.r[threadIdx.x]) + 2 * n_pad);
            st_release_sys(flags + R.rank, e);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
kc_peer_gather_adam_kernel(const __grid_constant__ PeerRegions R, const __grid_constant__ AdamTensors A, int64_t n, int64_t n_pad,
                           T* __restrict__ flat, int64_t* step_dev, const double* lr_dev, double b1d, double b2d, T eps, T wd,
                           int32_t* ticket, int adam_blocks) {
    __shared__ T s_lr1, s_bc2s;
    const uint32_t e = (uint32_t)(*step_dev + 1);
    if ((int)threadIdx.x < R.world) {
        const uint32_t* flags = reinterpret_cast<const uint32_t*>(reinterpret_cast<const T*>(R.r[R.rank]) + 2 * n_pad);
        while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - e) < 0) __nanosleep(20);
    }
    if (threadIdx.x == 32) {
        const double step = (double)(*step_dev + 1);
        s_lr1 = (T)(*lr_dev / (1.0 - pow(b1d, step)));
        s_bc2s = (T)sqrt(1.0 - pow(b2d, step));
    }
    __syncthreads();
    const size_t so = (size_t)(e & 1u) * n_pad;
    // all peers' values are requested before any is used (a loop of load -> add would pay one NVLink round trip per rank);
    // the sum is then formed in rank order, so every rank adds the same numbers in the same order
    auto gather = [&](int64_t fi) {
        T v[8];
#pragma unroll
        for (int p = 0; p < 8; ++p)
            v[p] = p < R.world ? *reinterpret_cast<const volatile T*>(reinterpret_cast<const T*>(R.r[p]) + so + fi) : T(0);
        T s = T(0);
#pragma unroll
        for (int p = 0; p < 8; ++p) if (p < R.world) s += v[p];
        flat[fi] = s;
        return s;
    };
    if ((int)blockIdx.x < adam_blocks) {
        int k = 0;
#pragma unroll
        for (int i = 1; i < 8; ++i) if (i < A.count && (int)blockIdx.x >= A.first_block[i]) k = i;
        const int64_t i = (int64_t)((int)blockIdx.x - A.first_block[k]) * 256 + threadIdx.x;
        if (i < A.n[k]) {
            T* p = (T*)A.p[k]; T* m = (T*)A.m[k]; T* v = (T*)A.v[k];
            const int64_t fi = (const T*)A.g[k] - flat + i;
            const T b1 = (T)b1d, b2 = (T)b2d;
            const T gi = gather(fi) + wd * p[i];
            const T mi = b1 * m[i] + (T(1) - b1) * gi;
            const T vi = b2 * v[i] + (T(1) - b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            const T denom = sqrt(vi) / s_bc2s + eps;
            T pi = p[i] - s_lr1 * (mi / denom);
            if (A.clamp[k] && pi < T(0)) pi = T(0);
            p[i] = pi;
        }
    } else {   // elements of the flat buffer that belong to no parameter tensor (the loss): sum only
        int64_t covered = 0;
        for (int i = 0; i < A.count; ++i) covered += A.n[i];
        const int64_t fi = covered + (int64_t)((int)blockIdx.x - adam_blocks) * 256 + threadIdx.x;
        if (fi < n) gather(fi);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1) == (int)gridDim.x - 1) {
            *step_dev += 1;
            *ticket = 0;
        }
    }
}

extern "C" int64_t kc_peer_region_bytes(int dtype, int64_t n_flat) {
    if ((dtype != KC_F32 && dtype != KC_F64) || n_flat < 1) return KC_EINVAL;
    return 2 * peer_npad(n_flat) * (dtype == KC_F32 ? 4 : 8) + 128;
}

static int peer_regions(int32_t world, int32_t rank, void* const* regions_host, PeerRegions& R) {
    KC_CHECK_ARG(world >= 1 && world <= 8 && rank >= 0 && rank < world && regions_host, "need 1 <= world <= 8, 0 <= rank < world");
    for (int i = 0; i < world; ++i) {
        KC_CHECK_ARG(regions_host[i], "peer region %d is NULL", i);
        R.r[i] = regions_host[i];
    }
    R.world = world; R.rank = rank;
    return KC_OK;
}

extern "C" int kc_peer_publish(int dtype, int32_t world, int32_t rank, void* const* regions_host, int64_t n_flat,
                               const void* flat, const int64_t* step_dev, int32_t* ticket_dev, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(n_flat >= 1 && flat && step_dev && ticket_dev, "NULL pointer or empty buffer");
    PeerRegions R{};
    int rc = peer_regions(world, rank, regions_host, R);
    if (rc) return rc;
    const int blocks = (int)((n_flat + 1023) / 1024);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        kc_peer_publish_kernel<float><<<blocks, 256, 0, st>>>(R, n_flat, peer_npad(n_flat), (const float*)flat, step_dev, ticket_dev);
    else
        kc_peer_publish_kernel<double><<<blocks, 256, 0, st>>>(R, n_flat, peer_npad(n_flat), (const double*)flat, step_dev, ticket_dev);
    KC_CHECK_LAUNCH("kc_peer_publish_kernel");
    return KC_OK;
}

extern "C" int kc_peer_gather_adam(int dtype, int32_t world, int32_t rank, void* const* regions_host, int64_t n_flat, void* flat,
                                   int32_t n_tensors, const kc_adam_tensor* t, int64_t* step_dev, const double* lr_dev,
                                   double beta1, double beta2, double eps, double weight_decay, int32_t* ticket_dev,
                                   void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(n_tensors >= 1 && n_tensors <= 8 && t, "1..8 tensors");
    KC_CHECK_ARG(n_flat >= 1 && flat && step_dev && lr_dev && ticket_dev, "NULL pointer or empty buffer");
    PeerRegions R{};
    int rc = peer_regions(world, rank, regions_host, R);
    if (rc) return rc;
    const size_t sz = dtype == KC_F32 ? 4 : 8;
    AdamTensors A{};
    A.count = n_tensors;
    int blocks = 0;
    int64_t covered = 0;
    for (int i = 0; i < n_tensors; ++i) {
        KC_CHECK_ARG(t[i].n >= 0 && (t[i].n == 0 || (t[i].param && t[i].grad && t[i].exp_avg && t[i].exp_avg_sq)),
                     "tensor %d: NULL data pointer or negative size", i);
        // the gradients must be the consecutive leading views of the flat buffer ([g0 | g1 | ... | rest])
        KC_CHECK_ARG((const char*)t[i].grad == (const char*)flat + (size_t)covered * sz,
                     "tensor %d: its gradient must be the view of the flat buffer that follows the previous tensor's", i);
        A.p[i] = t[i].param; A.g[i] = t[i].grad; A.m[i] = t[i].exp_avg; A.v[i] = t[i].exp_avg_sq;
        A.n[i] = t[i].n; A.clamp[i] = t[i].clamp_min_zero;
        A.first_block[i] = blocks;
        blocks += (int)((t[i].n + 255) / 256);
        covered += t[i].n;
    }
    A.first_block[n_tensors] = blocks;
    KC_CHECK_ARG(covered <= n_flat, "the tensors are larger than the flat buffer");
    const int adam_blocks = blocks;
    blocks += (int)((n_flat - covered + 255) / 256);
    if (blocks == 0) blocks = 1;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32)
        kc_peer_gather_adam_kernel<float><<<blocks, 256, 0, st>>>(R, A, n_flat, peer_npad(n_flat), (float*)flat, step_dev, lr_dev,
                                                                 beta1, beta2, (float)eps, (float)weight_decay, ticket_dev, adam_blocks);
    else
        kc_peer_gather_adam_kernel<double><<<blocks, 256, 0, st>>>(R, A, n_flat, peer_npad(n_flat), (double*)flat, step_dev, lr_dev,
                                                                  beta1, beta2, eps, weight_decay, ticket_dev, adam_blocks);
    KC_CHECK_LAUNCH("kc_peer_gather_adam_kernel");
    return KC_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// torch.optim.lr_scheduler.ReduceLROnPlateau('min', patience, factor) with its defaults (relative threshold, cooldown 0,
// min_lr 0) on the DEVICE — physics_train.py:206,297 steps it on the epoch loss every epoch.  state[6] = {best, bad epochs,
// patience, factor, threshold, eps}; *lr_dev is the learning rate kc_adam_clamp_multi / kc_peer_gather_adam read at their
// next launch.  One thread; launched after the optimiser inside the training step's CUDA graph, so the scheduler runs
// every step without a host read of the loss.
template <typename T>
__global__ void kc_plateau_step_kernel(const T* __restrict__ loss, double* __restrict__ state, double* __restrict__ lr) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double metric = (double)*loss;
    double best = state[0], bad = state[1];
    if (metric < best * (1.0 - state[4])) { best = metric; bad = 0.0; }
    else bad += 1.0;
    if (bad > state[2]) {
        const double old_lr = *lr, new_lr = old_lr * state[3];
        if (old_lr - new_lr > state[5]) *lr = new_lr;
        bad = 0.0;
    }
    state[0] = best; state[1] = bad;
}
extern "C" int kc_plateau_step(int dtype, const void* loss_dev, double* state_dev, double* lr_dev, void* stream) {
    KC_CHECK_ARG(dtype == KC_F32 || dtype == KC_F64, "dtype must be KC_F32 or KC_F64");
    KC_CHECK_ARG(loss_dev && state_dev && lr_dev, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == KC_F32) kc_plateau_step_kernel<float><<<1, 32, 0, st>>>((const float*)loss_dev, state_dev, lr_dev);
    else kc_plateau_step_kernel<double><<<1, 32, 0, st>>>((const double*)loss_dev, state_dev, lr_dev);
    KC_CHECK_LAUNCH("kc_plateau_step_kernel");
    return KC_OK;
}

#include "kc_ode_bwd.inl"
