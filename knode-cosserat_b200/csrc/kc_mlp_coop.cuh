// kc_mlp_coop.cuh — warp-cooperative KNODE MLP (device only).  One rod per WARP: all 32 lanes carry the same state and
// run the physics redundantly; the 28/53 -> H -> 25 MLP, which is 99 % of a KNODE node evaluation, is split over the
// lanes by hidden unit (unit i on lane i % 32) and the 25 outputs (or the input cotangent) are combined with a butterfly
// all-reduce, which leaves bit-identical values in every lane, so the lanes never diverge.  Used when the MLP sits inside
// the march (KNODE rollout, BPTT) and the batch is too small to fill the chip with one rod per thread: B rods give B
// warps instead of B/32, and the weights are streamed coalesced (unit index fastest) instead of broadcast.
#pragma once
#include "kc_rod.cuh"

// The weights of a hidden unit are fetched as ONE batch of independent loads into registers before any of them is
// used (the compiler otherwise rotates two load registers and every FMA waits out a full load latency: measured 1 430
// cycles per unit instead of ~200).
template <typename T, int IN>
__device__ __forceinline__ void mlp_eval(const MlpCoop<T>& M, const T* __restrict__ x, T* __restrict__ o) {
    constexpr int inP = (IN + 3) & ~3;
    const int lane = threadIdx.x & 31, Hp = M.Hp;
    const T* __restrict__ W1T = M.Wc;
    const T* __restrict__ b1 = W1T + (size_t)inP * Hp;
    const T* __restrict__ W2 = b1 + Hp;
    T acc[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) acc[c] = T(0);
#pragma unroll 1
    for (int i = lane; i < Hp; i += 32) {
        T w[IN], w2[25];
        const T bias = b1[i];
#pragma unroll
        for (int k = 0; k < IN; ++k) w[k] = W1T[(size_t)k * Hp + i];
#pragma unroll
        for (int c = 0; c < 25; ++c) w2[c] = W2[(size_t)c * Hp + i];
        asm volatile("" ::: "memory");   // all loads of the unit are issued before the first use (see above)
        T p[4] = {bias, T(0), T(0), T(0)};
#pragma unroll
        for (int k = 0; k < IN; ++k) p[k & 3] += w[k] * x[k];
        const T a = kc_elu((p[0] + p[1]) + (p[2] + p[3]));
#pragma unroll
        for (int c = 0; c < 25; ++c) acc[c] += w2[c] * a;
    }
#pragma unroll
    for (int c = 0; c < 25; ++c) {
        T v = acc[c];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        o[c] = M.b2[c] + v;
    }
}

template <typename T, int IN>
__device__ __forceinline__ void mlp_input_vjp(const MlpCoop<T>& M, const T* __restrict__ x, const T* __restrict__ go,
                                              T* __restrict__ gx) {
    constexpr int inP = (IN + 3) & ~3;
    const int lane = threadIdx.x & 31, Hp = M.Hp;
    const T* __restrict__ W1T = M.Wc;
    const T* __restrict__ b1 = W1T + (size_t)inP * Hp;
    const T* __restrict__ W2 = b1 + Hp;
    T acc[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) acc[k] = T(0);
#pragma unroll 1
    for (int i = lane; i < Hp; i += 32) {
        T w[IN], w2[25];
        const T bias = b1[i];
#pragma unroll
        for (int k = 0; k < IN; ++k) w[k] = W1T[(size_t)k * Hp + i];
#pragma unroll
        for (int c = 0; c < 25; ++c) w2[c] = W2[(size_t)c * Hp + i];
        asm volatile("" ::: "memory");   // all loads of the unit are issued before the first use (see above)
        T p[4] = {bias, T(0), T(0), T(0)};
#pragma unroll
        for (int k = 0; k < IN; ++k) p[k & 3] += w[k] * x[k];
        const T z1 = (p[0] + p[1]) + (p[2] + p[3]);
        T d[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int c = 0; c < 25; ++c) d[c & 3] += w2[c] * go[c];
        const T dz = ((d[0] + d[1]) + (d[2] + d[3])) * kc_elu_grad(z1);
#pragma unroll
        for (int k = 0; k < IN; ++k) acc[k] += dz * w[k];
    }
#pragma unroll
    for (int k = 0; k < IN; ++k) {
        T v = acc[k];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        gx[k] = v;
    }
}

// W1[H][in], b1[H], W2[25][H] -> Wc (see MlpCoop)
template <typename T>
__global__ void kc_pack_mlp_coop_kernel(const T* __restrict__ W1, const T* __restrict__ b1, const T* __restrict__ W2,
                                        T* __restrict__ Wc, int in_dim, int inP, int hidden, int Hp) {
    const int total = (inP + 1 + 25) * Hp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int r = e / Hp, i = e - r * Hp;
        T v = T(0);
        if (i < hidden) {
            if (r < inP) v = r < in_dim ? W1[(size_t)i * in_dim + r] : T(0);
            else if (r == inP) v = b1[i];
            else v = W2[(size_t)(r - inP - 1) * hidden + i];
        }
        Wc[e] = v;
    }
}
