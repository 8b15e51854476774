// kc_mlp_coop.cuh — warp-cooperative KNODE MLP (device only).  One rod per WARP: all 32 lanes carry the same state and
// run the physics redundantly; the 28/53 -> H -> 25 MLP, which is 99 % of a KNODE node evaluation, is split over the
// lanes by hidden unit (unit i on lane i % 32) and the 25 outputs (or the input cotangent) are combined with a butterfly
// all-reduce, which leaves bit-identical values in every lane, so the lanes never diverge.  Used when the MLP sits inside
// the march (KNODE rollout, BPTT) and the batch is too small to fill the chip with one rod per thread: B rods give B
// warps instead of B/32, and the weights are streamed coalesced (unit index fastest) instead of broadcast.
#pragma once
#include "kc_rod.cuh"

// The weights of a hidden unit (one row of Wc, see MlpCoop) are fetched as ONE batch of independent 16-byte loads
// before any of them is used: with scalar unit-fastest loads the kernel spent 57 % of its time on load instructions
// (LSU pipe 54 % busy, one load + one address IMAD per FMA), and before that the compiler rotated two load registers so
// that every FMA waited out a full load latency (1 430 cycles per unit instead of ~200).
// values [K0, K1) of unit i's row (K0, K1 multiples of 4) -> w[K0..K1)
template <typename T, int IN, int K0, int K1>
__device__ __forceinline__ void coop_load_row(const T* __restrict__ Wc, int i, T* __restrict__ w) {
    constexpr int ROW = kc_coop_row(IN);
    const T* __restrict__ row = Wc + (size_t)i * ROW;
#pragma unroll
    for (int k = K0; k < K1; k += 4) kc_ld4(row + k, w + k);
    asm volatile("" ::: "memory");   // every load of the batch is issued before the first use
}
// one batch when the row fits ~64 registers (fp32, 28 inputs), else first layer | second layer
template <typename T, int IN> struct CoopBatch {
    static constexpr int inP = (IN + 3) & ~3, ROW = kc_coop_row(IN);
    static constexpr bool two = sizeof(T) * ROW > 256;
    static constexpr int A1 = two ? inP + 4 : ROW;   // end of the first batch (covers W1 row and b1)
    static constexpr int B0 = two ? inP : ROW;       // start of the second batch (b1 again, then W2 column)
};

template <typename T, int IN>
__device__ __forceinline__ void mlp_eval(const MlpCoop<T>& M, const T* __restrict__ x, T* __restrict__ o) {
    constexpr int inP = (IN + 3) & ~3, ROW = kc_coop_row(IN);
    const int lane = threadIdx.x & 31, Hp = M.Hp;
    T acc[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) acc[c] = T(0);
#pragma unroll 1
    for (int i = lane; i < Hp; i += 32) {
        T w[ROW];
        coop_load_row<T, IN, 0, CoopBatch<T, IN>::A1>(M.Wc, i, w);
        T p[4] = {w[inP], T(0), T(0), T(0)};
#pragma unroll
        for (int k = 0; k < IN; ++k) p[k & 3] += w[k] * x[k];
        const T a = kc_elu((p[0] + p[1]) + (p[2] + p[3]));
        coop_load_row<T, IN, CoopBatch<T, IN>::B0, ROW>(M.Wc, i, w);
#pragma unroll
        for (int c = 0; c < 25; ++c) acc[c] += w[inP + 1 + c] * a;
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
    constexpr int inP = (IN + 3) & ~3, ROW = kc_coop_row(IN);
    const int lane = threadIdx.x & 31, Hp = M.Hp;
    T acc[IN];
#pragma unroll
    for (int k = 0; k < IN; ++k) acc[k] = T(0);
#pragma unroll 1
    for (int i = lane; i < Hp; i += 32) {
        T w[ROW];
        coop_load_row<T, IN, 0, CoopBatch<T, IN>::A1>(M.Wc, i, w);
        T p[4] = {w[inP], T(0), T(0), T(0)};
#pragma unroll
        for (int k = 0; k < IN; ++k) p[k & 3] += w[k] * x[k];
        const T z1 = (p[0] + p[1]) + (p[2] + p[3]);
        coop_load_row<T, IN, CoopBatch<T, IN>::B0, ROW>(M.Wc, i, w);
        T d[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int c = 0; c < 25; ++c) d[c & 3] += w[inP + 1 + c] * go[c];
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

// W1[H][in], b1[H], W2[25][H] -> Wc rows (see MlpCoop)
template <typename T>
__global__ void kc_pack_mlp_coop_kernel(const T* __restrict__ W1, const T* __restrict__ b1, const T* __restrict__ W2,
                                        T* __restrict__ Wc, int in_dim, int inP, int hidden, int Hp) {
    const int ROW = kc_coop_row(in_dim);
    const int total = ROW * Hp;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int i = e / ROW, r = e - i * ROW;
        T v = T(0);
        if (i < hidden) {
            if (r < in_dim) v = W1[(size_t)i * in_dim + r];
            else if (r == inP) v = b1[i];
            else if (r > inP && r <= inP + 25) v = W2[(size_t)(r - inP - 1) * hidden + i];
        }
        Wc[e] = v;
    }
}
