// kc_common.cuh — shared definitions for the sm_100a kernels of libknode_cosserat_b200.so.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define KC_HD __host__ __device__ __forceinline__
#define KC_D __device__ __forceinline__
#else
#define KC_HD inline
#define KC_D inline
#endif

#include "../../include/knode_cosserat.h"

// Rod constants in the arithmetic type of the kernel.  Passed by value as a __grid_constant__ kernel parameter, so
// every field is a uniform constant-bank operand of the FMA that uses it (no loads, no registers).
template <typename T>
struct RodC {
    T ds, c0, c1, c2, rhoA;
    T KseInv[9], KbtInv[9], Bse[9], Bbt[9], rhoJ[9];
    T KseVstar[3], rhoAg[3], C[3], Ftip[3], Mtip[3], p0[3], h0[4], q0[3], w0[3];
    T tdirs[12];
    int N;
    int diag;  // 1: KseInv, KbtInv, Bbt, rhoJ diagonal and Bse == 0 (every configuration the reference builds)
};

template <typename T>
inline RodC<T> make_rodc(const kc_rod_params& p) {
    RodC<T> c;
    c.ds = (T)p.ds; c.c0 = (T)p.c0; c.c1 = (T)p.c1; c.c2 = (T)p.c2; c.rhoA = (T)p.rhoA;
    bool diag = true;
    for (int i = 0; i < 9; ++i) {
        c.KseInv[i] = (T)p.Kse_c0Bse_inv[i]; c.KbtInv[i] = (T)p.Kbt_c0Bbt_inv[i];
        c.Bse[i] = (T)p.Bse[i]; c.Bbt[i] = (T)p.Bbt[i]; c.rhoJ[i] = (T)p.rhoJ[i];
        bool off = (i % 4) != 0;
        if (off && (p.Kse_c0Bse_inv[i] != 0 || p.Kbt_c0Bbt_inv[i] != 0 || p.Bbt[i] != 0 || p.rhoJ[i] != 0)) diag = false;
        if (p.Bse[i] != 0) diag = false;
    }
    for (int i = 0; i < 3; ++i) {
        c.KseVstar[i] = (T)p.Kse_vstar[i]; c.rhoAg[i] = (T)p.rhoAg[i]; c.C[i] = (T)p.C[i];
        c.Ftip[i] = (T)p.F_tip[i]; c.Mtip[i] = (T)p.M_tip[i]; c.p0[i] = (T)p.p0[i]; c.q0[i] = (T)p.q0[i]; c.w0[i] = (T)p.w0[i];
    }
    for (int i = 0; i < 4; ++i) c.h0[i] = (T)p.h0[i];
    for (int i = 0; i < 12; ++i) c.tdirs[i] = (T)p.tendon_dirs[i];
    c.N = p.N;
    c.diag = diag ? 1 : 0;
    return c;
}

// MLP view used by the SIMT evaluation paths.  Wp is a packed copy made by kc_pack_mlp: for hidden unit i,
// Wp[i*stride + 0..in) = W1[i][:], Wp[i*stride + inP] = b1[i], Wp[i*stride + inP + 4 + c] = W2[c][i] (c < 25),
// stride = inP + 32, inP = in_dim rounded up to 4 — so one hidden unit is a handful of 16-byte broadcast loads.
template <typename T>
struct MlpC {
    const T* Wp;
    const T* b2;
    int in_dim, inP, hidden, stride;
};

// Warp-cooperative MLP view (kc_mlp_coop.cuh): the 32 lanes of a warp hold the SAME rod and split the hidden units.
// Wc: one row of KC_COOP_ROW(in_dim) values per hidden unit (Hp = hidden rounded up to 32, zero rows as padding):
// [0, in_dim) W1[i][:], [inP] b1[i], [inP+1, inP+26) W2[:][i].  A lane reads its unit's row with 16-byte loads; the row
// length is a multiple of 4 with an ODD number of 16-byte words, so the 8 lanes of a 128-bit shared-memory phase hit
// distinct banks.  Passing an MlpCoop instead of an MlpC selects the cooperative evaluation by
// overload resolution; everything else of the per-rod code is unchanged (all lanes compute the physics redundantly).
constexpr int kc_coop_row(int in_dim) {
    const int n = ((((in_dim + 3) & ~3) + 26) + 3) & ~3;
    return ((n / 4) & 1) ? n : n + 4;
}
template <typename T>
struct MlpCoop : MlpC<T> {
    const T* Wc;
    int Hp;
};

void kc_set_error(const char* fmt, ...);
#define KC_CHECK_ARG(cond, ...)          \
    do {                                 \
        if (!(cond)) {                   \
            kc_set_error(__VA_ARGS__);   \
            return KC_EINVAL;            \
        }                                \
    } while (0)
#define KC_CHECK_LAUNCH(what)                                                         \
    do {                                                                              \
        cudaError_t e__ = cudaGetLastError();                                         \
        if (e__ != cudaSuccess) {                                                     \
            kc_set_error("%s: %s", what, cudaGetErrorString(e__));                    \
            return KC_ECUDA;                                                          \
        }                                                                             \
    } while (0)
