// kc_train_prep.cuh — one teacher-forced training sample from the ground-truth trajectory (physics_train.py:313-344): shared by
// the prep kernels of kc_train.cu and by the training kernel that forms its own samples (kc_train_tc3.cu).
#pragma once
#include "kc_rod.cuh"

struct KeyIdx64 { int32_t k[64]; };

// One (b, t, key) sample: ODE at node kn-1 of the NEXT ground-truth state -> x[XPG], phys[25], tgt[25] (unit stride).
// nxt / cur / prv: the [25][N] states t+1, t, t-1 of the trajectory (global or shared memory).
template <typename T, bool DIAG, int IN>
KC_D void prep_sample(const RodC<T>& P, int kn, const T* nxt, const T* cur, const T* prv, const T* tn4, T* x, T* ph, T* tg) {
    const int N = P.N, j = kn - 1;
    T y[19], hist[25], tn[4], tf[3], ys[19], z[6];
#pragma unroll
    for (int r = 0; r < 19; ++r) y[r] = nxt[r * N + j];
#pragma unroll
    for (int r = 0; r < 25; ++r) hist[r] = P.c1 * cur[r * N + j] + P.c2 * prv[r * N + j];
#pragma unroll
    for (int i = 0; i < 4; ++i) tn[i] = tn4[i];
    tendon_force(P, tn, tf);
    rod_ode<T, DIAG>(P, y, hist + 13, hist + 16, hist + 19, hist + 22, tf, ys, z);
    if (IN == 28) {
#pragma unroll
        for (int i = 0; i < 19; ++i) x[i] = y[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) x[19 + i] = z[i];
#pragma unroll
        for (int i = 0; i < 3; ++i) x[25 + i] = tf[i];
#pragma unroll
        for (int i = 28; i < 32; ++i) x[i] = T(0);
    } else {
#pragma unroll
        for (int i = 0; i < 19; ++i) { x[i] = y[i]; x[19 + i] = hist[i]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) { x[38 + i] = z[i]; x[44 + i] = hist[19 + i]; }
#pragma unroll
        for (int i = 0; i < 3; ++i) x[50 + i] = tf[i];
#pragma unroll
        for (int i = 53; i < 56; ++i) x[i] = T(0);
    }
#pragma unroll
    for (int r = 0; r < 19; ++r) { ph[r] = y[r] + P.ds * ys[r]; tg[r] = nxt[r * N + kn]; }
#pragma unroll
    for (int c = 0; c < 6; ++c) { ph[19 + c] = z[c]; tg[19 + c] = nxt[(19 + c) * N + kn - 1]; }
}

