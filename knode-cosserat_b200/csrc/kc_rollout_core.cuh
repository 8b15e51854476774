// kc_rollout_core.cuh — the per-rod time loop of knode.simulate (knode.py:55-102), written once and used by the
// rollout kernel (one rod per thread).  Layouts:
//   trajD : "device layout" trajectory [T][25*N][Bpad], rod index fastest, so every store of a warp is one coalesced
//           line; a transpose kernel (kc_rollout.cu) turns it into the reference's [B][T][rows][N].
//   Hs    : per-rod BDF2 history for the step being solved, [NH*(N-1)] values with element stride hs
//           (shared memory on the device: [slot*(N-1)+node][lane]).
#pragma once
#include "kc_rod.cuh"

template <int NH> KC_HD int slot_row(int s) { return NH == 12 ? s + 13 : s; }

template <typename T, int NH> struct HistView {
    const T* p; int hs; int Nm1;
    KC_HD void load(int j, T hist[NH]) const {
#pragma unroll
        for (int s = 0; s < NH; ++s) hist[s] = p[(size_t)(s * Nm1 + j) * hs];
    }
};

template <typename T> struct TrajSink {
    T* p; size_t os; int N;  // p -> element (t, k=0, b)
    KC_HD void put(int j, const T y[19]) {
#pragma unroll
        for (int r = 0; r < 19; ++r) p[(size_t)(r * N + j) * os] = y[r];
    }
    KC_HD void putz(int j, const T z[6]) {
#pragma unroll
        for (int c = 0; c < 6; ++c) p[(size_t)((19 + c) * N + j) * os] = z[c];
    }
};

// History for the next step from the two most recent states: H = c1*state[t] + c2*state[t-1] (knode.py:74-75).
template <typename T, int NH>
KC_HD void build_history(const RodC<T>& P, const T* cur, const T* prev, size_t os, T* Hs, int hs) {
    const int N = P.N, Nm1 = N - 1;
    for (int s = 0; s < NH; ++s) {
        const int row = slot_row<NH>(s);
        for (int j = 0; j < Nm1; ++j) {
            const size_t k = (size_t)(row * N + j) * os;
            Hs[(size_t)(s * Nm1 + j) * hs] = P.c1 * cur[k] + P.c2 * prev[k];
        }
    }
}

// Steps t = t_begin .. t_end-1 of one rod; step t reads tensions[t] and produces state t+1.
//   ten      : this rod's tensions, ten[t*4 + i]
//   trajD_b  : &trajD[0][0][b]; stride between k is os (= Bpad), between t is 25*N*os
//   Gout/iters: this rod's [T][6] / [T] output rows (reference layout) or nullptr
template <typename T, bool DIAG, int IN, int NH>
KC_HD void rollout_rod(const RodC<T>& P, const MlpC<T>& M, const ShootMem<T>& st, const T* __restrict__ ten,
                       T* trajD_b, size_t os, T* Hs, int hs, int t_begin, int t_end, T tol,
                       int max_iter, T fd_eps, T* Gout, int32_t* iters) {
    const int N = P.N;
    const size_t tstride = (size_t)25 * N * os;
    // (re)build the history of step t_begin from states t_begin and t_begin-1 (state[-1] := state[0], knode.py:65-66)
    {
        const T* cur = trajD_b + (size_t)t_begin * tstride;
        const T* prev = trajD_b + (size_t)(t_begin > 0 ? t_begin - 1 : 0) * tstride;
        build_history<T, NH>(P, cur, prev, os, Hs, hs);
    }
    for (int t = t_begin; t < t_end; ++t) {
        T tn[4], tf[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
        tendon_force(P, tn, tf);
        T* cur = trajD_b + (size_t)t * tstride;
        T* nxt = cur + tstride;
        HistView<T, NH> H{Hs, hs, N - 1};
        TrajSink<T> S{nxt, os, N};
        const int it = shoot_step<T, DIAG, IN, NH>(P, M, st, tf, H, S, tol, max_iter, fd_eps);
        // z[:, N-1] is never written by the march: it keeps its previous value (cosserat_ode.py:198-201)
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            const size_t k = (size_t)((19 + c) * N + (N - 1)) * os;
            nxt[k] = cur[k];
        }
        if (Gout) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Gout[(size_t)(t + 1) * 6 + i] = st.G(i);
        }
        if (iters) iters[t + 1] = it;
        build_history<T, NH>(P, nxt, cur, os, Hs, hs);
    }
}

// Initial state (index 0 of the trajectory): straight rod of knode.py:58-64 or the caller's y0/z0 ([19][N], [6][N]).
template <typename T>
KC_HD void rollout_init(const RodC<T>& P, const T* y0, const T* z0, T* trajD_b, size_t os) {
    const int N = P.N;
    for (int j = 0; j < N; ++j) {
        for (int r = 0; r < 19; ++r) {
            T v;
            if (y0) v = y0[r * N + j];
            else v = (r == 2) ? P.ds * T(j) : (r == 3 ? T(1) : T(0));  // linspace(0, L, N), L = ds*(N-1)
            trajD_b[(size_t)(r * N + j) * os] = v;
        }
        for (int c = 0; c < 6; ++c) {
            T v;
            if (z0) v = z0[c * N + j];
            else v = (c == 2) ? T(1) : T(0);
            trajD_b[(size_t)((19 + c) * N + j) * os] = v;
        }
    }
}
