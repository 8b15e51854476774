// kc_rollout_core.cuh — the per-rod time loop of knode.simulate (knode.py:55-102), written once and used by the
// rollout kernel (one rod per thread).  Every per-rod array has a COMPILE-TIME lane stride LS (32 on the device, 1 in the
// host test harness) and is node-major, so all loads/stores of the march use immediate offsets from one running pointer
// per node (no integer address arithmetic in the hot loop):
//   trajD : "device layout" trajectory, tiled by 32 rods: [tile][T][N][25][LS]; a rod's element (t, node j, row r) is
//           at ((t*N + j)*25 + r)*LS from its base, so each store of a warp is one coalesced 128-byte line; a transpose
//           kernel (kc_rollout.cu) turns it into the reference's [B][T][rows][N].
//   Hs    : per-rod BDF2 history for the step being solved, [N-1][NH][LS] (shared memory on the device).
#pragma once
#include "kc_rod.cuh"

template <int NH> KC_HD int slot_row(int s) { return NH == 12 ? s + 13 : s; }

template <typename T, int NH, int LS> struct HistView {
    const T* p;
    KC_HD void load(int j, T hist[NH]) const {
        const T* hn = p + (size_t)j * NH * LS;
#pragma unroll
        for (int s = 0; s < NH; ++s) hist[s] = hn[s * LS];
    }
};

template <typename T, int LS> struct TrajSink {
    T* p;  // -> element (t, node 0, row 0) of this rod
    KC_HD void put(int j, const T y[19]) {
        T* pn = p + (size_t)j * 25 * LS;
#pragma unroll
        for (int r = 0; r < 19; ++r) pn[r * LS] = y[r];
    }
    KC_HD void putz(int j, const T z[6]) {
        T* pn = p + (size_t)j * 25 * LS;
#pragma unroll
        for (int c = 0; c < 6; ++c) pn[(19 + c) * LS] = z[c];
    }
};

// History for the next step from the two most recent states: H = c1*state[t] + c2*state[t-1] (knode.py:74-75).
// LS = element stride of the trajectory, LSM = element stride of the history (they differ in the warp-cooperative kernels,
// where a warp owns ONE rod: device-layout trajectory, lane stride 32, but private per-warp scratch, stride 1)
// `nodes`: N-1 for the Euler march, N for RK4 (k4 of the last interval reads the history of node N-1).
template <typename T, int NH, int LS, int LSM = LS>
KC_HD void build_history(const RodC<T>& P, const T* cur, const T* prev, T* Hs, int nodes = -1) {
    const int Nm1 = nodes < 0 ? P.N - 1 : nodes;
    for (int j = 0; j < Nm1; ++j) {
        const T* cn = cur + (size_t)j * 25 * LS;
        const T* pn = prev + (size_t)j * 25 * LS;
        T* hn = Hs + (size_t)j * NH * LSM;
#pragma unroll
        for (int s = 0; s < NH; ++s) hn[s * LSM] = P.c1 * cn[slot_row<NH>(s) * LS] + P.c2 * pn[slot_row<NH>(s) * LS];
    }
}

// Steps t = t_begin .. t_end-1 of one rod; step t reads tensions[t] and produces state t+1.
//   ten      : this rod's tensions, ten[t*4 + i]
//   traj_b   : this rod's base in trajD (element t=0, node 0, row 0); time stride is N*25*LS
//   Gout/iters: this rod's [T][6] / [T] output rows (reference layout) or nullptr
template <typename T, bool DIAG, int IN, int NH, int LS, int LSM = LS, int METHOD = KC_MARCH_EULER, typename MLP>
KC_HD void rollout_rod(const RodC<T>& P, const MLP& M, const ShootMem<T, LSM>& st, const T* __restrict__ ten,
                       T* traj_b, T* Hs, int t_begin, int t_end, T tol, int max_iter, T fd_eps, T* Gout,
                       int32_t* iters) {
    const int N = P.N;
    const int hn = METHOD == KC_MARCH_RK4 ? N : N - 1;   // nodes with a history entry
    const size_t tstride = (size_t)25 * N * LS;
    // (re)build the history of step t_begin from states t_begin and t_begin-1 (state[-1] := state[0], knode.py:65-66)
    build_history<T, NH, LS, LSM>(P, traj_b + (size_t)t_begin * tstride,
                             traj_b + (size_t)(t_begin > 0 ? t_begin - 1 : 0) * tstride, Hs, hn);
    for (int t = t_begin; t < t_end; ++t) {
        T tn[4], tf[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
        tendon_force(P, tn, tf);
        T* cur = traj_b + (size_t)t * tstride;
        T* nxt = cur + tstride;
        HistView<T, NH, LSM> H{Hs};
        TrajSink<T, LS> S{nxt};
        const int it = shoot_step<T, DIAG, IN, NH, LSM, METHOD>(P, M, st, tf, H, S, tol, max_iter, fd_eps);
        // z[:, N-1] is never written by the march: it keeps its previous value (cosserat_ode.py:198-201)
        {
            const size_t o = (size_t)(N - 1) * 25 * LS;
#pragma unroll
            for (int c = 0; c < 6; ++c) nxt[o + (19 + c) * LS] = cur[o + (19 + c) * LS];
        }
        if (Gout) {
#pragma unroll
            for (int i = 0; i < 6; ++i) Gout[(size_t)(t + 1) * 6 + i] = st.G(i);
        }
        if (iters) iters[t + 1] = it;
        build_history<T, NH, LS, LSM>(P, nxt, cur, Hs, hn);
    }
}

// Initial state (index 0 of the trajectory): straight rod of knode.py:58-64 or the caller's y0/z0 ([19][N], [6][N]).
template <typename T, int LS>
KC_HD void rollout_init(const RodC<T>& P, const T* y0, const T* z0, T* traj_b) {
    const int N = P.N;
    for (int j = 0; j < N; ++j) {
        T* pn = traj_b + (size_t)j * 25 * LS;
        for (int r = 0; r < 19; ++r) {
            T v;
            if (y0) v = y0[r * N + j];
            else v = (r == 2) ? P.ds * T(j) : (r == 3 ? T(1) : T(0));  // linspace(0, L, N), L = ds*(N-1)
            pn[r * LS] = v;
        }
        for (int c = 0; c < 6; ++c) {
            T v;
            if (z0) v = z0[c * N + j];
            else v = (c == 2) ? T(1) : T(0);
            pn[(19 + c) * LS] = v;
        }
    }
}
