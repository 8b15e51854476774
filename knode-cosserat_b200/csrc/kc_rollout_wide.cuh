// kc_rollout_wide.cuh — "wide" shooting solve for SMALL batches: 8 lanes cooperate on one rod.
// Lane k = 0 marches the base point G, lanes k = 1..6 march G + eps_k e_k in the very same instruction stream, so one
// joint march yields the residual AND a fresh finite-difference Jacobian: Newton converges quadratically (≈3 joint
// marches per step on C2 instead of ≈7.5 sequential Broyden marches per warp).  It spends 8x the lanes per rod, which
// are idle anyway when B is a few thousand; for large B the one-rod-per-lane Broyden kernel is ≈3x cheaper and is used.
// The decision logic is __host__ __device__ (shared with tests/emul); only the lane exchange is device-specific.
#pragma once
#include "kc_rollout_core.cuh"

template <typename T> KC_HD void wide_eps(const T G[6], T fd_eps, T eps[6]) {
#pragma unroll
    for (int c = 0; c < 6; ++c) eps[c] = fd_eps * kc_max(T(1), kc_abs(G[c]));
}

// Fall[k][i]: residual i of lane k (k = 0 base, k = c+1 perturbed in component c).
// Returns 1 = converged at G (the base lane's march is the accepted state), 0 = G advanced by a Newton step,
// -1 = failure (NaN or singular Jacobian).
template <typename T> KC_HD int wide_decide(const T Fall[7][6], T G[6], const T eps[6], T tol) {
    const T fn = norm_inf6(Fall[0]);
    if (!(fn == fn)) return -1;
    if (fn <= tol * kc_max(T(1), norm_inf6(G))) return 1;
    T A[36], rhs[6], ie[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ie[c] = T(1) / eps[c];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        rhs[i] = -Fall[0][i];
#pragma unroll
        for (int c = 0; c < 6; ++c) A[i * 6 + c] = (Fall[c + 1][i] - Fall[0][i]) * ie[c];
    }
    // Gaussian elimination without pivoting (J = -[[I,0],[X,I]] + small, cond ~ 1.5), then back substitution
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const T piv = A[p * 6 + p];
        if (!(kc_abs(piv) > T(1e-30))) return -1;
        const T ip = T(1) / piv;
#pragma unroll
        for (int r = p + 1; r < 6; ++r) {
            const T f = A[r * 6 + p] * ip;
#pragma unroll
            for (int c = p + 1; c < 6; ++c) A[r * 6 + c] -= f * A[p * 6 + c];
            rhs[r] -= f * rhs[p];
        }
    }
    T dG[6];
#pragma unroll
    for (int r = 5; r >= 0; --r) {
        T s = rhs[r];
#pragma unroll
        for (int c = r + 1; c < 6; ++c) s -= A[r * 6 + c] * dG[c];
        dG[r] = s / A[r * 6 + r];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) G[i] += dG[i];
    return 0;
}

// Trajectory sink with a per-lane on/off switch (only the base lane of a not-yet-converged rod stores).  Besides the
// trajectory it keeps the history-relevant rows of the candidate state in shared memory (cand: [N-1][NH][CS]) so the next
// step's history is formed without reading the trajectory back from L2.
template <typename T, int LS, int NH, int CS> struct TrajSinkPred {
    T* p; T* cand; bool on; int nm1;  // cand covers nodes 0..nm1-1 (the tip node has no history)
    KC_HD void put(int j, const T y[19]) {
        if (on) {
            T* pn = p + (size_t)j * 25 * LS;
#pragma unroll
            for (int r = 0; r < 19; ++r) pn[r * LS] = y[r];
            if (cand && j < nm1) {
                T* cn = cand + (size_t)j * NH * CS;
                if (NH == 12) {
#pragma unroll
                    for (int s = 0; s < 6; ++s) cn[s * CS] = y[13 + s];
                } else {
#pragma unroll
                    for (int s = 0; s < 19; ++s) cn[(s % NH) * CS] = y[s];
                }
            }
        }
    }
    KC_HD void putz(int j, const T z[6]) {
        if (on) {
            T* pn = p + (size_t)j * 25 * LS;
#pragma unroll
            for (int c = 0; c < 6; ++c) pn[(19 + c) * LS] = z[c];
            if (cand) {
                T* cn = cand + (size_t)j * NH * CS;
#pragma unroll
                for (int c = 0; c < 6; ++c) cn[((NH == 12 ? 6 : 19) + c) % NH * CS] = z[c];
            }
        }
    }
};
