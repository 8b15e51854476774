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
// F2 (optional): a second right-hand side, solved with the same factorisation: X2 = -J^-1 (F2 - F(G)).
template <typename T>
KC_HD int wide_decide(const T Fall[7][6], T G[6], const T eps[6], T tol, const T* F2 = nullptr, T* X2 = nullptr) {
    const T fn = norm_inf6(Fall[0]);
    if (!(fn == fn)) return -1;
    if (fn <= tol * kc_max(T(1), norm_inf6(G))) return 1;
    T A[36], rhs[6], rhs2[6], ie[6], ip[6];
#pragma unroll
    for (int c = 0; c < 6; ++c) ie[c] = kc_rcp(eps[c]);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        rhs[i] = -Fall[0][i];
        rhs2[i] = F2 ? Fall[0][i] - F2[i] : T(0);
#pragma unroll
        for (int c = 0; c < 6; ++c) A[i * 6 + c] = (Fall[c + 1][i] - Fall[0][i]) * ie[c];
    }
    // Gaussian elimination without pivoting (J = -[[I,0],[X,I]] + small, cond ~ 1.5), then back substitution; branch
    // free (a bad pivot is only flagged) and with one fast reciprocal per pivot, reused by the back substitution
    bool bad = false;
#pragma unroll
    for (int p = 0; p < 6; ++p) {
        const T piv = A[p * 6 + p];
        bad = bad || !(kc_abs(piv) > T(1e-30));
        ip[p] = kc_rcp(piv);
#pragma unroll
        for (int r = p + 1; r < 6; ++r) {
            const T f = A[r * 6 + p] * ip[p];
#pragma unroll
            for (int c = p + 1; c < 6; ++c) A[r * 6 + c] -= f * A[p * 6 + c];
            rhs[r] -= f * rhs[p];
            if (F2) rhs2[r] -= f * rhs2[p];
        }
    }
    if (bad) return -1;
    T dG[6];
#pragma unroll
    for (int r = 5; r >= 0; --r) {
        T s = rhs[r], s2 = rhs2[r];
#pragma unroll
        for (int c = r + 1; c < 6; ++c) { s -= A[r * 6 + c] * dG[c]; if (F2) s2 -= A[r * 6 + c] * X2[c]; }
        dG[r] = s * ip[r];
        if (F2) X2[r] = s2 * ip[r];
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

// ---- wide mode with a linearised final correction ------------------------------------------------------------------
// Every joint march leaves the marched states of all 7 points in shared memory.  Once the Newton step dG is so small that
// the second-order remainder C*|dG|^2 is below the tolerance, the state at G + dG is formed as
//     state(G) + sum_c (state(G + eps_c e_c) - state(G)) * dG_c / eps_c
// instead of marching once more — the same first-order model Newton itself uses, applied to the whole rod.  C (the
// curvature of the residual map) is estimated from THIS step's own iterations: after a Newton step of size s the new
// residual is ~ C s^2 (so the correction can be used from the second joint march of a step on, never on stale data).  Returns 1 = converged at G (state = base lane's), 2 = accepted G + dG with the linearised state (w[c] =
// dG_c/eps_c returned), 0 = G advanced, keep marching, -1 = failure.
//
// Tension-aware predictor: on the FIRST joint march of a step the spare 8th lane marches the base point under the NEXT
// step's tendon load; Fnext is its residual.  D = -J^-1 (Fnext - F(G)) is the change of the root that the coming change
// of the tensions will cause (first order, this step's history) — the caller adds D(t) - D(t-1) to the linear
// extrapolation of G (a jump of the controls is anticipated instead of being discovered by Newton: 3.2 -> 3.0 joint
// marches per step on random tensions, 2.2 -> 2.0 on sines).  D is left untouched when Fnext == nullptr.
template <typename T>
KC_HD int wide_decide_lin(const T Fall[7][6], T G[6], const T eps[6], T tol, T& Cest, T& sprev, T w[6],
                          const T* Fnext = nullptr, T* D = nullptr) {
    const T fn = norm_inf6(Fall[0]);
    if (!(fn == fn)) return -1;
    const T scale = kc_max(T(1), norm_inf6(G));
    if (sprev > T(0)) Cest = kc_max(Cest, fn * kc_rcp(sprev * sprev));   // this residual is what the last step left behind
    if (!Fnext && fn <= tol * scale) return 1;
    T Gn[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) Gn[i] = G[i];
    const int r = wide_decide(Fall, Gn, eps, T(-1), Fnext, D);   // tol < 0: always factor and take the Newton step
    if (r < 0) return -1;
    if (fn <= tol * scale) return 1;                            // (first march: D was still wanted)
    T dG[6], s = T(0);
#pragma unroll
    for (int i = 0; i < 6; ++i) { dG[i] = Gn[i] - G[i]; s = kc_max(s, kc_abs(dG[i])); G[i] = Gn[i]; }
    sprev = s;
    if (Cest > T(0) && T(4) * Cest * s * s <= tol * scale) {
#pragma unroll
        for (int i = 0; i < 6; ++i) w[i] = dG[i] * kc_rcp(eps[i]);
        return 2;
    }
    return 0;
}

// Views in "output order" e = r*n + j (row r, node j: the order of one [25][N] slice of the reference trajectory).
// NC = compile-time node count (0: run-time n).  History element (slot s, node j) sits at s*n + j.
template <typename T, int NH, int LS, int NC> struct HistViewE {
    const T* p; int n;
    KC_HD void load(int j, T hist[NH]) const {
        const int N = NC ? NC : n;
        const T* hn = p + (size_t)j * LS;
#pragma unroll
        for (int s = 0; s < NH; ++s) hist[s] = hn[(size_t)s * N * LS];
    }
};

// Sink that keeps the marched state of one lane in shared memory, [25*n][slot stride SS] in output order.  Stores are
// unconditional: the kernel freezes the marched point of a finished rod, so its re-marches rewrite the same values.
template <typename T, int SS, int NC> struct SmemStateSinkE {
    T* p; int n;
    KC_HD void put(int j, const T y[19]) {
        const int N = NC ? NC : n;
        T* pn = p + (size_t)j * SS;
#pragma unroll
        for (int r = 0; r < 19; ++r) pn[(size_t)r * N * SS] = y[r];
    }
    KC_HD void putz(int j, const T z[6]) {
        const int N = NC ? NC : n;
        T* pn = p + (size_t)j * SS;
#pragma unroll
        for (int c = 0; c < 6; ++c) pn[(size_t)(19 + c) * N * SS] = z[c];
    }
};

struct WideNoSink {
    template <typename T> KC_HD void put(int, const T*) {}
    template <typename T> KC_HD void putz(int, const T*) {}
};
