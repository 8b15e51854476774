// kc_bptt_core.cuh — reverse mode THROUGH the time rollout (back-propagation through time), one rod per thread.
// This is the north-star extension K5: the reference never differentiates a rollout, so there is no
// reference implementation; the oracle is a finite-difference derivative of the fp64 rollout (tests/test_gpu_bptt.py).
//
// Forward step t -> t+1 (knode.py:70-100):  hist = c1 s_t + c2 s_{t-1};  find G with F(G; hist, tf, theta) = 0;  the march
// from base(G) IS the new state: y_j, z_j.  Reverse, given lambda = dL/ds_{t+1}:
//   1. adjoint march (nodes N-2..0) with cotangents lambda -> g_hist, g_tf, MLP samples (x, dL/do) and Gbar = dL/dG|explicit
//   2. implicit function theorem for the shooting solve: mu = -J^{-T} Gbar (J = dF/dG, central differences of the forward
//      march), then a second adjoint march with the tip cotangent -mu on (n_L, m_L) adds the implicit part
//   3. g_hist feeds lambda of the two previous states (c1, c2).
// The trajectory itself is the checkpoint (every state is stored), only per-node intermediates are recomputed.
// MLP weight gradients are NOT accumulated here: every node evaluation emits one (x, dL/do) sample per adjoint march and
// the sample-reduction kernels of the training step (kc_mlp_bwd) turn them into gW1, gb1, gW2, gb2.
#pragma once
#include "kc_rollout_core.cuh"
#include "kc_adjoint.cuh"

// Node evaluation in reverse: cotangents cys[19] (of ys incl. the MLP residual) and cz[6] (of the corrected z).
// hist layout as in node_eval (NH = 12: qh,wh,vh,uh; NH = 25: yh(19), zh(6)).  xs[IN] / gos[25] receive the MLP sample.
template <typename T, bool DIAG, int IN, int NH, typename MLP>
KC_HD void node_vjp(const RodC<T>& P, const MLP& M, const T* __restrict__ y, const T* __restrict__ hist,
                    const T tf[3], const T* __restrict__ cys, const T* __restrict__ cz, T* __restrict__ gy,
                    T* __restrict__ ghist, T gtf[3], T* __restrict__ xs, T* __restrict__ gos) {
    const T* qh = (NH == 12) ? hist : hist + 13;
    const T* wh = qh + 3;
    const T* vh = (NH == 12) ? hist + 6 : hist + 19;
    const T* uh = vh + 3;
    T czt[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) czt[i] = cz[i];
#pragma unroll
    for (int s = 0; s < NH; ++s) ghist[s] = T(0);
    T gy_nn[19], gtf_nn[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int i = 0; i < 19; ++i) gy_nn[i] = T(0);
    if (IN > 0) {
        constexpr int INX = IN > 0 ? IN : 28;
        T ys0[19], z0[6], x[INX], go[25], gx[INX];
        rod_ode<T, DIAG>(P, y, qh, wh, vh, uh, tf, ys0, z0);
        if (IN == 28) {
#pragma unroll
            for (int i = 0; i < 19; ++i) x[i] = y[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) x[19 + i] = z0[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) x[25 + i] = tf[i];
        } else {
#pragma unroll
            for (int i = 0; i < 19; ++i) { x[i] = y[i]; x[19 + i] = hist[i % NH]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { x[(38 + i) % INX] = z0[i]; x[(44 + i) % INX] = hist[(19 + i) % NH]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) x[(50 + i) % INX] = tf[i];
        }
#pragma unroll
        for (int i = 0; i < 19; ++i) go[i] = cys[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) go[19 + i] = cz[i];
        if (xs) {
#pragma unroll
            for (int i = 0; i < INX; ++i) xs[i] = x[i];
#pragma unroll
            for (int i = 0; i < 25; ++i) gos[i] = go[i];
        }
        mlp_input_vjp<T, INX>(M, x, go, gx);
        if (IN == 28) {
#pragma unroll
            for (int i = 0; i < 19; ++i) gy_nn[i] = gx[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) czt[i] += gx[19 + i];
#pragma unroll
            for (int i = 0; i < 3; ++i) gtf_nn[i] = gx[25 + i];
        } else {
#pragma unroll
            for (int i = 0; i < 19; ++i) { gy_nn[i] = gx[i]; ghist[i % NH] = gx[19 + i]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { czt[i] += gx[(38 + i) % INX]; ghist[(19 + i) % NH] = gx[(44 + i) % INX]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) gtf_nn[i] = gx[(50 + i) % INX];
        }
    }
    T gqh[3], gwh[3], gvh[3], guh[3];
    rod_ode_vjp<T, DIAG>(P, y, qh, wh, vh, uh, tf, cys, czt, gy, gqh, gwh, gvh, guh, gtf);
    const int oq = (NH == 12) ? 0 : 13, ov = (NH == 12) ? 6 : 19;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        ghist[oq + i] += gqh[i]; ghist[oq + 3 + i] += gwh[i]; ghist[ov + i] += gvh[i]; ghist[ov + 3 + i] += guh[i];
        gtf[i] += gtf_nn[i];
    }
#pragma unroll
    for (int i = 0; i < 19; ++i) gy[i] += gy_nn[i];
}

struct NullSink {
    template <typename T> KC_HD void put(int, const T*) {}
    template <typename T> KC_HD void putz(int, const T*) {}
};

// One rod, all steps in reverse.
//   traj_b, gtraj_b : this rod's [T][25][N] blocks (reference layout) of the forward trajectory and of dL/dtraj
//   ten             : tensions [T][4];  gten (may be null): dL/dtensions [T][4]
//   Hs              : per-rod scratch, 4 arrays of NH*(N-1) values with element stride LS: Hcur | Ha | Hb | Hc
//   xs, gos         : MLP samples of this rod: [(T-1)*(N-1)*2][IN] and [..][25] (null when IN == 0)
template <typename T, bool DIAG, int IN, int NH, int LS, typename MLP>
KC_HD void bptt_rod(const RodC<T>& P, const MLP& M, const T* __restrict__ traj_b, const T* __restrict__ gtraj_b,
                    const T* __restrict__ ten, T* __restrict__ gten, int T_, T* Hs, T* __restrict__ xs,
                    T* __restrict__ gos, T fd_eps) {
    const int N = P.N, Nm1 = N - 1, HN = NH * Nm1;
    T* Hcur = Hs;
    T* Ha = Hs + (size_t)HN * LS;       // g_hist of step t+1
    T* Hb = Hs + (size_t)2 * HN * LS;   // g_hist of step t+2
    T* Hc = Hs + (size_t)3 * HN * LS;   // g_hist of step t (being built)
    for (int e = 0; e < HN; ++e) { Ha[(size_t)e * LS] = T(0); Hb[(size_t)e * LS] = T(0); }
    const size_t ts = (size_t)25 * N;
    for (int t = T_ - 2; t >= 0; --t) {
        const T* s1 = traj_b + (size_t)(t + 1) * ts;              // new state (the march's y_j, z_j)
        const T* s0 = traj_b + (size_t)t * ts;
        const T* sm = traj_b + (size_t)(t > 0 ? t - 1 : 0) * ts;
        const T* lam = gtraj_b + (size_t)(t + 1) * ts;
        T tn[4], tf[3];
#pragma unroll
        for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
        tendon_force(P, tn, tf);
        for (int j = 0; j < Nm1; ++j) {   // history of this step: [node][slot]
#pragma unroll
            for (int s = 0; s < NH; ++s) {
                const int row = slot_row<NH>(s);
                Hcur[(size_t)(j * NH + s) * LS] = P.c1 * s0[row * N + j] + P.c2 * sm[row * N + j];
            }
        }
        HistView<T, NH, LS> H{Hcur};
        // ---- J = dF/dG by central differences of the forward march at the converged G ----
        T G[6], Jm[36];
#pragma unroll
        for (int i = 0; i < 6; ++i) G[i] = s1[(7 + i) * N];
#pragma unroll 1
        for (int c = 0; c < 6; ++c) {
            T Gp[6], Gn[6], Fp[6], Fn[6];
            T gc = T(0);
#pragma unroll
            for (int i = 0; i < 6; ++i) if (i == c) gc = G[i];
            const T e = fd_eps * kc_max(T(1), kc_abs(gc));
#pragma unroll
            for (int i = 0; i < 6; ++i) { Gp[i] = G[i] + (i == c ? e : T(0)); Gn[i] = G[i] - (i == c ? e : T(0)); }
            NullSink S0;
            rod_march<T, DIAG, IN, NH>(P, M, Gp, tf, H, S0, Fp);
            rod_march<T, DIAG, IN, NH>(P, M, Gn, tf, H, S0, Fn);
            const T ie = T(1) / (T(2) * e);
#pragma unroll
            for (int i = 0; i < 6; ++i) {
#pragma unroll
                for (int cc = 0; cc < 6; ++cc) if (cc == c) Jm[i * 6 + cc] = (Fp[i] - Fn[i]) * ie;
            }
        }
        inv6(Jm);   // Jm <- J^{-1}
        // ---- two adjoint marches ----
        T gtf_acc[3] = {T(0), T(0), T(0)};
        T mu[6] = {T(0), T(0), T(0), T(0), T(0), T(0)};
#pragma unroll 1
        for (int pass = 0; pass < 2; ++pass) {
            T yb[19];   // cotangent of y_{j+1}
#pragma unroll
            for (int r = 0; r < 19; ++r) yb[r] = pass == 0 ? lam[r * N + (N - 1)] : T(0);
            if (pass == 1) {
#pragma unroll
                for (int i = 0; i < 6; ++i) yb[7 + i] = -mu[i];
            }
            for (int j = Nm1 - 1; j >= 0; --j) {
                T y[19], hist[NH], cys[19], cz[6], gy[19], gh[NH], gtf[3];
#pragma unroll
                for (int r = 0; r < 19; ++r) { y[r] = s1[r * N + j]; cys[r] = P.ds * yb[r]; }
                H.load(j, hist);
#pragma unroll
                for (int c = 0; c < 6; ++c) {
                    T v = T(0);
                    if (pass == 0) {
                        v = lam[(19 + c) * N + j];
                        const int s = (NH == 12) ? 6 + c : 19 + c;   // z rows are history rows of later steps
                        v += P.c1 * Ha[(size_t)(j * NH + s) * LS] + P.c2 * Hb[(size_t)(j * NH + s) * LS];
                    }
                    cz[c] = v;
                }
                T* xq = nullptr;
                T* gq = nullptr;
                if (IN > 0 && xs) {
                    const size_t q = ((size_t)t * Nm1 + j) * 2 + pass;
                    xq = xs + q * (IN > 0 ? IN : 1);
                    gq = gos + q * 25;
                }
                node_vjp<T, DIAG, IN, NH>(P, M, y, hist, tf, cys, cz, gy, gh, gtf, xq, gq);
#pragma unroll
                for (int i = 0; i < 3; ++i) gtf_acc[i] += gtf[i];
#pragma unroll
                for (int s = 0; s < NH; ++s) {
                    T* hc = Hc + (size_t)(j * NH + s) * LS;
                    *hc = (pass == 0 ? T(0) : *hc) + gh[s];
                }
                // cotangent of y_j: direct (loss + later histories) + pass-through of the Euler update + node Jacobian
#pragma unroll
                for (int r = 0; r < 19; ++r) {
                    T d = T(0);
                    if (pass == 0) {
                        d = lam[r * N + j];
                        const int s = (NH == 12) ? r - 13 : r;
                        if (s >= 0) d += P.c1 * Ha[(size_t)(j * NH + s) * LS] + P.c2 * Hb[(size_t)(j * NH + s) * LS];
                    }
                    yb[r] = d + yb[r] + gy[r];
                }
            }
            if (pass == 0) {   // mu = -J^{-T} Gbar,  Gbar = cotangent of (n0, m0) = yb[7:13]
#pragma unroll
                for (int i = 0; i < 6; ++i) {
                    T a = T(0);
#pragma unroll
                    for (int kx = 0; kx < 6; ++kx) a += Jm[kx * 6 + i] * yb[7 + kx];
                    mu[i] = -a;
                }
            }
        }
        if (gten) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                gten[(size_t)t * 4 + i] = P.tdirs[i * 3] * gtf_acc[0] + P.tdirs[i * 3 + 1] * gtf_acc[1] + P.tdirs[i * 3 + 2] * gtf_acc[2];
        }
        // rotate the history-cotangent buffers: Hb <- Ha, Ha <- Hc
        T* tmp = Hb; Hb = Ha; Ha = Hc; Hc = tmp;
    }
    if (gten) {
#pragma unroll
        for (int i = 0; i < 4; ++i) gten[(size_t)(T_ - 1) * 4 + i] = T(0);   // the last control is never applied (knode.py:102)
    }
}
