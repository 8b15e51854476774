// kc_rod.cuh — per-rod arithmetic shared by every kernel: the Cosserat node ODE, the KNODE MLP residual (SIMT form),
// the spatial march and the quasi-Newton shooting step.  All functions are per-thread (one rod or one node sample
// per thread), hold their state in registers and are __host__ __device__ so tests/emul can run the very same solver
// logic on the CPU of the build container (there is no GPU there); the product only ever calls them from kernels.
//
// Reference semantics: knode_cosserat/cosserat_ode_torch.py:153-212 (== cosserat_ode.py:129-184).
#pragma once
#include "kc_common.cuh"

// |x| as fabs: an operand modifier on the device (the compare + select form cost two instructions, 2.8 % of the rollout kernel)
KC_HD float kc_abs(float x) { return fabsf(x); }
KC_HD double kc_abs(double x) { return fabs(x); }
template <typename T> KC_HD T kc_max(T a, T b) { return a > b ? a : b; }

// e^x on the device: ONE multiply and ONE MUFU.EX2 (ex2.approx.ftz, rel. error 2^-22).  __expf without -use_fast_math
// wraps the same instruction in denormal-range handling (~12 instructions): measured as half of all instructions of the
// tensor-core training kernel.
KC_HD float kc_exp_fast(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * 1.4426950408889634f));
    return r;
#else
    return expf(x);
#endif
}
KC_HD float kc_elu(float x) {
#if defined(__CUDA_ARCH__)
    return x > 0.f ? x : (kc_exp_fast(x) - 1.f);
#else
    return x > 0.f ? x : expm1f(x);
#endif
}
KC_HD double kc_elu(double x) { return x > 0.0 ? x : expm1(x); }
// ELU'(x) given x (pre-activation)
KC_HD float kc_elu_grad(float x) {
#if defined(__CUDA_ARCH__)
    return x > 0.f ? 1.f : kc_exp_fast(x);
#else
    return x > 0.f ? 1.f : expf(x);
#endif
}
KC_HD double kc_elu_grad(double x) { return x > 0.0 ? 1.0 : exp(x); }

template <typename T> KC_HD void kc_cross(const T a[3], const T b[3], T c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}
template <typename T> KC_HD void kc_mv(const T M[9], const T x[3], T r[3]) {
    r[0] = M[0] * x[0] + M[1] * x[1] + M[2] * x[2];
    r[1] = M[3] * x[0] + M[4] * x[1] + M[5] * x[2];
    r[2] = M[6] * x[0] + M[7] * x[1] + M[8] * x[2];
}
template <typename T> KC_HD void kc_mtv(const T M[9], const T x[3], T r[3]) {  // M^T x
    r[0] = M[0] * x[0] + M[3] * x[1] + M[6] * x[2];
    r[1] = M[1] * x[0] + M[4] * x[1] + M[7] * x[2];
    r[2] = M[2] * x[0] + M[5] * x[1] + M[8] * x[2];
}

// 2/x.  fp32 on the device: MUFU.RCP + one Newton step (≈1 ulp) instead of the IEEE division sequence, whose slow-path
// CALL splits the node evaluation's basic block in two and sits at the head of its critical path.
KC_HD float kc_two_over(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    r = fmaf(r, fmaf(-x, r, 1.0f), r);
    return r + r;
#else
    return 2.0f / x;
#endif
}
KC_HD double kc_two_over(double x) { return 2.0 / x; }

// 1/x for the small dense solves of the shooting iteration (same MUFU + Newton form; a Newton / Broyden step does not
// need a correctly rounded quotient and the IEEE sequence serialises the 6x6 elimination).
KC_HD float kc_rcp(float x) {
#if defined(__CUDA_ARCH__)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
    return 1.0f / x;
#endif
}
KC_HD double kc_rcp(double x) { return 1.0 / x; }

// Eq(10) quaternion (w,x,y,z) -> rotation, NOT normalised (cosserat_ode_torch.py:157-161).
template <typename T> KC_HD void kc_quat_R(const T h[4], T R[9]) {
    const T a = h[0], b = h[1], c = h[2], d = h[3];
    const T s = kc_two_over(a * a + b * b + c * c + d * d);
    R[0] = T(1) + s * (-c * c - d * d); R[1] = s * (b * c - d * a);        R[2] = s * (b * d + c * a);
    R[3] = s * (b * c + d * a);        R[4] = T(1) + s * (-b * b - d * d); R[5] = s * (c * d - b * a);
    R[6] = s * (b * d - c * a);        R[7] = s * (c * d + b * a);        R[8] = T(1) + s * (-b * b - c * c);
}

// One node of the rod ODE, physics only.  y[19] (p is not read), histories qh,wh (= yh[13:19]) and vh,uh (= zh),
// tf = global-frame distributed tendon force.  Outputs ys[19], z[6] = [v;u].
template <typename T, bool DIAG>
KC_HD void rod_ode(const RodC<T>& P, const T* __restrict__ y, const T qh[3], const T wh[3], const T vh[3],
                   const T uh[3], const T tf[3], T* __restrict__ ys, T* __restrict__ z) {
    const T* h = y + 3; const T* n = y + 7; const T* m = y + 10; const T* q = y + 13; const T* w = y + 16;
    T R[9];
    kc_quat_R(h, R);
    // Eq(6) solved constitutive law (:164-165)
    T t1[3], t2[3], v[3], u[3];
    kc_mtv(R, n, t1);
    kc_mtv(R, m, t2);
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            v[i] = P.KseInv[4 * i] * (t1[i] + P.KseVstar[i]);
            u[i] = P.KbtInv[4 * i] * (t2[i] - P.Bbt[4 * i] * uh[i]);
        }
    } else {
        T b1[3], b2[3];
        kc_mv(P.Bse, vh, b1);
        kc_mv(P.Bbt, uh, b2);
#pragma unroll
        for (int i = 0; i < 3; ++i) { t1[i] = t1[i] + P.KseVstar[i] - b1[i]; t2[i] = t2[i] - b2[i]; }
        kc_mv(P.KseInv, t1, v);
        kc_mv(P.KbtInv, t2, u);
    }
    // Eq(5) BDF2 time derivatives (:169-171)
    T vt[3], ut[3], qt[3], wt[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        vt[i] = P.c0 * v[i] + vh[i]; ut[i] = P.c0 * u[i] + uh[i];
        qt[i] = P.c0 * q[i] + qh[i]; wt[i] = P.c0 * w[i] + wh[i];
    }
    // Eq(3) weight + square-law drag + tendon load (:174)
    T dr[3], Rd[3], f[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) dr[i] = P.C[i] * q[i] * kc_abs(q[i]);
    kc_mv(R, dr, Rd);
#pragma unroll
    for (int i = 0; i < 3; ++i) f[i] = P.rhoAg[i] - Rd[i] + tf[i];
    // Eq(7) (:177-181)
    T* ps = ys; T* hs = ys + 3; T* ns = ys + 7; T* ms = ys + 10; T* qs = ys + 13; T* ws = ys + 16;
    kc_mv(R, v, ps);
    T wq[3], tmp[3], Rt[3];
    kc_cross(w, q, wq);
#pragma unroll
    for (int i = 0; i < 3; ++i) tmp[i] = wq[i] + qt[i];
    kc_mv(R, tmp, Rt);
#pragma unroll
    for (int i = 0; i < 3; ++i) ns[i] = P.rhoA * Rt[i] - f[i];
    T Jw[3], Jwt[3], wJw[3], psn[3];
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { Jw[i] = P.rhoJ[4 * i] * w[i]; Jwt[i] = P.rhoJ[4 * i] * wt[i]; }
    } else {
        kc_mv(P.rhoJ, w, Jw);
        kc_mv(P.rhoJ, wt, Jwt);
    }
    kc_cross(w, Jw, wJw);
#pragma unroll
    for (int i = 0; i < 3; ++i) tmp[i] = wJw[i] + Jwt[i];
    kc_mv(R, tmp, Rt);
    kc_cross(ps, n, psn);
#pragma unroll
    for (int i = 0; i < 3; ++i) ms[i] = Rt[i] - psn[i];
    T uq[3], wv[3], uw[3];
    kc_cross(u, q, uq);
    kc_cross(w, v, wv);
    kc_cross(u, w, uw);
#pragma unroll
    for (int i = 0; i < 3; ++i) { qs[i] = vt[i] - uq[i] + wv[i]; ws[i] = ut[i] - uw[i]; }
    // Eq(9) quaternion derivative (:185-189)
    hs[0] = T(0.5) * (-u[0] * h[1] - u[1] * h[2] - u[2] * h[3]);
    hs[1] = T(0.5) * (u[0] * h[0] + u[2] * h[2] - u[1] * h[3]);
    hs[2] = T(0.5) * (u[1] * h[0] - u[2] * h[1] + u[0] * h[3]);
    hs[3] = T(0.5) * (u[2] * h[0] + u[1] * h[1] - u[0] * h[2]);
#pragma unroll
    for (int i = 0; i < 3; ++i) { z[i] = v[i]; z[3 + i] = u[i]; }
}

// KNODE residual, SIMT form: o[25] = W2 ELU(W1 x + b1) + b2 with the packed weights of MlpC (uniform broadcast loads,
// 16-byte vectors).  The 28/53-term dot product of a hidden unit is split over 4 partial sums so that it is not one serial
// FMA chain, and two hidden units are in flight per iteration.
// 4 consecutive values from a 16-byte aligned address as ONE 128-bit load (fp32) / two (fp64)
#if defined(__CUDACC__)
KC_HD void kc_ld4(const float* __restrict__ p, float v[4]) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
KC_HD void kc_ld4(const double* __restrict__ p, double v[4]) {
    const double2 a = *reinterpret_cast<const double2*>(p);
    const double2 b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}
#else   // host build of the test harness (tests/emul): no CUDA vector types
template <typename T> inline void kc_ld4(const T* p, T v[4]) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3]; }
#endif
// bias + sum_k wrow[k] x[k], wrow 16-byte aligned and readable up to the next multiple of 4 (padding is zero-filled or
// multiplied by nothing: only k < IN is used); 4 partial sums break the serial FMA chain.
template <typename T, int IN>
KC_HD T mlp_unit_dot(const T* __restrict__ wrow, const T* __restrict__ x, T bias) {
    T p[4] = {bias, T(0), T(0), T(0)};
#pragma unroll
    for (int k = 0; k < IN; k += 4) {
        T w[4];
        kc_ld4(wrow + k, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (k + j < IN) p[j] += w[j] * x[k + j];
    }
    return (p[0] + p[1]) + (p[2] + p[3]);
}
// o[0..25) += wcol[0..25) * a with 128-bit loads (wcol 16-byte aligned, 28 readable)
template <typename T>
KC_HD void mlp_unit_axpy(const T* __restrict__ wcol, T a, T* __restrict__ o) {
#pragma unroll
    for (int c = 0; c < 28; c += 4) {
        T w[4];
        kc_ld4(wcol + c, w);
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c + j < 25) o[c + j] += w[j] * a;
    }
}
template <typename T, int IN>
KC_HD void mlp_eval(const MlpC<T>& M, const T* __restrict__ x, T* __restrict__ o) {
#pragma unroll
    for (int c = 0; c < 25; ++c) o[c] = M.b2[c];
    const int inP = (IN + 3) & ~3;
    int i = 0;
    for (; i + 1 < M.hidden; i += 2) {
        const T* __restrict__ w0 = M.Wp + (size_t)i * M.stride;
        const T* __restrict__ w1 = w0 + M.stride;
        const T a0 = kc_elu(mlp_unit_dot<T, IN>(w0, x, w0[inP]));
        const T a1 = kc_elu(mlp_unit_dot<T, IN>(w1, x, w1[inP]));
        mlp_unit_axpy<T>(w0 + inP + 4, a0, o);
        mlp_unit_axpy<T>(w1 + inP + 4, a1, o);
    }
    for (; i < M.hidden; ++i) {
        const T* __restrict__ w0 = M.Wp + (size_t)i * M.stride;
        const T a0 = kc_elu(mlp_unit_dot<T, IN>(w0, x, w0[inP]));
        mlp_unit_axpy<T>(w0 + inP + 4, a0, o);
    }
}

// Physics + optional MLP residual for one node.  NH = number of history rows carried per node:
// 12 -> hist = [qh,wh,vh,uh]; 25 -> hist = [yh(19); zh(6)] (needed when the MLP sees the history, IN == 53).
// The MLP sees the PRE-correction z; z is corrected after ys is formed (cosserat_ode_torch.py:192-212).
template <typename T, bool DIAG, int IN /*0 = physics only*/, int NH, typename MLP>
KC_HD void node_eval(const RodC<T>& P, const MLP& M, const T* __restrict__ y, const T* __restrict__ hist,
                     const T tf[3], T* __restrict__ ys, T* __restrict__ z) {
    const T* qh = (NH == 12) ? hist : hist + 13;
    const T* wh = qh + 3;
    const T* vh = (NH == 12) ? hist + 6 : hist + 19;
    const T* uh = vh + 3;
    rod_ode<T, DIAG>(P, y, qh, wh, vh, uh, tf, ys, z);
    if (IN > 0) {
        T x[IN > 0 ? IN : 1];
        T o[25];
        if (IN == 28) {
#pragma unroll
            for (int i = 0; i < 19; ++i) x[i] = y[i];
#pragma unroll
            for (int i = 0; i < 6; ++i) x[19 + i] = z[i];
#pragma unroll
            for (int i = 0; i < 3; ++i) x[25 + i] = tf[i];
        } else {  // 53: [y; yh; z; zh; tf] (:194-195)
#pragma unroll
            for (int i = 0; i < 19; ++i) { x[i] = y[i]; x[19 + i] = hist[i % NH]; }
#pragma unroll
            for (int i = 0; i < 6; ++i) { x[38 + i] = z[i]; x[44 + i] = hist[(19 + i) % NH]; }
#pragma unroll
            for (int i = 0; i < 3; ++i) x[50 + i] = tf[i];
        }
        mlp_eval<T, (IN > 0 ? IN : 28)>(M, x, o);
#pragma unroll
        for (int i = 0; i < 19; ++i) ys[i] += o[i];
#pragma unroll
        for (int i = 0; i < 6; ++i) z[i] += o[19 + i];
    }
}

template <typename T> KC_HD void tendon_force(const RodC<T>& P, const T ten[4], T tf[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
        tf[i] = ten[0] * P.tdirs[i] + ten[1] * P.tdirs[3 + i] + ten[2] * P.tdirs[6 + i] + ten[3] * P.tdirs[9 + i];
}

template <typename T> KC_HD void base_state(const RodC<T>& P, const T G[6], T y[19]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { y[i] = P.p0[i]; y[7 + i] = G[i]; y[10 + i] = G[3 + i]; y[13 + i] = P.q0[i]; y[16 + i] = P.w0[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) y[3 + i] = P.h0[i];
}

// Explicit-Euler shooting march (cosserat_ode.py:188-213).  Hist::load(j, hist[NH]) supplies node j's history,
// Sink::put(j, y[19]) receives the state entering node j (j = 0..N-1) and Sink::putz(j, z[6]) the z produced at node j
// (j = 0..N-2).  Returns res = [F_tip - n(L), M_tip - m(L)].
template <typename T, bool DIAG, int IN, int NH, typename Hist, typename Sink, typename MLP>
KC_HD void rod_march(const RodC<T>& P, const MLP& M, const T G[6], const T tf[3], const Hist& H, Sink& S,
                     T res[6]) {
    T y[19];
    base_state(P, G, y);
    const int N = P.N;
    // the state is handed to the sink right AFTER the Euler update that produced it: storing it at the top of the next
    // node instead makes every update wait for the store of the old value to release its register (measured: ~15 % of
    // the wide kernel's stall samples)
    S.put(0, y);
    for (int j = 0; j < N - 1; ++j) {
        T hist[NH], ys[19], z[6];
        H.load(j, hist);
        node_eval<T, DIAG, IN, NH>(P, M, y, hist, tf, ys, z);
        S.putz(j, z);
#pragma unroll
        for (int i = 0; i < 19; ++i) y[i] += P.ds * ys[i];
        S.put(j + 1, y);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { res[i] = P.Ftip[i] - y[7 + i]; res[3 + i] = P.Mtip[i] - y[10 + i]; }
}

// RK4 march (cosserat_ode.py:215-255): mid-point histories are linear interpolations (knode.py:80-81); only k1's z is kept.
// Hist must supply node N-1 as well (k4 of the last interval reads it).
template <typename T, bool DIAG, int IN, int NH, typename Hist, typename Sink, typename MLP>
KC_HD void rod_march_rk4(const RodC<T>& P, const MLP& M, const T G[6], const T tf[3], const Hist& H, Sink& S, T res[6]) {
    T y[19];
    base_state(P, G, y);
    const int N = P.N;
    T h0[NH], h1[NH], hm[NH];
    H.load(0, h0);
    for (int j = 0; j < N - 1; ++j) {
        H.load(j + 1, h1);
#pragma unroll
        for (int s = 0; s < NH; ++s) hm[s] = T(0.5) * (h0[s] + h1[s]);
        T k1[19], k2[19], k3[19], k4[19], z[6], zt[6], yt[19];
        S.put(j, y);
        node_eval<T, DIAG, IN, NH>(P, M, y, h0, tf, k1, z);
        S.putz(j, z);
#pragma unroll
        for (int i = 0; i < 19; ++i) yt[i] = y[i] + k1[i] * P.ds / T(2);
        node_eval<T, DIAG, IN, NH>(P, M, yt, hm, tf, k2, zt);
#pragma unroll
        for (int i = 0; i < 19; ++i) yt[i] = y[i] + k2[i] * P.ds / T(2);
        node_eval<T, DIAG, IN, NH>(P, M, yt, hm, tf, k3, zt);
#pragma unroll
        for (int i = 0; i < 19; ++i) yt[i] = y[i] + k3[i] * P.ds;
        node_eval<T, DIAG, IN, NH>(P, M, yt, h1, tf, k4, zt);
#pragma unroll
        for (int i = 0; i < 19; ++i) y[i] = y[i] + P.ds * (k1[i] + T(2) * (k2[i] + k3[i]) + k4[i]) / T(6);
#pragma unroll
        for (int s = 0; s < NH; ++s) h0[s] = h1[s];
    }
    S.put(N - 1, y);
#pragma unroll
    for (int i = 0; i < 3; ++i) { res[i] = P.Ftip[i] - y[7 + i]; res[3 + i] = P.Mtip[i] - y[10 + i]; }
}

// Spatial integrator chosen at compile time (KC_MARCH_EULER | KC_MARCH_RK4).
template <typename T, bool DIAG, int IN, int NH, int METHOD, typename Hist, typename Sink, typename MLP>
KC_HD void rod_march_m(const RodC<T>& P, const MLP& M, const T G[6], const T tf[3], const Hist& H, Sink& S, T res[6]) {
    if (METHOD == KC_MARCH_RK4) rod_march_rk4<T, DIAG, IN, NH>(P, M, G, tf, H, S, res);
    else rod_march<T, DIAG, IN, NH>(P, M, G, tf, H, S, res);
}

// ---- 6x6 helpers for the quasi-Newton shooting solve ------------------------------------------------------------
// In-place Gauss-Jordan inverse without pivoting: the shooting Jacobian is -[[I,0],[X,I]] + small (cond ~ 1.5,
// SURVEY §4), so pivots stay O(1).  Returns false if a pivot is tiny or not finite (caller flags the rod).
template <typename T> KC_HD bool inv6(T A[36]) {
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        const T p = A[k * 6 + k];
        if (!(kc_abs(p) > T(1e-30))) ok = false;
        const T ip = T(1) / p;
#pragma unroll
        for (int j = 0; j < 6; ++j) A[k * 6 + j] = (j == k) ? ip : A[k * 6 + j] * ip;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (i == k) continue;
            const T f = A[i * 6 + k];
#pragma unroll
            for (int j = 0; j < 6; ++j) A[i * 6 + j] = (j == k) ? -f * ip : A[i * 6 + j] - f * A[k * 6 + j];
        }
    }
    return ok;
}

template <typename T> KC_HD T norm_inf6(const T v[6]) {
    T m = kc_abs(v[0]);
#pragma unroll
    for (int i = 1; i < 6; ++i) m = kc_max(m, kc_abs(v[i]));
    return m;
}

// Per-rod solver memory carried across time steps, addressed with an element stride so that on the device it sits in
// shared memory as [slot][lane] (conflict-free) and costs no registers while a march is running:
//   slots 0..5 G (base reactions of the last solved step, warm start, knode.py:67,89), 6..11 the step before (linear
//   predictor), 12..47 inverse shooting Jacobian (row-major 6x6, Broyden-maintained across steps), 48 have_J flag.
constexpr int KC_SHOOT_SLOTS = 49;
template <typename T, int LS> struct ShootMem {
    T* p;
    KC_HD T& G(int i) const { return p[i * LS]; }
    KC_HD T& Gm1(int i) const { return p[(6 + i) * LS]; }
    KC_HD T& J(int i, int j) const { return p[(12 + i * 6 + j) * LS]; }
    KC_HD T& haveJ() const { return p[48 * LS]; }
    KC_HD void reset() const {
        for (int i = 0; i < KC_SHOOT_SLOTS; ++i) p[i * LS] = T(0);
    }
};

// "Good" Broyden update applied to the inverse (Sherman–Morrison): Jinv += (s - Jinv yv)(s^T Jinv) / (s^T Jinv yv).
template <typename T, int LS> KC_HD void broyden_update(const ShootMem<T, LS>& st, const T sv[6], const T yv[6]) {
    T Jy[6], sJ[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { Jy[i] = T(0); sJ[i] = T(0); }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const T a = st.J(i, j);
            Jy[i] += a * yv[j];
            sJ[j] += sv[i] * a;
        }
    }
    T den = T(0);
#pragma unroll
    for (int i = 0; i < 6; ++i) den += sv[i] * Jy[i];
    if (!(kc_abs(den) > T(1e-30))) return;
    const T iden = T(1) / den;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const T c = (sv[i] - Jy[i]) * iden;
#pragma unroll
        for (int j = 0; j < 6; ++j) st.J(i, j) += c * sJ[j];
    }
}

// Solve one time step for one rod: find G with residual(G) = 0 and leave the marched state in the Sink (the sink is
// written by every march; the last march is the accepted one — the reference likewise keeps "whatever the last
// residual evaluation left in y,z", knode.py:89,96).  Quasi-Newton: linear predictor from the two previous steps,
// Broyden-updated inverse Jacobian kept across steps, finite-difference (re)build of the Jacobian on the first step and
// whenever two consecutive iterations fail to halve the residual.  Written as a small state machine around ONE march
// call site so lanes in different phases (predictor / FD column / Broyden iterate) still execute the march together.
// Returns the number of marches (negative: tol not reached within max_iter, or a NaN / singular Jacobian appeared).
template <typename T, bool DIAG, int IN, int NH, int LS, int METHOD = KC_MARCH_EULER, typename Hist, typename Sink, typename MLP>
KC_HD int shoot_step(const RodC<T>& P, const MLP& M, const ShootMem<T, LS>& st, const T tf[3], const Hist& H, Sink& S,
                     T tol, int max_iter, T fd_eps) {
    enum { PRED = 0, FD = 1, BROY = 2 };
    T G[6], F[6], dG[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { G[i] = st.G(i) + (st.G(i) - st.Gm1(i)); F[i] = T(0); dG[i] = T(0); }
    int phase = PRED, k = 0, marches = 0, stalled = 0;
    T fprev = T(0), eps_k = T(0);
    bool converged = false, failed = false;
    while (true) {
        T Ge[6], Fn[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) Ge[i] = G[i];
        if (phase == FD) {
            T gk = T(0);
#pragma unroll
            for (int i = 0; i < 6; ++i) if (i == k) gk = G[i];
            eps_k = fd_eps * kc_max(T(1), kc_abs(gk));
#pragma unroll
            for (int i = 0; i < 6; ++i) if (i == k) Ge[i] += eps_k;
        }
        rod_march_m<T, DIAG, IN, NH, METHOD>(P, M, Ge, tf, H, S, Fn);
        ++marches;
        if (phase == FD) {
            const T ie = T(1) / eps_k;
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                const T d = (Fn[i] - F[i]) * ie;
#pragma unroll
                for (int c = 0; c < 6; ++c) if (c == k) st.J(i, c) = d;
            }
            if (++k < 6) continue;
            T A[36];
#pragma unroll
            for (int i = 0; i < 36; ++i) A[i] = st.J(i / 6, i % 6);
            if (!inv6(A)) { failed = true; break; }
#pragma unroll
            for (int i = 0; i < 36; ++i) st.J(i / 6, i % 6) = A[i];
            st.haveJ() = T(1);
            stalled = 0;
        } else {
            if (phase == BROY) {
                T yv[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) yv[i] = Fn[i] - F[i];
                broyden_update(st, dG, yv);
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) F[i] = Fn[i];
            const T fn = norm_inf6(F);
            if (!(fn == fn)) { failed = true; break; }
            if (phase == BROY) stalled = (fn > T(0.5) * fprev) ? stalled + 1 : 0;
            fprev = fn;
            converged = fn <= tol * kc_max(T(1), norm_inf6(G));
            if (converged) break;
            if (marches >= max_iter) break;
            if (st.haveJ() == T(0) || stalled >= 2) { phase = FD; k = 0; continue; }
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            T a = T(0);
#pragma unroll
            for (int j = 0; j < 6; ++j) a += st.J(i, j) * F[j];
            dG[i] = -a;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) G[i] += dG[i];
        phase = BROY;
    }
    if (!failed) {
#pragma unroll
        for (int i = 0; i < 6; ++i) { st.Gm1(i) = st.G(i); st.G(i) = G[i]; }
    }
    return (converged && !failed) ? marches : -marches;
}

// ---- loss helpers: Utils/transformations.py:3-31 (quaternion wxyz -> roll, pitch, yaw) and its reverse mode ---------
KC_HD float kc_sqrt(float x) { return sqrtf(x); }
KC_HD double kc_sqrt(double x) { return sqrt(x); }
KC_HD float kc_atan2(float a, float b) { return atan2f(a, b); }
KC_HD double kc_atan2(double a, double b) { return atan2(a, b); }
KC_HD float kc_asin(float a) { return asinf(a); }
KC_HD double kc_asin(double a) { return asin(a); }

template <typename T> KC_HD void quat_to_euler(const T q[4], T e[3]) {
    const T inv = T(1) / kc_sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const T w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    e[0] = kc_atan2(T(2) * (w * y + x * z), T(1) - T(2) * (y * y + z * z));
    T s = T(2) * (w * z - x * y);
    s = s < T(-1) ? T(-1) : (s > T(1) ? T(1) : s);
    e[1] = kc_asin(s);
    e[2] = kc_atan2(T(2) * (w * x + y * z), T(1) - T(2) * (x * x + z * z));
}

// cotangent g[3] of the Euler angles -> cotangent gq[4] of the (unnormalised) quaternion; the clamp passes gradient only
// strictly inside (-1, 1), as torch.clamp does.
template <typename T> KC_HD void quat_to_euler_vjp(const T q[4], const T g[3], T gq[4]) {
    const T n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    const T inv = T(1) / kc_sqrt(n2);
    const T w = q[0] * inv, x = q[1] * inv, y = q[2] * inv, z = q[3] * inv;
    T gw, gx, gy, gz;
    {
        const T a = T(2) * (w * y + x * z), b = T(1) - T(2) * (y * y + z * z);
        const T r = T(1) / (a * a + b * b);
        const T da = g[0] * b * r, db = -g[0] * a * r;
        gw = da * T(2) * y; gx = da * T(2) * z;
        gy = da * T(2) * w - db * T(4) * y; gz = da * T(2) * x - db * T(4) * z;
    }
    {
        const T s = T(2) * (w * z - x * y);
        if (s > T(-1) && s < T(1)) {
            const T d = g[1] / kc_sqrt(T(1) - s * s);
            gw += d * T(2) * z; gz += d * T(2) * w; gx -= d * T(2) * y; gy -= d * T(2) * x;
        }
    }
    {
        const T c = T(2) * (w * x + y * z), d = T(1) - T(2) * (x * x + z * z);
        const T r = T(1) / (c * c + d * d);
        const T dc = g[2] * d * r, dd = -g[2] * c * r;
        gw += dc * T(2) * x; gx += dc * T(2) * w - dd * T(4) * x;
        gy += dc * T(2) * z; gz += dc * T(2) * y - dd * T(4) * z;
    }
    const T dot = gw * w + gx * x + gy * y + gz * z;
    gq[0] = (gw - w * dot) * inv; gq[1] = (gx - x * dot) * inv;
    gq[2] = (gy - y * dot) * inv; gq[3] = (gz - z * dot) * inv;
}
