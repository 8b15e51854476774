// kc_adjoint.cuh — hand-written reverse mode (vector-Jacobian product) of the physics node ODE `rod_ode` (kc_rod.cuh).
// This is what torch.autograd computes through CosseratRodTorch.ODE_parallel (cosserat_ode_torch.py:217-306) with
// respect to its tensor inputs; the forward intermediates are recomputed here (cheaper than storing them).
#pragma once
#include "kc_rod.cuh"

// c = a x b  =>  ga += b x gc,  gb += gc x a
template <typename T> KC_HD void cross_vjp(const T a[3], const T b[3], const T gc[3], T ga[3], T gb[3]) {
    ga[0] += b[1] * gc[2] - b[2] * gc[1];
    ga[1] += b[2] * gc[0] - b[0] * gc[2];
    ga[2] += b[0] * gc[1] - b[1] * gc[0];
    gb[0] += gc[1] * a[2] - gc[2] * a[1];
    gb[1] += gc[2] * a[0] - gc[0] * a[2];
    gb[2] += gc[0] * a[1] - gc[1] * a[0];
}
// r = R x  =>  gR += gr (x) x,  gx += R^T gr
template <typename T> KC_HD void mv_vjp(const T R[9], const T x[3], const T gr[3], T gR[9], T gx[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { gR[i * 3 + j] += gr[i] * x[j]; gx[j] += R[i * 3 + j] * gr[i]; }
    }
}
// r = R^T x  =>  gR[i][j] += x[i] gr[j],  gx += R gr
template <typename T> KC_HD void mtv_vjp(const T R[9], const T x[3], const T gr[3], T gR[9], T gx[3]) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { gR[i * 3 + j] += x[i] * gr[j]; gx[i] += R[i * 3 + j] * gr[j]; }
    }
}

// Inputs as rod_ode; cotangents g_ys[19] (of ys) and g_z[6] (of the PRE-correction z = [v;u]).
// Outputs (overwritten): gy[19] (gy[0:3] = 0: p is not read), gqh[3], gwh[3], gvh[3], guh[3], gtf[3].
template <typename T, bool DIAG>
KC_HD void rod_ode_vjp(const RodC<T>& P, const T* __restrict__ y, const T qh[3], const T wh[3], const T vh[3],
                       const T uh[3], const T tf[3], const T* __restrict__ g_ys, const T* __restrict__ g_z,
                       T* __restrict__ gy, T gqh[3], T gwh[3], T gvh[3], T guh[3], T gtf[3]) {
    const T* h = y + 3; const T* n = y + 7; const T* m = y + 10; const T* q = y + 13; const T* w = y + 16;
    // ---- forward recompute ----
    T R[9];
    kc_quat_R(h, R);
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
        T b1[3], b2[3], a1[3], a2[3];
        kc_mv(P.Bse, vh, b1);
        kc_mv(P.Bbt, uh, b2);
#pragma unroll
        for (int i = 0; i < 3; ++i) { a1[i] = t1[i] + P.KseVstar[i] - b1[i]; a2[i] = t2[i] - b2[i]; }
        kc_mv(P.KseInv, a1, v);
        kc_mv(P.KbtInv, a2, u);
    }
    T qt[3], wt[3], dr[3], ps[3], wq[3], e1[3], Jw[3], Jwt[3], wJw[3], e2[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { qt[i] = P.c0 * q[i] + qh[i]; wt[i] = P.c0 * w[i] + wh[i]; dr[i] = P.C[i] * q[i] * kc_abs(q[i]); }
    kc_mv(R, v, ps);
    kc_cross(w, q, wq);
#pragma unroll
    for (int i = 0; i < 3; ++i) e1[i] = wq[i] + qt[i];
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { Jw[i] = P.rhoJ[4 * i] * w[i]; Jwt[i] = P.rhoJ[4 * i] * wt[i]; }
    } else {
        kc_mv(P.rhoJ, w, Jw);
        kc_mv(P.rhoJ, wt, Jwt);
    }
    kc_cross(w, Jw, wJw);
#pragma unroll
    for (int i = 0; i < 3; ++i) e2[i] = wJw[i] + Jwt[i];
    // ---- reverse ----
    const T* gps = g_ys; const T* ghs = g_ys + 3; const T* gns = g_ys + 7; const T* gms = g_ys + 10;
    const T* gqs = g_ys + 13; const T* gws = g_ys + 16;
    T gR[9], gh[4], gn[3], gm[3], gq[3], gw[3], gv[3], gu[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) gR[i] = T(0);
#pragma unroll
    for (int i = 0; i < 3; ++i) { gn[i] = gm[i] = gq[i] = gw[i] = T(0); gv[i] = g_z[i]; gu[i] = g_z[3 + i]; gtf[i] = T(0); }
    // hs = 0.5 Omega(u) h
    gu[0] += T(0.5) * (-ghs[0] * h[1] + ghs[1] * h[0] + ghs[2] * h[3] - ghs[3] * h[2]);
    gu[1] += T(0.5) * (-ghs[0] * h[2] - ghs[1] * h[3] + ghs[2] * h[0] + ghs[3] * h[1]);
    gu[2] += T(0.5) * (-ghs[0] * h[3] + ghs[1] * h[2] - ghs[2] * h[1] + ghs[3] * h[0]);
    gh[0] = T(0.5) * (ghs[1] * u[0] + ghs[2] * u[1] + ghs[3] * u[2]);
    gh[1] = T(0.5) * (-ghs[0] * u[0] - ghs[2] * u[2] + ghs[3] * u[1]);
    gh[2] = T(0.5) * (-ghs[0] * u[1] + ghs[1] * u[2] - ghs[3] * u[0]);
    gh[3] = T(0.5) * (-ghs[0] * u[2] - ghs[1] * u[1] + ghs[2] * u[0]);
    // ws = ut - u x w ; qs = vt - u x q + w x v
    T gut[3], gvt[3], neg[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { gut[i] = gws[i]; gvt[i] = gqs[i]; neg[i] = -gws[i]; }
    cross_vjp(u, w, neg, gu, gw);
#pragma unroll
    for (int i = 0; i < 3; ++i) neg[i] = -gqs[i];
    cross_vjp(u, q, neg, gu, gq);
    cross_vjp(w, v, gqs, gw, gv);
    // ms = R e2 - ps x n
    T gps_t[3], ge[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { gps_t[i] = gps[i]; neg[i] = -gms[i]; ge[i] = T(0); }
    cross_vjp(ps, n, neg, gps_t, gn);
    mv_vjp(R, e2, gms, gR, ge);               // ge = cotangent of e2 = of wJw and of Jwt
    T gJw[3] = {T(0), T(0), T(0)}, gwt[3];
    cross_vjp(w, Jw, ge, gw, gJw);
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < 3; ++i) { gw[i] += P.rhoJ[4 * i] * gJw[i]; gwt[i] = P.rhoJ[4 * i] * ge[i]; }
    } else {
        T tmp[3];
        kc_mtv(P.rhoJ, gJw, tmp);
        kc_mtv(P.rhoJ, ge, gwt);
#pragma unroll
        for (int i = 0; i < 3; ++i) gw[i] += tmp[i];
    }
    // ns = rhoA R e1 - f
    T gRe1[3], gqt[3] = {T(0), T(0), T(0)};
#pragma unroll
    for (int i = 0; i < 3; ++i) gRe1[i] = P.rhoA * gns[i];
    mv_vjp(R, e1, gRe1, gR, gqt);             // gqt = cotangent of e1 = of wq and of qt
    cross_vjp(w, q, gqt, gw, gq);
    // ps = R v
    mv_vjp(R, v, gps_t, gR, gv);
    // f = rhoAg - R dr + tf, gf = -gns
    T gdr[3] = {T(0), T(0), T(0)};
    mv_vjp(R, dr, gns, gR, gdr);              // gRd = -gf = gns
#pragma unroll
    for (int i = 0; i < 3; ++i) { gtf[i] = -gns[i]; gq[i] += T(2) * P.C[i] * kc_abs(q[i]) * gdr[i]; }
    // BDF2 time derivatives
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        gv[i] += P.c0 * gvt[i]; gvh[i] = gvt[i];
        gu[i] += P.c0 * gut[i]; guh[i] = gut[i];
        gq[i] += P.c0 * gqt[i]; gqh[i] = gqt[i];
        gw[i] += P.c0 * gwt[i]; gwh[i] = gwt[i];
    }
    // constitutive law
    T gt1[3], gt2[3];
    if (DIAG) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            gt1[i] = P.KseInv[4 * i] * gv[i];
            gt2[i] = P.KbtInv[4 * i] * gu[i];
            guh[i] -= P.Bbt[4 * i] * gt2[i];
        }
    } else {
        T tmp[3];
        kc_mtv(P.KseInv, gv, gt1);
        kc_mtv(P.KbtInv, gu, gt2);
        kc_mtv(P.Bse, gt1, tmp);
#pragma unroll
        for (int i = 0; i < 3; ++i) gvh[i] -= tmp[i];
        kc_mtv(P.Bbt, gt2, tmp);
#pragma unroll
        for (int i = 0; i < 3; ++i) guh[i] -= tmp[i];
    }
    mtv_vjp(R, n, gt1, gR, gn);
    mtv_vjp(R, m, gt2, gR, gm);
    // R = I + s M(h), s = 2/(h.h)
    {
        const T a = h[0], b = h[1], c = h[2], d = h[3];
        const T s = T(2) / (a * a + b * b + c * c + d * d);
        const T is = T(1) / s;
        T dot = T(0);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
#pragma unroll
            for (int j = 0; j < 3; ++j) dot += gR[i * 3 + j] * ((R[i * 3 + j] - (i == j ? T(1) : T(0))) * is);
        }
        const T ga_ = -d * gR[1] + c * gR[2] + d * gR[3] - b * gR[5] - c * gR[6] + b * gR[7];
        const T gb_ = c * gR[1] + d * gR[2] + c * gR[3] - T(2) * b * gR[4] - a * gR[5] + d * gR[6] + a * gR[7] - T(2) * b * gR[8];
        const T gc_ = -T(2) * c * gR[0] + b * gR[1] + a * gR[2] + b * gR[3] + d * gR[5] - a * gR[6] + d * gR[7] - T(2) * c * gR[8];
        const T gd_ = -T(2) * d * gR[0] - a * gR[1] + b * gR[2] + a * gR[3] - T(2) * d * gR[4] + c * gR[5] + b * gR[6] + c * gR[7];
        const T k = s * s * dot;
        gh[0] += s * ga_ - k * a;
        gh[1] += s * gb_ - k * b;
        gh[2] += s * gc_ - k * c;
        gh[3] += s * gd_ - k * d;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) { gy[i] = T(0); gy[7 + i] = gn[i]; gy[10 + i] = gm[i]; gy[13 + i] = gq[i]; gy[16 + i] = gw[i]; }
#pragma unroll
    for (int i = 0; i < 4; ++i) gy[3 + i] = gh[i];
}

// Reverse mode of the MLP with respect to its INPUT: gx = W1^T ((W2^T go) * ELU'(W1 x + b1)), packed weights (MlpC).
template <typename T, int IN>
KC_HD void mlp_input_vjp(const MlpC<T>& M, const T* __restrict__ x, const T* __restrict__ go, T* __restrict__ gx) {
    const int inP = (IN + 3) & ~3;
#pragma unroll
    for (int k = 0; k < IN; ++k) gx[k] = T(0);
    for (int i = 0; i < M.hidden; ++i) {
        const T* __restrict__ wrow = M.Wp + (size_t)i * M.stride;
        const T z1 = mlp_unit_dot<T, IN>(wrow, x, wrow[inP]);
        const T da = mlp_unit_dot<T, 25>(wrow + inP + 4, go, T(0));
        const T dz = da * kc_elu_grad(z1);
#pragma unroll
        for (int k = 0; k < IN; k += 4) {
            T w[4];
            kc_ld4(wrow + k, w);
#pragma unroll
            for (int j = 0; j < 4; ++j) if (k + j < IN) gx[k + j] += dz * w[j];
        }
    }
}

