// kc_emul.cpp — TEST-ONLY host harness: compiles the per-rod solver headers of knode-cosserat_b200/csrc with g++ and
// runs them one rod at a time on the CPU, so the quasi-Newton / history / layout logic of the rollout kernel can be
// checked against the oracle in the GPU-less build container.  Not linked into the product library, not reachable
// from the Python drop-ins.
#include <vector>
#include <cstring>
#include "../../knode-cosserat_b200/csrc/kc_rollout_wide.cuh"
#include "../../knode-cosserat_b200/csrc/kc_bptt_core.cuh"

void kc_set_error(const char*, ...) {}

template <typename T>
static void pack_mlp(const T* W1, const T* b1, const T* W2, int in_dim, int hidden, std::vector<T>& Wp, int& inP, int& stride) {
    inP = (in_dim + 3) & ~3;
    stride = inP + 32;
    Wp.assign((size_t)hidden * stride, T(0));
    for (int i = 0; i < hidden; ++i) {
        for (int k = 0; k < in_dim; ++k) Wp[(size_t)i * stride + k] = W1[(size_t)i * in_dim + k];
        Wp[(size_t)i * stride + inP] = b1[i];
        for (int c = 0; c < 25; ++c) Wp[(size_t)i * stride + inP + 4 + c] = W2[(size_t)c * hidden + i];
    }
}

template <typename T, bool DIAG, int IN, int NH, int METHOD = KC_MARCH_EULER>
static void run(const RodC<T>& P, const MlpC<T>& M, int64_t B, int64_t T_, const T* ten, T* traj, int32_t* iters,
                T* Gout, T tol, int max_iter, T fd_eps) {
    const int N = P.N;
    std::vector<T> trajD((size_t)T_ * 25 * N), Hs((size_t)NH * N);   // N history nodes cover the RK4 march too
    T* out = traj;
    for (int64_t b = 0; b < B; ++b) {
        T stmem[KC_SHOOT_SLOTS];
        ShootMem<T, 1> st{stmem};
        st.reset();
        rollout_init<T, 1>(P, nullptr, nullptr, trajD.data());
        if (iters) iters[b * T_] = 0;
        rollout_rod<T, DIAG, IN, NH, 1, 1, METHOD>(P, M, st, ten + b * T_ * 4, trajD.data(), Hs.data(), 0, (int)T_ - 1, tol,
                                        max_iter, fd_eps, Gout ? Gout + b * T_ * 6 : nullptr, iters ? iters + b * T_ : nullptr);
        // device layout per rod is [T][N][25] -> reference [T][25][N]
        for (int64_t t = 0; t < T_; ++t)
            for (int j = 0; j < N; ++j)
                for (int r = 0; r < 25; ++r)
                    out[(((size_t)b * T_ + t) * 25 + r) * N + j] = trajD[((size_t)t * N + j) * 25 + r];
    }
}

// wide mode: the 7 lanes of a rod are emulated one after the other; the decision logic (wide_decide) is the kernel's
template <typename T, bool DIAG, int IN, int NH>
static void run_wide(const RodC<T>& P, const MlpC<T>& M, int64_t B, int64_t T_, const T* ten, T* traj, int32_t* iters,
                     T* Gout, T tol, int max_iter, T fd_eps) {
    const int N = P.N;
    std::vector<T> trajD((size_t)T_ * 25 * N), Hs((size_t)NH * (N - 1)), scratch((size_t)25 * N);
    for (int64_t b = 0; b < B; ++b) {
        rollout_init<T, 1>(P, nullptr, nullptr, trajD.data());
        if (iters) iters[b * T_] = 0;
        build_history<T, NH, 1>(P, trajD.data(), trajD.data(), Hs.data());
        T G[6] = {0, 0, 0, 0, 0, 0}, Gm1[6] = {0, 0, 0, 0, 0, 0};
        const size_t ts = (size_t)25 * N;
        for (int t = 0; t < T_ - 1; ++t) {
            T tn[4], tf[3];
            for (int i = 0; i < 4; ++i) tn[i] = ten[(b * T_ + t) * 4 + i];
            tendon_force(P, tn, tf);
            T* cur = trajD.data() + (size_t)t * ts;
            T* nxt = cur + ts;
            T Gp[6];
            for (int i = 0; i < 6; ++i) { Gp[i] = G[i]; G[i] = G[i] + (G[i] - Gm1[i]); }
            int marches = 0, status = 0;
            HistView<T, NH, 1> H{Hs.data()};
            while (true) {
                T eps[6], Fall[7][6];
                wide_eps(G, fd_eps, eps);
                for (int k = 0; k < 7; ++k) {
                    T Ge[6];
                    for (int i = 0; i < 6; ++i) Ge[i] = G[i] + ((k == i + 1) ? eps[i] : T(0));
                    TrajSinkPred<T, 1, NH, 1> S{nxt, nullptr, k == 0, N - 1};
                    rod_march<T, DIAG, IN, NH>(P, M, Ge, tf, H, S, Fall[k]);
                }
                ++marches;
                const int r = wide_decide(Fall, G, eps, tol);
                if (r != 0) { status = r; break; }
                if (marches >= max_iter) { status = -1; break; }
            }
            for (int i = 0; i < 6; ++i) Gm1[i] = Gp[i];
            for (int c = 0; c < 6; ++c) nxt[((size_t)(N - 1) * 25 + 19 + c)] = cur[((size_t)(N - 1) * 25 + 19 + c)];
            if (Gout) for (int i = 0; i < 6; ++i) Gout[(b * T_ + t + 1) * 6 + i] = G[i];
            if (iters) iters[b * T_ + t + 1] = status > 0 ? marches : -marches;
            build_history<T, NH, 1>(P, nxt, cur, Hs.data());
        }
        for (int64_t t = 0; t < T_; ++t)
            for (int j = 0; j < N; ++j)
                for (int r = 0; r < 25; ++r)
                    traj[(((size_t)b * T_ + t) * 25 + r) * N + j] = trajD[((size_t)t * N + j) * 25 + r];
    }
}

// wide mode with the linearised final correction (wide_decide_lin), emulated lane by lane
template <typename T, bool DIAG, int IN, int NH>
static void run_wide_lin(const RodC<T>& P, const MlpC<T>& M, int64_t B, int64_t T_, const T* ten, T* traj, int32_t* iters,
                         T* Gout, T tol, int max_iter, T fd_eps) {
    const int N = P.N, NV = 25 * N;
    std::vector<T> A(NV), Hs((size_t)NH * (N - 1)), S((size_t)7 * NV), out_t(NV);
    for (int64_t b = 0; b < B; ++b) {
        rollout_init<T, 1>(P, nullptr, nullptr, A.data());
        auto emit = [&](int64_t t, const std::vector<T>& st) {
            for (int j = 0; j < N; ++j)
                for (int r = 0; r < 25; ++r) traj[(((size_t)b * T_ + t) * 25 + r) * N + j] = st[(size_t)j * 25 + r];
        };
        emit(0, A);
        if (iters) iters[b * T_] = 0;
        build_history<T, NH, 1>(P, A.data(), A.data(), Hs.data());
        T zlast[6];
        for (int c = 0; c < 6; ++c) zlast[c] = A[(size_t)(N - 1) * 25 + 19 + c];
        T G[6] = {0, 0, 0, 0, 0, 0}, Gm1[6] = {0, 0, 0, 0, 0, 0}, Cest = 0;
        T D[6] = {0, 0, 0, 0, 0, 0}, Dm1[6] = {0, 0, 0, 0, 0, 0};
        for (int t = 0; t < T_ - 1; ++t) {
            T tn[4], tf[3], tfn[3];
            for (int i = 0; i < 4; ++i) tn[i] = ten[(b * T_ + t) * 4 + i];
            tendon_force(P, tn, tf);
            for (int i = 0; i < 4; ++i) tn[i] = ten[(b * T_ + t + 1) * 4 + i];
            tendon_force(P, tn, tfn);
            T Gp[6], w[6] = {0, 0, 0, 0, 0, 0};
            for (int i = 0; i < 6; ++i) {
                Gp[i] = G[i];
                G[i] = G[i] + ((G[i] - Gm1[i]) + (D[i] - Dm1[i]));
                Dm1[i] = D[i];
            }
            int marches = 0, status = 0;
            T sprev = 0;
            Cest = T(0);
            HistView<T, NH, 1> H{Hs.data()};
            while (true) {
                T eps[6], Fall[7][6];
                wide_eps(G, fd_eps, eps);
                for (int k = 0; k < 7; ++k) {
                    T Ge[6];
                    for (int i = 0; i < 6; ++i) Ge[i] = G[i] + ((k == i + 1) ? eps[i] : T(0));
                    SmemStateSinkE<T, 1, 0> Sk{S.data() + (size_t)k * NV, N};   // output order e = r*N + j
                    rod_march<T, DIAG, IN, NH>(P, M, Ge, tf, H, Sk, Fall[k]);
                }
                ++marches;
                int r;
                if (marches == 1) {   // lane 7 of the first joint march: base point under the next step's tendon load
                    T Fnext[6];
                    NullSink S0;
                    rod_march<T, DIAG, IN, NH>(P, M, G, tfn, H, S0, Fnext);
                    r = wide_decide_lin(Fall, G, eps, tol, Cest, sprev, w, Fnext, D);
                } else {
                    r = wide_decide_lin(Fall, G, eps, tol, Cest, sprev, w);
                }
                if (r != 0) { status = r; break; }
                if (marches >= max_iter) { status = -1; break; }
            }
            for (int i = 0; i < 6; ++i) Gm1[i] = Gp[i];
            for (int e = 0; e < NV; ++e) {
                T v = S[e];
                if (status == 2) { const T v0 = v; for (int c = 0; c < 6; ++c) v += (S[(size_t)(c + 1) * NV + e] - v0) * w[c]; }
                const int r = e / N, j = e - r * N;
                if (j == N - 1 && r >= 19) v = zlast[r - 19];
                out_t[(size_t)j * 25 + r] = v;
            }
            build_history<T, NH, 1>(P, out_t.data(), A.data(), Hs.data());
            A = out_t;
            emit(t + 1, A);
            if (Gout) for (int i = 0; i < 6; ++i) Gout[(b * T_ + t + 1) * 6 + i] = G[i];
            if (iters) iters[b * T_ + t + 1] = status > 0 ? marches : -marches;
        }
    }
}

template <typename T>
static int emul(const kc_rod_params* p, int in_dim, int hidden, const void* W1, const void* b1, const void* W2,
                const void* b2, int64_t B, int64_t T_, const void* ten, void* traj, int32_t* iters, void* Gout,
                double tol, int max_iter, int wide) {
    RodC<T> P = make_rodc<T>(*p);
    MlpC<T> M{};
    std::vector<T> Wp;
    if (in_dim) {
        int inP, stride;
        pack_mlp<T>((const T*)W1, (const T*)b1, (const T*)W2, in_dim, hidden, Wp, inP, stride);
        M.Wp = Wp.data(); M.b2 = (const T*)b2; M.in_dim = in_dim; M.inP = inP; M.hidden = hidden; M.stride = stride;
    }
    const T fd_eps = sizeof(T) == 4 ? T(1e-2) : T(1e-6);
    const T tl = tol > 0 ? T(tol) : (sizeof(T) == 4 ? T(2e-6) : T(1e-11));
#define GO(D, I, H)                                                                                             \
    do {                                                                                                        \
        if (wide == 3) run<T, D, I, H, KC_MARCH_RK4>(P, M, B, T_, (const T*)ten, (T*)traj, iters, (T*)Gout, tl, max_iter, fd_eps); \
        else if (wide == 2) run_wide_lin<T, D, I, H>(P, M, B, T_, (const T*)ten, (T*)traj, iters, (T*)Gout, tl, max_iter, fd_eps); \
        else if (wide) run_wide<T, D, I, H>(P, M, B, T_, (const T*)ten, (T*)traj, iters, (T*)Gout, tl, max_iter, fd_eps); \
        else run<T, D, I, H>(P, M, B, T_, (const T*)ten, (T*)traj, iters, (T*)Gout, tl, max_iter, fd_eps);      \
    } while (0)
    if (P.diag) { if (in_dim == 0) GO(true, 0, 12); else if (in_dim == 28) GO(true, 28, 12); else GO(true, 53, 25); }
    else        { if (in_dim == 0) GO(false, 0, 12); else if (in_dim == 28) GO(false, 28, 12); else GO(false, 53, 25); }
    return 0;
}

extern "C" int kc_emul_rollout(int dtype, const kc_rod_params* p, int in_dim, int hidden, const void* W1, const void* b1,
                               const void* W2, const void* b2, int64_t B, int64_t T_, const void* ten, void* traj,
                               int32_t* iters, void* Gout, double tol, int max_iter, int wide) {
    if (dtype == KC_F32) return emul<float>(p, in_dim, hidden, W1, b1, W2, b2, B, T_, ten, traj, iters, Gout, tol, max_iter, wide);
    return emul<double>(p, in_dim, hidden, W1, b1, W2, b2, B, T_, ten, traj, iters, Gout, tol, max_iter, wide);
}

// ---- BPTT (reverse mode through the rollout), one rod after the other ----------------------------------------------
template <typename T, bool DIAG, int IN, int NH>
static void run_bptt(const RodC<T>& P, const MlpC<T>& M, int64_t B, int64_t T_, const T* ten, const T* traj, const T* gtraj,
                     T* gten, T* xs, T* gos, T fd_eps) {
    const int N = P.N;
    std::vector<T> Hs((size_t)4 * NH * (N - 1));
    const size_t per_rod = (size_t)(T_ - 1) * (N - 1) * 2;
    for (int64_t b = 0; b < B; ++b)
        bptt_rod<T, DIAG, IN, NH, 1>(P, M, traj + (size_t)b * T_ * 25 * N, gtraj + (size_t)b * T_ * 25 * N, ten + b * T_ * 4,
                                     gten ? gten + b * T_ * 4 : nullptr, (int)T_, Hs.data(),
                                     IN > 0 ? xs + b * per_rod * (IN > 0 ? IN : 1) : nullptr, IN > 0 ? gos + b * per_rod * 25 : nullptr, fd_eps);
}

template <typename T>
static int emul_bptt(const kc_rod_params* p, int in_dim, int hidden, const void* W1, const void* b1, const void* W2,
                     const void* b2, int64_t B, int64_t T_, const void* ten, const void* traj, const void* gtraj, void* gten,
                     void* xs, void* gos) {
    RodC<T> P = make_rodc<T>(*p);
    MlpC<T> M{};
    std::vector<T> Wp;
    if (in_dim) {
        int inP, stride;
        pack_mlp<T>((const T*)W1, (const T*)b1, (const T*)W2, in_dim, hidden, Wp, inP, stride);
        M.Wp = Wp.data(); M.b2 = (const T*)b2; M.in_dim = in_dim; M.inP = inP; M.hidden = hidden; M.stride = stride;
    }
    const T fd_eps = sizeof(T) == 4 ? T(1e-2) : T(1e-6);
#define GOB(D, I, H) run_bptt<T, D, I, H>(P, M, B, T_, (const T*)ten, (const T*)traj, (const T*)gtraj, (T*)gten, (T*)xs, (T*)gos, fd_eps)
    if (P.diag) { if (in_dim == 0) GOB(true, 0, 12); else if (in_dim == 28) GOB(true, 28, 12); else GOB(true, 53, 25); }
    else        { if (in_dim == 0) GOB(false, 0, 12); else if (in_dim == 28) GOB(false, 28, 12); else GOB(false, 53, 25); }
    return 0;
}

extern "C" int kc_emul_bptt(int dtype, const kc_rod_params* p, int in_dim, int hidden, const void* W1, const void* b1,
                            const void* W2, const void* b2, int64_t B, int64_t T_, const void* ten, const void* traj,
                            const void* gtraj, void* gten, void* xs, void* gos) {
    if (dtype == KC_F32) return emul_bptt<float>(p, in_dim, hidden, W1, b1, W2, b2, B, T_, ten, traj, gtraj, gten, xs, gos);
    return emul_bptt<double>(p, in_dim, hidden, W1, b1, W2, b2, B, T_, ten, traj, gtraj, gten, xs, gos);
}
