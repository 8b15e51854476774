// kc_knode_tc.cu — the KNODE residual MLP INSIDE the shooting march on the Blackwell tensor cores (tcgen05 + TMEM):
// north-star subsystem 2 for the rollout (knode.simulate with CosseratRod.get_nn_output in the ODE, cosserat_ode.py:169-184,
// cosserat_ode_torch.py:192-212) and for back-propagation through it (subsystem 3).  fp32 model, 28 inputs, hidden <= 512.
//
// One CTA = 128 "rows" = the 128 lanes of TMEM.  A row is one marched point of one rod: 16 rods x 8 points per CTA
// (forward: base point + 6 finite-difference points of the Newton shooting solve, as kc_rollout_wide_kernel; backward:
// the loss-seeded adjoint march + 6 unit-seeded adjoint marches, see below).  All 128 rows step through the nodes of the
// march together, so every node evaluation is ONE [128 x 32] x [32 x H] x [H x 32] MLP for the whole CTA:
//
//   warps 0-3  (128 threads) physics of their row on the FP32 pipes, then each writes its 28 inputs (+1 for the bias) as
//              a bf16 hi/lo row of the X tile in shared memory and arrives on barX;
//   warp 8     one elected thread issues the MMAs: Z_c = X W1_c^T (kind::f16, 3-pass bf16 hi/lo split = fp32-grade
//              products: hi*hi + lo*hi + hi*lo) into a ring of TMEM buffers, commits barZ[buf];
//   warps 0-7  epilogue: tcgen05.ld Z_c -> ELU -> bf16 hi/lo -> tcgen05.st back INTO THE SAME TMEM COLUMNS (the activation
//              tile never touches shared memory), arrive on barA[buf];
//   warp 8     O += A_c W2_c^T with the A operand read from TMEM (".ts" MMA), B = the resident W2 image; then the next
//              Z chunk into the freed buffer; after the last chunk commits barO;
//   warps 0-3  tcgen05.ld their row of O (25 outputs) and continue the physics (ys += o[0:19], z += o[19:25], Euler).
//
// Both weight images (bf16 hi|lo, K-major, no swizzle: the layouts pinned by kc_umma_selftest) stay resident in shared
// memory for the whole kernel: 128 KB forward, 192 KB backward (W1, W2^T, W1^T).
//
// Backward (kc_rollout_bwd for this shape): per time step ONE joint adjoint march.  Row 0 of a rod carries the cotangent of
// the loss, rows 1..6 carry unit cotangents on the tip's (n, m): their result at the base is the EXACT transposed shooting
// Jacobian (no finite differences), the implicit-function multiplier mu solves a 6x6 system per rod, and because the
// adjoint march is linear in its seed the implicit part is the mu-weighted sum of rows 1..6 — combined with warp shuffles
// afterwards.  The MLP input-VJP per node is three GEMMs: Z = X W1^T (recompute), dA = GO W2, gX = (dA * ELU'(Z)) W1, the
// last one again with its A operand in TMEM.  Weight gradients: one (x, dL/do) sample per node goes to the tensor-core
// sample reduction of the training step (kc_mlp_bwd).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdlib>
#include "kc_rollout_wide.cuh"
#include "kc_adjoint.cuh"
#include "kc_umma.cuh"

// Development aid (make EXTRA=-DKC_TC_TRACE): clock64 time stamps of the three roles of CTA 0 for a few node evaluations,
// read back by tools/trace_knode_tc.py.  Not compiled into the product build.
#ifdef KC_TC_TRACE
__device__ long long* g_ktc_trace = nullptr;   // [3 roles][8 evaluations][32 events]
extern "C" int kc_knode_tc_set_trace(long long* p) { return (int)cudaMemcpyToSymbol(g_ktc_trace, &p, sizeof(p)); }
#define KTC_TRACE(on, role, ev, id)                                                                                    \
    do {                                                                                                               \
        if ((on) && g_ktc_trace && (ev) >= 100 && (ev) < 108) g_ktc_trace[((role) * 8 + ((ev) - 100)) * 32 + (id)] = clock64(); \
    } while (0)
#else
#define KTC_TRACE(on, role, ev, id) do {} while (0)
#endif

namespace ktc {
constexpr int THREADS = 288;          // warps 0-3 physics + epilogue, 4-7 epilogue, 8 MMA issue
constexpr int RPC = 16;               // rods per CTA (8 rows each)
constexpr int NBUF = 3;               // TMEM ring: 3 x 128 columns + 32 columns of output
constexpr int COL_O = 384;
constexpr int IMG = 32768;            // one bf16 image (hi or lo) of a [512 x 32] / [32 x 512] weight matrix
// forward shared memory
constexpr int F_W1 = 0, F_W2 = 2 * IMG, F_X = 4 * IMG, F_MISC = F_X + 16384, F_HIST = F_MISC + 256;
// backward shared memory
constexpr int B_W1 = 0, B_W2T = 2 * IMG, B_W1T = 4 * IMG, B_X = 6 * IMG, B_GO = B_X + 16384, B_MISC = B_GO + 16384,
              B_BYTES = B_MISC + 256;
constexpr int SV = 37;                // scratch values per row and node in the backward: 12 history + 25 output cotangents

struct Bars {
    uint64_t X, O, Z[NBUF], A[NBUF];
    uint32_t tmem_slot;
    uint32_t exit_flag;
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// two values -> one packed pair of bf16 "hi" (even element in the low half) and one packed pair of bf16 "lo" = x - hi.
// hi is the fp32 word rounded to its upper half with integer arithmetic (nothing on the XU pipe, cf. kc_train_tc.cu).
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const uint32_t a = (__float_as_uint(x0) + 0x8000u) & 0xffff0000u;
    const uint32_t b = (__float_as_uint(x1) + 0x8000u) & 0xffff0000u;
    hi = __byte_perm(a, b, 0x7632);
    lo = pack_bf16x2(x0 - __uint_as_float(a), x1 - __uint_as_float(b));
}
// The split as the epilogues use it (measured pipe rates, tools/micro/pipe_bench.cu: FMA 128, ALU 64, XU 16 lanes/clk/SM;
// F2FP / PRMT / LOP3 / SHF / FSETP are ALU-pipe instructions): hi = one packed round-to-nearest conversion for the pair,
// unpacked again with a shift / a mask, lo = x - hi exactly, stored truncated to bf16 (|lo| <= 2^-9 |x|, so the truncation
// is 2^-17 relative and its sign is random).  3 instructions per value.
__device__ __forceinline__ void split_pair_fast(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = __byte_perm(__float_as_uint(x0 - h0), __float_as_uint(x1 - h1), 0x7632);
}
// Packed FP32 (FMUL2 / FADD2 / FFMA2: two values per issue slot).  The epilogues are ISSUE bound (ncu: the two warps of
// a scheduler alternate, "selected" + "not selected" = 60 % of their samples), so halving the issue slots of the plain
// arithmetic is what shortens them.
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ float ex2f(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// (x0, x1) -> packed bf16 hi pair, packed bf16 lo pair (as split_pair_fast), 5 issue slots per pair
__device__ __forceinline__ void split_pair2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    float l0, l1;
    upk2(fma2(pk2(h0, h1), pk2(-1.f, -1.f), pk2(x0, x1)), l0, l1);      // x - hi, exact
    lo = __byte_perm(__float_as_uint(l0), __float_as_uint(l1), 0x7632);
}
// ELU of a pair
__device__ __forceinline__ void elu_pair(float z0, float z1, float& a0, float& a1) {
    float t0, t1, e0, e1;
    upk2(mul2(pk2(z0, z1), pk2(1.4426950408889634f, 1.4426950408889634f)), t0, t1);
    upk2(add2(pk2(ex2f(t0), ex2f(t1)), pk2(-1.f, -1.f)), e0, e1);
    a0 = z0 > 0.f ? z0 : e0;
    a1 = z1 > 0.f ? z1 : e1;
}
// dA * ELU'(z) of a pair
__device__ __forceinline__ void dz_pair(float z0, float z1, float d0, float d1, float& r0, float& r1) {
    float t0, t1;
    upk2(mul2(pk2(z0, z1), pk2(1.4426950408889634f, 1.4426950408889634f)), t0, t1);
    const float e0 = z0 > 0.f ? 1.f : ex2f(t0), e1 = z1 > 0.f ? 1.f : ex2f(t1);
    upk2(mul2(pk2(d0, d1), pk2(e0, e1)), r0, r1);
}
// 32 values of one row -> the row's bf16 hi/lo entries of a [128 x 32] K-major tile (4 x 16-byte stores each)
__device__ __forceinline__ void store_row32(unsigned char* tile_hi, int row, const float v[32]) {
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        uint32_t h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) split_pair(v[8 * g + 2 * i], v[8 * g + 2 * i + 1], h[i], l[i]);
        const uint32_t o = umma::kmajor_off_b16(row, 8 * g, 32);
        *reinterpret_cast<uint4*>(tile_hi + o) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4*>(tile_hi + 8192 + o) = make_uint4(l[0], l[1], l[2], l[3]);
    }
}
// all 128 physics threads (warps 0-3): logical AND of a predicate, named barrier 1
__device__ __forceinline__ bool all128(bool p) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred q, o;\n\t"
        "setp.ne.u32 q, %1, 0;\n\t"
        "bar.red.and.pred o, 1, 128, q;\n\t"
        "selp.u32 %0, 1, 0, o;\n\t}" : "=r"(r) : "r"((uint32_t)p) : "memory");
    return r != 0;
}
__device__ __forceinline__ void sync128() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// ---- epilogues (warps 0-7; `half` = 0 for warps 0-3, 1 for warps 4-7) -------------------------------------------------
// forward: a Z chunk is 128 columns; this thread turns its 64 columns into ELU(z) as bf16 hi | lo IN PLACE: the 32 columns
// [s, s+32) of a sub-block become 16 packed hi columns [s, s+16) and 16 packed lo columns [s+16, s+32).
__device__ __forceinline__ void fwd_epilogue(uint32_t taddr) {   // taddr -> this thread's lane block, first of its 64 columns
    uint32_t v0[32], v1[32];
    umma::ld32(taddr, v0);
    umma::ld32(taddr + 32, v1);
    umma::wait_ld();
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a0, a1;
        elu_pair(__uint_as_float(v0[2 * i]), __uint_as_float(v0[2 * i + 1]), a0, a1);
        split_pair2(a0, a1, hi[i], lo[i]);
    }
    umma::st16(taddr, hi);
    umma::st16(taddr + 16, lo);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float a0, a1;
        elu_pair(__uint_as_float(v1[2 * i]), __uint_as_float(v1[2 * i + 1]), a0, a1);
        split_pair2(a0, a1, hi[i], lo[i]);
    }
    umma::st16(taddr + 32, hi);
    umma::st16(taddr + 48, lo);
    umma::wait_st();
}
// backward: a buffer is Z chunk (64 columns) | dA chunk (64 columns); this thread's 32 units: dz = dA * ELU'(z) as bf16
// hi | lo over its own 32 Z columns.
__device__ __forceinline__ void bwd_epilogue(uint32_t taddr) {   // taddr -> lane block, buffer base + half*32
    uint32_t z[32], d[32];
    umma::ld32(taddr, z);
    umma::ld32(taddr + 64, d);
    umma::wait_ld();
    uint32_t hi[16], lo[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        float r0, r1;
        dz_pair(__uint_as_float(z[2 * i]), __uint_as_float(z[2 * i + 1]), __uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1]), r0, r1);
        split_pair2(r0, r1, hi[i], lo[i]);
    }
    umma::st16(taddr, hi);
    umma::st16(taddr + 16, lo);
    umma::wait_st();
}
}  // namespace ktc

// MLP views handed to node_eval / the adjoint node step: overload resolution on these types selects the CTA-cooperative
// tensor-core evaluation (as MlpCoop selects the warp-cooperative one).
struct MlpTCF : MlpC<float> {
    unsigned char* sm;
    ktc::Bars* bars;
    uint32_t tbase, laneblk;
    int nch, row;
    mutable uint32_t ph;   // bit b: parity of barZ[b]; bit 3: parity of barO
    mutable int ev;        // node evaluations so far (trace builds)
    bool tr;
};

// forward: o[25] = W2 ELU(W1 x + b1) + b2 for this thread's row; called by all 128 physics threads together
template <typename T, int IN>
__device__ __forceinline__ void mlp_eval(const MlpTCF& M, const float* __restrict__ x, float* __restrict__ o) {
    static_assert(IN == 28, "tensor-core march: 28 inputs");
    KTC_TRACE(M.tr, 0, M.ev, 0);
#ifdef KC_TC_TRACE
    if (M.tr && g_ktc_trace && M.ev >= 100 && M.ev < 228) g_ktc_trace[3 * 8 * 32 + (M.ev - 100)] = clock64();
#endif
    float xv[32];
#pragma unroll
    for (int i = 0; i < 28; ++i) xv[i] = x[i];
    xv[28] = 1.f; xv[29] = 0.f; xv[30] = 0.f; xv[31] = 0.f;      // column 28 carries b1
    ktc::store_row32(M.sm + ktc::F_X, M.row, xv);
    umma::fence_async_smem();
    umma::fence_before();                                         // (orders this thread's last TMEM read of O before the next MMAs)
    umma::mbar_arrive(&M.bars->X);
    KTC_TRACE(M.tr, 0, M.ev, 1);
    for (int c = 0; c < M.nch; ++c) {
        const int buf = c % ktc::NBUF;
        umma::mbar_wait(&M.bars->Z[buf], (M.ph >> buf) & 1u);
        M.ph ^= 1u << buf;
        umma::fence_after();
        KTC_TRACE(M.tr, 0, M.ev, 2 + 2 * c);
        ktc::fwd_epilogue(M.tbase + M.laneblk + buf * 128);
        umma::fence_before();
        umma::mbar_arrive(&M.bars->A[buf]);
        KTC_TRACE(M.tr, 0, M.ev, 3 + 2 * c);
    }
    umma::mbar_wait(&M.bars->O, (M.ph >> 3) & 1u);
    M.ph ^= 8u;
    umma::fence_after();
    KTC_TRACE(M.tr, 0, M.ev, 20);
    uint32_t v[32];
    umma::ld32(M.tbase + M.laneblk + ktc::COL_O, v);
    umma::wait_ld();
#pragma unroll
    for (int c = 0; c < 25; ++c) o[c] = __uint_as_float(v[c]) + M.b2[c];
    KTC_TRACE(M.tr, 0, M.ev, 21);
    ++M.ev;
}

struct MlpTCB : MlpTCF {};
// backward: gx[28] = W1^T ((W2^T go) * ELU'(W1 x + b1)) for this thread's row
template <typename T, int IN>
__device__ __forceinline__ void mlp_input_vjp(const MlpTCB& M, const float* __restrict__ x, const float* __restrict__ go,
                                              float* __restrict__ gx) {
    static_assert(IN == 28, "tensor-core march: 28 inputs");
    float xv[32];
#pragma unroll
    for (int i = 0; i < 28; ++i) xv[i] = x[i];
    xv[28] = 1.f; xv[29] = 0.f; xv[30] = 0.f; xv[31] = 0.f;
    ktc::store_row32(M.sm + ktc::B_X, M.row, xv);
#pragma unroll
    for (int i = 0; i < 25; ++i) xv[i] = go[i];
#pragma unroll
    for (int i = 25; i < 32; ++i) xv[i] = 0.f;
    ktc::store_row32(M.sm + ktc::B_GO, M.row, xv);
    umma::fence_async_smem();
    umma::fence_before();
    umma::mbar_arrive(&M.bars->X);
    for (int c = 0; c < M.nch; ++c) {      // nch = chunks of 64 units here
        const int buf = c % ktc::NBUF;
        umma::mbar_wait(&M.bars->Z[buf], (M.ph >> buf) & 1u);
        M.ph ^= 1u << buf;
        umma::fence_after();
        ktc::bwd_epilogue(M.tbase + M.laneblk + buf * 128);
        umma::fence_before();
        umma::mbar_arrive(&M.bars->A[buf]);
    }
    umma::mbar_wait(&M.bars->O, (M.ph >> 3) & 1u);
    M.ph ^= 8u;
    umma::fence_after();
    uint32_t v[32];
    umma::ld32(M.tbase + M.laneblk + ktc::COL_O, v);
    umma::wait_ld();
#pragma unroll
    for (int k = 0; k < 28; ++k) gx[k] = __uint_as_float(v[k]);
}

namespace ktc {
// ---- the helper warps (4-7): second half of every epilogue, until the exit flag is raised ------------------------------
template <bool BWD>
__device__ __forceinline__ void helper_loop(Bars* bars, uint32_t tbase, uint32_t laneblk, int nch) {
    uint32_t ph = 0;
    int ev = 0;
    const bool tr = threadIdx.x == 128 && blockIdx.x == 0;
    (void)ev; (void)tr;
    while (true) {
        for (int c = 0; c < nch; ++c) {
            const int buf = c % NBUF;
            umma::mbar_wait(&bars->Z[buf], (ph >> buf) & 1u);
            ph ^= 1u << buf;
            if (*reinterpret_cast<volatile uint32_t*>(&bars->exit_flag)) return;
            umma::fence_after();
            KTC_TRACE(tr, 1, ev, 2 + 2 * c);
            if (BWD) bwd_epilogue(tbase + laneblk + buf * 128 + 32);
            else fwd_epilogue(tbase + laneblk + buf * 128 + 64);
            umma::fence_before();
            umma::mbar_arrive(&bars->A[buf]);
            KTC_TRACE(tr, 1, ev, 3 + 2 * c);
        }
        ++ev;
    }
}

// ---- the MMA warp (warp 8): all 32 lanes run the loop, the instructions are issued by one elected lane ------------------
// Descriptors: no swizzle, LBO 128 B; the 14-bit address field counts 16-byte units, so stepping through a tile is an add
// on the low word.  X / GO tiles and the W1 / W2^T images: [rows x 32] K-major, SBO 512; W2 / W1^T images:
// [32 x 512] K-major, SBO 8192.
__device__ __forceinline__ uint64_t desc_step(uint64_t d, uint32_t bytes) { return d + (bytes >> 4); }

// forward: Z chunk c = 128 units.  W1 image rows = units (chunk c at c*8192); in the W2 image 16 units = 256 bytes.
__device__ __forceinline__ void mma_loop_fwd(unsigned char* sm, Bars* bars, uint32_t tbase, int nch) {
    const uint32_t idZ = umma::make_idesc_bf16(128, 128), idO = umma::make_idesc_bf16(128, 32);
    const uint64_t dXh = umma::make_desc(umma::smem_u32(sm + F_X), 128, 512), dXl = desc_step(dXh, 8192);
    const uint64_t dW1h = umma::make_desc(umma::smem_u32(sm + F_W1), 128, 512), dW1l = desc_step(dW1h, IMG);
    const uint64_t dW2h = umma::make_desc(umma::smem_u32(sm + F_W2), 128, 8192), dW2l = desc_step(dW2h, IMG);
    uint32_t phX = 0, phA = 0;
    int evm = 0;
    const bool trm = (threadIdx.x & 31) == 0 && blockIdx.x == 0;
    (void)evm; (void)trm;
    auto gemm1 = [&](int c) {
        const uint32_t d = tbase + (c % NBUF) * 128, wo = c * 8192;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint64_t a = p == 1 ? dXl : dXh, b = desc_step(p == 2 ? dW1l : dW1h, wo);
#pragma unroll
            for (int kk = 0; kk < 2; ++kk)
                umma::mma_bf16_w(d, desc_step(a, kk * 256), desc_step(b, kk * 256), idZ, (p | kk) ? 1u : 0u);
        }
        umma::commit_w(&bars->Z[c % NBUF]);
    };
    while (true) {
        umma::mbar_wait(&bars->X, phX);
        phX ^= 1;
        if (*reinterpret_cast<volatile uint32_t*>(&bars->exit_flag)) {
            umma::mbar_arrive_w(&bars->Z[0]);      // releases the helper warps, which then see the flag
            return;
        }
        umma::fence_after();
        KTC_TRACE(trm, 2, evm, 0);
        for (int c = 0; c < nch && c < NBUF; ++c) gemm1(c);
        KTC_TRACE(trm, 2, evm, 1);
        for (int c = 0; c < nch; ++c) {
            const int buf = c % NBUF;
            umma::mbar_wait(&bars->A[buf], (phA >> buf) & 1u);
            phA ^= 1u << buf;
            umma::fence_after();
            KTC_TRACE(trm, 2, evm, 2 + 2 * c);
            const uint32_t ab = tbase + buf * 128;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint64_t b = desc_step(p == 2 ? dW2l : dW2h, c * 2048);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint32_t a = ab + (kk >> 2) * 64 + ((kk >> 1) & 1) * 32 + (kk & 1) * 8 + (p == 1 ? 16 : 0);
                    umma::mma_bf16_ts_w(tbase + COL_O, a, desc_step(b, kk * 256), idO, (c | p | kk) ? 1u : 0u);
                }
            }
            if (c + NBUF < nch) gemm1(c + NBUF);
            KTC_TRACE(trm, 2, evm, 3 + 2 * c);
        }
        umma::commit_w(&bars->O);
        KTC_TRACE(trm, 2, evm, 20);
        ++evm;
    }
}
// backward: chunk c = 64 units; buffer = Z (64 columns) | dA (64 columns).  W1 / W2^T images: rows = units (chunk c at
// c*4096); in the W1^T image 16 units = 256 bytes.
__device__ __forceinline__ void mma_loop_bwd(unsigned char* sm, Bars* bars, uint32_t tbase, int nch) {
    const uint32_t idZ = umma::make_idesc_bf16(128, 64), idO = umma::make_idesc_bf16(128, 32);
    const uint64_t dXh = umma::make_desc(umma::smem_u32(sm + B_X), 128, 512), dXl = desc_step(dXh, 8192);
    const uint64_t dGh = umma::make_desc(umma::smem_u32(sm + B_GO), 128, 512), dGl = desc_step(dGh, 8192);
    const uint64_t dW1h = umma::make_desc(umma::smem_u32(sm + B_W1), 128, 512), dW1l = desc_step(dW1h, IMG);
    const uint64_t dW2Th = umma::make_desc(umma::smem_u32(sm + B_W2T), 128, 512), dW2Tl = desc_step(dW2Th, IMG);
    const uint64_t dW1Th = umma::make_desc(umma::smem_u32(sm + B_W1T), 128, 8192), dW1Tl = desc_step(dW1Th, IMG);
    uint32_t phX = 0, phA = 0;
    auto gemm13 = [&](int c) {
        const uint32_t d = tbase + (c % NBUF) * 128, wo = c * 4096;
#pragma unroll
        for (int which = 0; which < 2; ++which) {
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint64_t a = which ? (p == 1 ? dGl : dGh) : (p == 1 ? dXl : dXh);
                const uint64_t b = desc_step(which ? (p == 2 ? dW2Tl : dW2Th) : (p == 2 ? dW1l : dW1h), wo);
#pragma unroll
                for (int kk = 0; kk < 2; ++kk)
                    umma::mma_bf16_w(d + which * 64, desc_step(a, kk * 256), desc_step(b, kk * 256), idZ, (p | kk) ? 1u : 0u);
            }
        }
        umma::commit_w(&bars->Z[c % NBUF]);
    };
    while (true) {
        umma::mbar_wait(&bars->X, phX);
        phX ^= 1;
        if (*reinterpret_cast<volatile uint32_t*>(&bars->exit_flag)) {
            umma::mbar_arrive_w(&bars->Z[0]);
            return;
        }
        umma::fence_after();
        for (int c = 0; c < nch && c < NBUF; ++c) gemm13(c);
        for (int c = 0; c < nch; ++c) {
            const int buf = c % NBUF;
            umma::mbar_wait(&bars->A[buf], (phA >> buf) & 1u);
            phA ^= 1u << buf;
            umma::fence_after();
            const uint32_t ab = tbase + buf * 128;
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint64_t b = desc_step(p == 2 ? dW1Tl : dW1Th, c * 1024);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t a = ab + (kk >> 1) * 32 + (kk & 1) * 8 + (p == 1 ? 16 : 0);
                    umma::mma_bf16_ts_w(tbase + COL_O, a, desc_step(b, kk * 256), idO, (c | p | kk) ? 1u : 0u);
                }
            }
            if (c + NBUF < nch) gemm13(c + NBUF);
        }
        umma::commit_w(&bars->O);
    }
}

// common prologue: TMEM, barriers, resident weight images
__device__ __forceinline__ Bars* cta_setup(unsigned char* sm, int misc_off, const unsigned char* __restrict__ img, int img_bytes) {
    Bars* bars = reinterpret_cast<Bars*>(sm + misc_off);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) umma::tmem_alloc(&bars->tmem_slot, 512);
    if (tid == 32) {
        umma::mbar_init(&bars->X, 128);
        umma::mbar_init(&bars->O, 1);
        for (int i = 0; i < NBUF; ++i) { umma::mbar_init(&bars->Z[i], 1); umma::mbar_init(&bars->A[i], 256); }
        bars->exit_flag = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    const uint4* src = reinterpret_cast<const uint4*>(img);
    uint4* dst = reinterpret_cast<uint4*>(sm);
    for (int e = tid; e < img_bytes / 16; e += THREADS) dst[e] = src[e];
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    return bars;
}
__device__ __forceinline__ void cta_teardown(Bars* bars, uint32_t tbase) {
    // physics threads: raise the flag, then a last arrival on barX wakes the MMA thread, which releases the helpers
    const int tid = threadIdx.x;
    if (tid < 128) {
        if (tid == 0) *reinterpret_cast<volatile uint32_t*>(&bars->exit_flag) = 1u;
        sync128();
        __threadfence_block();
        umma::fence_before();
        umma::mbar_arrive(&bars->X);
    }
    umma::fence_before();
    __syncthreads();
    if ((tid >> 5) == 0) umma::tmem_dealloc(tbase, 512);
}

template <typename T> __device__ __forceinline__ T* rod_base_tc(T* trajD, int64_t b, int T_, int N) {
    return trajD + (size_t)(b >> 5) * ((size_t)T_ * N * 25 * 32) + (b & 31);
}
}  // namespace ktc

// =========================================================================================================================
// Forward: KNODE rollout, persistent CTAs over groups of 16 rods, Newton shooting solve with a per-march finite-difference
// Jacobian and the linearised final correction of kc_rollout_wide_lin_kernel (the marched states of the 7 shooting points
// are kept — here in an L2-resident scratch, [25*N][128 rows] per CTA — and once the Newton step is so small that its
// second-order remainder is below the tolerance, the state at G + dG is formed from them instead of marching again: about
// 2.5 instead of 3.6 joint marches per step), MLP on tcgen05.  Trajectory in the device layout [tile][T][N][25][32].
struct ScrSink {   // marched state of this row, "output order" e = r*N + j, element stride 128 (rows of the CTA)
    float* p; int n; bool on;
    __device__ __forceinline__ void put(int j, const float y[19]) {
        if (on) {
#pragma unroll
            for (int r = 0; r < 19; ++r) p[(size_t)(r * n + j) * 128] = y[r];
        }
    }
    __device__ __forceinline__ void putz(int j, const float z[6]) {
        if (on) {
#pragma unroll
            for (int c = 0; c < 6; ++c) p[(size_t)((19 + c) * n + j) * 128] = z[c];
        }
    }
};

template <bool DIAG>
__global__ void __launch_bounds__(ktc::THREADS, 1)
kc_knode_tc_fwd_kernel(const __grid_constant__ RodC<float> P, const unsigned char* __restrict__ img, const float* __restrict__ b2,
                       int hidden, int64_t B, int T_, const float* __restrict__ tensions, const float* __restrict__ y0,
                       const float* __restrict__ z0, float* trajD, float tol, int max_iter, float fd_eps, float* Gout,
                       int32_t* iters, float* __restrict__ stscr_all) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int NH = 12, WG = ktc::RPC, LS = 32;
    const int N = P.N, tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // (provably warp-uniform)
    ktc::Bars* bars = ktc::cta_setup(sm, ktc::F_MISC, img, 4 * ktc::IMG);
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
    const int nch = (hidden + 127) / 128;
    if (warp == 8) {
        ktc::mma_loop_fwd(sm, bars, tbase, nch);
    } else if (warp >= 4) {
        ktc::helper_loop<false>(bars, tbase, laneblk, nch);
    } else {
        MlpTCF M{};
        M.sm = sm; M.bars = bars; M.tbase = tbase; M.laneblk = laneblk; M.nch = nch; M.row = tid; M.ph = 0; M.b2 = b2;
        M.ev = 0; M.tr = tid == 0 && blockIdx.x == 0;
        M.hidden = hidden; M.in_dim = 28;
        const int lane = tid & 31, g = tid >> 3, k = tid & 7;
        const unsigned full = 0xffffffffu;
        const int NV = 25 * N;
        float* Hs = reinterpret_cast<float*>(sm + ktc::F_HIST) + g;                 // history of the step being solved
        float* As = Hs + (size_t)(N - 1) * NH * WG;                                  // previous accepted state (history rows)
        float* stscr = stscr_all + (size_t)blockIdx.x * NV * 128;      // this CTA's marched states
        const size_t tstride = (size_t)25 * N * LS;
        const int64_t ngroups = (B + WG - 1) / WG;
        for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            const int64_t b_raw = grp * WG + g;
            const bool valid = b_raw < B;
            const int64_t b = valid ? b_raw : B - 1;   // surplus groups shadow the last rod and never store
            float* traj_b = ktc::rod_base_tc(trajD, b, T_, N);
            if (k == 0 && valid) {
                rollout_init<float, LS>(P, y0 ? y0 + (size_t)b * 19 * N : nullptr, z0 ? z0 + (size_t)b * 6 * N : nullptr, traj_b);
                if (Gout) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = 0.f;
                }
                if (iters) iters[(size_t)b * T_] = 0;
            }
            __threadfence_block();
            ktc::sync128();   // (surplus groups read the last rod's initial state written by another warp)
            // history of step 0 and "previous accepted state" from the initial state: state[-1] := state[0] (knode.py:65-66)
            for (int e = k; e < (N - 1) * NH; e += 8) {
                const int j = e / NH, s = e - j * NH;
                const float v = traj_b[((size_t)j * 25 + slot_row<NH>(s)) * LS];
                As[(size_t)e * WG] = v;
                Hs[(size_t)e * WG] = (P.c1 + P.c2) * v;
            }
            float zlast[6];   // z[:, N-1] is never written by the march (cosserat_ode.py:198-201)
#pragma unroll
            for (int c = 0; c < 6; ++c) zlast[c] = traj_b[((size_t)(N - 1) * 25 + 19 + c) * LS];
            __syncwarp();
            float G[6], Gm1[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) { G[i] = 0.f; Gm1[i] = 0.f; }
            const float* ten = tensions + (size_t)b * T_ * 4;
            for (int t = 0; t < T_ - 1; ++t) {
                float tn[4], tf[3];
#pragma unroll
                for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
                tendon_force(P, tn, tf);
                float* nxt = traj_b + (size_t)(t + 1) * tstride;
                float Gp[6], w[6], Gm[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) { Gp[i] = G[i]; G[i] = G[i] + (G[i] - Gm1[i]); w[i] = 0.f; Gm[i] = 0.f; }   // linear predictor
                bool done = false, first = true;
                int status = 0, marches = 0;
                float Cest = 0.f, sprev = 0.f;   // curvature estimate: from THIS step's own iterations only
                HistView<float, NH, WG> H{Hs};
                while (true) {
                    float eps[6], Ge[6], F[6];
                    // the marched point is frozen once the rod is done: its re-marches (while other rods of the CTA still
                    // iterate) reproduce the stored states bit for bit
#pragma unroll
                    for (int i = 0; i < 6; ++i) Gm[i] = done ? Gm[i] : G[i];
                    wide_eps(Gm, fd_eps, eps);
#pragma unroll
                    for (int i = 0; i < 6; ++i) Ge[i] = Gm[i] + ((k == i + 1) ? eps[i] : 0.f);
                    // the first joint march of a step is never the accepted one: it stores nothing
                    ScrSink S{stscr + tid, N, !first && k < 7};
                    rod_march<float, DIAG, 28, NH>(P, M, Ge, tf, H, S, F);
                    float Fall[7][6];
#pragma unroll
                    for (int c = 0; c < 7; ++c) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) Fall[c][i] = __shfl_sync(full, F[i], (lane & ~7) | c);
                    }
                    if (!done) {
                        ++marches;
                        const int r = wide_decide_lin(Fall, G, eps, tol, Cest, sprev, w);
                        if (r != 0) { done = true; status = r; }
                        else if (marches >= max_iter) { done = true; status = -1; }
                    }
                    // a rod that converged on the first march is re-marched once at its frozen point so that its state exists
                    if (ktc::all128(done) && !first) break;
                    first = false;
                }
#pragma unroll
                for (int i = 0; i < 6; ++i) { Gm1[i] = Gp[i]; w[i] = (status == 2) ? w[i] : 0.f; }
                __syncwarp();
                // accepted state = base state + first-order correction (w = 0 unless status == 2); the 8 lanes of the rod
                // split the elements of the time slice (e = r*N + j), four elements per lane in flight.  The next step's
                // history is formed here from the accepted values and the previous accepted state kept in shared memory —
                // nothing is read back from the trajectory, and a shadow group recomputes its rod's values itself, so the
                // step needs no CTA-wide barrier.
                {
                    const float* __restrict__ sp0 = stscr + (tid & ~7);
                    float* __restrict__ out_t = nxt;
                    for (int e0 = k; e0 < NV; e0 += 32) {
                        float s0[4], sc[4][6];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int e = e0 + 8 * u < NV ? e0 + 8 * u : NV - 1;
                            const float* sp = sp0 + (size_t)e * 128;
                            s0[u] = sp[0];
#pragma unroll
                            for (int c = 0; c < 6; ++c) sc[u][c] = sp[c + 1];
                        }
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int e = e0 + 8 * u;
                            const int r = e / N, j = e - r * N;
                            if (e >= NV || (r >= 19 && j == N - 1)) continue;    // (tip z: never marched, zlast below)
                            float va = (sc[u][0] - s0[u]) * w[0], vb = (sc[u][1] - s0[u]) * w[1];
                            va += (sc[u][2] - s0[u]) * w[2]; vb += (sc[u][3] - s0[u]) * w[3];
                            va += (sc[u][4] - s0[u]) * w[4]; vb += (sc[u][5] - s0[u]) * w[5];
                            const float v = s0[u] + (va + vb);
                            if (valid) out_t[((size_t)j * 25 + r) * LS] = v;
                            if (r >= 13 && j < N - 1) {
                                const size_t hi = (size_t)(j * NH + r - 13) * WG;
                                Hs[hi] = P.c1 * v + P.c2 * As[hi];
                                As[hi] = v;
                            }
                        }
                    }
                }
                if (k == 0 && valid) {
                    const size_t o = (size_t)(N - 1) * 25 * LS;
#pragma unroll
                    for (int c = 0; c < 6; ++c) nxt[o + (19 + c) * LS] = zlast[c];
                    if (Gout) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) Gout[((size_t)b * T_ + t + 1) * 6 + i] = G[i];
                    }
                    if (iters) iters[(size_t)b * T_ + t + 1] = status > 0 ? marches : -marches;
                }
                __syncwarp();
            }
        }
    }
    ktc::cta_teardown(bars, tbase);
}

// =========================================================================================================================
// Forward, LARGE batches: one row per rod (up to 128 rods per CTA), quasi-Newton shooting solve of the one-rod-per-lane kernel
// (shoot_step, kc_rod.cuh: linear predictor, Broyden-updated inverse Jacobian kept across time steps, finite-difference
// rebuild on the first step or after two stalled iterations) run in LOCK STEP over the CTA — every march is one MLP per node
// for all rods, so a rod that has converged keeps marching its frozen point (without storing) until the slowest rod of the
// CTA is done.  Per rod-step this evaluates the MLP ~7 times instead of 8 rows x ~3 joint marches: 3-4x the throughput of
// the 8-rows-per-rod kernel once the batch fills the chip with 128-row CTAs; the per-march latency is the same, so small
// batches (training, <= 8k rods) stay with the Newton kernel above.
struct LsState {   // per-rod state of the lock-step solve (registers)
    float G[6], F[6], dG[6], fprev, eps_k;
    int phase, k, marches, stalled;
    bool fin, converged, failed;
};

template <bool DIAG>
__global__ void __launch_bounds__(ktc::THREADS, 1)
kc_knode_tc_fwd1_kernel(const __grid_constant__ RodC<float> P, const unsigned char* __restrict__ img, const float* __restrict__ b2,
                        int hidden, int64_t B, int T_, int rpc, const float* __restrict__ tensions, const float* __restrict__ y0,
                        const float* __restrict__ z0, float* trajD, float tol, int max_iter, float fd_eps, float* Gout,
                        int32_t* iters, float* __restrict__ hist_all, int hist_in_smem) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int NH = 12, LS = 32, CS = 128;
    enum { PRED = 0, FD = 1, BROY = 2 };
    const int N = P.N, tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    ktc::Bars* bars = ktc::cta_setup(sm, ktc::F_MISC, img, 4 * ktc::IMG);
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
    const int nch = (hidden + 127) / 128;
    if (warp == 8) {
        ktc::mma_loop_fwd(sm, bars, tbase, nch);
    } else if (warp >= 4) {
        ktc::helper_loop<false>(bars, tbase, laneblk, nch);
    } else {
        MlpTCF M{};
        M.sm = sm; M.bars = bars; M.tbase = tbase; M.laneblk = laneblk; M.nch = nch; M.row = tid; M.ph = 0; M.b2 = b2;
        M.ev = 0; M.tr = false;
        M.hidden = hidden; M.in_dim = 28;
        float* stm = reinterpret_cast<float*>(sm + ktc::F_HIST) + tid;                       // ShootMem [49][128]
        float* Hs = hist_in_smem ? stm - tid + (size_t)KC_SHOOT_SLOTS * CS + tid             // history [N-1][12][128]
                                 : hist_all + (size_t)blockIdx.x * (N - 1) * NH * CS + tid;
        const ShootMem<float, CS> st{stm};
        const size_t tstride = (size_t)25 * N * LS;
        const int64_t ngroups = (B + rpc - 1) / rpc;
        for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            // rows beyond the group's rods shadow one of them (same values, no stores): every row takes part in every MLP
            const int64_t b_raw = grp * rpc + (tid % rpc);
            const bool valid = tid < rpc && b_raw < B;
            const int64_t b = b_raw < B ? b_raw : B - 1;
            float* traj_b = ktc::rod_base_tc(trajD, b, T_, N);
            st.reset();
            if (valid) {
                rollout_init<float, LS>(P, y0 ? y0 + (size_t)b * 19 * N : nullptr, z0 ? z0 + (size_t)b * 6 * N : nullptr, traj_b);
                if (Gout) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) Gout[(size_t)b * T_ * 6 + i] = 0.f;
                }
                if (iters) iters[(size_t)b * T_] = 0;
            }
            __threadfence_block();
            ktc::sync128();
            build_history<float, NH, LS, CS>(P, traj_b, traj_b, Hs);
            const float* ten = tensions + (size_t)b * T_ * 4;
            for (int t = 0; t < T_ - 1; ++t) {
                float tn[4], tf[3];
#pragma unroll
                for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
                tendon_force(P, tn, tf);
                float* cur = traj_b + (size_t)t * tstride;
                float* nxt = cur + tstride;
                HistView<float, NH, CS> H{Hs};
                // ---- shoot_step (kc_rod.cuh) in lock step ----
                LsState q;
#pragma unroll
                for (int i = 0; i < 6; ++i) { q.G[i] = st.G(i) + (st.G(i) - st.Gm1(i)); q.F[i] = 0.f; q.dG[i] = 0.f; }
                q.phase = PRED; q.k = 0; q.marches = 0; q.stalled = 0; q.fprev = 0.f; q.eps_k = 0.f;
                q.fin = false; q.converged = false; q.failed = false;
                while (true) {
                    float Ge[6], Fn[6];
#pragma unroll
                    for (int i = 0; i < 6; ++i) Ge[i] = q.G[i];
                    if (!q.fin && q.phase == FD) {
                        float gk = 0.f;
#pragma unroll
                        for (int i = 0; i < 6; ++i) if (i == q.k) gk = q.G[i];
                        q.eps_k = fd_eps * kc_max(1.f, kc_abs(gk));
#pragma unroll
                        for (int i = 0; i < 6; ++i) if (i == q.k) Ge[i] += q.eps_k;
                    }
                    TrajSinkPred<float, LS, NH, CS> S{nxt, nullptr, valid && !q.fin, N - 1};
                    rod_march<float, DIAG, 28, NH>(P, M, Ge, tf, H, S, Fn);
                    if (!q.fin) {
                        ++q.marches;
                        bool take_step = false;
                        if (q.phase == FD) {
                            const float ie = 1.f / q.eps_k;
#pragma unroll
                            for (int i = 0; i < 6; ++i) {
                                const float d = (Fn[i] - q.F[i]) * ie;
#pragma unroll
                                for (int c = 0; c < 6; ++c) if (c == q.k) st.J(i, c) = d;
                            }
                            if (++q.k == 6) {
                                float A[36];
#pragma unroll
                                for (int i = 0; i < 36; ++i) A[i] = st.J(i / 6, i % 6);
                                if (!inv6(A)) { q.failed = true; q.fin = true; }
                                else {
#pragma unroll
                                    for (int i = 0; i < 36; ++i) st.J(i / 6, i % 6) = A[i];
                                    st.haveJ() = 1.f;
                                    q.stalled = 0;
                                    take_step = true;
                                }
                            }
                        } else {
                            if (q.phase == BROY) {
                                float yv[6];
#pragma unroll
                                for (int i = 0; i < 6; ++i) yv[i] = Fn[i] - q.F[i];
                                broyden_update(st, q.dG, yv);
                            }
#pragma unroll
                            for (int i = 0; i < 6; ++i) q.F[i] = Fn[i];
                            const float fn = norm_inf6(q.F);
                            if (!(fn == fn)) { q.failed = true; q.fin = true; }
                            else {
                                if (q.phase == BROY) q.stalled = (fn > 0.5f * q.fprev) ? q.stalled + 1 : 0;
                                q.fprev = fn;
                                q.converged = fn <= tol * kc_max(1.f, norm_inf6(q.G));
                                if (q.converged || q.marches >= max_iter) q.fin = true;
                                else if (st.haveJ() == 0.f || q.stalled >= 2) { q.phase = FD; q.k = 0; }
                                else take_step = true;
                            }
                        }
                        if (take_step) {
#pragma unroll
                            for (int i = 0; i < 6; ++i) {
                                float a = 0.f;
#pragma unroll
                                for (int j = 0; j < 6; ++j) a += st.J(i, j) * q.F[j];
                                q.dG[i] = -a;
                            }
#pragma unroll
                            for (int i = 0; i < 6; ++i) q.G[i] += q.dG[i];
                            q.phase = BROY;
                        }
                    }
                    if (ktc::all128(q.fin)) break;
                }
                if (!q.failed) {
#pragma unroll
                    for (int i = 0; i < 6; ++i) { st.Gm1(i) = st.G(i); st.G(i) = q.G[i]; }
                }
                if (valid) {
                    const size_t o = (size_t)(N - 1) * 25 * LS;
#pragma unroll
                    for (int c = 0; c < 6; ++c) nxt[o + (19 + c) * LS] = cur[o + (19 + c) * LS];   // z[:, N-1] never changes
                    if (Gout) {
#pragma unroll
                        for (int i = 0; i < 6; ++i) Gout[((size_t)b * T_ + t + 1) * 6 + i] = st.G(i);
                    }
                    if (iters) iters[(size_t)b * T_ + t + 1] = (q.converged && !q.failed) ? q.marches : -q.marches;
                }
                __threadfence_block();
                ktc::sync128();       // shadow rows read the state their rod's owner has just stored
                build_history<float, NH, LS, CS>(P, nxt, cur, Hs);
            }
        }
    }
    ktc::cta_teardown(bars, tbase);
}

// =========================================================================================================================
// Backward through the rollout for the same shape: persistent CTAs over groups of 16 rods, steps in reverse, ONE joint
// adjoint march per step (see the header).  traj / gtraj in the reference layout [B][T][25][N].
//   rowscr : per CTA [(N-1)*SV][128] floats — per node and row: 12 history cotangents, 25 output cotangents
//   hscr   : per CTA [3][(N-1)*12][16] floats — history cotangents of steps t+1, t+2 and the one being built
template <bool DIAG>
__global__ void __launch_bounds__(ktc::THREADS, 1)
kc_knode_tc_bwd_kernel(const __grid_constant__ RodC<float> P, const unsigned char* __restrict__ img, int hidden, int64_t B,
                       int T_, const float* __restrict__ tensions, const float* __restrict__ traj,
                       const float* __restrict__ gtraj, float* __restrict__ gten, float* __restrict__ xs,
                       float* __restrict__ gos, float* __restrict__ rowscr_all, float* __restrict__ hscr_all) {
    extern __shared__ __align__(1024) unsigned char sm[];
    constexpr int NH = 12, WG = ktc::RPC, SV = ktc::SV;
    const int N = P.N, Nm1 = N - 1, tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    ktc::Bars* bars = ktc::cta_setup(sm, ktc::B_MISC, img + 4 * ktc::IMG, 6 * ktc::IMG);
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
    const int nch = (hidden + 63) / 64;
    if (warp == 8) {
        ktc::mma_loop_bwd(sm, bars, tbase, nch);
    } else if (warp >= 4) {
        ktc::helper_loop<true>(bars, tbase, laneblk, nch);
    } else {
        MlpTCB M{};
        M.sm = sm; M.bars = bars; M.tbase = tbase; M.laneblk = laneblk; M.nch = nch; M.row = tid; M.ph = 0;
        M.hidden = hidden; M.in_dim = 28;
        const int lane = tid & 31, g = tid >> 3, k = tid & 7;
        const unsigned full = 0xffffffffu;
        float* rowscr = rowscr_all + (size_t)blockIdx.x * Nm1 * SV * 128 + tid;
        float* hscr = hscr_all + (size_t)blockIdx.x * 3 * Nm1 * NH * WG + g;
        const size_t harr = (size_t)Nm1 * NH * WG;
        const size_t ts = (size_t)25 * N;
        const int64_t ngroups = (B + WG - 1) / WG;
        for (int64_t grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
            const int64_t b_raw = grp * WG + g;
            const bool valid = b_raw < B;
            const int64_t b = valid ? b_raw : B - 1;
            const float* traj_b = traj + (size_t)b * T_ * ts;
            const float* gtraj_b = gtraj + (size_t)b * T_ * ts;
            const float* ten = tensions + (size_t)b * T_ * 4;
            int ia = 0, ib = 1, ic = 2;   // hscr arrays: g_hist of step t+1, of step t+2, of step t (being built)
            for (int e = k; e < Nm1 * NH; e += 8) { hscr[ia * harr + (size_t)e * WG] = 0.f; hscr[ib * harr + (size_t)e * WG] = 0.f; }
            __syncwarp();
            for (int t = T_ - 2; t >= 0; --t) {
                const float* s1 = traj_b + (size_t)(t + 1) * ts;
                const float* s0 = traj_b + (size_t)t * ts;
                const float* sm1 = traj_b + (size_t)(t > 0 ? t - 1 : 0) * ts;
                const float* lam = gtraj_b + (size_t)(t + 1) * ts;
                const float* Ha = hscr + ia * harr;
                const float* Hb = hscr + ib * harr;
                float tn[4], tf[3];
#pragma unroll
                for (int i = 0; i < 4; ++i) tn[i] = ten[(size_t)t * 4 + i];
                tendon_force(P, tn, tf);
                float yb[19], gtf_acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
                for (int r = 0; r < 19; ++r) yb[r] = k == 0 ? lam[r * N + (N - 1)] : ((k >= 1 && k <= 6 && r == 6 + k) ? 1.f : 0.f);
                for (int j = Nm1 - 1; j >= 0; --j) {
                    float y[19], hist[NH], cys[19], cz[6];
#pragma unroll
                    for (int r = 0; r < 19; ++r) { y[r] = s1[r * N + j]; cys[r] = P.ds * yb[r]; }
#pragma unroll
                    for (int s = 0; s < NH; ++s) hist[s] = P.c1 * s0[(13 + s) * N + j] + P.c2 * sm1[(13 + s) * N + j];
#pragma unroll
                    for (int c = 0; c < 6; ++c) {
                        float v = 0.f;
                        if (k == 0) v = lam[(19 + c) * N + j] + P.c1 * Ha[(size_t)(j * NH + 6 + c) * WG] + P.c2 * Hb[(size_t)(j * NH + 6 + c) * WG];
                        cz[c] = v;
                    }
                    // MLP part: x = [y; z_pre; tf], go = [cys; cz]
                    float ys0[19], zp[6], x[28], go[25], gx[28];
                    rod_ode<float, DIAG>(P, y, hist, hist + 3, hist + 6, hist + 9, tf, ys0, zp);
#pragma unroll
                    for (int i = 0; i < 19; ++i) { x[i] = y[i]; go[i] = cys[i]; }
#pragma unroll
                    for (int i = 0; i < 6; ++i) { x[19 + i] = zp[i]; go[19 + i] = cz[i]; }
#pragma unroll
                    for (int i = 0; i < 3; ++i) x[25 + i] = tf[i];
                    if (k == 0 && valid) {
                        float* xq = xs + (((size_t)b * (T_ - 1) + t) * Nm1 + j) * 28;
#pragma unroll
                        for (int i = 0; i < 28; i += 4) *reinterpret_cast<float4*>(xq + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
                    }
                    float* rs = rowscr + (size_t)j * SV * 128;
#pragma unroll
                    for (int c = 0; c < 25; ++c) rs[(size_t)(NH + c) * 128] = go[c];
                    mlp_input_vjp<float, 28>(M, x, go, gx);
                    float czt[6], gy[19], gqh[3], gwh[3], gvh[3], guh[3], gtf[3];
#pragma unroll
                    for (int i = 0; i < 6; ++i) czt[i] = cz[i] + gx[19 + i];
                    rod_ode_vjp<float, DIAG>(P, y, hist, hist + 3, hist + 6, hist + 9, tf, cys, czt, gy, gqh, gwh, gvh, guh, gtf);
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        rs[(size_t)i * 128] = gqh[i]; rs[(size_t)(3 + i) * 128] = gwh[i];
                        rs[(size_t)(6 + i) * 128] = gvh[i]; rs[(size_t)(9 + i) * 128] = guh[i];
                        gtf_acc[i] += gtf[i] + gx[25 + i];
                    }
                    // cotangent of y_j: direct (loss + later histories) + pass-through of the Euler update + node Jacobian
#pragma unroll
                    for (int r = 0; r < 19; ++r) {
                        float d = 0.f;
                        if (k == 0) {
                            d = lam[r * N + j];
                            if (r >= 13) d += P.c1 * Ha[(size_t)(j * NH + r - 13) * WG] + P.c2 * Hb[(size_t)(j * NH + r - 13) * WG];
                        }
                        yb[r] = d + yb[r] + gy[r] + gx[r];
                    }
                }
                // implicit function theorem: rows 1..6 hold d y_N[7+i] / d G at the base -> A mu = Gbar, A[kx][i] = yb_i[7+kx]
                float A[36], mu[6];
#pragma unroll
                for (int kx = 0; kx < 6; ++kx) {
                    mu[kx] = __shfl_sync(full, yb[7 + kx], lane & ~7);
#pragma unroll
                    for (int i = 0; i < 6; ++i) A[kx * 6 + i] = __shfl_sync(full, yb[7 + kx], (lane & ~7) | (i + 1));
                }
#pragma unroll
                for (int p = 0; p < 6; ++p) {      // Gaussian elimination, no pivoting (A = [[I, X^T], [0, I]] + small)
                    const float ip = 1.f / A[p * 6 + p];
#pragma unroll
                    for (int r = p + 1; r < 6; ++r) {
                        const float f = A[r * 6 + p] * ip;
#pragma unroll
                        for (int c = p + 1; c < 6; ++c) A[r * 6 + c] -= f * A[p * 6 + c];
                        mu[r] -= f * mu[p];
                    }
                }
#pragma unroll
                for (int r = 5; r >= 0; --r) {
                    float s = mu[r];
#pragma unroll
                    for (int c = r + 1; c < 6; ++c) s -= A[r * 6 + c] * mu[c];
                    mu[r] = s / A[r * 6 + r];
                }
                // total cotangent = row 0 + sum_i (-mu_i) row_i: weight of this lane's row
                float wk = k == 0 ? 1.f : 0.f;
#pragma unroll
                for (int i = 0; i < 6; ++i) if (k == i + 1) wk = -mu[i];
                float* Hc = hscr + ic * harr;
                __syncwarp();
                for (int v0 = 0; v0 < Nm1 * SV; v0 += 32) {     // 32 loads in flight (the scratch sits in L2)
                    float a[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int v = v0 + q;
                        a[q] = v < Nm1 * SV ? rowscr[(size_t)v * 128] * wk : 0.f;
                    }
#pragma unroll
                    for (int q8 = 0; q8 < 4; ++q8) {
                        float mine = 0.f;
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            float t = a[q8 * 8 + q];
                            t += __shfl_xor_sync(full, t, 1);
                            t += __shfl_xor_sync(full, t, 2);
                            t += __shfl_xor_sync(full, t, 4);
                            if (q == k) mine = t;
                        }
                        const int v = v0 + q8 * 8 + k;
                        if (v < Nm1 * SV) {
                            const int j = v / SV, s = v - j * SV;
                            if (s < NH) Hc[(size_t)(j * NH + s) * WG] = mine;
                            else if (valid) gos[(((size_t)b * (T_ - 1) + t) * Nm1 + j) * 25 + (s - NH)] = mine;
                        }
                    }
                }
                if (gten) {
                    float gt[3];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        float a = gtf_acc[i] * wk;
                        a += __shfl_xor_sync(full, a, 1);
                        a += __shfl_xor_sync(full, a, 2);
                        a += __shfl_xor_sync(full, a, 4);
                        gt[i] = a;
                    }
                    if (k < 4 && valid)
                        gten[((size_t)b * T_ + t) * 4 + k] = P.tdirs[k * 3] * gt[0] + P.tdirs[k * 3 + 1] * gt[1] + P.tdirs[k * 3 + 2] * gt[2];
                }
                __syncwarp();
                const int tmp = ib; ib = ia; ia = ic; ic = tmp;   // Hb <- Ha, Ha <- Hc
            }
            if (gten && k < 4 && valid) gten[((size_t)b * T_ + (T_ - 1)) * 4 + k] = 0.f;   // the last control is never applied
        }
    }
    ktc::cta_teardown(bars, tbase);
}

// ---- weight images ------------------------------------------------------------------------------------------------------
// img: W1 hi|lo [512 units x 32] (column 28 = b1), W2 hi|lo [32 outs x 512 units], then for the backward W1 again, W2^T hi|lo
// [512 units x 32 outs] and W1^T hi|lo [32 inputs x 512 units] (columns >= 28 zero); all bf16, K-major, no swizzle.
__global__ void kc_knode_tc_prep_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                                        int hidden, unsigned char* __restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 512 * 32; e += gridDim.x * blockDim.x) {
        const int u = e >> 5, k = e & 31;
        const float w1 = u < hidden ? (k < 28 ? W1[(size_t)u * 28 + k] : (k == 28 ? b1[u] : 0.f)) : 0.f;
        const float w2 = (u < hidden && k < 25) ? W2[(size_t)k * hidden + u] : 0.f;
        const float w1t = (u < hidden && k < 28) ? W1[(size_t)u * 28 + k] : 0.f;
        auto put = [&](int image, uint32_t off, float w) {
            const __nv_bfloat16 h = __float2bfloat16_rn(w);
            const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
            *reinterpret_cast<__nv_bfloat16*>(img + (size_t)image * 2 * ktc::IMG + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(img + (size_t)image * 2 * ktc::IMG + ktc::IMG + off) = l;
        };
        // forward set (images 0, 1) and backward set (images 2, 3, 4): each is copied to shared memory as one block
        put(0, umma::kmajor_off_b16(u, k, 32), w1);      // W1   : rows = units, k = inputs
        put(1, umma::kmajor_off_b16(k, u, 512), w2);     // W2   : rows = outputs, k = units
        put(2, umma::kmajor_off_b16(u, k, 32), w1);      // W1 again
        put(3, umma::kmajor_off_b16(u, k, 32), w2);      // W2^T : rows = units, k = outputs
        put(4, umma::kmajor_off_b16(k, u, 512), w1t);    // W1^T : rows = inputs, k = units
    }
}

// ---- host side (called by kc_rollout.cu / kc_bptt.cu) ----------------------------------------------------------------
bool kc_knode_tc_eligible(int dtype, const kc_mlp* mlp, int N, int method) {
    if (!mlp || dtype != KC_F32 || mlp->in_dim != 28 || mlp->hidden > 512 || method != KC_MARCH_EULER) return false;
    if ((size_t)ktc::F_HIST + (size_t)2 * (N - 1) * 12 * ktc::RPC * 4 > 227 * 1024) return false;
    const char* e = getenv("KC_ROLLOUT_TC");
    if (e && e[0] == '0') return false;
    return true;
}
size_t kc_knode_tc_img_bytes() { return (size_t)10 * ktc::IMG; }
static int tc_bwd_grid(int64_t B) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t ng = (B + ktc::RPC - 1) / ktc::RPC;
    return (int)(ng < sms ? (ng > 0 ? ng : 1) : sms);
}
size_t kc_knode_tc_bwd_scratch_bytes(int N, int64_t B) {
    const size_t per_cta = ((size_t)(N - 1) * ktc::SV * 128 + (size_t)3 * (N - 1) * 12 * ktc::RPC) * sizeof(float);
    return per_cta * (size_t)tc_bwd_grid(B) + 256;
}
static int tc_prep(const kc_mlp* mlp, unsigned char* img, cudaStream_t st) {
    kc_knode_tc_prep_kernel<<<32, 256, 0, st>>>((const float*)mlp->W1, (const float*)mlp->b1, (const float*)mlp->W2, mlp->hidden, img);
    KC_CHECK_LAUNCH("kc_knode_tc_prep_kernel");
    return KC_OK;
}
static int tc_sms() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}
// rods per CTA of the one-row-per-rod kernel: a round of CTAs costs the same whether they hold 32 or 128 rods, so use the
// fewest rounds and spread the rods evenly over them (multiples of a warp)
static int tc_fwd1_rpc(int64_t B) {
    const int64_t sms = tc_sms();
    const int64_t rounds = (B + 128 * sms - 1) / (128 * sms);
    int64_t rpc = (B + rounds * sms - 1) / (rounds * sms);
    rpc = (rpc + 31) / 32 * 32;
    return (int)(rpc < 32 ? 32 : (rpc > 128 ? 128 : rpc));
}
static bool tc_fwd_one_row(int64_t B) {
    // 8 rows per rod (Newton, ~3 joint marches per step, 16 rods per CTA) for small batches; one row per rod (Broyden in
    // lock step, ~8 marches per step, 128 rods per CTA) once the 8-row kernel would need more than 3 rounds of CTAs.
    // Measured (H = 512, T = 30): 4096 rods 8.2 vs 12.0 ms, 8192 rods 16.0 vs 12.6 ms, 18 944 rods 31.5 vs 12.6 ms
    bool one = B > (int64_t)3 * tc_sms() * ktc::RPC;
    if (const char* e = getenv("KC_ROLLOUT_TC_ROWS")) { if (e[0] == '1') one = true; if (e[0] == '8') one = false; }
    return one;
}
size_t kc_knode_tc_fwd_scratch_bytes(int N, int64_t B) {
    const size_t states = (size_t)tc_bwd_grid(B) * 25 * N * 128 * sizeof(float);             // 8-row kernel: marched states
    const size_t hist = (size_t)148 * 2 * (N - 1) * 12 * 128 * sizeof(float);                // 1-row kernel: history, if not in smem
    return (states > hist ? states : hist) + 256;
}
int kc_knode_tc_fwd(const RodC<float>& P, const kc_mlp* mlp, int64_t B, int T_, const float* tensions, const float* y0,
                    const float* z0, float* trajD, float tol, int max_iter, float fd_eps, float* Gout, int32_t* iters,
                    unsigned char* img, unsigned char* scratch, cudaStream_t st) {
    int rc = tc_prep(mlp, img, st);
    if (rc) return rc;
    if (tc_fwd_one_row(B)) {
        const int rpc = tc_fwd1_rpc(B);
        const int64_t ngroups = (B + rpc - 1) / rpc;
        const int sms = tc_sms();
        const unsigned grid = (unsigned)(ngroups < sms ? ngroups : sms);
        const size_t hist_b = (size_t)(P.N - 1) * 12 * 128 * sizeof(float);
        const size_t base = (size_t)ktc::F_HIST + (size_t)KC_SHOOT_SLOTS * 128 * sizeof(float);
        const int hist_in_smem = base + hist_b <= 227 * 1024 ? 1 : 0;
        const size_t smem = base + (hist_in_smem ? hist_b : 0);
#define KC_GO1(D)                                                                                                      \
    do {                                                                                                               \
        cudaFuncSetAttribute(kc_knode_tc_fwd1_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);      \
        kc_knode_tc_fwd1_kernel<D><<<grid, ktc::THREADS, smem, st>>>(P, img, (const float*)mlp->b2, mlp->hidden, B, T_,\
            rpc, tensions, y0, z0, trajD, tol, max_iter, fd_eps, Gout, iters, reinterpret_cast<float*>(scratch),       \
            hist_in_smem);                                                                                             \
    } while (0)
        if (P.diag) KC_GO1(true); else KC_GO1(false);
#undef KC_GO1
        KC_CHECK_LAUNCH("kc_knode_tc_fwd1_kernel");
        return KC_OK;
    }
    const size_t smem = (size_t)ktc::F_HIST + (size_t)2 * (P.N - 1) * 12 * ktc::RPC * sizeof(float);
    const unsigned grid = (unsigned)tc_bwd_grid(B);
#define KC_GO(D)                                                                                                       \
    do {                                                                                                               \
        cudaFuncSetAttribute(kc_knode_tc_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        kc_knode_tc_fwd_kernel<D><<<grid, ktc::THREADS, smem, st>>>(P, img, (const float*)mlp->b2, mlp->hidden, B, T_, \
            tensions, y0, z0, trajD, tol, max_iter, fd_eps, Gout, iters, reinterpret_cast<float*>(scratch));           \
    } while (0)
    if (P.diag) KC_GO(true); else KC_GO(false);
#undef KC_GO
    KC_CHECK_LAUNCH("kc_knode_tc_fwd_kernel");
    return KC_OK;
}
int kc_knode_tc_bwd(const RodC<float>& P, const kc_mlp* mlp, int64_t B, int T_, const float* tensions, const float* traj,
                    const float* gtraj, float* gten, float* xs, float* gos, unsigned char* img, unsigned char* scratch,
                    cudaStream_t st) {
    int rc = tc_prep(mlp, img, st);
    if (rc) return rc;
    const int grid = tc_bwd_grid(B);
    float* rowscr = reinterpret_cast<float*>(scratch);
    float* hscr = rowscr + (size_t)grid * (P.N - 1) * ktc::SV * 128;
#define KC_GO(D)                                                                                                       \
    do {                                                                                                               \
        cudaFuncSetAttribute(kc_knode_tc_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, ktc::B_BYTES);    \
        kc_knode_tc_bwd_kernel<D><<<grid, ktc::THREADS, ktc::B_BYTES, st>>>(P, img, mlp->hidden, B, T_, tensions, traj,\
            gtraj, gten, xs, gos, rowscr, hscr);                                                                       \
    } while (0)
    if (P.diag) KC_GO(true); else KC_GO(false);
#undef KC_GO
    KC_CHECK_LAUNCH("kc_knode_tc_bwd_kernel");
    return KC_OK;
}
