// kc_train_tc2.cu — the teacher-forced KNODE training step on tcgen05 / TMEM, second generation: the same maths and operand
// formats as kc_train_tc.cu (physics_train.py:313-368; fp32 model, 28 inputs, hidden <= 512), re-organised as a
// warp-specialised pipeline so that the tensor pipe and the SIMT epilogues overlap (in kc_train_tc_kernel every stage was
// MMA -> wait -> epilogue -> barrier -> MMA: 45 % of its warp samples sat in mbarrier waits, profiles/r02_ncu_prof_train_tc.csv).
//
// One persistent CTA per SM, tiles of 128 samples, hidden units in SUB-CHUNKS of 64.  NG x 128 + 64 threads (NG = 2 is built):
//   warps 0..4NG-1  epilogue: thread = sample row (tid & 127), group = tid >> 7 owns 64/NG of a sub-chunk's 64 units
//   next warp  MMA issue, warp-uniform (all lanes run the loop, the instruction is predicated on elect.sync)
//   last warp  weight loader: TMA bulk copies of the next sub-chunks' operand images into two 4-slot rings (W1 | W2 images:
//              they are released at different times - W1 right after the Z GEMM, issued three sub-chunks ahead)
// TMEM (512 columns): gW1 accumulators 0..127 (32 per 128-unit chunk), gW2^T 128..255, working area 256..511:
//   forward : ring of three 64-column Z buffers (256, 320, 384) and O at 448..479
//   backward: two buffers of [Z 64 | dA 64] (256, 384)
// Forward per sub-chunk s:  Z = X W1_s^T (kind::f16, bf16 hi/lo in 3 passes = fp32-grade; column 28 of X carries b1)
//   -> epilogue: a = ELU(z) written back INTO the Z columns as packed bf16 hi | lo (tcgen05.st) -> O += A W2_s^T with the A
//   operand read from TMEM (".ts" MMA): the forward activations never touch shared memory.
// Loss: the rows of half 0 read O, form pred, the 4-term loss and dL/do; dO goes to shared memory (one tile, viewed K-major
//   by the dA GEMM and MN-major by the gW2 GEMM, as X is by GEMM1 and the gW1 GEMM).
// Backward per sub-chunk s: Z again and dA = dO W2_s (one commit) -> epilogue: tcgen05.ld, buffer released at once (the MMA
//   warp refills it with sub-chunk s+2 while the epilogue computes), a = ELU(z), dz = dA ELU'(z) -> bf16 hi/lo, MN-major
//   shared tiles of a 128-unit PAIR; after the second sub-chunk of a pair: gW1_c += dZ^T X, gW2_c^T += A^T dO (M = 128
//   units, K = 128 samples), accumulating in TMEM over all tiles of the CTA.
// At the end every CTA writes one partial-gradient slice in the layout kc_train_reduce_kernel sums.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kc_rod.cuh"
#include <cstdlib>
#include "kc_umma.cuh"

// Development aid (make EXTRA=-DKC_TC2_TRACE): clock64 stamps of CTA 0's epilogue warps 0 / 4 and MMA warp, tools/trace_train_tc2.py.
#ifdef KC_TC2_TRACE
__device__ long long* g_tc2_trace = nullptr;
extern "C" int kc_train_tc2_set_trace(long long* p) { return (int)cudaMemcpyToSymbol(g_tc2_trace, &p, sizeof(p)); }
#define TC2_TR(id) do { if (tr_role >= 0 && lane == 0 && g_tc2_trace && tr_n < 2048)                                         \
        g_tc2_trace[tr_role * 2048 + tr_n++] = ((long long)(id) << 48) | (clock64() & 0xffffffffffffll); } while (0)
#else
#define TC2_TR(id) do {} while (0)
#endif

namespace tc2 {
constexpr int STAGES = 4;
constexpr int OFF_X = 0;            // X^T hi | lo, bf16 MN-major [32 inputs x 128 samples]   2 x 8192
constexpr int OFF_DO = 16384;       // dO^T hi | lo, bf16 MN-major [32 outputs x 128 samples] 2 x 8192
constexpr int OFF_DZ = 32768;       // dZ^T hi | lo, bf16 MN-major [128 units x 128 samples]  2 x 32768
constexpr int OFF_A = 98304;        // A^T hi | lo                                             2 x 32768
constexpr int OFF_W1 = 163840;      // W1 ring: 4 slots x 8192 (hi | lo of a [64 units x 32] image)
constexpr int OFF_W2 = 196608;      // W2 ring: 4 slots x 8192 (forward image [32 outs x 64 units] or W2^T [64 units x 32 outs])
constexpr int OFF_MISC = 229376;
constexpr int SMEM_BYTES = OFF_MISC + 1024;
constexpr int SUB_IMG = 24576;      // per sub-chunk: W1 hi|lo (2 x 4096), W2^T hi|lo (2 x 4096), W2 forward image hi|lo (2 x 4096)
constexpr int COL_GW1 = 0, COL_GW2 = 128, COL_W = 256, COL_O = 448;

struct Bars {
    uint64_t w1full[STAGES], w1free[STAGES], w2full[STAGES], w2free[STAGES];
    uint64_t xrdy, ordy, dordy, gdone;
    uint64_t zf_rdy[3], zf_used[3];
    uint64_t zb_rdy[2], zb_used[2];
    uint64_t tile_rdy[4];   // indexed by sub-chunk & 3: the MMA warp may lag the epilogues by up to two sub-chunks (depth of the Z ring)
    uint32_t tmem_slot;
    double redd[8];
    float redb[4 * 25];
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// (x0, x1) -> packed bf16 hi pair (round to nearest) and packed bf16 lo pair (x - hi, truncated)
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(x0, x1);
    const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
    lo = __byte_perm(__float_as_uint(x0 - h0), __float_as_uint(x1 - h1), 0x7632);
}
__device__ __forceinline__ void split8(const float x[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_pair(x[2 * i], x[2 * i + 1], h[i], l[i]);
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint64_t dstep(uint64_t d, uint32_t bytes) { return d + (bytes >> 4); }
}  // namespace tc2

// W1[H][28], b1[H], W2[25][H] -> per sub-chunk of 64 units (zero rows beyond H): W1 hi|lo [64 units x 32] K-major (column 28 =
// b1), W2^T hi|lo [64 units x 32 outs] K-major (B of the dA GEMM), W2 hi|lo [32 outs x 64 units] K-major (B of the forward GEMM2)
__global__ void kc_tc2_prep_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                           const float* __restrict__ W2, int hidden, unsigned char* __restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 512 * 32; e += gridDim.x * blockDim.x) {
        const int u = e >> 5, k = e & 31, s = u >> 6, ul = u & 63;
        const float w1 = u < hidden ? (k < 28 ? W1[(size_t)u * 28 + k] : (k == 28 ? b1[u] : 0.f)) : 0.f;
        const float w2 = (u < hidden && k < 25) ? W2[(size_t)k * hidden + u] : 0.f;
        unsigned char* base = img + (size_t)s * tc2::SUB_IMG;
        auto put = [&](int image, uint32_t off, float w) {
            const __nv_bfloat16 h = __float2bfloat16_rn(w);
            const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
            *reinterpret_cast<__nv_bfloat16*>(base + image * 8192 + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(base + image * 8192 + 4096 + off) = l;
        };
        put(0, umma::kmajor_off_b16(ul, k, 32), w1);
        put(1, umma::kmajor_off_b16(ul, k, 32), w2);
        put(2, umma::kmajor_off_b16(k, ul, 64), w2);
    }
}

// NG = epilogue warp groups (each = 4 warps covering the 128 rows); a thread owns 64 / NG units of every sub-chunk.
template <int NG>
__global__ void __launch_bounds__(NG * 128 + 64, 1)
kc_train_tc2_kernel(int hidden, int nsub, const unsigned char* __restrict__ img, const float* __restrict__ b2, float ds, int64_t Q,
                    int T_, int K, const float* __restrict__ X, const float* __restrict__ PHYS, const float* __restrict__ TGT,
                    float* __restrict__ partial, int64_t NP, double* __restrict__ loss_part, float* __restrict__ pred_out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    using namespace tc2;
    Bars* bars = reinterpret_cast<Bars*>(sm + OFF_MISC);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    if (warp == 0) umma::tmem_alloc(&bars->tmem_slot, 512);
    if (tid == 32) {
        for (int i = 0; i < STAGES; ++i) {
            umma::mbar_init(&bars->w1full[i], 1); umma::mbar_init(&bars->w1free[i], 1);
            umma::mbar_init(&bars->w2full[i], 1); umma::mbar_init(&bars->w2free[i], 1);
        }
        constexpr int NE = NG * 128;   // epilogue threads
        umma::mbar_init(&bars->xrdy, NE); umma::mbar_init(&bars->ordy, 1); umma::mbar_init(&bars->dordy, NE);
        umma::mbar_init(&bars->gdone, 1);
        for (int i = 0; i < 4; ++i) umma::mbar_init(&bars->tile_rdy[i], NE);
        for (int i = 0; i < 3; ++i) { umma::mbar_init(&bars->zf_rdy[i], 1); umma::mbar_init(&bars->zf_used[i], NE); }
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->zb_rdy[i], 1); umma::mbar_init(&bars->zb_used[i], NE); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const int64_t ntiles = (Q + 127) / 128;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const int npair = nsub / 2;
#ifdef KC_TC2_TRACE
    const int tr_role = blockIdx.x != 0 ? -1 : (warp == 0 ? 0 : (warp == 4 ? 1 : (warp == NG * 4 ? 2 : -1)));
    int tr_n = 0;
#endif

    constexpr int CPT = 64 / NG;          // units (TMEM columns) of a sub-chunk per epilogue thread
    if (warp == NG * 4 + 1) {
        // ---- weight loader: unit u of a ring = sub-chunk-stage u (per tile: nsub forward stages, then nsub backward stages);
        // lane 0 feeds the W1 ring (the same image in both phases), lane 1 the W2 ring (forward image / W2^T) ----
        const int64_t total = my_tiles * 2 * nsub;
        if (lane < 2) {
            uint64_t* full = lane == 0 ? bars->w1full : bars->w2full;
            uint64_t* freeb = lane == 0 ? bars->w1free : bars->w2free;
            const int ring = lane == 0 ? OFF_W1 : OFF_W2;
            uint32_t phfree = 0;
            for (int64_t u = 0; u < total; ++u) {
                const int slot = (int)(u % STAGES), j = (int)(u % (2 * nsub)), sc = j % nsub;
                if (u >= STAGES) { umma::mbar_wait(&freeb[slot], (phfree >> slot) & 1u); phfree ^= 1u << slot; }
                const unsigned char* src = img + (size_t)sc * SUB_IMG + (lane == 0 ? 0 : (j < nsub ? 16384 : 8192));
                mbar_expect_tx(&full[slot], 8192);
                bulk_g2s(umma::smem_u32(sm + ring + slot * 8192), src, 8192, &full[slot]);
            }
        }
    } else if (warp == NG * 4) {
        // ---- MMA issue ----
        const uint32_t idZ = umma::make_idesc_bf16(128, 64), idO = umma::make_idesc_bf16(128, 32);
        const uint32_t idG = umma::make_idesc_bf16(128, 32, 1, 1);
        const uint64_t dXh = umma::make_desc(umma::smem_u32(sm + OFF_X), 2048, 128), dXl = dstep(dXh, 8192);        // K-major views
        const uint64_t dOh = umma::make_desc(umma::smem_u32(sm + OFF_DO), 2048, 128), dOl = dstep(dOh, 8192);
        const uint64_t mXh = umma::make_desc(umma::smem_u32(sm + OFF_X), 128, 2048), mXl = dstep(mXh, 8192);        // MN-major views
        const uint64_t mOh = umma::make_desc(umma::smem_u32(sm + OFF_DO), 128, 2048), mOl = dstep(mOh, 8192);
        const uint64_t mZh = umma::make_desc(umma::smem_u32(sm + OFF_DZ), 128, 2048), mZl = dstep(mZh, 32768);
        const uint64_t mAh = umma::make_desc(umma::smem_u32(sm + OFF_A), 128, 2048), mAl = dstep(mAh, 32768);
        uint32_t phf1 = 0, phf2 = 0, phx = 0, phdo = 0, phzfu = 0, phzbu = 0, phtile = 0;
        int64_t q = 0;   // running sub-chunk-stage = unit index of both weight rings
        auto wait_w1 = [&](int slot) { umma::mbar_wait(&bars->w1full[slot], (phf1 >> slot) & 1u); phf1 ^= 1u << slot; umma::fence_after(); };
        auto wait_w2 = [&](int slot) { umma::mbar_wait(&bars->w2full[slot], (phf2 >> slot) & 1u); phf2 ^= 1u << slot; umma::fence_after(); };
        // D[128 x 64] = A[128 samples x 32] (tile viewed K-major) * Bimg[64 units x 32]^T, 3 passes
        auto gemm_k32 = [&](uint32_t d, uint64_t ah, uint64_t al, uint32_t wsm) {
            const uint64_t bh = umma::make_desc(wsm, 128, 512), bl = dstep(bh, 4096);
#pragma unroll
            for (int p = 0; p < 3; ++p) {
                const uint64_t a = p == 1 ? al : ah, b = p == 2 ? bl : bh;
#pragma unroll
                for (int kk = 0; kk < 2; ++kk) umma::mma_bf16_w(d, dstep(a, kk * 4096), dstep(b, kk * 256), idZ, (p | kk) ? 1u : 0u);
            }
        };
        for (int64_t t = 0; t < my_tiles; ++t) {
            const bool first_tile = t == 0;
            umma::mbar_wait(&bars->xrdy, phx); phx ^= 1;
            umma::fence_after();
            TC2_TR(100);
            // ---------------- forward ----------------
            auto fwd_gemm1 = [&](int s) {
                const int slot = (int)((q + s) % STAGES);
                wait_w1(slot);
                gemm_k32(tbase + COL_W + (s % 3) * 64, dXh, dXl, umma::smem_u32(sm + OFF_W1 + slot * 8192));
                umma::commit_w(&bars->zf_rdy[s % 3]);
                umma::commit_w(&bars->w1free[slot]);
            };
            for (int s = 0; s < nsub && s < 3; ++s) fwd_gemm1(s);
            for (int s = 0; s < nsub; ++s) {
                const int b = s % 3, slot = (int)((q + s) % STAGES);
                umma::mbar_wait(&bars->zf_used[b], (phzfu >> b) & 1u); phzfu ^= 1u << b;
                umma::fence_after();
                TC2_TR(120 + s);
                wait_w2(slot);
                const uint64_t wh = umma::make_desc(umma::smem_u32(sm + OFF_W2 + slot * 8192), 128, 1024), wl = dstep(wh, 4096);
                const uint32_t ab = tbase + COL_W + b * 64;
#pragma unroll
                for (int p = 0; p < 3; ++p) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        // 16-unit group kk lives in the column block of the thread group that owns it: hi | lo halves
                        const uint32_t a = ab + (kk / (CPT / 16)) * CPT + (kk % (CPT / 16)) * 8 + (p == 1 ? CPT / 2 : 0);
                        umma::mma_bf16_ts_w(tbase + COL_O, a, dstep(p == 2 ? wl : wh, kk * 256), idO, (s | p | kk) ? 1u : 0u);
                    }
                }
                umma::commit_w(&bars->w2free[slot]);
                if (s + 3 < nsub) fwd_gemm1(s + 3);
                TC2_TR(130 + s);
            }
            umma::commit_w(&bars->ordy);
            q += nsub;
            // ---------------- backward ----------------
            umma::mbar_wait(&bars->dordy, phdo); phdo ^= 1;
            umma::fence_after();
            TC2_TR(140);
            auto bwd_gemm13 = [&](int s) {
                const int slot = (int)((q + s) % STAGES);
                wait_w1(slot);
                wait_w2(slot);
                const uint32_t d = tbase + COL_W + (s & 1) * 128;
                gemm_k32(d, dXh, dXl, umma::smem_u32(sm + OFF_W1 + slot * 8192));          // Z  = X W1_s^T
                gemm_k32(d + 64, dOh, dOl, umma::smem_u32(sm + OFF_W2 + slot * 8192));     // dA = dO W2_s
                umma::commit_w(&bars->zb_rdy[s & 1]);
                umma::commit_w(&bars->w1free[slot]);
                umma::commit_w(&bars->w2free[slot]);
            };
            for (int s = 0; s < nsub && s < 2; ++s) bwd_gemm13(s);
            for (int s = 0; s < nsub; ++s) {
                const int b = s & 1;
                umma::mbar_wait(&bars->zb_used[b], (phzbu >> b) & 1u); phzbu ^= 1u << b;   // both Z and dA are in registers
                umma::fence_after();
                TC2_TR(150 + s);
                if (s + 2 < nsub) bwd_gemm13(s + 2);
                TC2_TR(160 + s);
                umma::mbar_wait(&bars->tile_rdy[s & 3], (phtile >> (s & 3)) & 1u); phtile ^= 1u << (s & 3);   // a, dz of sub-chunk s are in shared memory
                umma::fence_after();
                TC2_TR(170 + s);
                if (s & 1) {
                    const int c = s >> 1;
                    const uint32_t d1 = tbase + COL_GW1 + 32 * c, d2 = tbase + COL_GW2 + 32 * c;
                    const uint32_t acc0 = first_tile ? 0u : 1u;
#pragma unroll
                    for (int p = 0; p < 3; ++p) {      // gW1_c += dZ^T X  (hi*hi, lo*hi, hi*lo)
                        const uint64_t a = p == 1 ? mZl : mZh, bb = p == 2 ? mXl : mXh;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) umma::mma_bf16_w(d1, dstep(a, kk * 256), dstep(bb, kk * 256), idG, (p | kk) ? 1u : acc0);
                    }
#pragma unroll
                    for (int p = 0; p < 3; ++p) {      // gW2_c^T += A^T dO
                        const uint64_t a = p == 1 ? mAl : mAh, bb = p == 2 ? mOl : mOh;
#pragma unroll
                        for (int kk = 0; kk < 8; ++kk) umma::mma_bf16_w(d2, dstep(a, kk * 256), dstep(bb, kk * 256), idG, (p | kk) ? 1u : acc0);
                    }
                    umma::commit_w(&bars->gdone);
                    TC2_TR(180 + s);
                }
            }
            q += nsub;
        }
    } else {
        // ---- epilogue warps ----
        const int row = tid & 127, grp = tid >> 7;     // grp: which CPT-wide column block of every sub-chunk this thread owns
        const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t phzf = 0, phzb = 0, pho = 0, phg = 0;
        float gb2acc[25];
#pragma unroll
        for (int c = 0; c < 25; ++c) gb2acc[c] = 0.f;
        double lossacc = 0.0;
        int64_t pairs_done = 0;      // gradient-GEMM completions consumed so far
        constexpr int IPT = 32 / NG;            // inputs per thread (16 or 8)
        float xv[IPT];
        // this thread's slice of a sample's inputs (columns 28..31 of X are never read from memory; column 28 := 1)
        auto load_x = [&](int64_t tile_) {
            const int64_t qr = tile_ * 128 + row;
            const bool ok = qr < Q;
            const float4* xs = reinterpret_cast<const float4*>(X + (size_t)(ok ? qr : 0) * 32 + grp * IPT);
#pragma unroll
            for (int g = 0; g < IPT / 4; ++g) {
                const int k0 = grp * IPT + 4 * g;
                const float4 v4 = (ok && k0 < 28) ? xs[g] : make_float4(0.f, 0.f, 0.f, 0.f);
                xv[4 * g] = v4.x; xv[4 * g + 1] = v4.y; xv[4 * g + 2] = v4.z; xv[4 * g + 3] = v4.w;
            }
            if (grp == NG - 1) { xv[IPT - 4] = ok ? 1.f : 0.f; xv[IPT - 3] = 0.f; xv[IPT - 2] = 0.f; xv[IPT - 1] = 0.f; }
        };
        if (my_tiles > 0) load_x(blockIdx.x);
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int64_t tile = blockIdx.x + t * gridDim.x;
            const int64_t qrow = tile * 128 + row;
            const bool valid = qrow < Q;
            TC2_TR(0);
            // ---- X tile: bf16 hi/lo, MN-major [32 inputs x 128 samples]; column 28 = 1 carries b1 / yields gb1 (xv was loaded
            // before the previous tile's last gradient wait) ----
#pragma unroll
            for (int g = 0; g < IPT / 8; ++g) {
                uint4 hi, lo;
                split8(xv + 8 * g, hi, lo);
                const uint32_t o = umma::mnmajor_off_b16(grp * IPT + 8 * g, row, 128);
                *reinterpret_cast<uint4*>(sm + OFF_X + o) = hi;
                *reinterpret_cast<uint4*>(sm + OFF_X + 8192 + o) = lo;
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&bars->xrdy);
            TC2_TR(1);
            // ---- forward epilogues ----
            for (int s = 0; s < nsub; ++s) {
                const int b = s % 3;
                umma::mbar_wait(&bars->zf_rdy[b], (phzf >> b) & 1u); phzf ^= 1u << b;
                umma::fence_after();
                TC2_TR(10 + s);
                const uint32_t ta = tbase + laneblk + COL_W + b * 64 + grp * CPT;
                uint32_t z[CPT], hi[CPT / 2], lo[CPT / 2];
                if (CPT == 32) umma::ld32(ta, z); else umma::ld16(ta, z);
                umma::wait_ld();
#pragma unroll
                for (int i = 0; i < CPT / 2; ++i) split_pair(kc_elu(__uint_as_float(z[2 * i])), kc_elu(__uint_as_float(z[2 * i + 1])), hi[i], lo[i]);
                if (CPT == 32) { umma::st16(ta, hi); umma::st16(ta + 16, lo); }
                else { umma::st8(ta, hi); umma::st8(ta + 8, lo); }
                umma::wait_st();
                umma::fence_before();
                umma::mbar_arrive(&bars->zf_used[b]);
                TC2_TR(20 + s);
            }
            // ---- loss and dL/do (the threads of group 0 own the sample) ----
            float ph[25], tg[25];    // physics prediction and target of this sample: in flight while the last GEMM2 drains
            if (grp == 0 && valid) {
#pragma unroll
                for (int r = 0; r < 25; ++r) { ph[r] = PHYS[(size_t)qrow * 25 + r]; tg[r] = TGT[(size_t)qrow * 25 + r]; }
            }
            umma::mbar_wait(&bars->ordy, pho); pho ^= 1;
            umma::fence_after();
            TC2_TR(30);
            if (grp == 0) {
                float o[25], g[25];
                {
                    uint32_t v[32];
                    umma::ld32(tbase + laneblk + COL_O, v);
                    umma::wait_ld();
#pragma unroll
                    for (int c = 0; c < 25; ++c) { o[c] = __uint_as_float(v[c]) + b2[c]; g[c] = 0.f; }
                }
                if (valid) {
                    float pred[25];
#pragma unroll
                    for (int r = 0; r < 19; ++r) pred[r] = ph[r] + ds * o[r];
#pragma unroll
                    for (int c = 19; c < 25; ++c) pred[c] = ph[c] + o[c];
                    const float S = float(T_ - 1);
                    const float wp = 1.f / (float(3 * K) * S), wf = 1.f / (float(12 * K) * S), wz = 1.f / (float(6 * K) * S);
                    float acc = 0.f;
#pragma unroll
                    for (int r = 0; r < 3; ++r) { const float e = pred[r] - tg[r]; acc += wp * e * e; g[r] = 2.f * wp * e * ds; }
#pragma unroll
                    for (int r = 7; r < 19; ++r) { const float e = pred[r] - tg[r]; acc += wf * e * e; g[r] = 2.f * wf * e * ds; }
#pragma unroll
                    for (int r = 19; r < 25; ++r) { const float e = pred[r] - tg[r]; acc += wz * e * e; g[r] = 2.f * wz * e; }
                    float ep[3], et[3], ge[3], gq[4];
                    quat_to_euler(pred + 3, ep);
                    quat_to_euler(tg + 3, et);
#pragma unroll
                    for (int i = 0; i < 3; ++i) { const float e = ep[i] - et[i]; acc += wp * e * e; ge[i] = 2.f * wp * e; }
                    quat_to_euler_vjp(pred + 3, ge, gq);
#pragma unroll
                    for (int i = 0; i < 4; ++i) g[3 + i] = gq[i] * ds;
                    lossacc += (double)acc;
                    if (pred_out) {
                        const int kk = (int)(qrow % K);
                        const int64_t bt = qrow / K;
                        float* po = pred_out + (size_t)bt * 25 * K + kk;
#pragma unroll
                        for (int r = 0; r < 25; ++r) po[r * K] = pred[r];
                    }
                }
#pragma unroll
                for (int c = 0; c < 25; ++c) gb2acc[c] += g[c];
#pragma unroll
                for (int gi = 0; gi < 4; ++gi) {
                    float g8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const int c = gi * 8 + j; g8[j] = c < 25 ? g[c < 25 ? c : 0] : 0.f; }
                    uint4 hi, lo;
                    split8(g8, hi, lo);
                    const uint32_t off = umma::mnmajor_off_b16(gi * 8, row, 128);
                    *reinterpret_cast<uint4*>(sm + OFF_DO + off) = hi;
                    *reinterpret_cast<uint4*>(sm + OFF_DO + 8192 + off) = lo;
                }
            }
            umma::fence_async_smem();
            umma::fence_before();
            umma::mbar_arrive(&bars->dordy);
            TC2_TR(31);
            // ---- backward epilogues ----
            for (int s = 0; s < nsub; ++s) {
                const int b = s & 1;
                umma::mbar_wait(&bars->zb_rdy[b], (phzb >> b) & 1u); phzb ^= 1u << b;
                umma::fence_after();
                TC2_TR(40 + s);
                const uint32_t ta = tbase + laneblk + COL_W + b * 128 + grp * CPT;
                uint32_t z[CPT], d[CPT];
                if (CPT == 32) { umma::ld32(ta, z); umma::ld32(ta + 64, d); }
                else { umma::ld16(ta, z); umma::ld16(ta + 64, d); }
                umma::wait_ld();
                umma::fence_before();
                umma::mbar_arrive(&bars->zb_used[b]);                  // the buffer may be refilled while this thread computes
                TC2_TR(50 + s);
                // a = ELU(z), dz = dA ELU'(z) as bf16 hi/lo, held in registers ...
                uint4 ah[CPT / 8], al[CPT / 8], dh[CPT / 8], dl[CPT / 8];
#pragma unroll
                for (int g8 = 0; g8 < CPT / 8; ++g8) {
                    float a8[8], d8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float zz = __uint_as_float(z[g8 * 8 + j]);
                        const float a = kc_elu(zz);
                        a8[j] = a;
                        d8[j] = __uint_as_float(d[g8 * 8 + j]) * (zz > 0.f ? 1.f : a + 1.f);   // ELU'(z) = e^z = ELU(z) + 1
                    }
                    split8(a8, ah[g8], al[g8]);
                    split8(d8, dh[g8], dl[g8]);
                }
                TC2_TR(60 + s);
                // ... until the shared tiles of this pair are free: the gradient MMAs of the previous pair are done (they were
                // issued when the previous sub-chunk was stored; the arithmetic above ran while they executed)
                if ((s & 1) == 0 && pairs_done < t * npair + (s >> 1)) {
                    umma::mbar_wait(&bars->gdone, phg); phg ^= 1;
                    ++pairs_done;
                }
                TC2_TR(70 + s);
                const int u0 = (s & 1) * 64 + grp * CPT;
#pragma unroll
                for (int g8 = 0; g8 < CPT / 8; ++g8) {
                    const uint32_t off = umma::mnmajor_off_b16(u0 + g8 * 8, row, 128);
                    *reinterpret_cast<uint4*>(sm + OFF_A + off) = ah[g8];
                    *reinterpret_cast<uint4*>(sm + OFF_A + 32768 + off) = al[g8];
                    *reinterpret_cast<uint4*>(sm + OFF_DZ + off) = dh[g8];
                    *reinterpret_cast<uint4*>(sm + OFF_DZ + 32768 + off) = dl[g8];
                }
                umma::fence_async_smem();
                umma::mbar_arrive(&bars->tile_rdy[s & 3]);
                TC2_TR(80 + s);
            }
            // X / dO of the next tile overwrite what the last pair's gradient MMAs read: consume its completion now (the next
            // tile's inputs are fetched into registers first: their latency hides behind this wait)
            if (t + 1 < my_tiles) load_x(tile + gridDim.x);
            while (pairs_done < (t + 1) * npair) {
                umma::mbar_wait(&bars->gdone, phg); phg ^= 1;
                ++pairs_done;
            }
            umma::fence_after();
            TC2_TR(90);
        }
        // ---- this CTA's partial gradients ----
        float* out = partial + (size_t)blockIdx.x * NP;
        const int64_t ob1 = (int64_t)hidden * 28, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
        if (my_tiles == 0) {   // (grid <= ntiles, so this does not happen; keep the slice defined anyway)
            for (int64_t i = tid; i < NP; i += NG * 128) out[i] = 0.f;
        } else if (grp == 0) {
            for (int c = 0; c < npair; ++c) {
                const int u = c * 128 + row;
                uint32_t v[32];
                umma::ld32(tbase + laneblk + COL_GW1 + 32 * c, v);
                umma::wait_ld();
                if (u < hidden) {
#pragma unroll
                    for (int k = 0; k < 28; ++k) out[(size_t)u * 28 + k] = __uint_as_float(v[k]);
                    out[ob1 + u] = __uint_as_float(v[28]);
                }
                umma::ld32(tbase + laneblk + COL_GW2 + 32 * c, v);
                umma::wait_ld();
                if (u < hidden) {
#pragma unroll
                    for (int co = 0; co < 25; ++co) out[oW2 + (size_t)co * hidden + u] = __uint_as_float(v[co]);
                }
            }
        }
#pragma unroll
        for (int c = 0; c < 25; ++c) {
            float s = gb2acc[c];
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_down_sync(0xffffffffu, s, o2);
            if (lane == 0 && warp < 4) bars->redb[warp * 25 + c] = s;
        }
        {
            double s = lossacc;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_down_sync(0xffffffffu, s, o2);
            if (lane == 0) bars->redd[warp] = s;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(NG * 128) : "memory");
        if (tid < 25) out[ob2 + tid] = bars->redb[tid] + bars->redb[25 + tid] + bars->redb[50 + tid] + bars->redb[75 + tid];
        if (tid == 0) loss_part[blockIdx.x] = bars->redd[0] + bars->redd[1] + bars->redd[2] + bars->redd[3];
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Host side: same contract as kc_train_tc_launch (kc_train_tc.cu); `img`: 8 x 24 KB of workspace.
int kc_train_tc2_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS, const float* TGT,
                        unsigned char* img, float* partial, int64_t NP, double* loss_part, float* pred_out, int grid,
                        cudaStream_t st) {
    const int nsub = 2 * ((mlp->hidden + 127) / 128);
    kc_tc2_prep_weights_kernel<<<32, 256, 0, st>>>((const float*)mlp->W1, (const float*)mlp->b1, (const float*)mlp->W2,
                                                   mlp->hidden, img);
    KC_CHECK_LAUNCH("kc_tc2_prep_weights_kernel");
    // two epilogue warp groups (8 epilogue warps, 32 units per thread).  The kernel is written over NG, but only NG = 2 is
    // instantiated: a 16-warp variant (96 registers, spills) measured slower (0.254 vs 0.240 ms per kc_train_step at 1024x29).
    cudaFuncSetAttribute(kc_train_tc2_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES);
    kc_train_tc2_kernel<2><<<grid, 2 * 128 + 64, tc2::SMEM_BYTES, st>>>(mlp->hidden, nsub, img, (const float*)mlp->b2, ds, Q, T_, K,
                                                                      X, PHYS, TGT, partial, NP, loss_part, pred_out);
    KC_CHECK_LAUNCH("kc_train_tc2_kernel");
    return KC_OK;
}
