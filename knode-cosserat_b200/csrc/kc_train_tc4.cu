// kc_train_tc4.cu — the teacher-forced KNODE training step on tcgen05 / TMEM, fourth generation: kc_train_tc3.cu (transposed
// backward, activations in TMEM, all operand images resident in shared memory; see there for the maths and the operand
// formats) with the epilogue warps split into TWO TEAMS that work half a step out of phase.
//
// In kc_train_tc4_kernel all eight epilogue warps process the same 64-column step: they wait on the same mbarrier, on the same
// TMEM loads and on the same TMEM stores at the same time, and its timeline (tools/trace_train_tc2.py) shows a 40k-cycle tile
// against 17.6k cycles of epilogue issue.  Here a step is 32 columns wide and belongs to ONE team (team = step & 1; warps 0..3 /
// 4..7, thread = TMEM lane, all 32 columns of the step in its registers), with four TMEM buffers per phase (two per team:
// one being processed, one already filled): while one team sits in a barrier / TMEM round trip the other team's warp on the
// same scheduler issues.
//
// One persistent CTA per SM, tiles of 128 samples, 288 threads (8 epilogue warps + MMA-issue warp).
// TMEM (512 columns): gW1 accumulators 0..127 (32 per 128-unit chunk), gW2^T 128..255, working area 256..511:
//   forward : four 32-column Z buffers (256 + 32 k), O at 480..511;   step i = 32 hidden units
//   backward: four buffers of [Z^T 32 | dA^T 32] (256 + 64 k);         step j = (chunk j >> 2, 32-sample quarter j & 3)
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kc_rod.cuh"
#include <cstdlib>
#include "kc_umma.cuh"

// Development aid (make EXTRA=-DKC_TC4_TRACE): clock64 stamps of CTA 0's epilogue warps 0 / 4 and MMA warp, tools/trace_train_tc2.py (KC_TRACE_GEN=4).
#ifdef KC_TC4_TRACE
__device__ long long* g_tc4_trace = nullptr;
extern "C" int kc_train_tc4_set_trace(long long* p) { return (int)cudaMemcpyToSymbol(g_tc4_trace, &p, sizeof(p)); }
#define TC4_TR(id) do { if (tr_role >= 0 && lane == 0 && g_tc4_trace && tr_n < 2048)                                         \
        g_tc4_trace[tr_role * 2048 + tr_n++] = ((long long)(id) << 48) | (clock64() & 0xffffffffffffll); } while (0)
#else
#define TC4_TR(id) do {} while (0)
#endif

namespace tc4 {
constexpr int OFF_X = 0;            // X^T hi | lo, bf16 [32 inputs x 128 samples], 8 inputs contiguous        2 x 8192
constexpr int OFF_DO = 16384;       // dO^T hi | lo, same layout [32 outputs x 128 samples]                    2 x 8192
constexpr int OFF_W1 = 32768;       // per 128-unit chunk: W1 hi | lo, K-major [128 units x 32] (col 28 = b1)  4 x 16384
constexpr int OFF_W2T = 98304;      // per chunk: W2^T hi | lo, K-major [128 units x 32 outs]                  4 x 16384
constexpr int OFF_W2F = 163840;     // per 64-unit sub-chunk: W2 hi | lo, K-major [32 outs x 64 units]         8 x 8192
constexpr int OFF_MISC = 229376;
constexpr int SMEM_BYTES = OFF_MISC + 1024;
constexpr int COL_GW1 = 0, COL_GW2 = 128, COL_W = 256, COL_O = 480;

struct Bars {
    uint64_t wfull, xrdy, ordy, dordy, gdone;
    uint64_t zf_rdy[4], zf_used[4];
    uint64_t zb_rdy[4], zb_done[4];
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
}  // namespace tc4

// W1[H][28], b1[H], W2[25][H] -> the three operand images in shared-memory order (zero rows beyond H)
__global__ void kc_tc4_prep_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                           const float* __restrict__ W2, int hidden, unsigned char* __restrict__ img) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 512 * 32; e += gridDim.x * blockDim.x) {
        const int u = e >> 5, k = e & 31;
        const float w1 = u < hidden ? (k < 28 ? W1[(size_t)u * 28 + k] : (k == 28 ? b1[u] : 0.f)) : 0.f;
        const float w2 = (u < hidden && k < 25) ? W2[(size_t)k * hidden + u] : 0.f;
        auto put = [&](uint32_t off, uint32_t lo_off, float w) {
            const __nv_bfloat16 h = __float2bfloat16_rn(w);
            const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
            *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
            *reinterpret_cast<__nv_bfloat16*>(img + off + lo_off) = l;
        };
        put((u >> 7) * 16384 + umma::kmajor_off_b16(u & 127, k, 32), 8192, w1);
        put(65536 + (u >> 7) * 16384 + umma::kmajor_off_b16(u & 127, k, 32), 8192, w2);
        put(131072 + (u >> 6) * 8192 + umma::kmajor_off_b16(k, u & 63, 64), 4096, w2);
    }
}

__global__ void __launch_bounds__(288, 1)
kc_train_tc4_kernel(int hidden, int nsub, const unsigned char* __restrict__ img, const float* __restrict__ b2, float ds, int64_t Q,
                    int T_, int K, const float* __restrict__ X, const float* __restrict__ PHYS, const float* __restrict__ TGT,
                    float* __restrict__ partial, int64_t NP, double* __restrict__ loss_part, float* __restrict__ pred_out) {
    extern __shared__ __align__(1024) unsigned char sm[];
    using namespace tc4;
    Bars* bars = reinterpret_cast<Bars*>(sm + OFF_MISC);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int nchunk = nsub / 2;
    if (warp == 0) umma::tmem_alloc(&bars->tmem_slot, 512);
    if (tid == 32) {
        umma::mbar_init(&bars->wfull, 1);
        umma::mbar_init(&bars->xrdy, 256); umma::mbar_init(&bars->ordy, 1); umma::mbar_init(&bars->dordy, 128);
        umma::mbar_init(&bars->gdone, 1);
        for (int i = 0; i < 4; ++i) {     // a buffer belongs to one team: 128 arrivals
            umma::mbar_init(&bars->zf_rdy[i], 1); umma::mbar_init(&bars->zf_used[i], 128);
            umma::mbar_init(&bars->zb_rdy[i], 1); umma::mbar_init(&bars->zb_done[i], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const int64_t ntiles = (Q + 127) / 128;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
#ifdef KC_TC4_TRACE
    const int tr_role = blockIdx.x != 0 ? -1 : (warp == 0 ? 0 : (warp == 4 ? 1 : (warp == 8 ? 2 : -1)));
    int tr_n = 0;
#endif

    if (warp == 8) {
        // ---- all operand images -> shared memory, once (three regions of nchunk x 16 KB) ----
        if (lane == 0 && my_tiles > 0) {
            mbar_expect_tx(&bars->wfull, (uint32_t)nchunk * 49152u);
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < nchunk; ++c)
                    bulk_g2s(umma::smem_u32(sm + OFF_W1 + r * 65536 + c * 16384), img + (size_t)r * 65536 + (size_t)c * 16384, 16384,
                             &bars->wfull);
        }
        __syncwarp();
        // ---- MMA issue ----
        const uint32_t idZ = umma::make_idesc_bf16(128, 32), idO = umma::make_idesc_bf16(128, 32);
        const uint32_t idG = umma::make_idesc_bf16(128, 32, 0, 1);      // A from TMEM, B = X / dO tile viewed MN-major
        const uint32_t sX = umma::smem_u32(sm + OFF_X), sO = umma::smem_u32(sm + OFF_DO);
        const uint64_t dXh = umma::make_desc(sX, 2048, 128), dXl = dstep(dXh, 8192);        // [samples x 32] K-major views
        const uint64_t dOh = umma::make_desc(sO, 2048, 128), dOl = dstep(dOh, 8192);
        const uint64_t mXh = umma::make_desc(sX, 128, 2048), mXl = dstep(mXh, 8192);        // [32 x samples] MN-major views
        const uint64_t mOh = umma::make_desc(sO, 128, 2048), mOl = dstep(mOh, 8192);
        const int nst = 2 * nsub;          // steps per phase (32 units forward / 32 samples of a 128-unit chunk backward)
        uint32_t phx = 0, phdo = 0, phzfu = 0, phzbd = 0;
        // D[128 x 32] = A[128 x 32] * B[32 x 32]^T, both K-major, 3 passes (hi hi, lo hi, hi lo) x 2 k-steps: one issue batch
        auto gemm_k32 = [&](uint32_t d, uint64_t ah, uint64_t al, uint32_t astep, uint64_t bh, uint64_t bl, uint32_t bstep) {
            umma::mma_bf16_ss_3x2_w(d, ah, al, astep >> 4, bh, bl, bstep >> 4, idZ);
        };
        if (my_tiles > 0) { umma::mbar_wait(&bars->wfull, 0); umma::fence_after(); }
        for (int64_t t = 0; t < my_tiles; ++t) {
            const bool first_tile = t == 0;
            umma::mbar_wait(&bars->xrdy, phx); phx ^= 1;
            umma::fence_after();
            TC4_TR(100);
            // ---------------- forward: Z_i = X W1_i^T (A = X tile, B = 32 rows of the W1 image), buffer i & 3 ----------------
            auto fwd_gemm1 = [&](int i) {
                const uint32_t w = umma::smem_u32(sm + OFF_W1 + (i >> 2) * 16384 + (i & 3) * 2048);
                const uint64_t bh = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (i & 3) * 32, dXh, dXl, 4096, bh, dstep(bh, 8192), 256);
                umma::commit_w(&bars->zf_rdy[i & 3]);
            };
            for (int i = 0; i < nst && i < 4; ++i) fwd_gemm1(i);
            for (int i = 0; i < nst; ++i) {
                const int b = i & 3;
                umma::mbar_wait(&bars->zf_used[b], (phzfu >> b) & 1u); phzfu ^= 1u << b;
                umma::fence_after();
                TC4_TR(110 + i);
                // O += A_i W2_i^T: A = the step's 32 columns as [hi 16 | lo 16] packed pairs, B = 32 units of the W2 image
                const uint64_t wh = umma::make_desc(umma::smem_u32(sm + OFF_W2F + (i >> 1) * 8192 + (i & 1) * 512), 128, 1024);
                umma::mma_bf16_ts_3x2_w<8, 16>(tbase + COL_O, tbase + COL_W + b * 32, wh, dstep(wh, 4096), idO, i ? 1u : 0u);
                if (i + 4 < nst) fwd_gemm1(i + 4);
            }
            umma::commit_w(&bars->ordy);
            // ---------------- backward ----------------
            // step j: chunk c = j >> 2, samples 32 r .. 32 r + 31 (r = j & 3), buffer j & 3 = [Z^T 32 | dA^T 32]
            auto bwd_z = [&](int j) {       // Z^T = W1_c X_r^T
                const uint32_t w = umma::smem_u32(sm + OFF_W1 + (j >> 2) * 16384);
                const uint64_t ah = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (j & 3) * 64, ah, dstep(ah, 8192), 256, dstep(dXh, (j & 3) * 512), dstep(dXl, (j & 3) * 512), 4096);
            };
            auto bwd_da = [&](int j) {      // dA^T = W2_c^T dO_r^T
                const uint32_t w = umma::smem_u32(sm + OFF_W2T + (j >> 2) * 16384);
                const uint64_t ah = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (j & 3) * 64 + 32, ah, dstep(ah, 8192), 256, dstep(dOh, (j & 3) * 512), dstep(dOl, (j & 3) * 512), 4096);
                umma::commit_w(&bars->zb_rdy[j & 3]);
            };
            // the Z halves of the first four steps do not need dO: they run while the loss is being formed (their columns
            // 256.., 320.., 384.., 448..479 are free: every forward MMA that read them was issued before; O sits at 480..511,
            // which only dA of step 3 overwrites - after dO exists, i.e. after the loss has read O)
            for (int j = 0; j < nst && j < 4; ++j) bwd_z(j);
            umma::mbar_wait(&bars->dordy, phdo); phdo ^= 1;
            umma::fence_after();
            TC4_TR(140);
            for (int j = 0; j < nst && j < 4; ++j) bwd_da(j);
            for (int j = 0; j < nst; ++j) {
                const int b = j & 3, c = j >> 2, r = j & 3;
                umma::mbar_wait(&bars->zb_done[b], (phzbd >> b) & 1u); phzbd ^= 1u << b;   // a, dz are in TMEM as bf16 hi | lo
                umma::fence_after();
                TC4_TR(150 + j);
                const uint32_t d1 = tbase + COL_GW1 + 32 * c, d2 = tbase + COL_GW2 + 32 * c;
                const uint32_t acc0 = (first_tile && r == 0) ? 0u : 1u;
                const uint32_t za = tbase + COL_W + b * 64, da = za + 32;
                // gW1_c += dZ^T X, gW2_c^T += A^T dO (hi*hi, lo*hi, hi*lo) over this step's 32 samples; packed layout [hi 8 | lo 8]
                umma::mma_bf16_ts_3x2_w<16, 8>(d1, da, dstep(mXh, r * 512), dstep(mXl, r * 512), idG, acc0);
                umma::mma_bf16_ts_3x2_w<16, 8>(d2, za, dstep(mOh, r * 512), dstep(mOl, r * 512), idG, acc0);
                if (j + 4 < nst) { bwd_z(j + 4); bwd_da(j + 4); }   // same buffer: executes after the MMAs above (issue order)
                TC4_TR(170 + j);
            }
            umma::commit_w(&bars->gdone);    // every MMA that read the X / dO tiles of this tile is done
        }
    } else {
        // ---- epilogue warps ----
        const int row = tid & 127, grp = tid >> 7;     // grp: the team (steps grp, grp + 2, ... of either phase)
        const int nst = 2 * nsub;
        const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t phzf = 0, phzb = 0, pho = 0, phg = 0;
        double lossacc = 0.0;
        if (tid < 100) bars->redb[tid] = 0.f;     // gb2 partial sums, one row of 25 per warp of group 0 (only that warp touches it)
        float xv[16];
        // this thread's 16 inputs of a sample (columns 28..31 of X are never read from memory; column 28 := 1)
        auto load_x = [&](int64_t tile_) {
            const int64_t qr = tile_ * 128 + row;
            const bool ok = qr < Q;
            const float4* xs = reinterpret_cast<const float4*>(X + (size_t)(ok ? qr : 0) * 32 + grp * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const int k0 = grp * 16 + 4 * g;
                const float4 v4 = (ok && k0 < 28) ? xs[g] : make_float4(0.f, 0.f, 0.f, 0.f);
                xv[4 * g] = v4.x; xv[4 * g + 1] = v4.y; xv[4 * g + 2] = v4.z; xv[4 * g + 3] = v4.w;
            }
            if (grp == 1) { xv[12] = ok ? 1.f : 0.f; xv[13] = 0.f; xv[14] = 0.f; xv[15] = 0.f; }
        };
        if (my_tiles > 0) load_x(blockIdx.x);
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int64_t tile = blockIdx.x + t * gridDim.x;
            const int64_t qrow = tile * 128 + row;
            const bool valid = qrow < Q;
            TC4_TR(0);
            // ---- X tile: bf16 hi/lo [32 inputs x 128 samples]; column 28 = 1 carries b1 / yields gb1 ----
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint4 hi, lo;
                split8(xv + 8 * g, hi, lo);
                const uint32_t o = umma::mnmajor_off_b16(grp * 16 + 8 * g, row, 128);
                *reinterpret_cast<uint4*>(sm + OFF_X + o) = hi;
                *reinterpret_cast<uint4*>(sm + OFF_X + 8192 + o) = lo;
            }
            umma::fence_async_smem();
            umma::mbar_arrive(&bars->xrdy);
            TC4_TR(1);
            // ---- forward epilogues (this team's steps i = grp, grp + 2, ...): a = ELU(z) back into the Z columns as packed bf16
            // [hi 16 | lo 16].  The Z of the team's NEXT step is loaded before this one is processed ----
            auto fwd_step = [&](int i, uint32_t (&zc)[32], uint32_t (&zn)[32]) {
                const int b = i & 3;
                umma::wait_ld();                                   // zc holds step i
                if (i + 2 < nst) {
                    const int bn = (i + 2) & 3;
                    umma::mbar_wait(&bars->zf_rdy[bn], (phzf >> bn) & 1u); phzf ^= 1u << bn;
                    umma::fence_after();
                    umma::ld32(tbase + laneblk + COL_W + bn * 32, zn);
                }
                TC4_TR(10 + i);
                const uint32_t ta = tbase + laneblk + COL_W + b * 32;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) split_pair(kc_elu(__uint_as_float(zc[2 * k])), kc_elu(__uint_as_float(zc[2 * k + 1])), hi[k], lo[k]);
                umma::st16(ta, hi);
                umma::st16(ta + 16, lo);
                umma::wait_st();
                umma::fence_before();
                umma::mbar_arrive(&bars->zf_used[b]);
                TC4_TR(30 + i);
            };
            {
                uint32_t zA[32], zB[32];
                umma::mbar_wait(&bars->zf_rdy[grp], (phzf >> grp) & 1u); phzf ^= 1u << grp;
                umma::fence_after();
                umma::ld32(tbase + laneblk + COL_W + grp * 32, zA);
                for (int i = grp; i < nst; i += 4) { fwd_step(i, zA, zB); fwd_step(i + 2, zB, zA); }
            }
            // ---- loss and dL/do: team 0 (its threads own the sample rows); team 1 goes straight on to its backward steps ----
            if (grp == 0) {
            // physics prediction and target of this sample, and the target's Euler angles: formed while the last GEMM2 drains
            float ph[25], tg[25], et[3];
            if (valid) {
#pragma unroll
                for (int r = 0; r < 25; ++r) { ph[r] = PHYS[(size_t)qrow * 25 + r]; tg[r] = TGT[(size_t)qrow * 25 + r]; }
                quat_to_euler(tg + 3, et);
            }
            umma::mbar_wait(&bars->ordy, pho); pho ^= 1;
            umma::fence_after();
            TC4_TR(50);
            {
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
                    float ep[3], ge[3], gq[4];
                    quat_to_euler(pred + 3, ep);
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
                // gb2 += sum over the warp's 32 samples (kept in shared memory: 25 live registers less in the epilogue loops)
#pragma unroll
                for (int c = 0; c < 25; ++c) {
                    float sg = g[c];
#pragma unroll
                    for (int o2 = 16; o2 > 0; o2 >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o2);
                    if (lane == 0) bars->redb[warp * 25 + c] += sg;
                }
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
            TC4_TR(51);
            }   // team 0
            // ---- backward epilogues (this team's steps j = grp, grp + 2, ...): thread = unit `row` of chunk j >> 2, columns =
            // samples 32 (j & 3) + k.  a = ELU(z), dz = dA ELU'(z) go back in place as packed bf16 [hi 8 | lo 8] per 16 samples (the
            // A operands of the gradient MMAs).  Samples beyond Q and units beyond H carry z = dA = 0 (zero X rows / zero weight
            // rows): they contribute nothing ----
            for (int j = grp; j < nst; j += 2) {
                const int b = j & 3;
                umma::mbar_wait(&bars->zb_rdy[b], (phzb >> b) & 1u); phzb ^= 1u << b;
                umma::fence_after();
                TC4_TR(60 + j);
                const uint32_t ta = tbase + laneblk + COL_W + b * 64;
                uint32_t z[32], d[32];
                umma::ld32(ta, z);
                umma::ld32(ta + 32, d);
                umma::wait_ld();
                uint32_t ah[16], al[16], dh[16], dl[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) {
                    const float z0 = __uint_as_float(z[2 * k]), z1 = __uint_as_float(z[2 * k + 1]);
                    const float a0 = kc_elu(z0), a1 = kc_elu(z1);
                    const float g0 = __uint_as_float(d[2 * k]) * (z0 > 0.f ? 1.f : a0 + 1.f);       // ELU'(z) = e^z = ELU(z) + 1
                    const float g1 = __uint_as_float(d[2 * k + 1]) * (z1 > 0.f ? 1.f : a1 + 1.f);
                    split_pair(a0, a1, ah[k], al[k]);
                    split_pair(g0, g1, dh[k], dl[k]);
                }
                umma::st8(ta, ah); umma::st8(ta + 8, al); umma::st8(ta + 16, ah + 8); umma::st8(ta + 24, al + 8);
                umma::st8(ta + 32, dh); umma::st8(ta + 40, dl); umma::st8(ta + 48, dh + 8); umma::st8(ta + 56, dl + 8);
                umma::wait_st();
                umma::fence_before();
                umma::mbar_arrive(&bars->zb_done[b]);
                TC4_TR(80 + j);
            }
            // X / dO of the next tile overwrite what this tile's gradient MMAs read (the next tile's inputs are fetched into
            // registers first: their latency hides behind this wait)
            if (t + 1 < my_tiles) load_x(tile + gridDim.x);
            umma::mbar_wait(&bars->gdone, phg); phg ^= 1;
            umma::fence_after();
            TC4_TR(99);
        }
        // ---- this CTA's partial gradients ----
        float* out = partial + (size_t)blockIdx.x * NP;
        const int64_t ob1 = (int64_t)hidden * 28, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
        if (my_tiles == 0) {   // (grid <= ntiles, so this does not happen; keep the slice defined anyway)
            for (int64_t i = tid; i < NP; i += 256) out[i] = 0.f;
        } else {
            for (int c = grp; c < nchunk; c += 2) {
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
        {
            double s = lossacc;
#pragma unroll
            for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_down_sync(0xffffffffu, s, o2);
            if (lane == 0) bars->redd[warp] = s;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        if (tid < 25) out[ob2 + tid] = bars->redb[tid] + bars->redb[25 + tid] + bars->redb[50 + tid] + bars->redb[75 + tid];
        if (tid == 0) loss_part[blockIdx.x] = bars->redd[0] + bars->redd[1] + bars->redd[2] + bars->redd[3];
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Host side: same contract as kc_train_tc_launch (kc_train_tc.cu); `img`: 192 KB of workspace.
int kc_train_tc4_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS, const float* TGT,
                        unsigned char* img, float* partial, int64_t NP, double* loss_part, float* pred_out, int grid,
                        cudaStream_t st) {
    const int nsub = 2 * ((mlp->hidden + 127) / 128);
    kc_tc4_prep_weights_kernel<<<32, 256, 0, st>>>((const float*)mlp->W1, (const float*)mlp->b1, (const float*)mlp->W2,
                                                   mlp->hidden, img);
    KC_CHECK_LAUNCH("kc_tc4_prep_weights_kernel");
    cudaFuncSetAttribute(kc_train_tc4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc4::SMEM_BYTES);
    kc_train_tc4_kernel<<<grid, 288, tc4::SMEM_BYTES, st>>>(mlp->hidden, nsub, img, (const float*)mlp->b2, ds, Q, T_, K, X, PHYS, TGT,
                                                            partial, NP, loss_part, pred_out);
    KC_CHECK_LAUNCH("kc_train_tc4_kernel");
    return KC_OK;
}
