// kc_train_tc3.cu — the teacher-forced KNODE training step on tcgen05 / TMEM, third generation.  Same maths and operand
// formats as kc_train_tc2.cu (physics_train.py:313-368; fp32 model, 28 inputs, hidden <= 512; bf16 hi/lo in 3 passes =
// fp32-grade), with the BACKWARD half transposed so that no activation ever goes through shared memory:
//
// The clock64 timeline of kc_train_tc2_kernel (tools/trace_train_tc2.py, KC_TRACE_GEN=2) showed a tile of 128 samples taking 46.8k
// cycles of which 24k were the backward epilogues (3.0k per 64-unit sub-chunk: 1.4k of issue, 0.7k of shared-memory stores of the a / dz
// tiles - 64 KB per sub-chunk, bandwidth bound - and the rest barrier / fence latency), while the gradient MMAs re-read those
// tiles three times (240 KB per 128 units: 40 cycles per N = 32 MMA, shared-memory bound).
//
// Here the backward GEMMs produce Z^T and dA^T: D[128 UNITS x 64 samples] = W_c[128 units x 32] * tile[64 samples x 32]^T (the
// weight image is the A operand, the X / dO tile - the very same bytes as in the forward - the B operand).  The epilogue
// thread is a UNIT row, a = ELU(z) and dz = dA ELU'(z) are written back in place as packed bf16 hi | lo, and the gradient
// GEMMs gW1_c += dZ^T X, gW2_c^T += A^T dO read them straight from TMEM (".ts" MMA: M = 128 units, K = 64 samples, 17 cycles
// per N = 32 MMA).  Shared memory then holds only the X / dO tiles and ALL operand images (192 KB, loaded once per CTA by bulk
// copies: no weight ring, no loader warp).
//
// One persistent CTA per SM, tiles of 128 samples.  288 threads:
//   warps 0..7  epilogue: forward  thread = sample row (tid & 127), group = tid >> 7 owns 32 of a sub-chunk's 64 units;
//                         backward thread = unit row of the 128-unit chunk, group owns 32 of a step's 64 samples
//   warp 8      MMA issue, warp-uniform (the instruction is predicated on elect.sync); its lane 0 also starts the weight copies
// TMEM (512 columns): gW1 accumulators 0..127 (32 per 128-unit chunk), gW2^T 128..255, working area 256..511:
//   forward : ring of three 64-column Z buffers (256, 320, 384) and O at 448..479 (as in kc_train_tc2.cu)
//   backward: two buffers of [Z^T 64 | dA^T 64] (256, 384); step j = (chunk j >> 1, sample half j & 1) uses buffer j & 1
// Backward per step: Z^T and dA^T (one commit) -> epilogue in place -> 24 gradient MMAs from TMEM, then the refill of the same
// buffer for step j + 2 (the tensor pipe executes in issue order, so no barrier is needed between the two).
// Packed layouts inside a thread group's 32 columns: forward [hi 16 | lo 16] (two 16-column stores), backward [hi 8 | lo 8] per
// 16 samples (8-column stores: measured faster).  A 3-pass product is ONE issue batch (one elect.sync per 6 / 12 MMAs).
// The loss of a sample row is split over the two groups (outputs 0..7 with the Euler angles | outputs 8..24).
// At the end every CTA writes one partial-gradient slice in the layout kc_train_reduce_kernel sums.  A tile now takes 34.5k
// cycles (DESIGN.md section 5 has the timeline and the measured history 199 -> 148 us).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kc_rod.cuh"
#include "kc_train_prep.cuh"
#include <cstdlib>
#include "kc_umma.cuh"

// Development aid (make EXTRA=-DKC_TC3_TRACE): clock64 stamps of CTA 0's epilogue warps 0 / 4 and MMA warp, tools/trace_train_tc2.py (KC_TRACE_GEN=3).
#ifdef KC_TC3_TRACE
__device__ long long* g_tc3_trace = nullptr;
extern "C" int kc_train_tc3_set_trace(long long* p) { return (int)cudaMemcpyToSymbol(g_tc3_trace, &p, sizeof(p)); }
#define TC3_TR(id) do { if (tr_role >= 0 && lane == 0 && g_tc3_trace && tr_n < 2048)                                         \
        g_tc3_trace[tr_role * 2048 + tr_n++] = ((long long)(id) << 48) | (clock64() & 0xffffffffffffll); } while (0)
#else
#define TC3_TR(id) do {} while (0)
#endif

namespace tc3 {
constexpr int OFF_X = 0;            // X^T hi | lo, bf16 [32 inputs x 128 samples], 8 inputs contiguous        2 x 8192
constexpr int OFF_DO = 16384;       // dO^T hi | lo, same layout [32 outputs x 128 samples]                    2 x 8192
constexpr int OFF_W1 = 32768;       // per 128-unit chunk: W1 hi | lo, K-major [128 units x 32] (col 28 = b1)  4 x 16384
constexpr int OFF_W2T = 98304;      // per chunk: W2^T hi | lo, K-major [128 units x 32 outs]                  4 x 16384
constexpr int OFF_W2F = 163840;     // per 64-unit sub-chunk: W2 hi | lo, K-major [32 outs x 64 units]         8 x 8192
constexpr int OFF_MISC = 229376;
constexpr int SMEM_BYTES = OFF_MISC + 1024;
constexpr int COL_GW1 = 0, COL_GW2 = 128, COL_W = 256, COL_O = 448;

struct Bars {
    uint64_t wfull, xrdy, ordy, dordy, gdone;
    uint64_t zf_rdy[3], zf_used[3];
    uint64_t zb_rdy[2], zb_done[2];
    uint32_t tmem_slot;
    double redd[8];
    float redb[4 * 25];
    float b2s[32];
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
}  // namespace tc3

// W1[H][28], b1[H], W2[25][H] -> the three operand images in shared-memory order (zero rows beyond H)
__global__ void kc_tc3_prep_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
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

// FUSE: the kernel forms its own samples.  The thread group that has nothing to do while the loss of a tile is being formed
// (group 1; the tensor pipe idles there too) computes X | phys | tgt of the CTA's NEXT tile from the trajectory (prep_sample:
// gathers + the rod ODE at the key node) and leaves them in the L2-resident sample arrays, from where all threads read them
// a phase later, as they would read the output of the stand-alone prep kernel (26 us, its own launch) otherwise.
struct Tc3Src { const float* traj; const float* controls; int64_t B; };
template <bool FUSE>
__global__ void __launch_bounds__(288, 1)
kc_train_tc3_kernel(int hidden, int nsub, const unsigned char* __restrict__ img, const float* __restrict__ b2, float ds, int64_t Q,
                    int T_, int K, float* X, float* PHYS, float* TGT,
                    float* __restrict__ partial, int64_t NP, double* __restrict__ loss_part, float* __restrict__ pred_out,
                    const __grid_constant__ RodC<float> RP, const __grid_constant__ KeyIdx64 key, const Tc3Src src) {
    extern __shared__ __align__(1024) unsigned char sm[];
    using namespace tc3;
    Bars* bars = reinterpret_cast<Bars*>(sm + OFF_MISC);
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int nchunk = nsub / 2;
    if (warp == 0) umma::tmem_alloc(&bars->tmem_slot, 512);
    if (tid == 32) {
        umma::mbar_init(&bars->wfull, 1);
        umma::mbar_init(&bars->xrdy, 256); umma::mbar_init(&bars->ordy, 1); umma::mbar_init(&bars->dordy, 256);
        umma::mbar_init(&bars->gdone, 1);
        for (int i = 0; i < 3; ++i) { umma::mbar_init(&bars->zf_rdy[i], 1); umma::mbar_init(&bars->zf_used[i], 256); }
        for (int i = 0; i < 2; ++i) { umma::mbar_init(&bars->zb_rdy[i], 1); umma::mbar_init(&bars->zb_done[i], 256); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    const uint32_t tbase = __shfl_sync(0xffffffffu, bars->tmem_slot, 0);
    const int64_t ntiles = (Q + 127) / 128;
    const int64_t my_tiles = ntiles > blockIdx.x ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
#ifdef KC_TC3_TRACE
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
        const uint32_t idZ = umma::make_idesc_bf16(128, 64), idO = umma::make_idesc_bf16(128, 32);
        const uint32_t idG = umma::make_idesc_bf16(128, 32, 0, 1);      // A from TMEM, B = X / dO tile viewed MN-major
        const uint32_t sX = umma::smem_u32(sm + OFF_X), sO = umma::smem_u32(sm + OFF_DO);
        const uint64_t dXh = umma::make_desc(sX, 2048, 128), dXl = dstep(dXh, 8192);        // [samples x 32] K-major views
        const uint64_t dOh = umma::make_desc(sO, 2048, 128), dOl = dstep(dOh, 8192);
        const uint64_t mXh = umma::make_desc(sX, 128, 2048), mXl = dstep(mXh, 8192);        // [32 x samples] MN-major views
        const uint64_t mOh = umma::make_desc(sO, 128, 2048), mOl = dstep(mOh, 8192);
        uint32_t phx = 0, phdo = 0, phzfu = 0, phzbd = 0;
        // D[128 x 64] = A[128 x 32] * B[64 x 32]^T, both K-major with 16 inputs per instruction, 3 passes (hi hi, lo hi, hi lo)
        auto gemm_k32 = [&](uint32_t d, uint64_t ah, uint64_t al, uint32_t astep, uint64_t bh, uint64_t bl, uint32_t bstep) {
            umma::mma_bf16_ss_3x2_w(d, ah, al, astep >> 4, bh, bl, bstep >> 4, idZ);
        };
        if (my_tiles > 0) { umma::mbar_wait(&bars->wfull, 0); umma::fence_after(); }
        for (int64_t t = 0; t < my_tiles; ++t) {
            const bool first_tile = t == 0;
            umma::mbar_wait(&bars->xrdy, phx); phx ^= 1;
            umma::fence_after();
            TC3_TR(100);
            // ---------------- forward: Z_s = X W1_s^T (A = X tile, B = 64 rows of the W1 image) ----------------
            auto fwd_gemm1 = [&](int s) {
                const uint32_t w = umma::smem_u32(sm + OFF_W1 + (s >> 1) * 16384 + (s & 1) * 4096);
                const uint64_t bh = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (s % 3) * 64, dXh, dXl, 4096, bh, dstep(bh, 8192), 256);
                umma::commit_w(&bars->zf_rdy[s % 3]);
            };
            for (int s = 0; s < nsub && s < 3; ++s) fwd_gemm1(s);
            for (int s = 0; s < nsub; ++s) {
                const int b = s % 3;
                umma::mbar_wait(&bars->zf_used[b], (phzfu >> b) & 1u); phzfu ^= 1u << b;
                umma::fence_after();
                TC3_TR(120 + s);
                const uint64_t wh = umma::make_desc(umma::smem_u32(sm + OFF_W2F + s * 8192), 128, 1024), wl = dstep(wh, 4096);
                const uint32_t ab = tbase + COL_W + b * 64;
                // 16-unit group kk lives in the 32-column block of the thread group that owns it: [hi 16 | lo 16] columns
                umma::mma_bf16_ts_3x4_w(tbase + COL_O, ab, wh, wl, idO, s ? 1u : 0u);
                if (s + 3 < nsub) fwd_gemm1(s + 3);
                TC3_TR(130 + s);
            }
            umma::commit_w(&bars->ordy);
            // ---------------- backward ----------------
            // step j: chunk c = j >> 1, samples 64 h .. 64 h + 63 (h = j & 1), buffer j & 1
            auto bwd_z = [&](int j) {       // Z^T = W1_c X_h^T
                const uint32_t w = umma::smem_u32(sm + OFF_W1 + (j >> 1) * 16384);
                const uint64_t ah = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (j & 1) * 128, ah, dstep(ah, 8192), 256, dstep(dXh, (j & 1) * 1024), dstep(dXl, (j & 1) * 1024), 4096);
            };
            auto bwd_da = [&](int j) {      // dA^T = W2_c^T dO_h^T
                const uint32_t w = umma::smem_u32(sm + OFF_W2T + (j >> 1) * 16384);
                const uint64_t ah = umma::make_desc(w, 128, 512);
                gemm_k32(tbase + COL_W + (j & 1) * 128 + 64, ah, dstep(ah, 8192), 256, dstep(dOh, (j & 1) * 1024), dstep(dOl, (j & 1) * 1024), 4096);
                umma::commit_w(&bars->zb_rdy[j & 1]);
            };
            // the Z halves of the first two steps do not need dO: they run while the loss is being formed (their columns are
            // free: every forward MMA that read them was issued before, and dA of step 1 - which overlaps O - waits for dO)
            for (int j = 0; j < nsub && j < 2; ++j) bwd_z(j);
            umma::mbar_wait(&bars->dordy, phdo); phdo ^= 1;
            umma::fence_after();
            TC3_TR(140);
            for (int j = 0; j < nsub && j < 2; ++j) bwd_da(j);
            for (int j = 0; j < nsub; ++j) {
                const int b = j & 1, c = j >> 1, h = j & 1;
                umma::mbar_wait(&bars->zb_done[b], (phzbd >> b) & 1u); phzbd ^= 1u << b;   // a, dz are in TMEM as bf16 hi | lo
                umma::fence_after();
                TC3_TR(150 + j);
                const uint32_t d1 = tbase + COL_GW1 + 32 * c, d2 = tbase + COL_GW2 + 32 * c;
                const uint32_t acc0 = (first_tile && h == 0) ? 0u : 1u;
                const uint32_t za = tbase + COL_W + b * 128, da = za + 64;
                // gW1_c += dZ^T X, gW2_c^T += A^T dO (hi*hi, lo*hi, hi*lo) over this step's 64 samples
                umma::mma_bf16_ts_3x4_w<16, 32, 48, 8>(d1, da, dstep(mXh, h * 1024), dstep(mXl, h * 1024), idG, acc0);
                umma::mma_bf16_ts_3x4_w<16, 32, 48, 8>(d2, za, dstep(mOh, h * 1024), dstep(mOl, h * 1024), idG, acc0);
                TC3_TR(160 + j);
                if (j + 2 < nsub) { bwd_z(j + 2); bwd_da(j + 2); }   // same buffer: executes after the MMAs above (issue order)
                TC3_TR(170 + j);
            }
            umma::commit_w(&bars->gdone);    // every MMA that read the X / dO tiles of this tile is done
        }
    } else {
        // ---- epilogue warps ----
        const int row = tid & 127, grp = tid >> 7;     // grp: which 32-column block of every 64-column buffer this thread owns
        const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t phzf = 0, phzb = 0, pho = 0, phg = 0;
        double lossacc = 0.0;
        if (tid < 100) bars->redb[tid] = 0.f;     // gb2 partial sums: row (warp & 3), columns 0..7 by group 0, 8..24 by group 1
        if (tid >= 128 && tid < 160) bars->b2s[tid - 128] = tid - 128 < 25 ? b2[tid - 128] : 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
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
        // sample `row` of a tile straight from the trajectory -> its rows of X | PHYS | TGT (FUSE only)
        auto prep_tile = [&](int64_t tile_) {
            const int64_t qr = tile_ * 128 + row;
            if (qr >= Q) return;
            const int N = RP.N, kk = (int)(qr % K);
            const int64_t bt = qr / K;
            const int tt = (int)(bt % (T_ - 1));
            const int64_t bb = bt / (T_ - 1);
            const float* tb = src.traj + (size_t)bb * T_ * 25 * N;
            float x[32];
            float* ph = PHYS + (size_t)qr * 25;
            float* tg = TGT + (size_t)qr * 25;
            if (RP.diag)
                prep_sample<float, true, 28>(RP, key.k[kk], tb + (size_t)(tt + 1) * 25 * N, tb + (size_t)tt * 25 * N,
                                             tb + (size_t)(tt > 0 ? tt - 1 : 0) * 25 * N, src.controls + (size_t)(bb * T_ + tt) * 4, x, ph, tg);
            else
                prep_sample<float, false, 28>(RP, key.k[kk], tb + (size_t)(tt + 1) * 25 * N, tb + (size_t)tt * 25 * N,
                                              tb + (size_t)(tt > 0 ? tt - 1 : 0) * 25 * N, src.controls + (size_t)(bb * T_ + tt) * 4, x, ph, tg);
            float4* xr = reinterpret_cast<float4*>(X + (size_t)qr * 32);
#pragma unroll
            for (int i = 0; i < 8; ++i) xr[i] = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
        };
        if (FUSE && my_tiles > 0) {
            if (grp == 1) prep_tile(blockIdx.x);
            asm volatile("bar.sync 1, 256;" ::: "memory");      // the samples of the first tile are visible to the whole CTA
        }
        if (my_tiles > 0) load_x(blockIdx.x);
        for (int64_t t = 0; t < my_tiles; ++t) {
            const int64_t tile = blockIdx.x + t * gridDim.x;
            const int64_t qrow = tile * 128 + row;
            const bool valid = qrow < Q;
            TC3_TR(0);
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
            TC3_TR(1);
            // ---- forward epilogues: a = ELU(z) back into the Z columns as packed bf16 hi | lo.  The Z of the NEXT sub-chunk is
            // loaded (it was produced two sub-chunks ago) before this one is processed: its TMEM latency and the barrier round
            // trip hide behind the arithmetic ----
            auto fwd_step = [&](int s, uint32_t (&zc)[32], uint32_t (&zn)[32]) {
                const int b = s % 3;
                umma::wait_ld();                                   // zc holds sub-chunk s
                if (s + 1 < nsub) {
                    const int bn = (s + 1) % 3;
                    umma::mbar_wait(&bars->zf_rdy[bn], (phzf >> bn) & 1u); phzf ^= 1u << bn;
                    umma::fence_after();
                    umma::ld32(tbase + laneblk + COL_W + bn * 64 + grp * 32, zn);
                }
                TC3_TR(10 + s);
                const uint32_t ta = tbase + laneblk + COL_W + b * 64 + grp * 32;
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) split_pair(kc_elu(__uint_as_float(zc[2 * i])), kc_elu(__uint_as_float(zc[2 * i + 1])), hi[i], lo[i]);
                umma::st16(ta, hi);
                umma::st16(ta + 16, lo);
                umma::wait_st();
                umma::fence_before();
                umma::mbar_arrive(&bars->zf_used[b]);
                TC3_TR(20 + s);
            };
            {
                uint32_t zA[32], zB[32];
                umma::mbar_wait(&bars->zf_rdy[0], phzf & 1u); phzf ^= 1u;
                umma::fence_after();
                umma::ld32(tbase + laneblk + COL_W + grp * 32, zA);
                for (int s = 0; s < nsub; s += 2) { fwd_step(s, zA, zB); fwd_step(s + 1, zB, zA); }
            }
            // ---- loss and dL/do, split over the two thread groups of a sample row: group 0 takes outputs 0..7 (position,
            // quaternion -> Euler angles, n_x), group 1 outputs 8..24 (n_y, n_z, m, q, w, v, u).  The physics prediction, the
            // target and the target's Euler angles are fetched / formed while the last GEMM2 drains ----
            float ph[17], tg[17], et[3];
            if (valid) {
                if (grp == 0) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) { ph[r] = PHYS[(size_t)qrow * 25 + r]; tg[r] = TGT[(size_t)qrow * 25 + r]; }
                    quat_to_euler(tg + 3, et);
                } else {
#pragma unroll
                    for (int r = 0; r < 17; ++r) { ph[r] = PHYS[(size_t)qrow * 25 + 8 + r]; tg[r] = TGT[(size_t)qrow * 25 + 8 + r]; }
                }
            }
            umma::mbar_wait(&bars->ordy, pho); pho ^= 1;
            umma::fence_after();
            TC3_TR(30);
            {
                const float S = float(T_ - 1);
                const float wp = 1.f / (float(3 * K) * S), wf = 1.f / (float(12 * K) * S), wz = 1.f / (float(6 * K) * S);
                uint32_t v[32];
                umma::ld32(tbase + laneblk + COL_O, v);
                umma::wait_ld();
                float acc = 0.f;
                float* po = nullptr;
                if (pred_out && valid) po = pred_out + (size_t)(qrow / K) * 25 * K + (int)(qrow % K);
                if (grp == 0) {
                    float g[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) g[c] = 0.f;
                    if (valid) {
                        float pred[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) pred[r] = ph[r] + ds * (__uint_as_float(v[r]) + bars->b2s[r]);
#pragma unroll
                        for (int r = 0; r < 3; ++r) { const float e = pred[r] - tg[r]; acc += wp * e * e; g[r] = 2.f * wp * e * ds; }
                        { const float e = pred[7] - tg[7]; acc += wf * e * e; g[7] = 2.f * wf * e * ds; }
                        float ep[3], ge[3], gq[4];
                        quat_to_euler(pred + 3, ep);
#pragma unroll
                        for (int i = 0; i < 3; ++i) { const float e = ep[i] - et[i]; acc += wp * e * e; ge[i] = 2.f * wp * e; }
                        quat_to_euler_vjp(pred + 3, ge, gq);
#pragma unroll
                        for (int i = 0; i < 4; ++i) g[3 + i] = gq[i] * ds;
                        if (po) {
#pragma unroll
                            for (int r = 0; r < 8; ++r) po[r * K] = pred[r];
                        }
                    }
                    // gb2 += sum over the warp's 32 samples (kept in shared memory: 25 live registers less in the epilogue loops)
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float sg = g[c];
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o2);
                        if (lane == 0) bars->redb[(warp & 3) * 25 + c] += sg;
                    }
                    uint4 hi, lo;
                    split8(g, hi, lo);
                    const uint32_t off = umma::mnmajor_off_b16(0, row, 128);
                    *reinterpret_cast<uint4*>(sm + OFF_DO + off) = hi;
                    *reinterpret_cast<uint4*>(sm + OFF_DO + 8192 + off) = lo;
                } else {
                    float g[24];     // outputs 8..31 (25..31 are padding)
#pragma unroll
                    for (int c = 0; c < 24; ++c) g[c] = 0.f;
                    if (valid) {
#pragma unroll
                        for (int r = 0; r < 17; ++r) {
                            const int c = 8 + r;
                            const float o = __uint_as_float(v[c]) + bars->b2s[c];
                            const float pred = ph[r] + (c < 19 ? ds * o : o);
                            const float e = pred - tg[r];
                            const float wgt = c < 19 ? wf : wz;
                            acc += wgt * e * e;
                            g[r] = c < 19 ? 2.f * wgt * e * ds : 2.f * wgt * e;
                            if (po) po[c * K] = pred;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 17; ++r) {
                        float sg = g[r];
#pragma unroll
                        for (int o2 = 16; o2 > 0; o2 >>= 1) sg += __shfl_xor_sync(0xffffffffu, sg, o2);
                        if (lane == 0) bars->redb[(warp & 3) * 25 + 8 + r] += sg;
                    }
#pragma unroll
                    for (int gi = 0; gi < 3; ++gi) {
                        uint4 hi, lo;
                        split8(g + 8 * gi, hi, lo);
                        const uint32_t off = umma::mnmajor_off_b16(8 + gi * 8, row, 128);
                        *reinterpret_cast<uint4*>(sm + OFF_DO + off) = hi;
                        *reinterpret_cast<uint4*>(sm + OFF_DO + 8192 + off) = lo;
                    }
                }
                if (valid) lossacc += (double)acc;
            }
            umma::fence_async_smem();
            umma::fence_before();
            umma::mbar_arrive(&bars->dordy);
            TC3_TR(31);
            if (FUSE && grp == 1 && t + 1 < my_tiles) prep_tile(tile + gridDim.x);   // off the critical path: dO does not wait for it
            // ---- backward epilogues: thread = unit `row` of chunk j >> 1, columns = samples 64 (j & 1) + 32 grp + i.  a = ELU(z),
            // dz = dA ELU'(z) go back in place as packed bf16 hi | lo (the A operands of the gradient MMAs).  Samples beyond Q
            // and units beyond H carry z = dA = 0 (zero X rows / zero weight rows): they contribute nothing ----
            for (int j = 0; j < nsub; ++j) {
                const int b = j & 1;
                umma::mbar_wait(&bars->zb_rdy[b], (phzb >> b) & 1u); phzb ^= 1u << b;
                umma::fence_after();
                TC3_TR(40 + j);
                const uint32_t ta = tbase + laneblk + COL_W + b * 128 + grp * 32;
                uint32_t z[32], d[32];
                umma::ld32(ta, z);
                umma::ld32(ta + 64, d);
                umma::wait_ld();
                TC3_TR(50 + j);
                uint32_t ah[16], al[16], dh[16], dl[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float z0 = __uint_as_float(z[2 * i]), z1 = __uint_as_float(z[2 * i + 1]);
                    const float a0 = kc_elu(z0), a1 = kc_elu(z1);
                    const float g0 = __uint_as_float(d[2 * i]) * (z0 > 0.f ? 1.f : a0 + 1.f);       // ELU'(z) = e^z = ELU(z) + 1
                    const float g1 = __uint_as_float(d[2 * i + 1]) * (z1 > 0.f ? 1.f : a1 + 1.f);
                    split_pair(a0, a1, ah[i], al[i]);
                    split_pair(g0, g1, dh[i], dl[i]);
                }
                // packed layout per thread group: [hi 8 | lo 8] columns per 16 samples (8-column stores: measured faster than
                // [hi 16 | lo 16] with 16-column stores, 0.186 vs 0.201 ms per step)
                umma::st8(ta, ah); umma::st8(ta + 8, al); umma::st8(ta + 16, ah + 8); umma::st8(ta + 24, al + 8);
                umma::st8(ta + 64, dh); umma::st8(ta + 72, dl); umma::st8(ta + 80, dh + 8); umma::st8(ta + 88, dl + 8);
                umma::wait_st();
                umma::fence_before();
                umma::mbar_arrive(&bars->zb_done[b]);
                TC3_TR(60 + j);
            }
            // X / dO of the next tile overwrite what this tile's gradient MMAs read (the next tile's inputs are fetched into
            // registers first: their latency hides behind this wait)
            if (FUSE) asm volatile("bar.sync 1, 256;" ::: "memory");   // group 1's samples of the next tile are visible
            if (t + 1 < my_tiles) load_x(tile + gridDim.x);
            umma::mbar_wait(&bars->gdone, phg); phg ^= 1;
            umma::fence_after();
            TC3_TR(90);
        }
        // ---- this CTA's partial gradients ----
        float* out = partial + (size_t)blockIdx.x * NP;
        const int64_t ob1 = (int64_t)hidden * 28, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
        if (my_tiles == 0) {   // (grid <= ntiles, so this does not happen; keep the slice defined anyway)
            for (int64_t i = tid; i < NP; i += 256) out[i] = 0.f;
        } else if (grp == 0) {
            for (int c = 0; c < nchunk; ++c) {
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
        if (tid == 0)
            loss_part[blockIdx.x] = ((bars->redd[0] + bars->redd[1]) + (bars->redd[2] + bars->redd[3])) +
                                    ((bars->redd[4] + bars->redd[5]) + (bars->redd[6] + bars->redd[7]));
    }
    umma::fence_before();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Host side: same contract as kc_train_tc_launch (kc_train_tc.cu); `img`: 192 KB of workspace.  With `src` (trajectory,
// controls, rod constants, key nodes) the kernel forms its own samples and the caller skips the prep kernel.
int kc_train_tc3_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, float* X, float* PHYS, float* TGT,
                        unsigned char* img, float* partial, int64_t NP, double* loss_part, float* pred_out, int grid,
                        cudaStream_t st, const RodC<float>* RP, const KeyIdx64* key, const float* traj, const float* controls,
                        int64_t B) {
    const int nsub = 2 * ((mlp->hidden + 127) / 128);
    kc_tc3_prep_weights_kernel<<<32, 256, 0, st>>>((const float*)mlp->W1, (const float*)mlp->b1, (const float*)mlp->W2,
                                                   mlp->hidden, img);
    KC_CHECK_LAUNCH("kc_tc3_prep_weights_kernel");
    if (RP) {
        const Tc3Src src{traj, controls, B};
        cudaFuncSetAttribute(kc_train_tc3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::SMEM_BYTES);
        kc_train_tc3_kernel<true><<<grid, 288, tc3::SMEM_BYTES, st>>>(mlp->hidden, nsub, img, (const float*)mlp->b2, ds, Q, T_, K, X, PHYS,
                                                                      TGT, partial, NP, loss_part, pred_out, *RP, *key, src);
    } else {
        const RodC<float> none{};
        const KeyIdx64 nokey{};
        const Tc3Src src{nullptr, nullptr, 0};
        cudaFuncSetAttribute(kc_train_tc3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::SMEM_BYTES);
        kc_train_tc3_kernel<false><<<grid, 288, tc3::SMEM_BYTES, st>>>(mlp->hidden, nsub, img, (const float*)mlp->b2, ds, Q, T_, K, X, PHYS,
                                                                       TGT, partial, NP, loss_part, pred_out, none, nokey, src);
    }
    KC_CHECK_LAUNCH("kc_train_tc3_kernel");
    return KC_OK;
}
