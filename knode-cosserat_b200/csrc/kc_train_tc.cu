// kc_train_tc.cu — tensor-core (tcgen05 / TMEM) version of the teacher-forced KNODE training step, fp32 model, 28 inputs,
// hidden <= 512.  Replaces kc_train_fwd_kernel + kc_train_bwd_kernel (kc_train.cu), which stay as the SIMT reference path
// (other dtypes / shapes, and KC_TRAIN_MODE=simt).  Same maths (physics_train.py:313-368), see kc_train.cu header.
//
// One persistent CTA per SM, 256 threads, tiles of 128 samples (TMEM lane = sample row for the activations, = hidden unit
// for the gradient accumulators).  Per tile and per chunk of 128 hidden units:
//   forward : Z = X W1c^T            tcgen05.mma kind::tf32, 3-pass split (hi*hi + lo*hi + hi*lo: fp32-accurate), K = 32
//             (column 28 of X is 1 and carries b1) -> TMEM; epilogue: a = ELU(z) -> shared memory as bf16 hi/lo pairs;
//             O += A W2c^T            kind::f16 (bf16 3-pass split), K = 128 units, accumulated in TMEM over the chunks
//   loss    : pred, 4-term loss, dL/do per sample row (same code as the SIMT kernel)
//   backward: Z again (recompute) and dA = dO W2c (bf16 3-pass, K = 32); epilogue: dz = dA * ELU'(z); a and dz go to
//             shared memory as bf16 hi/lo; gW1c += dZ^T X and gW2c^T += A^T dO (bf16 3-pass), K = the tile's 128
//             samples, accumulating in TMEM across all tiles of the CTA.
// The activation tile A is written ONCE per chunk, thread = sample row, 8 units per 16-byte store; the same bytes are the
// K-major A operand of the forward GEMM (K = units) and the MN-major A operand of the gradient GEMM (K = samples) - only
// LBO/SBO swap in the descriptor.  The same holds for dO (A of the dA GEMM, B of the gW2 GEMM).
// At the end the accumulators are written as one partial-gradient slice per CTA; kc_train_reduce_kernel sums the slices.
// Operand layouts / descriptors are the ones pinned by kc_umma_selftest (tests/test_gpu_tensorcore.py).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "kc_rod.cuh"
#include "kc_umma.cuh"

namespace tc {
constexpr int TS = 128, HC = 128, THREADS = 256;
constexpr int OFF_XH = 0;        // X hi  [128 x 32] tf32 K-major                      16384
constexpr int OFF_XL = 16384;    // X lo                                               16384
constexpr int OFF_W2 = 32768;    // W2 chunk as a bf16 B operand, hi | lo (2 x 8192): pass 1 [32 outs x 128 units] K-major,
                                 // pass 2 [128 units x 32 outs] K-major                                      16384
constexpr int OFF_DZ = 49152;    // dZ^T hi | lo, bf16 MN-major [128 units x 128 samples] 2 x 32768;
                                 // W1 chunk hi | lo (tf32 K-major [128 x 32], 2 x 16384) ALIASES the LO half: the gradient
                                 // MMAs that read dZ lo are issued (and signalled) first, so the next W1 chunk can land
                                 // while the remaining gradient MMAs still run
constexpr int OFF_A = 114688;    // A^T hi | lo, bf16 MN-major                           2 x 32768
constexpr int OFF_XT = 180224;   // X^T hi | lo, bf16 MN-major [32 inputs x 128 samples]  2 x 8192
constexpr int OFF_DO = 196608;   // dO^T hi | lo, bf16 MN-major [32 outputs x 128 samples] 2 x 8192
constexpr int OFF_MISC = 212992; // tmem slot, mbarriers, reduction scratch
constexpr int OFF_W2B = OFF_MISC + 1024;   // second W2 image buffer (the images of consecutive stages alternate)   16384
constexpr int SMEM_BYTES = OFF_W2B + 16384;
constexpr int W1_TILE_FLOATS = 128 * 32;          // one hi or lo tile
constexpr int W2B_CHUNK_BYTES = 32768;            // per chunk: W2 fwd image hi|lo (16 KB) then W2^T bwd image hi|lo (16 KB)
// TMEM columns
constexpr int COL_Z = 0, COL_GW1 = 128, COL_GW2 = 256, COL_O = 384;
}  // namespace tc

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&v);
}
// x = hi + lo with hi, lo bf16 (about 16 mantissa bits together)
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
    hi = __bfloat162float(__float2bfloat16_rn(x));
    lo = x - hi;
}

// global -> shared bulk copy by the TMA engine (no registers, no threads): completes on an mbarrier by byte count
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(umma::smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}

// Eight values -> eight bf16 hi | eight bf16 lo (x = hi + lo to ~16 mantissa bits), ready for one 16-byte store each.
// hi = the fp32 word rounded to its upper half with integer arithmetic (add half an ulp, mask; a byte permute packs two
// of them), lo = x - hi exactly, rounded to bf16 by the packing conversion: nothing on the XU pipe (cvt.rn.bf16.f32 is an
// XU instruction and the kernel's exponentials live there too).  Plain truncation of hi would save one more instruction
// per value but doubles |lo|; the first Adam step (g / (|g| + eps)) is sensitive enough to see that.
__device__ __forceinline__ void split_pack8(const float x[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t a = (__float_as_uint(x[2 * i]) + 0x8000u) & 0xffff0000u;
        const uint32_t b = (__float_as_uint(x[2 * i + 1]) + 0x8000u) & 0xffff0000u;
        h[i] = __byte_perm(a, b, 0x7632);
        l[i] = pack_bf16x2(x[2 * i] - __uint_as_float(a), x[2 * i + 1] - __uint_as_float(b));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// K-major, no swizzle, 2-byte elements: element (r, k) of a tile with KT k-values per row
__device__ __forceinline__ uint32_t kmajor_off_b16(int r, int k, int KT) {
    return (uint32_t)(((r >> 3) * (KT >> 3) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

// W1[H][28], b1[H], W2[25][H] -> per chunk c: W1hl[c] = {hi tile, lo tile} (tf32 values, K-major interleaved, column 28
// = b1, rows >= H zero); W2b[c] = bf16 hi|lo images of W2 as the B operand of the forward GEMM ([32 outs][128 units],
// K-major) followed by hi|lo images of W2^T as the B operand of the dA GEMM ([128 units][32 outs], K-major).
__global__ void kc_tc_prep_weights_kernel(const float* __restrict__ W1, const float* __restrict__ b1,
                                          const float* __restrict__ W2, int hidden, float* __restrict__ W1hl,
                                          unsigned char* __restrict__ W2b) {
    const int c = blockIdx.x;   // grid = (chunks, 8): 8 CTAs share a chunk
    for (int e = blockIdx.y * blockDim.x + threadIdx.x; e < 128 * 32; e += gridDim.y * blockDim.x) {
        const int ul = e >> 5, k = e & 31, u = c * 128 + ul;
        float v = 0.f;
        if (u < hidden) v = k < 28 ? W1[(size_t)u * 28 + k] : (k == 28 ? b1[u] : 0.f);
        const float hi = umma::to_tf32_rn(v), lo = umma::to_tf32_rn(v - hi);
        const uint32_t o = umma::kmajor_off(ul, k, 32) >> 2;
        W1hl[(size_t)c * 2 * tc::W1_TILE_FLOATS + o] = hi;
        W1hl[(size_t)c * 2 * tc::W1_TILE_FLOATS + tc::W1_TILE_FLOATS + o] = lo;
        // the same (unit ul, output co = k) pair feeds both W2 images
        const int co = k;
        const float w = (u < hidden && co < 25) ? W2[(size_t)co * hidden + u] : 0.f;
        float wh, wl;
        split_bf16(w, wh, wl);
        unsigned char* base = W2b + (size_t)c * tc::W2B_CHUNK_BYTES;
        const uint32_t of = kmajor_off_b16(co, ul, 128);   // forward: rows = outputs, k = units
        *reinterpret_cast<__nv_bfloat16*>(base + of) = __float2bfloat16_rn(wh);
        *reinterpret_cast<__nv_bfloat16*>(base + 8192 + of) = __float2bfloat16_rn(wl);
        const uint32_t ob = kmajor_off_b16(ul, co, 32);    // backward: rows = units, k = outputs
        *reinterpret_cast<__nv_bfloat16*>(base + 16384 + ob) = __float2bfloat16_rn(wh);
        *reinterpret_cast<__nv_bfloat16*>(base + 16384 + 8192 + ob) = __float2bfloat16_rn(wl);
        // W1 chunk as the B operand of the input-gradient GEMM gX += dZ W1c: [32 inputs x 128 units] K-major (the same
        // format as the forward W2 image); input columns 28..31 (bias, padding) are zero
        {
            const float w1 = (u < hidden && k < 28) ? W1[(size_t)u * 28 + k] : 0.f;
            float h1, l1;
            split_bf16(w1, h1, l1);
            unsigned char* b1t = W2b + (size_t)4 * tc::W2B_CHUNK_BYTES + (size_t)c * 16384;
            *reinterpret_cast<__nv_bfloat16*>(b1t + of) = __float2bfloat16_rn(h1);
            *reinterpret_cast<__nv_bfloat16*>(b1t + 8192 + of) = __float2bfloat16_rn(l1);
        }
    }
}

// MODE 3: INPUT gradients only, gX[Q][gxpitch] = ((dOin W2) * ELU'(X W1^T + b1)) W1 (gxpitch passed in K, output in
// pred_out) — the MLP part of kc_ode_bwd / kc_mlp_bwd.
// MODE 0: the full training step.  MODE 1: forward only — pred_out[Q][25] = PHYS + s*o (s = ds for rows < 19, 1 for the
// rest; kc_ode_fwd / kc_mlp_fwd pass ds = 1 and PHYS = the physics part or zeros).  MODE 2: parameter gradients only from
// given samples (X, dOin[Q][32]) — kc_mlp_bwd / kc_ode_bwd and the sample reduction of kc_rollout_bwd.
template <int MODE>
__global__ void __launch_bounds__(tc::THREADS, 1)
kc_train_tc_kernel(int hidden, int nch, const float* __restrict__ W1hl, const unsigned char* __restrict__ W2b,
                   const float* __restrict__ b2, float ds, int64_t Q, int T_, int K, const float* __restrict__ X,
                   const float* __restrict__ PHYS, const float* __restrict__ TGT, float* __restrict__ partial,
                   int64_t NP, double* __restrict__ loss_part, float* __restrict__ pred_out,
                   const float* __restrict__ dOin, int xpitch, int dopitch) {
    extern __shared__ __align__(1024) unsigned char sm[];
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sm + tc::OFF_MISC);
    uint64_t* barZ = reinterpret_cast<uint64_t*>(sm + tc::OFF_MISC + 8);
    uint64_t* barG = reinterpret_cast<uint64_t*>(sm + tc::OFF_MISC + 16);
    uint64_t* barO = reinterpret_cast<uint64_t*>(sm + tc::OFF_MISC + 24);
    uint64_t* barW = reinterpret_cast<uint64_t*>(sm + tc::OFF_MISC + 96);   // weights of the next stage have landed
    uint64_t* barL = reinterpret_cast<uint64_t*>(sm + tc::OFF_MISC + 104);  // gradient MMAs reading the dZ lo tile done
    double* redd = reinterpret_cast<double*>(sm + tc::OFF_MISC + 32);    // [8]
    float* redb = reinterpret_cast<float*>(sm + tc::OFF_MISC + 128);     // [4][25]
    float* sO = reinterpret_cast<float*>(sm + tc::OFF_MISC + 640);       // unused now (kept for layout stability)
    (void)sO;
    const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;   // (warp: provably uniform)
    const int row = tid & 127, half = tid >> 7;
    const uint32_t laneblk = (uint32_t)((warp & 3) * 32) << 16;
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    if (tid == 0) {
        umma::mbar_init(barZ, 1);
        umma::mbar_init(barG, 1);
        umma::mbar_init(barO, 1);
        umma::mbar_init(barW, 1);
        umma::mbar_init(barL, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    umma::fence_async_smem();
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    // MMAs are issued by ALL lanes of warp 0 (descriptors in uniform registers, the instruction predicated on elect.sync):
    // issuing from the single-thread region `if (tid == 0)` made the compiler wrap every tcgen05.mma in an R2UR / ELECT loop
    const uint32_t tbase = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    uint32_t phZ = 0, phG = 0, phO = 0, phW = 0, phL = 0;
    const uint32_t idescZ = umma::make_idesc_tf32(128, 128);
    const uint32_t idescG = umma::make_idesc_bf16(128, 32, 1, 1);   // gradient GEMMs: both operands MN-major
    const uint32_t idescO = umma::make_idesc_bf16(128, 32, 0, 0);   // O  = A W2c^T   (K-major views)
    const uint32_t idescD = umma::make_idesc_bf16(128, 128, 0, 0);  // dA = dO W2c    (K-major views)
    const uint32_t aWB0 = umma::smem_u32(sm + tc::OFF_W2), aWB1 = umma::smem_u32(sm + tc::OFF_W2B);
    const uint32_t aXh = umma::smem_u32(sm + tc::OFF_XH), aXl = umma::smem_u32(sm + tc::OFF_XL);
    const uint32_t aW1h = umma::smem_u32(sm + tc::OFF_DZ + 32768), aW1l = aW1h + 16384;   // aliases the dZ LO tile
    const uint32_t aDZh = umma::smem_u32(sm + tc::OFF_DZ), aDZl = aDZh + 32768;
    const uint32_t aAh = umma::smem_u32(sm + tc::OFF_A), aAl = aAh + 32768;
    const uint32_t aXth = umma::smem_u32(sm + tc::OFF_XT), aXtl = aXth + 8192;
    const uint32_t aDOh = umma::smem_u32(sm + tc::OFF_DO), aDOl = aDOh + 8192;

    // Stages of a tile: s = 0..nch-1 forward chunks, s = nch..2nch-1 backward chunks.  The weights of a stage arrive by TMA
    // bulk copies issued by thread 0 as early as their buffers are free (barW counts the bytes):
    //   W1 chunk hi|lo (32 KB) -> OFF_DZ (free once the previous user of that region has completed),
    //   W2 image hi|lo (16 KB: forward [outs x units] or backward [units x outs]) -> OFF_W2 / OFF_W2B alternating by stage.
    auto fetch_w1 = [&](int c) {
        bulk_g2s(aW1h, W1hl + (size_t)c * 2 * tc::W1_TILE_FLOATS, 2 * tc::W1_TILE_FLOATS * 4, barW);
    };
    auto fetch_w2 = [&](int c, int pass, int stage) {
        bulk_g2s((stage & 1) ? aWB1 : aWB0, W2b + (size_t)c * tc::W2B_CHUNK_BYTES + (pass == 2 ? 16384 : 0), 16384, barW);
    };
    auto fetch_w1t = [&](int c, int stage) {   // MODE 3: W1^T image -> OFF_A (+16 KB for odd stages)
        bulk_g2s(aAh + ((stage & 1) ? 16384u : 0u), W2b + (size_t)4 * tc::W2B_CHUNK_BYTES + (size_t)c * 16384, 16384, barW);
    };
    constexpr uint32_t WBYTES = 32768 + 16384 + (MODE == 3 ? 16384 : 0);   // bytes per stage on barW
    auto issue_gemm1 = [&]() {  // Z[128 samples x 128 units] = X W1c^T, 3-pass tf32 split, K = 32 (4 x K8)
        uint32_t acc = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t a = (p == 1) ? aXl : aXh, b = (p == 2) ? aW1l : aW1h;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
                umma::mma_tf32_w(tbase + tc::COL_Z, umma::make_desc(a + kk * 256, 128, 1024), umma::make_desc(b + kk * 256, 128, 1024), idescZ, acc);
                acc = 1;
            }
        }
    };
    // O[128 samples x 32 outs] (+)= A[samples x 128 units] W2c^T: A = the activation tile viewed K-major (LBO 2048, SBO 128),
    // B = W2 image [32 outs x 128 units] K-major (LBO 128, SBO 2048); K = 128 units = 8 x K16
    auto issue_gemm2 = [&](bool first_chunk, uint32_t aWB) {
        uint32_t acc = first_chunk ? 0u : 1u;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t a = (p == 1) ? aAl : aAh, b = (p == 2) ? aWB + 8192 : aWB;
            for (int kk = 0; kk < 8; ++kk) {
                umma::mma_bf16_w(tbase + tc::COL_O, umma::make_desc(a + kk * 4096, 2048, 128), umma::make_desc(b + kk * 256, 128, 2048), idescO, acc);
                acc = 1;
            }
        }
        umma::commit_w(barO);
    };
    // dA[128 samples x 128 units] = dO[samples x 32 outs] W2c: A = the dO tile viewed K-major (LBO 2048, SBO 128),
    // B = W2^T image [128 units x 32 outs] K-major (LBO 128, SBO 512); K = 32 = 2 x K16
    auto issue_gemm3 = [&](uint32_t aWB) {
        uint32_t acc = 0;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t a = (p == 1) ? aDOl : aDOh, b = (p == 2) ? aWB + 8192 : aWB;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                umma::mma_bf16_w(tbase + tc::COL_O, umma::make_desc(a + kk * 4096, 2048, 128), umma::make_desc(b + kk * 256, 128, 512), idescD, acc);
                acc = 1;
            }
        }
    };
    // gW1c += dZ^T X ; gW2c^T += A^T dO ; K = 128 samples (8 x K16).  The pass that reads the dZ LO tile goes first and is
    // committed on barL (its shared memory is the landing zone of the next W1 chunk); everything else commits on barG.
    auto issue_grads = [&](int c, bool first_tile) {
        const uint32_t d1 = tbase + tc::COL_GW1 + 32 * c, d2 = tbase + tc::COL_GW2 + 32 * c;
        uint32_t acc = first_tile ? 0u : 1u;
        for (int kk = 0; kk < 8; ++kk) {   // dZ lo * X hi
            umma::mma_bf16_w(d1, umma::make_desc(aDZl + kk * 256, 128, 2048), umma::make_desc(aXth + kk * 256, 128, 2048), idescG, acc);
            acc = 1;
        }
        umma::commit_w(barL);
#pragma unroll
        for (int p = 0; p < 2; ++p) {      // dZ hi * X hi, dZ hi * X lo
            const uint32_t b = p == 0 ? aXth : aXtl;
            for (int kk = 0; kk < 8; ++kk)
                umma::mma_bf16_w(d1, umma::make_desc(aDZh + kk * 256, 128, 2048), umma::make_desc(b + kk * 256, 128, 2048), idescG, 1u);
        }
        acc = first_tile ? 0u : 1u;
#pragma unroll
        for (int p = 0; p < 3; ++p) {
            const uint32_t a = (p == 1) ? aAl : aAh, b = (p == 2) ? aDOl : aDOh;
            for (int kk = 0; kk < 8; ++kk) {
                umma::mma_bf16_w(d2, umma::make_desc(a + kk * 256, 128, 2048), umma::make_desc(b + kk * 256, 128, 2048), idescG, acc);
                acc = 1;
            }
        }
        umma::commit_w(barG);
    };

    // MODE 3: gX[128 samples x 32 inputs] += dZ[samples x 128 units] W1c: A = the dZ tile viewed K-major (as the activation
    // tile in GEMM2), B = the W1^T image; the pass that reads dZ lo goes first and commits barL (its bytes take the next W1)
    auto issue_gemm4 = [&](bool first_chunk, uint32_t aW1T) {
        uint32_t acc = first_chunk ? 0u : 1u;
        for (int kk = 0; kk < 8; ++kk) {
            umma::mma_bf16_w(tbase + tc::COL_GW1, umma::make_desc(aDZl + kk * 4096, 2048, 128), umma::make_desc(aW1T + kk * 256, 128, 2048), idescO, acc);
            acc = 1;
        }
        umma::commit_w(barL);
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            const uint32_t b = p == 0 ? aW1T : aW1T + 8192;
            for (int kk = 0; kk < 8; ++kk)
                umma::mma_bf16_w(tbase + tc::COL_GW1, umma::make_desc(aDZh + kk * 4096, 2048, 128), umma::make_desc(b + kk * 256, 128, 2048), idescO, 1u);
        }
        umma::commit_w(barG);
    };
    float gb2acc[25];
#pragma unroll
    for (int c = 0; c < 25; ++c) gb2acc[c] = 0.f;
    double lossacc = 0.0;
    const int64_t ntiles = (Q + tc::TS - 1) / tc::TS;
    bool first_tile = true;
    int gs = 0;   // running stage count: the W2 image buffers alternate by it (any number of stages per tile)
    if (tid == 0 && blockIdx.x < ntiles) {   // stage 0 of the first tile
        mbar_expect_tx(barW, WBYTES);
        fetch_w1(0);
        fetch_w2(0, MODE >= 2 ? 2 : 1, 0);
        if (MODE == 3) fetch_w1t(0, 0);
    }
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t q = tile * tc::TS + row;
        const bool valid = q < Q;
        // ---- stage the X tile: tf32 hi/lo K-major (A of GEMM1) and bf16 hi/lo MN-major (B of the gW1 GEMM) ----
        {
            float xv[16];
            // rows of X are xpitch floats apart (32, or 28 for an unpadded [Q][28] input); columns 28..31 are never read
            const float4* xs = reinterpret_cast<const float4*>(X + (size_t)(valid ? q : 0) * xpitch + half * 16);
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float4 t = (valid && !(half == 1 && g == 3)) ? xs[g] : make_float4(0.f, 0.f, 0.f, 0.f);
                xv[4 * g] = t.x; xv[4 * g + 1] = t.y; xv[4 * g + 2] = t.z; xv[4 * g + 3] = t.w;
            }
            if (half == 1) {   // column 28: carries b1 in GEMM1 and yields gb1 in the gW1 GEMM; 29..31: padding (must be finite)
                xv[12] = valid ? 1.f : 0.f;
                xv[13] = 0.f; xv[14] = 0.f; xv[15] = 0.f;
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                float4 h, l;
                h.x = umma::to_tf32_rn(xv[4 * g]);     l.x = umma::to_tf32_rn(xv[4 * g] - h.x);
                h.y = umma::to_tf32_rn(xv[4 * g + 1]); l.y = umma::to_tf32_rn(xv[4 * g + 1] - h.y);
                h.z = umma::to_tf32_rn(xv[4 * g + 2]); l.z = umma::to_tf32_rn(xv[4 * g + 2] - h.z);
                h.w = umma::to_tf32_rn(xv[4 * g + 3]); l.w = umma::to_tf32_rn(xv[4 * g + 3] - h.w);
                const uint32_t o = umma::kmajor_off(row, half * 16 + 4 * g, 32);
                *reinterpret_cast<float4*>(sm + tc::OFF_XH + o) = h;
                *reinterpret_cast<float4*>(sm + tc::OFF_XL + o) = l;
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                uint4 hi, lo;
                split_pack8(xv + 8 * g, hi, lo);
                const uint32_t o = umma::mnmajor_off_b16(half * 16 + 8 * g, row, 128);
                *reinterpret_cast<uint4*>(sm + tc::OFF_XT + o) = hi;
                *reinterpret_cast<uint4*>(sm + tc::OFF_XT + 8192 + o) = lo;
            }
        }
        const bool more_tiles = tile + gridDim.x < ntiles;
        // ---- pass 1: forward ----
        umma::fence_async_smem();       // the X tile
        umma::fence_before();
        __syncthreads();
        if (MODE < 2) {
        for (int c = 0; c < nch; ++c, ++gs) {
            const int stage = gs;
            umma::mbar_wait(barW, phW); phW ^= 1;                     // W1(c), W2 fwd image(c) have landed
            if (warp == 0) { umma::fence_after(); issue_gemm1(); umma::commit_w(barZ); }
            if (c > 0) { umma::mbar_wait(barO, phO); phO ^= 1; }      // chunk c-1's O GEMM done: A tile and its W2 buffer free
            umma::mbar_wait(barZ, phZ); phZ ^= 1;
            umma::fence_after();
            if (tid == 0 && (MODE == 0 || c + 1 < nch || more_tiles)) {
                // next stage's weights: W1 region free (GEMM1 done), other W2 buffer free (see above)
                mbar_expect_tx(barW, 32768 + 16384);
                if (c + 1 < nch) { fetch_w1(c + 1); fetch_w2(c + 1, 1, stage + 1); }
                else             { fetch_w1(0);     fetch_w2(0, MODE == 0 ? 2 : 1, stage + 1); }
            }
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t v[32];
                umma::ld32(tbase + laneblk + tc::COL_Z + half * 64 + cc * 32, v);
                umma::wait_ld();
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float a8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) a8[j] = kc_elu(__uint_as_float(v[g8 * 8 + j]));
                    uint4 ah, al;
                    split_pack8(a8, ah, al);
                    const uint32_t off = umma::mnmajor_off_b16(half * 64 + cc * 32 + g8 * 8, row, 128);
                    *reinterpret_cast<uint4*>(sm + tc::OFF_A + off) = ah;
                    *reinterpret_cast<uint4*>(sm + tc::OFF_A + 32768 + off) = al;
                }
            }
            umma::fence_async_smem();
            umma::fence_before();
            __syncthreads();
            if (warp == 0) { umma::fence_after(); issue_gemm2(c == 0, (stage & 1) ? aWB1 : aWB0); }
        }
        umma::mbar_wait(barO, phO); phO ^= 1;
        umma::fence_after();
        }
        // ---- loss and dL/do (rows of half 0 own the sample) ----
        if (half == 0) {
            float o[25], g[25];
            if (MODE < 2) {
                uint32_t v[32];
                umma::ld32(tbase + laneblk + tc::COL_O, v);
                umma::wait_ld();
#pragma unroll
                for (int c = 0; c < 25; ++c) { o[c] = __uint_as_float(v[c]) + b2[c]; g[c] = 0.f; }
            } else {
#pragma unroll
                for (int c = 0; c < 25; ++c) { o[c] = 0.f; g[c] = valid ? dOin[(size_t)q * dopitch + c] : 0.f; }
            }
            if (MODE == 1) {   // pred = PHYS (or 0) + s*o -> pred_out[Q][25], or split as ys[Q][19] | z[Q][6] (z = `partial`)
                if (valid) {
                    float pr[25];
#pragma unroll
                    for (int r = 0; r < 25; ++r)
                        pr[r] = (PHYS ? PHYS[(size_t)q * 25 + r] : 0.f) + (r < 19 ? ds : 1.f) * o[r];
                    if (partial) {
#pragma unroll
                        for (int r = 0; r < 19; ++r) pred_out[(size_t)q * 19 + r] = pr[r];
#pragma unroll
                        for (int c = 0; c < 6; ++c) partial[(size_t)q * 6 + c] = pr[19 + c];
                    } else {
#pragma unroll
                        for (int r = 0; r < 25; ++r) pred_out[(size_t)q * 25 + r] = pr[r];
                    }
                }
            }
            if (MODE == 0 && valid) {
                float pred[25], tg[25];
#pragma unroll
                for (int r = 0; r < 19; ++r) pred[r] = PHYS[(size_t)q * 25 + r] + ds * o[r];
#pragma unroll
                for (int c = 19; c < 25; ++c) pred[c] = PHYS[(size_t)q * 25 + c] + o[c];
#pragma unroll
                for (int r = 0; r < 25; ++r) tg[r] = TGT[(size_t)q * 25 + r];
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
                    const int kk = (int)(q % K);
                    const int64_t bt = q / K;
                    float* po = pred_out + (size_t)bt * 25 * K + kk;
#pragma unroll
                    for (int r = 0; r < 25; ++r) po[r * K] = pred[r];
                }
            }
#pragma unroll
            for (int c = 0; c < 25; ++c) gb2acc[c] += g[c];
            // dO^T as bf16 hi/lo, MN-major [32 outputs x 128 samples] (B operand of the gW2 GEMM)
#pragma unroll
            for (int gi = 0; gi < 4; ++gi) {
                float g8v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = gi * 8 + j;
                    g8v[j] = c < 25 ? g[c < 25 ? c : 0] : 0.f;
                }
                uint4 hi, lo;
                split_pack8(g8v, hi, lo);
                const uint32_t off = umma::mnmajor_off_b16(gi * 8, row, 128);
                *reinterpret_cast<uint4*>(sm + tc::OFF_DO + off) = hi;
                *reinterpret_cast<uint4*>(sm + tc::OFF_DO + 8192 + off) = lo;
            }
        }
        umma::fence_async_smem();       // the dO tile
        umma::fence_before();
        __syncthreads();
        // ---- pass 2: backward ----
        if (MODE != 1) {
        for (int c = 0; c < nch; ++c, ++gs) {
            const int stage = gs;
            const uint32_t aWB = (stage & 1) ? aWB1 : aWB0;
            umma::mbar_wait(barW, phW); phW ^= 1;                     // W1(c), W2 bwd image(c) have landed
            if (warp == 0) { umma::fence_after(); issue_gemm1(); issue_gemm3(aWB); umma::commit_w(barZ); }
            umma::mbar_wait(barZ, phZ); phZ ^= 1;                     // (in-order pipe: chunk c-1's gradient MMAs are done too)
            if (c > 0) { umma::mbar_wait(barG, phG); phG ^= 1; }      // A and dZ tiles free
            umma::fence_after();
            const bool prefetch = c + 1 < nch || more_tiles;
            if (tid == 0 && prefetch) {   // the other W2 buffer is free now; W1 follows once the dZ-lo MMAs of this chunk are done
                mbar_expect_tx(barW, WBYTES);
                if (c + 1 < nch) fetch_w2(c + 1, 2, stage + 1);
                else fetch_w2(0, MODE >= 2 ? 2 : 1, stage + 1);
                if (MODE == 3) fetch_w1t(c + 1 < nch ? c + 1 : 0, stage + 1);   // (the W1^T buffer of stage-1 is free: barG waited)
            }
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t v[32], d[32];
                umma::ld32(tbase + laneblk + tc::COL_Z + half * 64 + cc * 32, v);
                umma::ld32(tbase + laneblk + tc::COL_O + half * 64 + cc * 32, d);
                umma::wait_ld();
#pragma unroll
                for (int g8 = 0; g8 < 4; ++g8) {
                    float a8[8], d8[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float z = __uint_as_float(v[g8 * 8 + j]);
                        const float a = kc_elu(z);
                        a8[j] = a;
                        // ELU'(z) = 1 (z > 0) or e^z = ELU(z) + 1: no second exponential
                        d8[j] = __uint_as_float(d[g8 * 8 + j]) * (z > 0.f ? 1.f : a + 1.f);
                    }
                    uint4 ah, al, dh, dl;
                    split_pack8(d8, dh, dl);
                    const uint32_t off = umma::mnmajor_off_b16(half * 64 + cc * 32 + g8 * 8, row, 128);
                    if (MODE != 3) {   // (MODE 3 keeps the W1^T images in the activation-tile region)
                        split_pack8(a8, ah, al);
                        *reinterpret_cast<uint4*>(sm + tc::OFF_A + off) = ah;
                        *reinterpret_cast<uint4*>(sm + tc::OFF_A + 32768 + off) = al;
                    }
                    *reinterpret_cast<uint4*>(sm + tc::OFF_DZ + off) = dh;
                    *reinterpret_cast<uint4*>(sm + tc::OFF_DZ + 32768 + off) = dl;
                }
            }
            umma::fence_async_smem();
            umma::fence_before();
            __syncthreads();
            if (warp == 0) {
                umma::fence_after();
                if (MODE == 3) issue_gemm4(c == 0, aAh + ((stage & 1) ? 16384u : 0u));
                else issue_grads(c, first_tile);
                umma::mbar_wait(barL, phL);                           // dZ lo consumed: its bytes may take the next W1 chunk
                if (lane == 0 && prefetch) fetch_w1(c + 1 < nch ? c + 1 : 0);
            }
            phL ^= 1;
        }
        umma::mbar_wait(barG, phG); phG ^= 1;   // last chunk's gradient MMAs: they read X^T / dO^T which the next tile overwrites
        umma::fence_after();
        if (MODE == 3 && half == 0) {   // this tile's input gradients
            uint32_t v[32];
            umma::ld32(tbase + laneblk + tc::COL_GW1, v);
            umma::wait_ld();
            if (valid) {
#pragma unroll
                for (int k = 0; k < 28; ++k) pred_out[(size_t)q * K + k] = __uint_as_float(v[k]);
            }
        }
        if (MODE == 3) { umma::fence_before(); __syncthreads(); umma::fence_after(); }   // gX is overwritten by the next tile
        }
        first_tile = false;
    }
    if (MODE == 1 || MODE == 3) {   // no parameter gradients to write
        umma::fence_before();
        __syncthreads();
        if (warp == 0) umma::tmem_dealloc(tbase, 512);
        return;
    }
    // ---- write this CTA's partial gradients ----
    umma::fence_before();
    __syncthreads();
    umma::fence_after();
    float* out = partial + (size_t)blockIdx.x * NP;
    const int64_t ob1 = (int64_t)hidden * 28, oW2 = ob1 + hidden, ob2 = oW2 + (int64_t)25 * hidden;
    if (half == 0) {
        for (int c = 0; c < nch; ++c) {
            const int u = c * tc::HC + row;
            uint32_t v[32];
            umma::ld32(tbase + laneblk + tc::COL_GW1 + 32 * c, v);
            umma::wait_ld();
            if (u < hidden) {
#pragma unroll
                for (int k = 0; k < 28; ++k) out[(size_t)u * 28 + k] = __uint_as_float(v[k]);
                out[ob1 + u] = __uint_as_float(v[28]);
            }
            umma::ld32(tbase + laneblk + tc::COL_GW2 + 32 * c, v);
            umma::wait_ld();
            if (u < hidden) {
#pragma unroll
                for (int co = 0; co < 25; ++co) out[oW2 + (size_t)co * hidden + u] = __uint_as_float(v[co]);
            }
        }
    }
    // gb2 and loss: block reductions (fixed order: deterministic)
#pragma unroll
    for (int c = 0; c < 25; ++c) {
        float s = gb2acc[c];
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_down_sync(0xffffffffu, s, o2);
        if (lane == 0 && warp < 4) redb[warp * 25 + c] = s;
    }
    {
        double s = lossacc;
#pragma unroll
        for (int o2 = 16; o2 > 0; o2 >>= 1) s += __shfl_down_sync(0xffffffffu, s, o2);
        if (lane == 0) redd[warp] = s;
    }
    umma::fence_before();
    __syncthreads();
    if (tid < 25) out[ob2 + tid] = redb[tid] + redb[25 + tid] + redb[50 + tid] + redb[75 + tid];
    if (tid == 0) loss_part[blockIdx.x] = redd[0] + redd[1] + redd[2] + redd[3];
    if (warp == 0) umma::tmem_dealloc(tbase, 512);
}

// Host side: called by kc_train_step (kc_train.cu) when the shape is eligible.  W1hl: nch*2*4096 floats, W2b: nch*32768
// bytes (workspace), partial: grid*NP floats, loss_part: grid doubles.
int kc_tc_launch_mode(int mode, const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                      const float* TGT, float* W1hl, float* W2c_, float* partial, int64_t NP, double* loss_part,
                      float* pred_out, const float* dOin, int grid, cudaStream_t st, int xpitch = 32, int dopitch = 32);
int kc_train_tc_grid(int64_t Q) {
    const int64_t ntiles = (Q + tc::TS - 1) / tc::TS;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int)(ntiles < sms ? (ntiles > 0 ? ntiles : 1) : sms);
}

int kc_train_tc_launch(const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                       const float* TGT, float* W1hl, float* W2c_, float* partial, int64_t NP, double* loss_part,
                       float* pred_out, int grid, cudaStream_t st) {
    return kc_tc_launch_mode(0, mlp, ds, Q, T_, K, X, PHYS, TGT, W1hl, W2c_, partial, NP, loss_part, pred_out, nullptr,
                             grid, st, 32);
}

// mode 0: training step; 1: forward only (pred_out[Q][25] = PHYS + s*o); 2: parameter gradients from (X, dOin) samples
int kc_tc_launch_mode(int mode, const kc_mlp* mlp, float ds, int64_t Q, int T_, int K, const float* X, const float* PHYS,
                      const float* TGT, float* W1hl, float* W2c_, float* partial, int64_t NP, double* loss_part,
                      float* pred_out, const float* dOin, int grid, cudaStream_t st, int xpitch, int dopitch) {
    const int nch = (mlp->hidden + tc::HC - 1) / tc::HC;
    unsigned char* W2b = reinterpret_cast<unsigned char*>(W2c_);
    kc_tc_prep_weights_kernel<<<dim3(nch, 8), 256, 0, st>>>((const float*)mlp->W1, (const float*)mlp->b1,
                                                          (const float*)mlp->W2, mlp->hidden, W1hl, W2b);
    KC_CHECK_LAUNCH("kc_tc_prep_weights_kernel");
#define KC_TC_GO(M)                                                                                                    \
    do {                                                                                                               \
        cudaFuncSetAttribute(kc_train_tc_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES);      \
        kc_train_tc_kernel<M><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(mlp->hidden, nch, W1hl, W2b,                  \
            (const float*)mlp->b2, ds, Q, T_, K, X, PHYS, TGT, partial, NP, loss_part, pred_out, dOin, xpitch, dopitch);\
    } while (0)
    if (mode == 0) KC_TC_GO(0); else if (mode == 1) KC_TC_GO(1); else if (mode == 2) KC_TC_GO(2); else KC_TC_GO(3);
#undef KC_TC_GO
    KC_CHECK_LAUNCH("kc_train_tc_kernel");
    return KC_OK;
}
