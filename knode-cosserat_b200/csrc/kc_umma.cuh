// kc_umma.cuh — thin inline-PTX wrappers for the Blackwell tensor-core path (tcgen05 + TMEM), sm_100a only.
// Operands are staged in shared memory in the canonical K-major, no-swizzle ("interleave") layout:
//   8-row x 16-byte core matrices; element (r, k) of a tile with K 4-byte elements per row lives at
//   ((r/8) * (K/4) + k/4) * 128 + (r%8) * 16 + (k%4) * 4 bytes,  i.e.  LBO = 128 B (next core matrix along K),
//   SBO = (K/4) * 128 B (next 8-row group).  One kind::tf32 instruction consumes K = 8 (two core matrices).
#pragma once
#include <stdint.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// byte offset of element (r, k) in a K-major interleaved tile with KT 4-byte elements per row
__device__ __forceinline__ uint32_t kmajor_off(int r, int k, int KT) {
    return (uint32_t)(((r >> 3) * (KT >> 2) + (k >> 2)) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version
    return d;
}

// instruction descriptor: D fp32, A/B tf32, M x N tile; a_mn / b_mn = 1 selects MN-major for that operand
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// MN-major, no swizzle, 4-byte elements: 8 (k) x 16-byte (4 mn elements) core matrices.  Element (mn, k) of a tile with
// KT k-values lives at (mn/4) * SBO + (k/8) * 128 + (k%8) * 16 + (mn%4) * 4 with LBO = 128 B (next group of 8 k),
// SBO = (KT/8) * 128 B (next group of 4 mn).  One kind::tf32 instruction consumes K = 8 = one k-group.
__device__ __forceinline__ uint32_t mnmajor_off(int mn, int k, int KT) {
    return (uint32_t)((mn >> 2) * ((KT >> 3) * 128) + (k >> 3) * 128 + (k & 7) * 16 + (mn & 3) * 4);
}

// bf16 operands (kind::f16), D fp32
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn = 0, int b_mn = 0) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// MN-major, no swizzle, 2-byte elements: 8 (k) x 16-byte (8 mn elements) core matrices.  Element (mn, k) of a tile with
// KT k-values: (mn/8) * SBO + (k/8) * 128 + (k%8) * 16 + (mn%8) * 2, LBO = 128 B, SBO = (KT/8) * 128 B.
// One kind::f16 instruction consumes K = 16 = two k-groups (256 B).
__device__ __forceinline__ uint32_t mnmajor_off_b16(int mn, int k, int KT) {
    return (uint32_t)((mn >> 3) * ((KT >> 3) * 128) + (k >> 3) * 128 + (k & 7) * 16 + (mn & 7) * 2);
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {  // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// make generic-proxy shared-memory writes visible to the async proxy (the tensor core reads operands through it)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// non-blocking phase test: true once the phase with this parity has completed
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T, issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// TMEM -> registers: this warp's 32 lanes (lane field of taddr must be 32*(warp%4)), 32 consecutive columns
__device__ __forceinline__ void ld32(uint32_t taddr, uint32_t v[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float to_tf32_rn(float x) {  // round-to-nearest tf32 (the MMA itself truncates)
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}


// ---- A operand from TMEM (".ts" form) ------------------------------------------------------------------------------------
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (bf16 operands): A is M = 128 rows x 16 k-values per instruction, row r in
// TMEM lane r, the k-values packed two per 32-bit column (even k in the low half), i.e. 8 columns per instruction
// (CuTe: SM100_MMA_F16BF16_TS, A fragment = tmem_frg_1sm<bf16, bf16>).  The activation tile of the KNODE MLP is written
// there by the epilogue threads with tcgen05.st and never touches shared memory.
__device__ __forceinline__ void mma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: this warp's 32 lanes, 16 consecutive columns
__device__ __forceinline__ void st16(uint32_t taddr, const uint32_t v[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
          "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void ld16(uint32_t taddr, uint32_t v[16]) {   // TMEM -> registers, 16 consecutive columns
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st8(uint32_t taddr, const uint32_t v[8]) {     // registers -> TMEM, 8 consecutive columns
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// plain arrive (count 1) on a CTA-local mbarrier, release semantics
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- warp-uniform issue: every lane of the MMA warp executes these (so descriptors and addresses stay in uniform
// registers), the instruction itself is predicated on elect.sync.  Issuing from a single-lane divergent region instead
// makes the compiler wrap every tcgen05.mma in an R2UR / ELECT loop (~15 instructions and several dependent latencies per MMA).
__device__ __forceinline__ void mma_bf16_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_tf32_w(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_bf16_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred e;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "@e mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
// ---- batched warp-uniform issue: ONE elect.sync for a whole 3-pass (hi*hi, lo*hi, hi*lo) product, so the per-instruction
// ELECT / VOTEU / R2UR chain (~11 instructions, 40-60 cycles per MMA when issued one by one) is paid once per batch ----
// TS form, 3 passes x 4 k-steps of 16: A in TMEM as packed bf16 pairs, laid out as two 32-column blocks (one per epilogue
// thread group), each block = [hi 16 columns | lo 16 columns]; k-step kk reads 8 columns at block (kk >> 1), offset (kk & 1) * 8.
// B: shared-memory descriptors of the hi and lo images, advancing 256 B (16 descriptor units) per k-step.
// O1..O3: column offsets of k-steps 1..3, LO: offset of the lo half (defaults: the layout described above; the transposed
// backward of kc_train_tc3.cu uses [hi 8 | lo 8] per k-step: O = 16, 32, 48, LO = 8).
// BSTEP: descriptor increment (bytes >> 4) of one k-step of the B images.
template <int O1 = 8, int O2 = 32, int O3 = 40, int LO = 16, int BSTEP = 16>
__device__ __forceinline__ void mma_bf16_ts_3x4_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bh, uint64_t bl, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, e, t;\n\t.reg .b32 a<4>, l<4>;\n\t.reg .b64 h<4>, g<4>;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.u32 t, 0, 0;\n\t"
        "mov.b32 a0, %1;\n\tadd.u32 a1, %1, %6;\n\tadd.u32 a2, %1, %7;\n\tadd.u32 a3, %1, %8;\n\t"
        "add.u32 l0, %1, %9;\n\tadd.u32 l1, %1, %10;\n\tadd.u32 l2, %1, %11;\n\tadd.u32 l3, %1, %12;\n\t"
        "mov.b64 h0, %2;\n\tadd.u64 h1, %2, %13;\n\tadd.u64 h2, %2, %14;\n\tadd.u64 h3, %2, %15;\n\t"
        "mov.b64 g0, %3;\n\tadd.u64 g1, %3, %13;\n\tadd.u64 g2, %3, %14;\n\tadd.u64 g3, %3, %15;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a0], h0, %4, p;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], h1, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], h2, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], h3, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [l0], h0, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [l1], h1, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [l2], h2, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [l3], h3, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a0], g0, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], g1, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], g2, %4, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], g3, %4, t;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bh), "l"(bl), "r"(idesc), "r"(accumulate), "n"(O1), "n"(O2), "n"(O3), "n"(LO), "n"(O1 + LO),
          "n"(O2 + LO), "n"(O3 + LO), "n"(BSTEP), "n"(2 * BSTEP), "n"(3 * BSTEP) : "memory");
}
// SS form, 3 passes x 2 k-steps (K = 32): D = A B^T from scratch (the first instruction overwrites D); astep / bstep are the
// descriptor increments (bytes >> 4) of one k-step of 16.
__device__ __forceinline__ void mma_bf16_ss_3x2_w(uint32_t d_tmem, uint64_t ah, uint64_t al, uint32_t astep, uint64_t bh, uint64_t bl,
                                                  uint32_t bstep, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred e, t, f;\n\t.reg .b64 ah1, al1, bh1, bl1, sa, sb;\n\t"
        "elect.sync _|e, 0xffffffff;\n\t"
        "setp.eq.u32 t, 0, 0;\n\tsetp.ne.u32 f, 0, 0;\n\t"
        "cvt.u64.u32 sa, %3;\n\tcvt.u64.u32 sb, %6;\n\t"
        "add.u64 ah1, %1, sa;\n\tadd.u64 al1, %2, sa;\n\tadd.u64 bh1, %4, sb;\n\tadd.u64 bl1, %5, sb;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %4, %7, f;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah1, bh1, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %4, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], al1, bh1, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %5, %7, t;\n\t"
        "@e tcgen05.mma.cta_group::1.kind::f16 [%0], ah1, bl1, %7, t;\n\t}"
        ::"r"(d_tmem), "l"(ah), "l"(al), "r"(astep), "l"(bh), "l"(bl), "r"(bstep), "r"(idesc) : "memory");
}
// K-major, no swizzle, 2-byte elements: element (r, k) of a tile with KT k-values per row (8 x 16-byte core matrices:
// LBO = 128 B to the next 8 k-values, SBO = (KT/8) * 128 B to the next 8 rows); one kind::f16 instruction consumes
// K = 16 = two core matrices (256 B)
__device__ __forceinline__ uint32_t kmajor_off_b16(int r, int k, int KT) {
    return (uint32_t)(((r >> 3) * (KT >> 3) + (k >> 3)) * 128 + (r & 7) * 16 + (k & 7) * 2);
}

}  // namespace umma
