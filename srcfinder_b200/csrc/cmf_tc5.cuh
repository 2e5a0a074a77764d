// tcgen05 / TMEM primitives shared by the 5th-generation tensor-core kernels (k_screen5.cu: TF32 screening pass,
// k_gram8.cu: integer Gram pass).  Descriptor encodings follow the public CUTLASS layout
// (cute/arch/mma_sm100_desc.hpp: UMMA::InstrDescriptor, UMMA::SmemDescriptor).
#pragma once

#include "cmf_common.cuh"

namespace cmf {

// ------------------------------------------------------------------ tcgen05 primitives
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// mbarrier wait that traps instead of hanging the GPU when a protocol error leaves it unsignalled
__device__ __forceinline__ void mbar_wait_guard(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

// D[tmem] (+)= A[tmem] . B[smem]^T, kind::tf32, issued by one thread for the CTA
__device__ __forceinline__ void mma_ts_tf32(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// instruction descriptor (cute::UMMA::InstrDescriptor): F32 accumulate, TF32 x TF32, K-major A and B, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// the same with the descriptor given as its two 32-bit words: issue loops that step through the k-steps of an operand
// add to the low word only (the start address field, 16-byte units), which keeps the arithmetic on 32-bit uniform
// registers instead of 64-bit values the compiler hoists and spills
__device__ __forceinline__ void mma_ts_tf32_w(uint32_t d_tmem, uint32_t a_tmem, uint32_t desc_lo, uint32_t desc_hi,
                                              uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n.reg .b64 bd;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "mov.b64 bd, {%2, %3};\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], bd, %4, p;\n}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(desc_lo), "r"(desc_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor), K-major, no swizzle: element (row, 16-byte
// K chunk c) lives at start + (row % 8) * 16 + (row / 8) * SBO + c * LBO
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t lo = ((addr >> 4) & 0x3fffu) | (((lbo_bytes >> 4) & 0x3fffu) << 16);
    const uint32_t hi = ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14);       // version 1 (Blackwell)
    return ((uint64_t)hi << 32) | lo;
}

__device__ __forceinline__ void tmem_ld8(uint32_t addr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(addr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t addr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(addr));
}

__device__ __forceinline__ void tmem_st8(uint32_t addr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(addr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// FP32 -> TF32 on the integer pipe (cvt.rna.tf32.f32 issues on the 16-lane XU pipe and was the busiest
// pipe of the first version): round-half-away on the 13 dropped bits for the hi part, truncation for the lo
// part (its error is 2^-22 of the value).  Inputs are finite.
__device__ __forceinline__ uint32_t tf32_round(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t tf32_trunc(float x) { return __float_as_uint(x) & 0xffffe000u; }


// Packed FP32 pairs (Blackwell FFMA2 / FMUL2 / FADD2): a three-register FFMA issues every second cycle per scheduler,
// the packed forms do two lanes' worth of arithmetic in the same slot.  Pairs live in 64-bit registers {lo, hi}.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ uint64_t f2_pack_u(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// One lane of a CONVERGED warp (elect.sync).  The MMA issuer warps run their loops with all 32 lanes and issue under
// this predicate: with uniform control flow the descriptors stay in uniform registers, whereas a loop entered by
// lane 0 alone makes the compiler wrap every tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY sequence
// (measured: ~108 cycles per MMA issued against ~45 cycles of tensor work, profiles/r02_screen5_issue.md).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n.reg .pred P1;\n"
        "elect.sync _|P1, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P1;\n}\n"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] . B[smem]^T, kind::i8 (signed 8-bit operands, 32-bit integer accumulators), one thread issues
__device__ __forceinline__ void mma_ss_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// instruction descriptor: S32 accumulate, S8 x S8, K-major A and B, M = 128
__host__ __device__ constexpr uint32_t idesc_i8(int n) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

}  // namespace cmf
