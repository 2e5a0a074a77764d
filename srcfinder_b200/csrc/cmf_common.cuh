// Shared device helpers for the B200 (sm_100a) columnwise matched filter kernels.
//
// Hardware notes that shape this file:
//  * FP64 tensor work on sm_100a is DMMA.8x8x4 (every larger mma.sync f64 shape is lowered
//    to it by ptxas), so mma884() is the only tensor primitive used for FP64 contractions.
//  * tcgen05 has no FP64 kind; SURVEY.md 7.3(1) shows FP32-class accumulation misses the
//    parity tolerance, so the statistics / LOO contractions stay on the FP64 tensor path.
//  * Bulk async copies (cp.async.bulk, SASS UBLKCP) + mbarrier feed the per-warp tile rings.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#define CMF_WARP 32

namespace cmf {

// ------------------------------------------------------------------ FP64 tensor MMA
// D(8x8) += A(8x4) * B(4x8).  Fragment layout (PTX ISA, mma.m8n8k4 .f64):
//   A: lane holds A[row = lane/4][col = lane%4]
//   B: lane holds B[row = lane%4][col = lane/4]
//   C: lane holds C[row = lane/4][col = 2*(lane%4) + {0,1}]
__device__ __forceinline__ void mma884(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// ------------------------------------------------------------------ TF32 warp MMA (screening pass)
// D(16x8) += A(16x8) * B(8x8), FP32 accumulate.  Fragments (PTX ISA, mma.m16n8k8 .tf32), g = lane/4, q = lane%4:
//   A: a0 (g, q)  a1 (g+8, q)  a2 (g, q+4)  a3 (g+8, q+4)        B: b0 (k=q, n=g)  b1 (k=q+4, n=g)
//   C: c0 (g, 2q)  c1 (g, 2q+1)  c2 (g+8, 2q)  c3 (g+8, 2q+1)
__device__ __forceinline__ void mma_tf32(float (&c)[4], const float4& a, uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
          "r"(__float_as_uint(a.w)), "r"(b0), "r"(b1));
}

// round-to-nearest FP32 -> TF32 (10-bit mantissa), returned in an FP32 container
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// ------------------------------------------------------------------ mbarrier + bulk copy
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}

__device__ __forceinline__ void fence_mbar_init() {
    // make barrier initialisation visible to the async proxy before the first bulk copy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP).  16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                         uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ misc
__device__ __forceinline__ bool pixel_value_ok(float x) {
    // the reference's useidx rule (cmf/robust_mf.py:282): finite and not negative (-0.0 passes)
    return !(x < 0.0f) && (fabsf(x) <= 3.402823466e38f);
}

// log(det G) as the reference sees it (cmf/robust_mf.py:111-117): G_det is a double, so it is 0 when the
// determinant is below half the smallest subnormal (2^-1075: the alpha is skipped), +inf above DBL_MAX
// (nll = inf), and carries only a few bits in the subnormal range (log of the rounded value is what enters nll).
// Returns -inf / +inf for the two excluded cases.
__device__ __forceinline__ double det_roundtrip(double logdet) {
    if (logdet > 709.782712893384) return __longlong_as_double(0x7ff0000000000000LL);
    if (logdet < -708.3964185322641) {               // below 2^-1022: the subnormal grid has spacing 2^-1074
        const double units = rint(exp(logdet + 744.4400719213812));   // det in units of 2^-1074, ties to even
        return units >= 1.0 ? log(units) - 744.4400719213812 : -__longlong_as_double(0x7ff0000000000000LL);
    }
    return logdet;
}

__device__ __forceinline__ float2 ldg_nc_f2(const float* p) {
    float2 v;
    asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}

__device__ __forceinline__ float ldg_nc_f1(const float* p) {
    float v;
    asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    return __shfl_xor_sync(0xffffffffu, v, m);
}

}  // namespace cmf
