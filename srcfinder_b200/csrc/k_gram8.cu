// Wide-window statistics pass on the 5th-generation tensor cores: the centred Gram matrix of every column
// (numpy.cov in the reference, cmf/robust_mf.py:52-70, a 416 x 416 contraction over up to 20 000 lines for the
// `-R` window, :186-187) as an EXACT integer contraction on tcgen05.mma kind::i8 with TMEM accumulators.
//
// Precision scheme (the reference accumulates in FP64; tcgen05 has no FP64 kind and FP32-class accumulation
// misses the tolerance): the centred radiances are fixed-point numbers
//     q = rint((x - ctr) 2^(30 - e)),  |q| <= 2^30          (ctr, e per column and band; wide_mean_kernel)
// which is exact for every FP32 radiance above 2^-7 of the band's range, written in balanced base-256 digits
//     q = d0 2^24 + d1 2^16 + d2 2^8 + d3,   d in [-128, 127]              (wide_pack_kernel)
// so that  sum_l q_i q_j = sum_{s,t} 2^(8 (6 - s - t)) sum_l d_s,i d_t,j  and every inner sum is an int8 x int8
// contraction with 32-bit accumulation that cannot overflow below 2^17 lines.  The 16 digit pairs of a 32 x 32
// band block are one 128 x 128 accumulator tile (row = digit * 32 + band); the digits are recombined in 64-bit
// integers (over t, exact) and FP64 (over s, smallest terms first) in the epilogue.
//
// Mapping: CTA = (column, 32-band row block i, group of <= 4 column blocks j <= i): the A tile (block i) and the
// B tiles (blocks j) of a 64-line step are ready-made shared-memory images in global memory (canonical K-major
// no-swizzle core matrices), so the producer moves them with two bulk copies per step; warp 0 = bulk-copy
// producer, warp 1 = MMA issuer (one thread, SS form: both operands from shared memory), warps 2-5 = epilogue
// (tcgen05.ld, digit recombination, symmetric store).  4 stages x 40 KB, all 512 TMEM columns when nj = 4.
#include "cmf_common.cuh"
#include "cmf_internal.h"
#include "cmf_tc5.cuh"

namespace cmf {

namespace {

constexpr int kG8Threads = 192;
constexpr int kG8Stages = 4;
constexpr int kG8MaxNJ = 4;
constexpr uint32_t kG8Tile = 8192;        // [4 chunks of 16 lines][128 rows][16 B]
constexpr uint32_t kG8Stage = (1 + kG8MaxNJ) * kG8Tile;

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

__global__ void __launch_bounds__(kG8Threads, 1)
    wide_gram8_kernel(const int8_t* __restrict__ img, int nkb, int nrb, int DP, double* __restrict__ gram) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + kG8Stages * kG8Stage);   // full[NS], empty[NS], done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kG8Stages + 1);
    const int s = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // work item -> (row block i, first column block j0, nj column blocks)
    int rbi = 0, w = blockIdx.x;
    for (;; ++rbi) {
        const int ng = (rbi + 1 + kG8MaxNJ - 1) / kG8MaxNJ;
        if (w < ng) break;
        w -= ng;
    }
    const int j0 = w * kG8MaxNJ, nj = min(kG8MaxNJ, rbi + 1 - j0);

    if (tid == 0) {
        for (int i = 0; i < kG8Stages; ++i) { mbar_init(&bars[i], 1); mbar_init(&bars[kG8Stages + i], 1); }
        mbar_init(&bars[2 * kG8Stages], 1);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            const int8_t* src = img + (long long)s * nkb * nrb * kG8Tile;
            for (int kb = 0; kb < nkb; ++kb) {
                const int slot = kb % kG8Stages, use = kb / kG8Stages;
                if (use > 0) mbar_wait_guard(&bars[kG8Stages + slot], (uint32_t)((use - 1) & 1));
                unsigned char* dst = smem_raw + slot * kG8Stage;
                const int8_t* blk = src + (long long)kb * nrb * kG8Tile;
                mbar_expect_tx(&bars[slot], (uint32_t)(1 + nj) * kG8Tile);
                bulk_g2s(dst, blk + (long long)rbi * kG8Tile, kG8Tile, &bars[slot]);
                bulk_g2s(dst + kG8Tile, blk + (long long)j0 * kG8Tile, (uint32_t)nj * kG8Tile, &bars[slot]);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // MMA issuer: the whole warp runs the loop, one elected lane issues (see elect_one())
        const uint32_t sbase = smem_u32(smem_raw);
        constexpr uint32_t lbo = 128 * 16, sbo = 128;
        const uint32_t id = idesc_i8(128);
        for (int kb = 0; kb < nkb; ++kb) {
            const int slot = kb % kG8Stages, use = kb / kG8Stages;
            mbar_wait_guard(&bars[slot], (uint32_t)(use & 1));
            tc_fence_after();
            const uint32_t st = sbase + slot * kG8Stage;
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {                 // 32 lines = two 16-line chunks per MMA
                const uint64_t da = smem_desc(st + ks * 2 * lbo, lbo, sbo);
#pragma unroll
                for (int jj = 0; jj < kG8MaxNJ; ++jj) {
                    const uint64_t db = smem_desc(st + (1 + jj) * kG8Tile + ks * 2 * lbo, lbo, sbo);
                    if (jj < nj && elect_one()) mma_ss_i8(tmem + 128 * jj, da, db, id, (kb > 0 || ks > 0) ? 1u : 0u);
                }
            }
            if (elect_one()) tc_commit(&bars[kG8Stages + slot]);     // the stage is free once these MMAs have read it
        }
        if (elect_one()) tc_commit(&bars[2 * kG8Stages]);
        __syncwarp();
    } else {
        // ---------------- epilogue: TMEM lane = digit s * 32 + band i, column = 128 jj + digit t * 32 + band j
        const int dg = warp & 3;                                  // the TMEM lane quarter this warp may read
        const uint32_t tl = tmem + ((uint32_t)(32 * dg) << 16);
        const int et = (warp - 2) * 32 + lane;                    // 0..127
        mbar_wait_guard(&bars[2 * kG8Stages], 0);
        tc_fence_after();
        double* xch = reinterpret_cast<double*>(smem_raw);        // [4 digits][32][33], the stages are idle now
        double* G = gram + (long long)s * DP * DP;
        const double wgt = (double)(1 << (8 * (3 - dg)));
        for (int jj = 0; jj < nj; ++jj) {
#pragma unroll 1
            for (int h = 0; h < 2; ++h) {
                uint32_t v0[16], v1[16], v2[16], v3[16];
                tmem_ld16(tl + 128 * jj + 0 * 32 + 16 * h, v0);
                tmem_ld16(tl + 128 * jj + 1 * 32 + 16 * h, v1);
                tmem_ld16(tl + 128 * jj + 2 * 32 + 16 * h, v2);
                tmem_ld16(tl + 128 * jj + 3 * 32 + 16 * h, v3);
                tc_wait_ld();
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const long long V = ((long long)(int)v0[e] << 24) + ((long long)(int)v1[e] << 16) +
                                        ((long long)(int)v2[e] << 8) + (long long)(int)v3[e];
                    xch[(dg * 32 + lane) * 33 + 16 * h + e] = __dmul_rn((double)V, wgt);
                }
            }
            epi_bar_sync();
            const int rbj = j0 + jj;
#pragma unroll 1
            for (int k = 0; k < 8; ++k) {
                const int idx = et + 128 * k, bi = idx >> 5, bj = idx & 31;
                const double x0 = xch[(0 * 32 + bi) * 33 + bj], x1 = xch[(1 * 32 + bi) * 33 + bj],
                             x2 = xch[(2 * 32 + bi) * 33 + bj], x3 = xch[(3 * 32 + bi) * 33 + bj];
                const double g = __dadd_rn(__dadd_rn(__dadd_rn(x3, x2), x1), x0);
                const int r = rbi * 32 + bi, c = rbj * 32 + bj;
                if (r < DP && c < DP && (rbj != rbi || bj <= bi)) {
                    G[(long long)r * DP + c] = g;
                    G[(long long)c * DP + r] = g;
                }
            }
            epi_bar_sync();
        }
        tc_fence_before();
    }
    __syncthreads();
    tc_fence_after();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

}  // namespace

size_t wide_img_bytes(const Dims& d) {
    const size_t nrb = (d.D + 31) / 32, nkb = (d.L + 63) / 64;
    return (size_t)d.S * nkb * nrb * kG8Tile;
}

void launch_wide_gram8(const Dims& d, const int8_t* img, double* gram, cudaStream_t st) {
    const int nrb = (d.D + 31) / 32, nkb = (d.L + 63) / 64;
    int items = 0;
    for (int i = 0; i < nrb; ++i) items += (i + 1 + kG8MaxNJ - 1) / kG8MaxNJ;
    const size_t smem = (size_t)kG8Stages * kG8Stage + (2 * kG8Stages + 1) * sizeof(uint64_t) + 16;
    cudaFuncSetAttribute(wide_gram8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(items, d.S);
    wide_gram8_kernel<<<grid, kG8Threads, smem, st>>>(img, nkb, nrb, d.DP, gram);
}

}  // namespace cmf
