// HBM-bound streaming kernels of the columnwise matched filter (sm_100a):
//   K0  repack_pair_kernel / repack_pipe_kernel   BIL active slab [L][D][S] -> column-major [S][L][DP] + valid mask
//       + column sums (8-byte / 4-byte copy variants)
//   K0b mean_kernel     per-column mean / valid count from the K0 partials (fixed order)
//   K5  score_kernel    the scoring pass: one read of the slab, fused mask + mean removal + dot + scale
//   K6  colstats_kernel per-column npix / mean / std of the written scores
//
// Reference steps replaced (cmf/robust_mf.py): :282,:298-304 (valid mask + gather), :347 (mean),
// :376-386 (matched filter), :388-392 (column statistics).
#include <stdio.h>
#include <stdlib.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

// ---------------------------------------------------------------------------------------- K0 (4-byte copies)
// BIL active slab -> column-major copy + valid mask + column sums as a software pipeline (used when the rows are
// not 8-byte aligned, i.e. odd sample counts; repack_pair_kernel below is the kernel otherwise).  One CTA = CG columns x a line range, tiles of LT lines.
// Every radiance is moved global -> shared by a 4-byte LDGSTS (cp.async) straight into the TRANSPOSED tile
// [column][line][band], so the loads of tile i+1 are in flight while tile i is checked and written, and no
// register is held across the memory latency.  The write side is then pure 16-byte traffic: thread <->
// (column, 4 bands) reads LT float4 from shared memory, flags bad pixels, and stores LT float4 to the column's
// contiguous [LT][DP] block of xt; its four FP64 band sums live in registers for the whole line range.
// Shared layout: column stride CS = LT*DP + 4 floats (== 4 mod 32 banks) and the low two bits of the element
// index XORed with a per-column key, so that the 32 lanes of one LDGSTS (CG columns x 32/CG bands) hit 32
// different banks while the float4 reads stay aligned; the key is undone in registers.
__device__ __forceinline__ void cp_async4(float* dst_smem, const float* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int CG>
__device__ __forceinline__ int repack_key(int c) {
    return CG == 16 ? ((c >> 3) & 1) << 1 : (c >> 3) & 3;
}

// resident CTAs per SM the register allocation is held to: what shared memory (227 KB) and threads (2048) allow
constexpr int repack_pipe_minb(int nt, int cg, int lt, int ns) {
    const int by_smem = (227 * 1024) / (ns * cg * (lt * 8 * nt + 4) * 4 + 1024 + 256);
    const int by_thr = 2048 / (cg * 2 * nt);
    const int m = by_smem < by_thr ? by_smem : by_thr;
    return m < 1 ? 1 : (m > 4 ? 4 : m);
}

// validity of four radiances at once: as unsigned integers every finite value >= +0 is below 0x7f800000, so one
// 3-input max + compare decides the common case; -0.0 (valid: not < 0) is the only value the fast test rejects
// wrongly, and it is re-examined exactly (the reference's rule, cmf/robust_mf.py:282)
__device__ __forceinline__ bool quad_ok(const float4& x) {
    const uint32_t a = __float_as_uint(x.x), b = __float_as_uint(x.y), c = __float_as_uint(x.z),
                   d = __float_as_uint(x.w);
    const uint32_t m = max(max(a, b), max(c, d));
    if (m < 0x7f800000u) return true;
    return pixel_value_ok(x.x) && pixel_value_ok(x.y) && pixel_value_ok(x.z) && pixel_value_ok(x.w);
}

template <int KEY>
__device__ __forceinline__ float4 repack_unkey(const float4& x) {
    if (KEY == 0) return x;
    if (KEY == 1) return make_float4(x.y, x.x, x.w, x.z);
    if (KEY == 2) return make_float4(x.z, x.w, x.x, x.y);
    return make_float4(x.w, x.z, x.y, x.x);
}

// pass A of one tile for a thread whose column has key KEY: shared -> registers, flag bad pixels
template <int KEY, int LT, int DP, int CG>
__device__ __forceinline__ void repack_pass_a(const float* src, int nl, int c, uint8_t* bd, float4 (&v)[LT]) {
#pragma unroll
    for (int l = 0; l < LT; ++l) {
        if (l < nl) {
            const float4 x = repack_unkey<KEY>(*reinterpret_cast<const float4*>(src + l * DP));
            v[l] = x;
            if (!quad_ok(x)) bd[l * CG + c] = 1;
        }
    }
}

template <int NT, int CG, int LT, int NS>
__global__ void __launch_bounds__(CG * 2 * NT, repack_pipe_minb(NT, CG, LT, NS)) repack_pipe_kernel(
    const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S, int D,
    float* __restrict__ xt, uint8_t* __restrict__ mask, double* __restrict__ colsum_part,
    int* __restrict__ colcnt_part, int lines_per_split, int line_base, int line_limit, int split_base,
    const uint8_t* __restrict__ sel, int write_mask, const int32_t* __restrict__ rowidx) {
    constexpr int DP = 8 * NT, Q = 2 * NT, NTH = CG * Q, CS = LT * DP + 4;
    extern __shared__ __align__(16) float tile[];   // [NS][CG][CS]: ring of NS tiles
    __shared__ uint8_t bad[2][LT * CG];

    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * CG;
    const int split = split_base + blockIdx.y;
    const int l_begin = line_base + blockIdx.y * lines_per_split;
    const int l_end = min(line_limit, l_begin + lines_per_split);
    // load side: lane <-> column (coalesced runs of CG floats per (line, band) row); a thread owns bands
    // r_ld + k*Q, k = 0..3, of every line of the tile
    const int c_ld = tid % CG, r_ld = tid / CG;
    const bool ld_ok = s0 + c_ld < S;
    const int key_ld = repack_key<CG>(c_ld);
    uint32_t sdst[4];
    long long goff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = r_ld + k * Q;
        sdst[k] = smem_u32(tile + (size_t)c_ld * CS + (b ^ key_ld));
        goff[k] = (ld_ok && b < D) ? (long long)b * band_pitch : -1;
    }
    // store side: thread <-> (column, 4 consecutive bands)
    const int c = tid / Q, q = tid % Q;
    const bool col_ok = s0 + c < S;
    const int key = repack_key<CG>(c);

    for (int i = tid; i < NS * CG * CS; i += NTH) tile[i] = 0.0f;      // padded bands stay zero for good
    for (int i = tid; i < 2 * LT * CG; i += NTH) (&bad[0][0])[i] = 0;
    __syncthreads();

    auto issue = [&](int buf, int l0, int nl) {
        const float* src = slab + (long long)l0 * line_pitch + s0 + c_ld;
        const uint32_t boff = (uint32_t)(buf * CG * CS * sizeof(float));
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            if (l < nl) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (goff[k] >= 0)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sdst[k] + boff +
                                                                                      (uint32_t)(l * DP * 4)),
                                     "l"(src + goff[k])
                                     : "memory");
            }
            src += line_pitch;
        }
        cp_async_commit();
    };

    double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0, acc3 = 0.0;
    int cnt = 0;
    const int ntiles = (l_end - l_begin + LT - 1) / LT;
    // ring of NS tiles: tiles it .. it+NS-1 are in flight while tile `it` is consumed.  One commit group per
    // tile slot (empty past the end), so "all but the NS-1 newest groups complete" always means tile `it`.
#pragma unroll
    for (int p = 0; p < NS - 1; ++p) {
        if (p < ntiles) issue(p, l_begin + p * LT, min(LT, l_end - (l_begin + p * LT)));
        else cp_async_commit();
    }
    int buf = 0, nbuf = NS - 1;
    for (int it = 0; it < ntiles; ++it) {
        const int l0 = l_begin + it * LT;
        const int nl = min(LT, l_end - l0);
        if (it + NS - 1 < ntiles) issue(nbuf, l0 + (NS - 1) * LT, min(LT, l_end - l0 - (NS - 1) * LT));
        else cp_async_commit();
        cp_async_wait<NS - 1>();
        __syncthreads();                                   // tile `it` has landed for every thread
        uint8_t* bd = bad[it & 1];
        float4 v[LT];
        if (col_ok) {
            const float* src = tile + (size_t)(buf * CG + c) * CS + 4 * q;
            // the key changes every 8 columns, so this branch is uniform for all but a few warps
            switch (key) {
                case 0: repack_pass_a<0, LT, DP, CG>(src, nl, c, bd, v); break;
                case 1: repack_pass_a<1, LT, DP, CG>(src, nl, c, bd, v); break;
                case 2: repack_pass_a<2, LT, DP, CG>(src, nl, c, bd, v); break;
                default: repack_pass_a<3, LT, DP, CG>(src, nl, c, bd, v); break;
            }
        }
        __syncthreads();                                   // flags complete; this tile buffer may be refilled
        // validity is a property of the pixel (cmf/robust_mf.py:282); a background-mode pass (sel != NULL)
        // additionally drops the valid pixels that are not members of the mode being fitted (:341)
        for (int i = tid; i < LT * CG; i += NTH) {
            const int l = i / CG, cc = i % CG;
            if (l < nl && s0 + cc < S) {
                const long long o = (long long)(l0 + l) * S + s0 + cc;
                if (write_mask) mask[o] = bd[i] ? 0 : 1;
                if (sel != nullptr && sel[o] == 0) bd[i] = 1;
            }
            bad[(it & 1) ^ 1][i] = 0;                       // flags of the next tile (last read one tile ago)
        }
        if (sel != nullptr) __syncthreads();
        if (col_ok) {
            float* colbase = xt + (long long)(s0 + c) * L * DP + 4 * q;
            float* dst = colbase + (long long)l0 * DP;
            const float qnan = __int_as_float(0x7fc00000);
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                if (l < nl) {
                    float4 x = v[l];
                    const bool drop = bd[l * CG + c] != 0;
                    if (drop) {
                        x = make_float4(qnan, qnan, qnan, qnan);
                    } else {
                        acc0 += (double)x.x; acc1 += (double)x.y; acc2 += (double)x.z; acc3 += (double)x.w;
                        ++cnt;
                    }
                    if (rowidx == nullptr) *reinterpret_cast<float4*>(dst + l * DP) = x;
                    else if (!drop)      // compacted: the member pixels of the column, in line order
                        *reinterpret_cast<float4*>(colbase + (long long)rowidx[(long long)(l0 + l) * S + s0 + c] * DP) = x;
                }
            }
        }
        nbuf = buf;
        buf = buf + 1 == NS ? 0 : buf + 1;
    }
    if (col_ok) {
        double* o = colsum_part + ((long long)split * S + s0 + c) * DP + 4 * q;
        o[0] = acc0; o[1] = acc1; o[2] = acc2; o[3] = acc3;
        if (q == 0) colcnt_part[split * S + s0 + c] = cnt;
    }
}

// ---------------------------------------------------------------------------------------- K0 (pair copies)
// The same pipeline with 8-byte copies.  Measured in isolation on B200 (cmf_microbench kinds 30-33): the slab read
// through 4-byte LDGSTS runs at 2.1 TB/s -- it alone took 1.6 of the pass's 1.9 ms -- through 8-byte LDGSTS at
// 3.9 TB/s, and the 16-byte write side at 5.3 TB/s.  So whenever the rows are 8-byte aligned (even sample count
// and pitches: d.vec2) a copy moves the radiances of TWO neighbouring columns, and the shared tile is
// [column pair][line][band] in float2 units: pair stride 2 mod 16 units plus an XOR of the lowest unit index bit
// for pairs 8..15 keeps the 16 lanes of a half-warp copy on 16 different 8-byte bank pairs, and a thread's four
// bands x two columns are two aligned 16-byte reads.  thread <-> (column pair, 4 bands): 2 LDS.128 in, two
// validity tests, 8 FP64 column-sum adds and 2 STG.128 out per line.  Always 32 columns per CTA.
template <int NT, int LT, int NS>
__global__ void __launch_bounds__(32 * NT, repack_pipe_minb(NT, 32, LT, NS)) repack_pair_kernel(
    const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S, int D,
    float* __restrict__ xt, uint8_t* __restrict__ mask, double* __restrict__ colsum_part,
    int* __restrict__ colcnt_part, int lines_per_split, int line_base, int line_limit, int split_base,
    const uint8_t* __restrict__ sel, int write_mask, const int32_t* __restrict__ rowidx) {
    constexpr int DP = 8 * NT, Q = 2 * NT, NP = 16, CG = 32, NTH = NP * Q, CS2 = LT * DP + 2;   // float2 units
    extern __shared__ __align__(16) float2 tile2[];   // [NS][NP][CS2]
    __shared__ uint8_t bad[2][LT * CG];

    const int tid = threadIdx.x;
    const int s0 = blockIdx.x * CG;
    const int split = split_base + blockIdx.y;
    const int l_begin = line_base + blockIdx.y * lines_per_split;
    const int l_end = min(line_limit, l_begin + lines_per_split);
    // load side: half-warp <-> the 16 column pairs of one (line, band) row; a thread owns bands r_ld + k*Q
    const int p_ld = tid % NP, r_ld = tid / NP;
    const bool ld_ok = s0 + 2 * p_ld < S;
    const int key_ld = (p_ld >> 3) & 1;
    uint32_t sdst[4];
    long long goff[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int b = r_ld + k * Q;
        sdst[k] = smem_u32(tile2 + (size_t)p_ld * CS2 + (b ^ key_ld));
        goff[k] = (ld_ok && b < D) ? (long long)b * band_pitch : -1;
    }
    // store side: thread <-> (column pair, 4 consecutive bands)
    const int cp = tid / Q, q = tid % Q;
    const bool col_ok = s0 + 2 * cp < S;
    const int key = (cp >> 3) & 1;

    for (int i = tid; i < NS * NP * CS2; i += NTH) tile2[i] = make_float2(0.f, 0.f);   // padded bands stay zero
    for (int i = tid; i < 2 * LT * CG; i += NTH) (&bad[0][0])[i] = 0;
    __syncthreads();

    auto issue = [&](int buf, int l0, int nl) {
        const float* src = slab + (long long)l0 * line_pitch + s0 + 2 * p_ld;
        const uint32_t boff = (uint32_t)(buf * NP * CS2 * sizeof(float2));
#pragma unroll
        for (int l = 0; l < LT; ++l) {
            if (l < nl) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (goff[k] >= 0)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst[k] + boff +
                                                                                      (uint32_t)(l * DP * 8)),
                                     "l"(src + goff[k])
                                     : "memory");
            }
            src += line_pitch;
        }
        cp_async_commit();
    };

    double acc[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.0;
    int cnt0 = 0, cnt1 = 0;
    const int ntiles = (l_end - l_begin + LT - 1) / LT;
#pragma unroll
    for (int p = 0; p < NS - 1; ++p) {
        if (p < ntiles) issue(p, l_begin + p * LT, min(LT, l_end - (l_begin + p * LT)));
        else cp_async_commit();
    }
    int buf = 0, nbuf = NS - 1;
    for (int it = 0; it < ntiles; ++it) {
        const int l0 = l_begin + it * LT;
        const int nl = min(LT, l_end - l0);
        if (it + NS - 1 < ntiles) issue(nbuf, l0 + (NS - 1) * LT, min(LT, l_end - l0 - (NS - 1) * LT));
        else cp_async_commit();
        cp_async_wait<NS - 1>();
        __syncthreads();                                   // tile `it` has landed for every thread
        uint8_t* bd = bad[it & 1];
        float4 v0[LT], v1[LT];                             // the two columns of the pair
        if (col_ok) {
            const float4* src = reinterpret_cast<const float4*>(tile2 + (size_t)(buf * NP + cp) * CS2 + 4 * q);
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                if (l < nl) {
                    float4 a = src[l * (DP / 2)], b = src[l * (DP / 2) + 1];   // units (4q, 4q+1), (4q+2, 4q+3)
                    if (key) {   // pairs 8..15 hold neighbouring bands swapped (uniform for all but a few warps)
                        a = make_float4(a.z, a.w, a.x, a.y);
                        b = make_float4(b.z, b.w, b.x, b.y);
                    }
                    const float4 x0 = make_float4(a.x, a.z, b.x, b.z), x1 = make_float4(a.y, a.w, b.y, b.w);
                    v0[l] = x0;
                    v1[l] = x1;
                    if (!quad_ok(x0)) bd[l * CG + 2 * cp] = 1;
                    if (!quad_ok(x1)) bd[l * CG + 2 * cp + 1] = 1;
                }
            }
        }
        __syncthreads();                                   // flags complete; this tile buffer may be refilled
        for (int i = tid; i < LT * CG; i += NTH) {
            const int l = i / CG, cc = i % CG;
            if (l < nl && s0 + cc < S) {
                const long long o = (long long)(l0 + l) * S + s0 + cc;
                if (write_mask) mask[o] = bd[i] ? 0 : 1;
                if (sel != nullptr && sel[o] == 0) bd[i] = 1;
            }
            bad[(it & 1) ^ 1][i] = 0;                       // flags of the next tile (last read one tile ago)
        }
        if (sel != nullptr) __syncthreads();
        if (col_ok) {
            float* col0 = xt + (long long)(s0 + 2 * cp) * L * DP + 4 * q;
            float* col1 = col0 + (long long)L * DP;
            float* dst0 = col0 + (long long)l0 * DP;
            float* dst1 = col1 + (long long)l0 * DP;
            const float qnan = __int_as_float(0x7fc00000);
            const float4 nan4 = make_float4(qnan, qnan, qnan, qnan);
#pragma unroll
            for (int l = 0; l < LT; ++l) {
                if (l < nl) {
                    float4 x0 = v0[l], x1 = v1[l];
                    const bool drop0 = bd[l * CG + 2 * cp] != 0, drop1 = bd[l * CG + 2 * cp + 1] != 0;
                    if (drop0) {
                        x0 = nan4;
                    } else {
                        acc[0][0] += (double)x0.x; acc[0][1] += (double)x0.y; acc[0][2] += (double)x0.z;
                        acc[0][3] += (double)x0.w;
                        ++cnt0;
                    }
                    if (drop1) {
                        x1 = nan4;
                    } else {
                        acc[1][0] += (double)x1.x; acc[1][1] += (double)x1.y; acc[1][2] += (double)x1.z;
                        acc[1][3] += (double)x1.w;
                        ++cnt1;
                    }
                    if (rowidx == nullptr) {
                        *reinterpret_cast<float4*>(dst0 + l * DP) = x0;
                        *reinterpret_cast<float4*>(dst1 + l * DP) = x1;
                    } else {
                        // compacted: the member pixels of each column, in line order (rows past the member count are
                        // never read: the consumers stop at nrows[s])
                        const int2 r = *reinterpret_cast<const int2*>(rowidx + (long long)(l0 + l) * S + s0 + 2 * cp);
                        if (!drop0) *reinterpret_cast<float4*>(col0 + (long long)r.x * DP) = x0;
                        if (!drop1) *reinterpret_cast<float4*>(col1 + (long long)r.y * DP) = x1;
                    }
                }
            }
        }
        nbuf = buf;
        buf = buf + 1 == NS ? 0 : buf + 1;
    }
    if (col_ok) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            double* o = colsum_part + ((long long)split * S + s0 + 2 * cp + j) * DP + 4 * q;
            o[0] = acc[j][0]; o[1] = acc[j][1]; o[2] = acc[j][2]; o[3] = acc[j][3];
        }
        if (q == 0) {
            colcnt_part[split * S + s0 + 2 * cp] = cnt0;
            colcnt_part[split * S + s0 + 2 * cp + 1] = cnt1;
        }
    }
}

// ---------------------------------------------------------------------------------------- K0b
// Splits [0, nsplit) give the column mean; the same kernel over the first `nsplit` = pilot splits gives the
// pilot centre of the Gram pass (n == NULL: the count is not stored).
__global__ void mean_kernel(const double* __restrict__ colsum_part, const int* __restrict__ colcnt_part,
                            int nsplit, int S, int DP, double* __restrict__ mu, int* __restrict__ n) {
    const int s = blockIdx.x;
    int cnt = 0;
    for (int k = 0; k < nsplit; ++k) cnt += colcnt_part[k * S + s];
    for (int b = threadIdx.x; b < DP; b += blockDim.x) {
        double a = 0.0;
        for (int k = 0; k < nsplit; ++k) a += colsum_part[((long long)k * S + s) * DP + b];
        mu[(long long)s * DP + b] = cnt > 0 ? a / (double)cnt : 0.0;   // numpy mean: sum / n
    }
    if (threadIdx.x == 0 && n != nullptr) n[s] = cnt;
}

// ---------------------------------------------------------------------------------------- K5
// Scoring pass.  thread <-> (column pair, group of 8 lines); per band one 16-byte weight load shared by
// the 8 lines and eight 8-byte radiance loads; FP64 FMA on converted FP32 radiances so that the score is
// exact to FP64 rounding given the weights.  mf = sum_b x_b w_b - mu.w ; invalid pixels keep nodata.
template <int VEC>
__global__ void __launch_bounds__(256) score_kernel(const float* __restrict__ slab, long long line_pitch,
                                                    int band_pitch, int L, int S, int D,
                                                    const uint8_t* __restrict__ mask,
                                                    const double* __restrict__ wT, int Sp,
                                                    const double* __restrict__ c0,
                                                    const int* __restrict__ status, double nodata,
                                                    double* __restrict__ mf, double* __restrict__ stat_part,
                                                    int nlanes, const uint8_t* __restrict__ sel,
                                                    const int* __restrict__ mindex,
                                                    int16_t* __restrict__ alpha_img) {
    constexpr int NL = kScoreLines;
    const int SC = (S + VEC - 1) / VEC;               // column groups per line
    const int ngroups = (L + NL - 1) / NL;            // line groups
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (long long)SC * nlanes) return;
    const int sc = (int)(gid % SC);
    const int lane_id = (int)(gid / SC);
    const int col = sc * VEC;
    const bool has1 = (VEC == 2) && (col + 1 < S);

    double sum0 = 0.0, sum1 = 0.0, sq0 = 0.0, sq1 = 0.0;
    const double cz0 = c0[col], cz1 = has1 ? c0[col + 1] : 0.0;
    const bool zero0 = (status[col] & (kStatusSingular)) != 0;
    const bool zero1 = has1 ? ((status[col + 1] & (kStatusSingular)) != 0) : false;

    for (int grp = lane_id; grp < ngroups; grp += nlanes) {
        const int l0 = grp * NL;
        double a0[NL], a1[NL];
#pragma unroll
        for (int j = 0; j < NL; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
        const float* base = slab + (long long)l0 * line_pitch + col;
        if (l0 + NL <= L) {
#pragma unroll 2
            for (int b = 0; b < D; ++b) {
                double w0, w1;
                if (VEC == 2) {
                    const double2 w = *reinterpret_cast<const double2*>(wT + (long long)b * Sp + col);
                    w0 = w.x; w1 = w.y;
                } else {
                    w0 = wT[(long long)b * Sp + col]; w1 = 0.0;
                }
                const float* pb = base + (long long)b * band_pitch;
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    if (VEC == 2) {
                        const float2 v = ldg_nc_f2(pb + (long long)j * line_pitch);
                        a0[j] = fma((double)v.x, w0, a0[j]);
                        a1[j] = fma((double)v.y, w1, a1[j]);
                    } else {
                        const float v = ldg_nc_f1(pb + (long long)j * line_pitch);
                        a0[j] = fma((double)v, w0, a0[j]);
                    }
                }
            }
        } else {
            for (int b = 0; b < D; ++b) {
                const double w0 = wT[(long long)b * Sp + col];
                const double w1 = has1 ? wT[(long long)b * Sp + col + 1] : 0.0;
                const float* pb = base + (long long)b * band_pitch;
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    if (l0 + j < L) {
                        const float v = ldg_nc_f1(pb + (long long)j * line_pitch);
                        a0[j] = fma((double)v, w0, a0[j]);
                        if (has1) {
                            const float u = ldg_nc_f1(pb + (long long)j * line_pitch + 1);
                            a1[j] = fma((double)u, w1, a1[j]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NL; ++j) {
            const int l = l0 + j;
            if (l < L) {
                const long long o = (long long)l * S + col;
                if (sel != nullptr) {
                    // background-mode pass: only the members of this mode are (over)written (:383-386)
                    if (sel[o]) { mf[o] = zero0 ? 0.0 : a0[j] - cz0; alpha_img[o] = (int16_t)mindex[col]; }
                    if (has1 && sel[o + 1]) {
                        mf[o + 1] = zero1 ? 0.0 : a1[j] - cz1;
                        alpha_img[o + 1] = (int16_t)mindex[col + 1];
                    }
                    continue;
                }
                const bool ok0 = mask[o] != 0;
                const double v0 = ok0 ? (zero0 ? 0.0 : a0[j] - cz0) : nodata;
                if (ok0) { sum0 += v0; sq0 += v0 * v0; }
                if (has1) {
                    const bool ok1 = mask[o + 1] != 0;
                    const double v1 = ok1 ? (zero1 ? 0.0 : a1[j] - cz1) : nodata;
                    if (ok1) { sum1 += v1; sq1 += v1 * v1; }
                    if ((((uintptr_t)(mf + o)) & 15) == 0) {
                        *reinterpret_cast<double2*>(mf + o) = make_double2(v0, v1);
                    } else {
                        mf[o] = v0; mf[o + 1] = v1;
                    }
                } else {
                    mf[o] = v0;
                }
            }
        }
    }
    if (sel != nullptr) return;
    double* sp = stat_part + ((long long)lane_id * S + col) * 2;
    sp[0] = sum0; sp[1] = sq0;
    if (has1) { sp[2] = sum1; sp[3] = sq1; }
}

// ---------------------------------------------------------------------------------------- K5 (tiled)
// Scoring pass, HBM-bound form.  CTA = (tile of 128 columns) x (range of lines); the tile's FP64 weights
// (D x 64 column pairs x 16 B) stay in shared memory for the whole range, so a thread's registers hold
// nothing but radiances in flight: thread <-> (column pair, NL consecutive lines), BC bands per batch =
// NL*BC independent 8-byte loads issued back to back before the first FMA.  A warp reads 256 contiguous
// bytes of each (line, band) row and writes 512 contiguous bytes of scores per line.
constexpr int kScoreTile = 128;   // columns per CTA
constexpr int kScoreSlots = 4;    // line slots per CTA (256 threads = 64 column pairs x 4 slots)

template <int NL, int BC, int MINB>
__global__ void __launch_bounds__(256, MINB)
    score_tiled_kernel(const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S,
                       int D, const uint8_t* __restrict__ mask, const double* __restrict__ wT, int Sp,
                       const double* __restrict__ c0, const int* __restrict__ status, double nodata,
                       double* __restrict__ mf, double* __restrict__ stat_part, int lines_per_cta,
                       const uint8_t* __restrict__ sel, const int* __restrict__ mindex,
                       int16_t* __restrict__ alpha_img) {
    constexpr int CP = kScoreTile / 2;
    extern __shared__ double2 w_s[];   // [D][CP]
    const int tid = threadIdx.x;
    const int cp = tid & (CP - 1), slot = tid / CP;
    const int col0 = blockIdx.x * kScoreTile;
    for (int i = tid; i < D * CP; i += 256) {
        const int b = i / CP, c = col0 + 2 * (i % CP);
        w_s[i] = (c < S) ? *reinterpret_cast<const double2*>(wT + (long long)b * Sp + c) : make_double2(0.0, 0.0);
    }
    __syncthreads();
    const int col = col0 + 2 * cp;
    const int part = blockIdx.y * kScoreSlots + slot;             // index of this thread's statistics partial
    double sum0 = 0.0, sum1 = 0.0, sq0 = 0.0, sq1 = 0.0;
    if (col < S) {                                                // S is even on this path: col + 1 < S too
        const double cz0 = c0[col], cz1 = c0[col + 1];
        const bool zero0 = (status[col] & kStatusSingular) != 0, zero1 = (status[col + 1] & kStatusSingular) != 0;
        const int l_begin = blockIdx.y * lines_per_cta;
        const int l_end = min(L, l_begin + lines_per_cta);
        const double2* wp = w_s + cp;
        for (int l0 = l_begin + slot * NL; l0 < l_end; l0 += kScoreSlots * NL) {
            const float* base = slab + (long long)l0 * line_pitch + col;
            long long lo[NL];                                     // a missing line re-reads line l0
#pragma unroll
            for (int j = 0; j < NL; ++j) lo[j] = (l0 + j < l_end) ? (long long)j * line_pitch : 0;
            double a0[NL], a1[NL];
#pragma unroll
            for (int j = 0; j < NL; ++j) { a0[j] = 0.0; a1[j] = 0.0; }
            int b0 = 0;
            for (; b0 + BC <= D; b0 += BC) {
                float2 v[NL][BC];
                const float* pb = base + (long long)b0 * band_pitch;
#pragma unroll
                for (int k = 0; k < BC; ++k)
#pragma unroll
                    for (int j = 0; j < NL; ++j) v[j][k] = ldg_nc_f2(pb + (long long)k * band_pitch + lo[j]);
#pragma unroll
                for (int k = 0; k < BC; ++k) {
                    const double2 w = wp[(b0 + k) * CP];
#pragma unroll
                    for (int j = 0; j < NL; ++j) {
                        a0[j] = fma((double)v[j][k].x, w.x, a0[j]);
                        a1[j] = fma((double)v[j][k].y, w.y, a1[j]);
                    }
                }
            }
            for (; b0 < D; ++b0) {
                const float* pb = base + (long long)b0 * band_pitch;
                const double2 w = wp[b0 * CP];
#pragma unroll
                for (int j = 0; j < NL; ++j) {
                    const float2 u = ldg_nc_f2(pb + lo[j]);
                    a0[j] = fma((double)u.x, w.x, a0[j]);
                    a1[j] = fma((double)u.y, w.y, a1[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < NL; ++j) {
                if (l0 + j < l_end) {
                    const long long o = (long long)(l0 + j) * S + col;
                    if (sel == nullptr) {
                        const uchar2 ok = *reinterpret_cast<const uchar2*>(mask + o);
                        const double r0 = ok.x ? (zero0 ? 0.0 : a0[j] - cz0) : nodata;
                        const double r1 = ok.y ? (zero1 ? 0.0 : a1[j] - cz1) : nodata;
                        if (ok.x) { sum0 += r0; sq0 += r0 * r0; }
                        if (ok.y) { sum1 += r1; sq1 += r1 * r1; }
                        *reinterpret_cast<double2*>(mf + o) = make_double2(r0, r1);
                    } else {
                        // background-mode pass: only the members of this mode are (over)written (:383-386)
                        const uchar2 ok = *reinterpret_cast<const uchar2*>(sel + o);
                        if (ok.x) { mf[o] = zero0 ? 0.0 : a0[j] - cz0; alpha_img[o] = (int16_t)mindex[col]; }
                        if (ok.y) { mf[o + 1] = zero1 ? 0.0 : a1[j] - cz1; alpha_img[o + 1] = (int16_t)mindex[col + 1]; }
                    }
                }
            }
        }
        if (sel == nullptr) {
            double* sp = stat_part + ((long long)part * S + col) * 2;
            sp[0] = sum0; sp[1] = sq0; sp[2] = sum1; sp[3] = sq1;
        }
    }
}

// ---------------------------------------------------------------------------------------- K6
// colnum / colavg / colstd (np.mean, np.std ddof=0 of the written scores; cmf/robust_mf.py:388-391)
__global__ void colstats_kernel(const double* __restrict__ stat_part, int nlanes, int S,
                                const int* __restrict__ n, double nodata, double* __restrict__ colstats) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    const int cnt = n[s];
    if (cnt == 0) {
        colstats[s] = nodata; colstats[S + s] = nodata; colstats[2 * S + s] = nodata;
        return;
    }
    double sum = 0.0, sq = 0.0;
    for (int k = 0; k < nlanes; ++k) {
        sum += stat_part[((long long)k * S + s) * 2];
        sq += stat_part[((long long)k * S + s) * 2 + 1];
    }
    const double mean = sum / cnt;
    double var = sq / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    colstats[s] = (double)cnt;
    colstats[S + s] = mean;
    colstats[2 * S + s] = sqrt(var);
}

// ---------------------------------------------------------------------------------------- launchers
// Repack plan: tile shape (CG columns x LT lines) and ring depth NS of the pipelined kernel.  CG == 0 selects
// the older single-stage kernel (kept for A/B measurements through tools/).
struct RepackVariant { int cg, lt, ns; };

// the shipped library holds the measured-best shape and the small-tile fall-back; the sweep lives in the tools build
#ifdef CMF_TUNING_HOOKS
#define CMF_REPACK_VARIANTS(X) X(32, 8, 2) X(32, 4, 3) X(32, 4, 4) X(32, 2, 8) X(16, 4, 2)
#else
#define CMF_REPACK_VARIANTS(X) X(32, 4, 4) X(16, 4, 2)
#endif

static RepackVariant repack_variant_requested() {
    // tuning hook (tools/ only): CMF_REPACK_VARIANT=CG,LT,NS picks another instantiation
    static RepackVariant v = [] {
        const RepackVariant dflt{32, 4, 4};     // measured on B200 (profiles/r01g_ / r01i_tune_repack.json)
        RepackVariant r = dflt;
        if (const char* e = cmf_hook("CMF_REPACK_VARIANT")) sscanf(e, "%d,%d,%d", &r.cg, &r.lt, &r.ns);
        bool known = false;
#define CMF_RV(CGv, LTv, NSv) known = known || (r.cg == CGv && r.lt == LTv && r.ns == NSv);
        CMF_REPACK_VARIANTS(CMF_RV)
#undef CMF_RV
        return known ? r : dflt;
    }();
    return v;
}

// the variant that is launched for NT band tiles: a ring that does not fit shared memory at this band count
// falls back to the two-stage (16, 4) tile
static RepackVariant repack_variant(int nt) {
    const RepackVariant v = repack_variant_requested();
    if ((size_t)v.ns * v.cg * (v.lt * 8 * nt + 4) * sizeof(float) > 227 * 1024) return RepackVariant{16, 4, 2};
    return v;
}

template <int NT, int CG, int LT, int NS>
static size_t repack_pipe_smem() { return (size_t)NS * CG * (LT * 8 * NT + 4) * sizeof(float); }

template <int NT, int CG, int LT, int NS>
static int repack_pipe_resident() {
    const size_t smem = repack_pipe_smem<NT, CG, LT, NS>();
    if (smem > 227 * 1024) return 0;
    cudaFuncSetAttribute(repack_pipe_kernel<NT, CG, LT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, repack_pipe_kernel<NT, CG, LT, NS>, CG * 2 * NT, smem) !=
        cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    return nb;
}

// 8-byte copies (repack_pair_kernel) whenever the rows allow it; CMF_REPACK_PAIR=0 keeps the 4-byte kernel (tools/)
static bool repack_use_pair(const Dims& d, const RepackVariant& v) {
    static const bool enabled = [] { const char* e = cmf_hook("CMF_REPACK_PAIR"); return !(e && e[0] == '0'); }();
    return enabled && d.vec2 && v.cg == 32;
}

template <int NT, int LT, int NS>
static size_t repack_pair_smem() { return (size_t)NS * 16 * (LT * 8 * NT + 2) * sizeof(float2); }

template <int NT, int LT, int NS>
static int repack_pair_resident() {
    const size_t smem = repack_pair_smem<NT, LT, NS>();
    if (smem > 227 * 1024) return 0;
    cudaFuncSetAttribute(repack_pair_kernel<NT, LT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, repack_pair_kernel<NT, LT, NS>, 32 * NT, smem) != cudaSuccess) {
        cudaGetLastError();
        nb = 0;
    }
    return nb;
}

template <int NT>
static int repack_pair_resident_t(const RepackVariant& v) {
    if (v.lt == 4 && v.ns == 4) return repack_pair_resident<NT, 4, 4>();
#ifdef CMF_TUNING_HOOKS
    if (v.lt == 2 && v.ns == 8) return repack_pair_resident<NT, 2, 8>();
    if (v.lt == 4 && v.ns == 3) return repack_pair_resident<NT, 4, 3>();
    if (v.lt == 8 && v.ns == 2) return repack_pair_resident<NT, 8, 2>();
#endif
    return 0;
}

template <int NT>
static int repack_resident_t(const RepackVariant& v) {
#define CMF_RV(CGv, LTv, NSv) \
    if (v.cg == CGv && v.lt == LTv && v.ns == NSv) return repack_pipe_resident<NT, CGv, LTv, NSv>();
    CMF_REPACK_VARIANTS(CMF_RV)
#undef CMF_RV
    return 0;
}

#define CMF_NT_SWITCH(nt, CALL)                                                      \
    switch (nt) {                                                                    \
        case 1: { constexpr int NTc = 1; CALL; } break;                              \
        case 2: { constexpr int NTc = 2; CALL; } break;                              \
        case 3: { constexpr int NTc = 3; CALL; } break;                              \
        case 4: { constexpr int NTc = 4; CALL; } break;                              \
        case 5: { constexpr int NTc = 5; CALL; } break;                              \
        case 6: { constexpr int NTc = 6; CALL; } break;                              \
        case 7: { constexpr int NTc = 7; CALL; } break;                              \
        case 8: { constexpr int NTc = 8; CALL; } break;                              \
        case 9: { constexpr int NTc = 9; CALL; } break;                              \
        case 10: { constexpr int NTc = 10; CALL; } break;                            \
        case 11: { constexpr int NTc = 11; CALL; } break;                            \
        case 12: { constexpr int NTc = 12; CALL; } break;                            \
        default: break;                                                              \
    }

// Line ranges per column group.  The pipelined kernel is launched as (at most) one resident wave: every CTA
// streams its whole line range, so the ranges are sized to fill the SMs' resident slots once.
int repack_nsplit(const Dims& d) {
    const RepackVariant v = repack_variant(d.NT);
    int resident = 0, sms = 148, dev = 0;
    if (repack_use_pair(d, v)) { CMF_NT_SWITCH(d.NT, (resident = repack_pair_resident_t<NTc>(v))); }
    else { CMF_NT_SWITCH(d.NT, (resident = repack_resident_t<NTc>(v))); }
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (resident < 1) resident = 1;
    const int groups = (d.S + v.cg - 1) / v.cg;
    int ns = (sms * resident) / groups;
    if (const char* e = cmf_hook("CMF_REPACK_NSPLIT")) ns = atoi(e);      // tuning hook (tools/ only)
    const int maxsplit = (d.L + 4 * v.lt - 1) / (4 * v.lt);             // at least 4 tiles per CTA
    if (ns > maxsplit) ns = maxsplit;
    return ns < 1 ? 1 : ns;
}

template <int NT, int CG, int LT, int NS>
static void launch_repack_pipe(const Dims& d, const float* slab, float* xt, uint8_t* mask, double* colsum_part,
                               int* colcnt_part, int lps, int line_base, int line_limit, int split_base,
                               const uint8_t* sel, int write_mask, cudaStream_t st) {
    const size_t smem = repack_pipe_smem<NT, CG, LT, NS>();
    cudaFuncSetAttribute(repack_pipe_kernel<NT, CG, LT, NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int nblk = (line_limit - line_base + lps - 1) / lps;
    if (nblk <= 0) return;
    dim3 grid((d.S + CG - 1) / CG, nblk);
    repack_pipe_kernel<NT, CG, LT, NS><<<grid, CG * 2 * NT, smem, st>>>(slab, d.line_pitch, d.band_pitch, d.L, d.S,
                                                                        d.D, xt, mask, colsum_part, colcnt_part, lps,
                                                                        line_base, line_limit, split_base, sel,
                                                                        write_mask, d.rowidx);
}

template <int NT>
static void launch_repack_t(const Dims& d, const float* slab, float* xt, uint8_t* mask, double* colsum_part,
                            int* colcnt_part, int lps, int line_base, int line_limit, int split_base,
                            const uint8_t* sel, int write_mask, cudaStream_t st) {
    const RepackVariant v = repack_variant(NT);
    if (repack_use_pair(d, v)) {
        const int nblk = (line_limit - line_base + lps - 1) / lps;
        if (nblk <= 0) return;
        dim3 grid((d.S + 31) / 32, nblk);
#define CMF_RP(LTv, NSv)                                                                                          \
        if (v.lt == LTv && v.ns == NSv) {                                                                         \
            const size_t smem = repack_pair_smem<NT, LTv, NSv>();                                                 \
            cudaFuncSetAttribute(repack_pair_kernel<NT, LTv, NSv>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                 (int)smem);                                                                      \
            repack_pair_kernel<NT, LTv, NSv><<<grid, 32 * NT, smem, st>>>(                                        \
                slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, xt, mask, colsum_part, colcnt_part, lps,         \
                line_base, line_limit, split_base, sel, write_mask, d.rowidx);                                    \
            return;                                                                                               \
        }
        CMF_RP(4, 4)
#ifdef CMF_TUNING_HOOKS
        CMF_RP(4, 3) CMF_RP(8, 2) CMF_RP(2, 8)
#endif
#undef CMF_RP
    }
#define CMF_RV(CGv, LTv, NSv)                                                                                \
    if (v.cg == CGv && v.lt == LTv && v.ns == NSv)                                                   \
        return launch_repack_pipe<NT, CGv, LTv, NSv>(d, slab, xt, mask, colsum_part, colcnt_part, lps,       \
                                                     line_base, line_limit, split_base, sel, write_mask, st);
    CMF_REPACK_VARIANTS(CMF_RV)
#undef CMF_RV
    // repack_variant() only returns instantiated shapes
    launch_repack_pipe<NT, 16, 4, 2>(d, slab, xt, mask, colsum_part, colcnt_part, lps, line_base, line_limit,
                                     split_base, sel, write_mask, st);
}

int repack_lines_per_split(const Dims& d, int nsplit) {
    int lps = (d.L + nsplit - 1) / nsplit;
    return (lps + kGramTL - 1) / kGramTL * kGramTL;    // Gram chunks are whole splits made of whole Gram tiles
}

// Repack lines [line_base, line_limit); line_base must be a multiple of lines-per-split so that the
// partial-sum slots (split_base + i) are the same whether the cube is processed whole or in blocks.
void launch_repack(const Dims& d, const float* slab, float* xt, uint8_t* mask, double* colsum_part,
                   int* colcnt_part, int lps, int line_base, int line_limit, const uint8_t* sel, int write_mask,
                   cudaStream_t st) {
    const int split_base = line_base / lps;
    CMF_NT_SWITCH(d.NT, (launch_repack_t<NTc>(d, slab, xt, mask, colsum_part, colcnt_part, lps, line_base,
                                              line_limit, split_base, sel, write_mask, st)));
}

void launch_mean(const Dims& d, const double* colsum_part, const int* colcnt_part, int nsplit, double* mu,
                 int* n, cudaStream_t st) {
    mean_kernel<<<d.S, 128, 0, st>>>(colsum_part, colcnt_part, nsplit, d.S, d.DP, mu, n);
}

// Plan of the scoring pass: number of statistics partials per column (`nlanes`) and, on the tiled path,
// the lines each CTA covers.  The tiled grid is sized to one wave of MINB CTAs per SM.
struct ScoreVariant { int nl, bc, minb; };

static ScoreVariant score_variant() {
    // tuning hook (tools/ only): CMF_SCORE_VARIANT=NL,BC,MINB picks another instantiation
    static ScoreVariant v = [] {
        ScoreVariant r{2, 18, 2};   // measured best on B200: 0.59 ms per 20k-line flightline
        if (const char* e = cmf_hook("CMF_SCORE_VARIANT")) sscanf(e, "%d,%d,%d", &r.nl, &r.bc, &r.minb);
        return r;
    }();
    return v;
}

int score_plan(const Dims& d, int sm_count, int* lines_per_cta) {
    if (d.vec2 && d.D <= kScoreTiledMaxD) {
        const ScoreVariant v = score_variant();
        const int ntiles = (d.S + kScoreTile - 1) / kScoreTile;
        const int step = kScoreSlots * v.nl;
        int nranges = (sm_count * v.minb) / ntiles;
        if (nranges < 1) nranges = 1;
        int lpc = (d.L + nranges - 1) / nranges;
        lpc = (lpc + step - 1) / step * step;
        nranges = (d.L + lpc - 1) / lpc;
        *lines_per_cta = lpc;
        return nranges * kScoreSlots;
    }
    const int ngroups = (d.L + kScoreLines - 1) / kScoreLines;
    int want = (sm_count * 1536) / d.S;
    if (want < 1) want = 1;
    const int per = (ngroups + want - 1) / want;
    *lines_per_cta = 0;
    return (ngroups + per - 1) / per;
}

template <int NL, int BC, int MINB>
static void launch_score_tiled(const Dims& d, const float* slab, const uint8_t* mask, const double* wT,
                               const double* c0, const int* status, double nodata, double* mf,
                               double* stat_part, int nlanes, int lines_per_cta, const uint8_t* sel,
                               const int* mindex, int16_t* alpha_img, cudaStream_t st) {
    const int Sp = (d.S + 1) & ~1;
    const size_t smem = (size_t)d.D * (kScoreTile / 2) * sizeof(double2);
    cudaFuncSetAttribute(score_tiled_kernel<NL, BC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((d.S + kScoreTile - 1) / kScoreTile, nlanes / kScoreSlots);
    score_tiled_kernel<NL, BC, MINB><<<grid, 256, smem, st>>>(slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, mask,
                                                              wT, Sp, c0, status, nodata, mf, stat_part,
                                                              lines_per_cta, sel, mindex, alpha_img);
}

void launch_score(const Dims& d, const float* slab, const uint8_t* mask, const double* wT, const double* c0,
                  const int* status, double nodata, double* mf, double* stat_part, int nlanes,
                  int lines_per_cta, const uint8_t* sel, const int* mindex, int16_t* alpha_img,
                  cudaStream_t st) {
    const int Sp = (d.S + 1) & ~1;
    if (d.vec2 && d.D <= kScoreTiledMaxD) {
        const ScoreVariant v = score_variant();
#define CMF_SV(NL, BC, MB)                                                                              \
    if (v.nl == NL && v.bc == BC && v.minb == MB)                                                       \
        return launch_score_tiled<NL, BC, MB>(d, slab, mask, wT, c0, status, nodata, mf, stat_part, nlanes, \
                                              lines_per_cta, sel, mindex, alpha_img, st);
        CMF_SV(2, 18, 2)
#ifdef CMF_TUNING_HOOKS
        CMF_SV(2, 12, 2) CMF_SV(4, 8, 2) CMF_SV(2, 8, 3)
#endif
#undef CMF_SV
        return launch_score_tiled<2, 18, 2>(d, slab, mask, wT, c0, status, nodata, mf, stat_part, nlanes,
                                            lines_per_cta, sel, mindex, alpha_img, st);
    }
    if (d.vec2) {      // wide windows: weights from global memory / L1, column pairs
        const long long total = (long long)((d.S + 1) / 2) * nlanes;
        score_kernel<2><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
            slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, mask, wT, Sp, c0, status, nodata, mf, stat_part,
            nlanes, sel, mindex, alpha_img);
        return;
    }
    const long long total = (long long)d.S * nlanes;
    score_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(
        slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, mask, wT, Sp, c0, status, nodata, mf, stat_part,
        nlanes, sel, mindex, alpha_img);
}

void launch_colstats(const Dims& d, const double* stat_part, int nlanes, const int* n, double nodata,
                     double* colstats, cudaStream_t st) {
    colstats_kernel<<<(d.S + 127) / 128, 128, 0, st>>>(stat_part, nlanes, d.S, n, nodata, colstats);
}

// Number of line chunks per column so that S*nchunk CTAs fill whole waves of (sm_count*ctas_per_sm).
int pick_chunks(int S, int L, int min_lines, int sm_count, int ctas_per_sm) {
    const int slots = sm_count * ctas_per_sm;
    int maxc = L / min_lines;
    if (maxc < 1) maxc = 1;
    if (maxc > 64) maxc = 64;
    int best = 1;
    double best_eff = 0.0;
    for (int c = 1; c <= maxc; ++c) {
        const long long ctas = (long long)S * c;
        const long long waves = (ctas + slots - 1) / slots;
        const double eff = (double)ctas / (double)(waves * slots);
        if (eff > best_eff + 0.01) { best_eff = eff; best = c; }   // prefer fewer chunks unless >1 % better
        if (best_eff > 0.985) break;
    }
    return best;
}

}  // namespace cmf
