// K3a on the 5th-generation tensor cores: the screening pass of the leave-one-out alpha search
// (cmf/robust_mf.py:105-117) as two chained tcgen05.mma contractions with TMEM accumulators.
//
// Same arithmetic as k_screen.cu (which stays as the path for windows that do not fit TMEM):
//     GEMM1  Y = Xc . P                 3 x TF32 (xh.Ph + xl.Ph + xh.Pl), FP32 accumulate in TMEM
//     GEMM2  R = (Y*Y) . W              3 x TF32 (zh.Wh + zh.Wl + zl.Wh), FP32 accumulate in TMEM
//     h(r)   = log(1-u) + r u/(1-u), u = beta r, FP32 series, summed per alpha
// but organised for the Blackwell tensor pipe instead of per-warp mma.sync fragments:
//   * one CTA per (column, line chunk), 128-pixel tiles: the pixel is the TMEM lane (M = 128),
//   * the A operand never touches shared memory: the converting threads write xh|xl straight into
//     TMEM with tcgen05.st, GEMM1 leaves Y in TMEM, the squaring threads read it with tcgen05.ld and
//     write zh|zl back, GEMM2 takes them as its TMEM A operand (tcgen05.mma "TS" form),
//   * the B operands (P and W, hi and lo TF32 parts) sit in shared memory for the whole CTA in the
//     canonical K-major no-swizzle core-matrix layout, built once per column by screen5_tables_kernel,
//   * R is split into two alpha halves with their own full/empty barriers so that the FP32 epilogue
//     of one half overlaps the tensor work of the other half and of the next tile,
//   * warp roles: warp 0 bulk-copy producer, warp 1 MMA issuer (one thread), warps 4-7 convert and
//     square (thread = pixel), warps 8-15 epilogue (thread = pixel, 112 alphas each, accumulators in
//     registers; setmaxnreg moves registers from the control warps to the epilogue warps).
//
// TMEM columns (DP = padded bands, N1 = DP rounded to 16, NA = padded alphas):
//     [0, 2DP)            xh | xl          A of GEMM1
//     [2DP, 2DP+N1)       Y, then zh       D of GEMM1, A of GEMM2
//     [2DP+N1, 3DP+N1)    zl               A of GEMM2
//     [3DP+N1, .. + NA)   R                D of GEMM2
#include <algorithm>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "cmf_common.cuh"
#include "cmf_internal.h"
#include "cmf_tc5.cuh"

namespace cmf {

namespace {

constexpr int kT5Threads = 512;
// register budgets after setmaxnreg (4 control warps, 4 convert/square warps, 8 epilogue warps; the sum
// 4 a + 4 b + 8 c may not exceed 2048): the epilogue keeps 112 per-alpha accumulators per thread
#ifndef CMF_S5_REGS_CTL
#define CMF_S5_REGS_CTL 40
#define CMF_S5_REGS_CVT 104
#define CMF_S5_REGS_EPI 184
#endif
#define CMF_STR2(x) #x
#define CMF_STR(x) CMF_STR2(x)
constexpr int kT5Stages = 3;       // 64-row half tiles in flight
constexpr int kT5HalfRows = 64;

// The epilogue sums g = h + u per alpha, where h = log(1-u) + r u/(1-u) and u = beta r (log q + r/q = r + h):
//     g = sum_{k>=2} u^k (1/beta - 1/k)
// sum_k u_k = beta sum_k r_k is known in closed form (rsum), so K3b subtracts it in FP64.  d_k = 1/beta - 1/k
// are per-alpha constants: k = 2..5 come from a shared-memory table (terms to u^5: |u| <= 2^-6), k <= 10
// (|u| <= 2^-3) are formed on the fly; anything larger is the rare out-of-line slow path.
// Everything that is not the 5-term fast path, out of line so that the hot loop stays small: 10 terms while
// |u| <= 2^-3 (u^11 < 2^-33), log1p and a division up to u = 1/4, poison beyond.
__device__ __noinline__ void g_cold16(const float* __restrict__ rp, const float* __restrict__ beta,
                                      const float4* __restrict__ dp, float* __restrict__ g) {
    float r[16], u[16], umax = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { r[e] = rp[e]; u[e] = beta[e] * r[e]; umax = fmaxf(umax, fabsf(u[e])); }
    if (!(umax == umax)) umax = 1.0f;
    if (__ballot_sync(0xffffffffu, umax > 0x1p-3f) == 0u) {
        float pz[16], binv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { binv[e] = dp[e].x + 0.5f; pz[e] = binv[e] - 0.1f; }
#pragma unroll 1
        for (int kk = 9; kk >= 2; --kk) {
            const float ik = 1.0f / (float)kk;
#pragma unroll
            for (int e = 0; e < 16; ++e) pz[e] = fmaf(pz[e], u[e], binv[e] - ik);
        }
#pragma unroll
        for (int e = 0; e < 16; ++e) g[e] = u[e] * u[e] * pz[e];
    } else {
#pragma unroll 1
        for (int e = 0; e < 16; ++e) {
            const float uu = beta[e] * rp[e], au = fabsf(uu);
            // beyond u = 1/4 the term amplifies the TF32 error of r by u/(1-u) and more: such a pixel is an
            // extreme outlier (x'G^-1 x > n/4); poison the sum so that the column is searched in FP64
            g[e] = (au <= 0.25f) ? (log1pf(-uu) + rp[e] * uu / (1.0f - uu)) + uu : __int_as_float(0x7fc00000);
        }
    }
}

// ------------------------------------------------------------------ shared-memory plan
struct T5Plan {
    int DP, N1, NA, KC;
    uint32_t ph_off, pl_off, wh_off, wl_off, tab_bytes;      // B operand tables (one bulk copy)
    uint32_t ring_off, mu_off, beta_off, dtab_off, bar_off, total;
};

__host__ __device__ inline T5Plan t5_plan(int NT, int NT16) {
    T5Plan p;
    p.DP = 8 * NT; p.N1 = (p.DP + 15) / 16 * 16; p.NA = 16 * NT16; p.KC = p.DP / 4;
    const uint32_t pbytes = (uint32_t)p.KC * p.N1 * 16, wbytes = (uint32_t)p.KC * p.NA * 16;
    p.ph_off = 0; p.pl_off = pbytes; p.wh_off = 2 * pbytes; p.wl_off = 2 * pbytes + wbytes;
    p.tab_bytes = 2 * pbytes + 2 * wbytes;
    p.ring_off = p.tab_bytes;
    p.mu_off = p.ring_off + (uint32_t)kT5Stages * kT5HalfRows * p.DP * 4;
    p.beta_off = p.mu_off + (uint32_t)p.DP * 8;
    p.dtab_off = (p.beta_off + (uint32_t)p.NA * 4 + 15u) & ~15u;
    p.bar_off = p.dtab_off + (uint32_t)p.NA * 16;
    p.total = p.bar_off + 24 * 8;
    return p;
}

enum { B_TAB = 0, B_XFULL = 1, B_XEMPTY = 4, B_XREADY = 7, B_G1 = 8, B_ZREADY = 9, B_RFULL = 10, B_REMPTY = 12,
       B_COUNT = 14 };

// ------------------------------------------------------------------ B operand tables
// tab[s] = Ph | Pl | Wh | Wl, each [K chunk c][row][4]: P rows are eigen-directions j (K = band b),
// W rows are alphas i (K = eigen-direction j); hi = tf32(v), lo = tf32(v - hi).
__global__ void __launch_bounds__(256)
    screen5_tables_kernel(const int* __restrict__ n_g, const int* __restrict__ nloo_g,
                          const double* __restrict__ alphas, int A, int D, int NT, int NT16c, int nparts, int AP16,
                          const double* __restrict__ P_g, const double* __restrict__ lam_g,
                          float* __restrict__ tab_g, float* __restrict__ betaf_g) {
    // one table set per (column, alpha part): a part is the NT16c alpha tiles one CTA of the screen handles
    const T5Plan p = t5_plan(NT, NT16c);
    const int s = blockIdx.x, part = blockIdx.y, tid = threadIdx.x;
    const int a_off = part * p.NA;
    const int n = n_g[s];
    if (n < 2) return;
    const int DP = p.DP;
    float* tab = tab_g + ((size_t)s * nparts + part) * (p.tab_bytes / 4);
    const double* P = P_g + (long long)s * DP * DP;
    const double* lam = lam_g + (long long)s * DP;
    for (int idx = tid; idx < p.KC * p.N1 * 4; idx += blockDim.x) {
        const int e = idx & 3, j = (idx >> 2) % p.N1, c = (idx >> 2) / p.N1;
        const int b = 4 * c + e;
        const double v = (j < DP) ? P[b * DP + j] : 0.0;
        const float hi = to_tf32((float)v);
        tab[p.ph_off / 4 + idx] = hi;
        tab[p.pl_off / 4 + idx] = to_tf32((float)(v - (double)hi));
    }
    const double dn = (double)(nloo_g ? nloo_g[s] : n);
    if (part == 0)
        for (int i = tid; i < AP16; i += blockDim.x)
            betaf_g[(long long)s * AP16 + i] = (i < A) ? (float)((1.0 - alphas[i]) / (dn - 1.0)) : 0.f;
    for (int idx = tid; idx < p.KC * p.NA * 4; idx += blockDim.x) {
        const int e = idx & 3, i = a_off + (idx >> 2) % p.NA, c = (idx >> 2) / p.NA;
        const int j = 4 * c + e;
        double w = 0.0;
        if (j < D && i < A) {
            const double al = alphas[i];
            const double be = (1.0 - al) / (dn - 1.0);
            w = 1.0 / (dn * be * lam[j] + al);
        }
        const float hi = to_tf32((float)w);
        tab[p.wh_off / 4 + idx] = hi;
        tab[p.wl_off / 4 + idx] = to_tf32((float)(w - (double)hi));
    }
}

// ------------------------------------------------------------------ the screening kernel
template <int NT, int NH16>
__global__ void __launch_bounds__(kT5Threads, 1)
    loo_screen5_kernel(const float* __restrict__ xt, const double* __restrict__ mu_g,
                       const float* __restrict__ tab_g, const float* __restrict__ betaf_g,
                       const int* __restrict__ n_g, int L, int NT16, int AP16, int lines_per_chunk,
                       double* __restrict__ fscreen, const int* __restrict__ ncomp) {
    // NT16 = alpha tiles of THIS CTA (blockIdx.z selects the alpha part; windows of more than 72 bands split the
    // alphas over two CTAs so that tables and accumulators fit shared memory and TMEM), AP16 = padded alphas in all
    constexpr int DP = 8 * NT, N1 = (DP + 15) / 16 * 16;
    constexpr uint32_t C_XH = 0, C_XL = DP, C_Y = 2 * DP, C_ZL = 2 * DP + N1, C_R = 3 * DP + N1;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const T5Plan p = t5_plan(NT, NT16);
    const int NA = p.NA;
    const int nt_a = min(NH16, NT16), nt_b = NT16 - nt_a;       // 16-alpha tiles per half
    const int NA_a = 16 * nt_a, NA_b = 16 * nt_b;

    const int s = blockIdx.x, chunk = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int a_off = blockIdx.z * NA;                           // first alpha of this CTA's part
    const int na_out = min(NA, AP16 - a_off);                    // alphas of the part that exist
    double* out = fscreen + ((long long)s * gridDim.y + chunk) * AP16 + a_off;
    if (n_g[s] < 2) return;                                      // nothing to search (K4 handles n < 2)
    const int c_begin = chunk * lines_per_chunk;
    const int c_end = min(ncomp ? min(L, ncomp[s]) : L, c_begin + lines_per_chunk);   // compacted mode pass: ncomp[s] rows
    if (c_end <= c_begin) {                                      // empty tail chunk
        for (int i = tid; i < na_out; i += blockDim.x) out[i] = 0.0;
        return;
    }
    const int nrows = c_end - c_begin;
    const int ntiles = (nrows + 127) / 128;
    const int nhalf = (nrows + kT5HalfRows - 1) / kT5HalfRows;

    float* ring = reinterpret_cast<float*>(smem_raw + p.ring_off);
    float2* mu2 = reinterpret_cast<float2*>(smem_raw + p.mu_off);
    float* beta_s = reinterpret_cast<float*>(smem_raw + p.beta_off);
    float4* dtab = reinterpret_cast<float4*>(smem_raw + p.dtab_off);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + p.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    if (tid == 0) {
        mbar_init(&bars[B_TAB], 1);
        for (int i = 0; i < kT5Stages; ++i) { mbar_init(&bars[B_XFULL + i], 1); mbar_init(&bars[B_XEMPTY + i], 2); }
        mbar_init(&bars[B_XREADY], 128);
        mbar_init(&bars[B_G1], 1);
        mbar_init(&bars[B_ZREADY], 128);
        mbar_init(&bars[B_RFULL], 1); mbar_init(&bars[B_RFULL + 1], 1);
        mbar_init(&bars[B_REMPTY], 128); mbar_init(&bars[B_REMPTY + 1], 128);
        fence_mbar_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < DP; i += blockDim.x) {
        const double m = mu_g[(long long)s * DP + i];
        const float mh = (float)m;
        mu2[i] = make_float2(mh, (float)(m - (double)mh));
    }
    for (int i = tid; i < NA; i += blockDim.x) {
        const float b = (i < na_out) ? betaf_g[(long long)s * AP16 + a_off + i] : 0.f;
        const float binv = 1.0f / b;
        beta_s[i] = b;
        // beta == 0 (alpha == 1, padding): u == 0 and every term vanishes; keep the constants finite
        dtab[i] = (b > 0.f) ? make_float4(binv - 0.5f, binv - 1.0f / 3.0f, binv - 0.25f, binv - 0.2f)
                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    const int wg = warp >> 2;
    if (wg == 0) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_CTL) ";");
        if (warp == 0 && lane == 0) {
            // ---------------- producer: tables once, then 64-row half tiles through the ring
            mbar_expect_tx(&bars[B_TAB], p.tab_bytes);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(tab_g) +
                                       ((size_t)s * gridDim.z + blockIdx.z) * p.tab_bytes;
            for (uint32_t off = 0; off < p.tab_bytes; off += 32768u) {
                const uint32_t bytes = min(32768u, p.tab_bytes - off);
                bulk_g2s(smem_raw + off, src + off, bytes, &bars[B_TAB]);
            }
            const float* col_base = xt + ((long long)s * L + c_begin) * DP;
            for (int h = 0; h < nhalf; ++h) {
                const int slot = h % kT5Stages, use = h / kT5Stages;
                if (use > 0) mbar_wait_guard(&bars[B_XEMPTY + slot], (uint32_t)((use - 1) & 1));
                const int rows = min(kT5HalfRows, nrows - h * kT5HalfRows);
                const uint32_t bytes = (uint32_t)(rows * DP * sizeof(float));
                mbar_expect_tx(&bars[B_XFULL + slot], bytes);
                bulk_g2s(ring + slot * kT5HalfRows * DP, col_base + (long long)h * kT5HalfRows * DP, bytes,
                         &bars[B_XFULL + slot]);
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues (see elect_one())
            const uint32_t sbase = smem_u32(smem_raw);
            const uint32_t lbo1 = N1 * 16, lbo2 = (uint32_t)NA * 16;
            const uint32_t id1 = idesc_tf32(N1), id2a = idesc_tf32(NA_a), id2b = idesc_tf32(NA_b > 0 ? NA_b : 16);
            mbar_wait_guard(&bars[B_TAB], 0);
            for (int t = 0; t < ntiles; ++t) {
                const uint32_t ph = (uint32_t)(t & 1);
                mbar_wait_guard(&bars[B_XREADY], ph);
                tc_fence_after();
#pragma unroll 1
                for (int ks = 0; ks < NT; ++ks) {
                    const uint64_t dh = smem_desc(sbase + p.ph_off + 2 * ks * lbo1, lbo1, 128);
                    const uint64_t dl = smem_desc(sbase + p.pl_off + 2 * ks * lbo1, lbo1, 128);
                    if (elect_one()) {
                        mma_ts_tf32(tmem + C_Y, tmem + C_XH + 8 * ks, dh, id1, ks > 0);
                        mma_ts_tf32(tmem + C_Y, tmem + C_XL + 8 * ks, dh, id1, 1);
                        mma_ts_tf32(tmem + C_Y, tmem + C_XH + 8 * ks, dl, id1, 1);
                    }
                }
                if (elect_one()) tc_commit(&bars[B_G1]);
                mbar_wait_guard(&bars[B_ZREADY], ph);
                if (t > 0) mbar_wait_guard(&bars[B_REMPTY], ph ^ 1u);
                tc_fence_after();
#pragma unroll 1
                for (int ks = 0; ks < NT; ++ks) {
                    const uint64_t dh = smem_desc(sbase + p.wh_off + 2 * ks * lbo2, lbo2, 128);
                    const uint64_t dl = smem_desc(sbase + p.wl_off + 2 * ks * lbo2, lbo2, 128);
                    if (elect_one()) {
                        mma_ts_tf32(tmem + C_R, tmem + C_Y + 8 * ks, dh, id2a, ks > 0);
                        mma_ts_tf32(tmem + C_R, tmem + C_Y + 8 * ks, dl, id2a, 1);
                        mma_ts_tf32(tmem + C_R, tmem + C_ZL + 8 * ks, dh, id2a, 1);
                    }
                }
                if (elect_one()) tc_commit(&bars[B_RFULL]);
                if (NA_b > 0) {
                    if (t > 0) { mbar_wait_guard(&bars[B_REMPTY + 1], ph ^ 1u); tc_fence_after(); }
#pragma unroll 1
                    for (int ks = 0; ks < NT; ++ks) {
                        const uint64_t dh = smem_desc(sbase + p.wh_off + 2 * ks * lbo2 + NA_a * 16, lbo2, 128);
                        const uint64_t dl = smem_desc(sbase + p.wl_off + 2 * ks * lbo2 + NA_a * 16, lbo2, 128);
                        if (elect_one()) {
                            mma_ts_tf32(tmem + C_R + NA_a, tmem + C_Y + 8 * ks, dh, id2b, ks > 0);
                            mma_ts_tf32(tmem + C_R + NA_a, tmem + C_Y + 8 * ks, dl, id2b, 1);
                            mma_ts_tf32(tmem + C_R + NA_a, tmem + C_ZL + 8 * ks, dh, id2b, 1);
                        }
                    }
                    if (elect_one()) tc_commit(&bars[B_RFULL + 1]);
                }
            }
        }
        __syncwarp();
    } else if (wg == 1) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_CVT) ";");
        // ---------------- convert (x -> xh|xl) and square (y -> zh|zl): thread = pixel = TMEM lane
        const int q = warp & 3, row = 32 * q + lane, myhalf = q >> 1, rin = row & (kT5HalfRows - 1);
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        // loops stay rolled: one warp per scheduler runs this code, so its footprint has to sit in the
        // instruction cache (the unrolled first version spent most of its time in instruction fetch)
        auto square8 = [&](uint32_t (&y)[8], int c) {
            uint32_t lo[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float yv = __uint_as_float(y[e]);
                const float z = yv * yv;
                y[e] = tf32_round(z);
                lo[e] = tf32_trunc(z - __uint_as_float(y[e]));
            }
            tmem_st8(tl + C_Y + 8 * c, y);
            tmem_st8(tl + C_ZL + 8 * c, lo);
        };
#pragma unroll 1
        for (int t = 0; t <= ntiles; ++t) {
            if (t > 0) {
                // ---- y -> zh | zl of tile t-1
                mbar_wait_guard(&bars[B_G1], (uint32_t)((t - 1) & 1));
                tc_fence_after();
                uint32_t ya[8], yb[8];
                tmem_ld8(tl + C_Y, ya);
#pragma unroll 1
                for (int c = 0; c < NT; c += 2) {
                    tc_wait_ld();
                    if (c + 1 < NT) tmem_ld8(tl + C_Y + 8 * (c + 1), yb);
                    square8(ya, c);
                    if (c + 1 < NT) {
                        tc_wait_ld();
                        if (c + 2 < NT) tmem_ld8(tl + C_Y + 8 * (c + 2), ya);
                        square8(yb, c + 1);
                    }
                }
                tc_wait_st();
                tc_fence_before();
                mbar_arrive(&bars[B_ZREADY]);
            }
            if (t < ntiles) {
                // ---- x -> xh | xl of tile t
                const int h = 2 * t + myhalf, slot = h % kT5Stages, use = h / kT5Stages;
                const bool have = h < nhalf;
                if (have) mbar_wait_guard(&bars[B_XFULL + slot], (uint32_t)(use & 1));
                const bool row_ok = have && (128 * t + row < nrows);
                const float4* src = reinterpret_cast<const float4*>(ring + (slot * kT5HalfRows + rin) * DP);
#pragma unroll 1
                for (int c = 0; c < NT; ++c) {
                    float x[8];
                    {
                        const float4 a = src[2 * c], b = src[2 * c + 1];    // stale but in-bounds when !row_ok
                        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
                    }
                    uint32_t hi[8], lo[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float2 m = mu2[8 * c + e];
                        const float v = (row_ok && x[e] == x[e]) ? (x[e] - m.x) - m.y : 0.f;
                        hi[e] = tf32_round(v);
                        lo[e] = tf32_trunc(v - __uint_as_float(hi[e]));
                    }
                    tmem_st8(tl + C_XH + 8 * c, hi);
                    tmem_st8(tl + C_XL + 8 * c, lo);
                }
                tc_wait_st();
                if (have) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[B_XEMPTY + slot]);
                }
                tc_fence_before();
                mbar_arrive(&bars[B_XREADY]);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_EPI) ";");
        // ---------------- epilogue: thread = pixel, one half of the alphas, per-alpha sums in registers
        const int q = warp & 3, half = (warp - 8) >> 2;
        const int ntile = half ? nt_b : nt_a;
        const int cb = half ? NA_a : 0;
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16) + C_R + (uint32_t)cb;
        float acc[NH16][16];
#pragma unroll
        for (int k = 0; k < NH16; ++k)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[k][e] = 0.f;
        if (ntile > 0) {
            for (int t = 0; t < ntiles; ++t) {
                mbar_wait_guard(&bars[B_RFULL + half], (uint32_t)(t & 1));
                tc_fence_after();
                uint32_t rbuf[2][16];
                tmem_ld16(tl, rbuf[0]);
#pragma unroll
                for (int k = 0; k < NH16; ++k) {
                    if (k < ntile) {
                        tc_wait_ld();
                        if (k + 1 < NH16 && k + 1 < ntile) tmem_ld16(tl + 16 * (k + 1), rbuf[(k + 1) & 1]);
                        float r[16], u[16], umax = 0.f;
                        const float4* bp = reinterpret_cast<const float4*>(beta_s + cb + 16 * k);
#pragma unroll
                        for (int v4 = 0; v4 < 4; ++v4) {
                            const float4 b = bp[v4];
                            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                r[4 * v4 + e] = __uint_as_float(rbuf[k & 1][4 * v4 + e]);
                                u[4 * v4 + e] = bb[e] * r[4 * v4 + e];
                                umax = fmaxf(umax, fabsf(u[4 * v4 + e]));
                            }
                        }
                        if (!(umax == umax)) umax = 1.0f;                   // NaN -> slow path -> poison
                        const unsigned big = __ballot_sync(0xffffffffu, umax > 0x1p-6f);
                        const float4* dp = dtab + cb + 16 * k;
                        if (big == 0u) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const float4 d = dp[e];                     // d2..d5 of this alpha (broadcast)
                                const float pz = fmaf(fmaf(fmaf(d.w, u[e], d.z), u[e], d.y), u[e], d.x);
                                acc[k][e] = fmaf(u[e] * u[e], pz, acc[k][e]);
                            }
                        } else {
                            float g[16], rc[16];                // copies keep r itself in registers
#pragma unroll
                            for (int e = 0; e < 16; ++e) rc[e] = r[e];
                            g_cold16(rc, beta_s + cb + 16 * k, dp, g);
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[k][e] += g[e];
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars[B_REMPTY + half]);
            }
        }
        // every x tile has been consumed by now: the ring doubles as the reduction scratch [8 warps][NH16*16].
        // Butterfly with halving: after the five exchanges lane l holds the 32-lane total of value l.
        double* red = reinterpret_cast<double*>(ring) + (warp - 8) * (NH16 * 16);
#pragma unroll
        for (int g0 = 0; g0 < NH16 * 16; g0 += 32) {
            constexpr int kTot = NH16 * 16;
            double v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = (g0 + i < kTot) ? (double)acc[(g0 + i) >> 4][(g0 + i) & 15] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const bool up = (lane & o) != 0;
#pragma unroll
                for (int i = 0; i < o; ++i) {
                    const double send = up ? v[i] : v[i + o], keep = up ? v[i + o] : v[i];
                    v[i] = keep + shfl_xor_f64(send, o);
                }
            }
            if (g0 + lane < kTot) red[g0 + lane] = v[0];
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
    const double* red = reinterpret_cast<const double*>(ring);
    for (int i = tid; i < na_out; i += blockDim.x) {
        const int half = (i >= NA_a) ? 1 : 0, j = i - (half ? NA_a : 0);
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < 4; ++w) a += red[(4 * half + w) * (NH16 * 16) + j];
        out[i] = a;
    }
}

#ifdef CMF_TUNING_HOOKS
// ------------------------------------------------------------------ self test (cmf_microbench kinds 20..23)
// One 128 x N x K contraction with A written to TMEM by tcgen05.st and B in the canonical shared-memory
// layout: checks the descriptor encoding, the TMEM A layout and the ld/st lane mapping against the host.
__global__ void __launch_bounds__(128, 1)
    tc5_selftest_kernel(const float* __restrict__ A, const float* __restrict__ Bt, int N, int K, int row_off,
                        int swap_lbo_sbo, float* __restrict__ Dout) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    const int KC = K / 4, NB = N + row_off;                      // the table holds NB rows, the MMA uses rows row_off..
    float* bs = reinterpret_cast<float*>(smem_raw);              // [KC][NB][4]
    for (int idx = tid; idx < KC * NB * 4; idx += blockDim.x) {
        const int e = idx & 3, n = (idx >> 2) % NB, c = (idx >> 2) / NB;
        bs[idx] = Bt[n * K + 4 * c + e];
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
    for (int c = 0; c < K / 8; ++c) {
        uint32_t v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(A[tid * K + 8 * c + e]);
        tmem_st8(tl + 8 * c, v);
    }
    tc_wait_st();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) {
        const uint32_t lbo = (uint32_t)NB * 16, sbo = 128;
        const uint32_t id = idesc_tf32(N);
        for (int ks = 0; ks < K / 8; ++ks) {
            const uint32_t addr = smem_u32(bs) + 2 * ks * lbo + row_off * 16;
            const uint64_t d = swap_lbo_sbo ? smem_desc(addr, sbo, lbo) : smem_desc(addr, lbo, sbo);
            mma_ts_tf32(tmem + 256, tmem + 8 * ks, d, id, ks > 0);
        }
        tc_commit(&bar);
    }
    mbar_wait_guard(&bar, 0);
    tc_fence_after();
    for (int c = 0; c < N / 16; ++c) {
        uint32_t v[16];
        tmem_ld16(tl + 256 + 16 * c, v);
        tc_wait_ld();
#pragma unroll
        for (int e = 0; e < 16; ++e) Dout[tid * N + 16 * c + e] = __uint_as_float(v[e]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

#endif  // CMF_TUNING_HOOKS

}  // namespace

// ------------------------------------------------------------------ launchers
// alpha parts (CTAs per column and chunk along the alphas) the screen needs for this window: 1 up to 72 bands, 2 up to
// 88 bands (tables 146 KB + ring 68 KB of shared memory, 472 TMEM columns); 0 = does not fit
static int screen5_parts(const Dims& d) {
    if (d.NT < 1 || d.NT16 < 1 || d.NT16 > 14) return 0;
    for (int parts = 1; parts <= 2; ++parts) {
        if (parts == 1 && d.NT > 9) continue;
        if (parts == 2 && (d.NT < 10 || d.NT > 11)) continue;
        const int nt16c = (d.NT16 + parts - 1) / parts;
        const int nh16 = parts == 1 ? 7 : 4;
        if (nt16c > 2 * nh16) continue;
        const T5Plan p = t5_plan(d.NT, nt16c);
        if (3 * p.DP + p.N1 + p.NA > 512) continue;
        // reduction scratch (8 warps x nh16 x 16 doubles) reuses the ring
        if ((size_t)kT5Stages * kT5HalfRows * p.DP * 4 < (size_t)8 * nh16 * 16 * sizeof(double)) continue;
        if (p.total <= 227u * 1024u) return parts;
    }
    return 0;
}

bool screen5_supported(const Dims& d) { return screen5_parts(d) > 0; }

size_t screen5_table_floats(const Dims& d) {
    const int parts = std::max(1, screen5_parts(d));
    return (size_t)parts * (t5_plan(d.NT, (d.NT16 + parts - 1) / parts).tab_bytes / 4);
}

int screen5_lines_per_chunk(const Dims& d, int nchunk) {
    int lpc = (d.L + nchunk - 1) / nchunk;
    return (lpc + 127) / 128 * 128;
}

template <int NT, int NH16>
static void launch_screen5_t(const Dims& d, int parts, const float* xt, const double* mu, const float* tab,
                             const float* betaf, const int* n, int nchunk, double* fscreen, cudaStream_t st) {
    const int nt16c = (d.NT16 + parts - 1) / parts;
    const T5Plan p = t5_plan(d.NT, nt16c);
    cudaFuncSetAttribute(loo_screen5_kernel<NT, NH16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.total);
    dim3 grid(d.S, nchunk, parts);
    loo_screen5_kernel<NT, NH16><<<grid, kT5Threads, p.total, st>>>(xt, mu, tab, betaf, n, d.L, nt16c, d.AP16,
                                                                    screen5_lines_per_chunk(d, nchunk), fscreen,
                                                                    d.nrows);
}

void launch_screen5(const Dims& d, const float* xt, const double* mu, const int* n, const int* nloo,
                    const double* alphas, const double* P, const double* lam, float* tab, float* betaf,
                    int nchunk, double* fscreen, cudaStream_t st) {
    const int parts = screen5_parts(d);
    if (parts < 1) return;
    const int nt16c = (d.NT16 + parts - 1) / parts;
    screen5_tables_kernel<<<dim3(d.S, parts), 256, 0, st>>>(n, nloo, alphas, d.A, d.D, d.NT, nt16c, parts, d.AP16, P, lam,
                                                            tab, betaf);
    switch (d.NT) {
#define CMF_CASE(k) case k: launch_screen5_t<k, 7>(d, parts, xt, mu, tab, betaf, n, nchunk, fscreen, st); break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6) CMF_CASE(7) CMF_CASE(8) CMF_CASE(9)
#undef CMF_CASE
        case 10: launch_screen5_t<10, 4>(d, parts, xt, mu, tab, betaf, n, nchunk, fscreen, st); break;
        case 11: launch_screen5_t<11, 4>(d, parts, xt, mu, tab, betaf, n, nchunk, fscreen, st); break;
        default: break;
    }
}

#ifdef CMF_TUNING_HOOKS
// max |D - A.B^T| of one tcgen05 TS-form contraction against the host; < 0 on a CUDA error
double screen5_selftest(int N, int K, int row_off, int swap_lbo_sbo) {
    const int NB = N + row_off;
    std::vector<float> A(128 * K), B(NB * K), D(128 * N);
    uint32_t seed = 12345u;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return (float)((int)((seed >> 16) & 127) - 64) / 64.0f; };
    for (auto& v : A) v = rnd();
    for (auto& v : B) v = rnd();
    float *dA, *dB, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xff, D.size() * 4);
    const size_t smem = (size_t)(K / 4) * NB * 16;
    cudaFuncSetAttribute(tc5_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc5_selftest_kernel<<<1, 128, smem>>>(dA, dB, N, K, row_off, swap_lbo_sbo, dD);
    cudaError_t e = cudaDeviceSynchronize();
    double err = -1.0;
    if (e == cudaSuccess) {
        cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
        err = 0.0;
        for (int m = 0; m < 128; ++m)
            for (int nn = 0; nn < N; ++nn) {
                double ref = 0.0;
                for (int k = 0; k < K; ++k) ref += (double)A[m * K + k] * (double)B[(nn + row_off) * K + k];
                const double dv = fabs((double)D[m * N + nn] - ref);
                if (!(dv <= err)) err = dv;
            }
    } else {
        fprintf(stderr, "screen5_selftest: %s\n", cudaGetErrorString(e));
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
    return err;
}

#endif  // CMF_TUNING_HOOKS

}  // namespace cmf
