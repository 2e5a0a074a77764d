// K3a on the 5th-generation tensor cores: the screening pass of the leave-one-out alpha search
// (cmf/robust_mf.py:105-117) as two chained tcgen05.mma contractions with TMEM accumulators.
//
// Same arithmetic as k_screen.cu (which stays as the path for windows that do not fit TMEM):
//     GEMM1  Y = Xc . P                 3 x TF32 (xh.Ph + xl.Ph + xh.Pl), FP32 accumulate in TMEM
//     GEMM2  U = (Y*Y) . (beta W)       3 x TF32 (zh.Wh + zh.Wl + zl.Wh), FP32 accumulate in TMEM; u = beta r
//     h(r)   = log(1-u) + r u/(1-u), FP32 series in u (packed FP32 pairs), summed per alpha
// but organised for the Blackwell tensor pipe instead of per-warp mma.sync fragments:
//   * one CTA per (column, line chunk), 128-pixel tiles: the pixel is the TMEM lane (M = 128),
//   * the A operand never touches shared memory: the converting threads write xh|xl straight into
//     TMEM with tcgen05.st, GEMM1 leaves Y in TMEM, the squaring threads read it with tcgen05.ld and
//     write zh|zl back, GEMM2 takes them as its TMEM A operand (tcgen05.mma "TS" form),
//   * the B operands (P and W, hi and lo TF32 parts) sit in shared memory for the whole CTA in the
//     canonical K-major no-swizzle core-matrix layout, built once per column by screen5_tables_kernel,
//   * R is split into two alpha halves with their own full/empty barriers so that the FP32 epilogue
//     of one half overlaps the tensor work of the other half and of the next tile; the squares reach GEMM2 in
//     three parts (k-steps), so the tensor pipe does not wait for the whole squaring pass,
//   * the MMA issue path is uniform: bases through a lane-0 shuffle, descriptor arithmetic on the 32-bit start
//     address word, k-steps unrolled (about two uniform instructions per tcgen05.mma),
//   * warp roles: warps 0-7 epilogue (thread = pixel, 112 alphas each, accumulators in registers), warps 8-11
//     convert and square (thread = pixel), warp 12 bulk-copy producer, warp 13 MMA issuer (one elected lane);
//     setmaxnreg moves registers from the control warps to the epilogue warps.
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

#include <type_traits>
#include <vector>

#include "cmf_common.cuh"
#include "cmf_internal.h"
#include "cmf_tc5.cuh"

namespace cmf {

namespace {

constexpr int kT5Threads = 512;
// Warp roles.  The schedulers favour the higher warp of the warps that are ready (measured: with the issuer in warp 1 it
// waited several hundred cycles behind the epilogue warps of its scheduler before every GEMM), so the latency-critical
// roles sit at the top: warps 0-7 epilogue (two alpha halves x four lane quarters), 8-11 convert / square, 12 bulk-copy
// producer, 13 MMA issuer.
constexpr int kT5WarpProducer = 12, kT5WarpIssuer = 13;
// register budgets after setmaxnreg (4 control warps, 4 convert/square warps, 8 epilogue warps; the sum
// 4 a + 4 b + 8 c may not exceed 2048): the epilogue keeps 112 per-alpha accumulators per thread
#ifndef CMF_S5_REGS_CTL
#define CMF_S5_REGS_CTL 24
#define CMF_S5_REGS_CVT 72
#define CMF_S5_REGS_EPI 208
#endif
#ifndef CMF_S5_ZP
#define CMF_S5_ZP 3                 // hand-over parts of the squares (1..3)
#endif
#define CMF_STR2(x) #x
#define CMF_STR(x) CMF_STR2(x)
#ifdef CMF_TUNING_HOOKS
// per-phase timeline of one CTA (tools build): SM clock at the hand-offs of tiles 6..9 (tools/s5_timeline.py)
__device__ long long g_s5_tl[4][32];
#define S5_TL(cond, tile, ev)                                                                             \
    do {                                                                                                  \
        if ((cond) && lane == 0 && s == 5 && chunk == 1 && blockIdx.z == 0 && (tile) >= 6 && (tile) < 10) \
            g_s5_tl[(tile) - 6][ev] = clock64();                                                          \
    } while (0)
#define S5_CTA(cond, ev)                                                         \
    do {                                                                         \
        if ((cond) && blockIdx.x == 5 && blockIdx.y == 1 && blockIdx.z == 0)     \
            g_s5_tl[3][ev] = clock64();                                          \
    } while (0)
#else
#define S5_TL(cond, tile, ev) do {} while (0)
#define S5_CTA(cond, ev) do {} while (0)
#endif
constexpr int kT5Stages = 3;       // 64-row half tiles in flight
constexpr int kT5HalfRows = 64;

// The epilogue sums g = h + u per alpha, where h = log(1-u) + r u/(1-u) and u = beta r (log q + r/q = r + h):
//     g = sum_{k>=2} u^k (1/beta - 1/k)
// sum_k u_k = beta sum_k r_k is known in closed form (rsum), so K3b subtracts it in FP64.  d_k = 1/beta - 1/k
// are per-alpha constants: d2 comes from a shared-memory table, the others from it (terms to u^5: |u| <= 2^-6), k <= 10
// (|u| <= 2^-3) are formed on the fly; anything larger is the rare out-of-line slow path.
// Everything that is not the 5-term fast path, out of line so that the hot loop stays small: 10 terms while
// |u| <= 2^-3 (u^11 < 2^-33), log1p and a division up to u = 1/4, poison beyond.
// up: u = beta r of these 16 alphas (GEMM2 delivers it, beta is folded into W), dp: d2 = 1/beta - 1/2
__device__ __noinline__ void g_cold16(const float* __restrict__ up, const float* __restrict__ dp, float* __restrict__ g) {
    float u[16], umax = 0.f;
#pragma unroll
    for (int e = 0; e < 16; ++e) { u[e] = up[e]; umax = fmaxf(umax, fabsf(u[e])); }
    if (!(umax == umax)) umax = 1.0f;
    if (__ballot_sync(0xffffffffu, umax > 0x1p-3f) == 0u) {
        float pz[16], binv[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) { binv[e] = dp[e] + 0.5f; pz[e] = binv[e] - 0.1f; }
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
            const float uu = up[e], au = fabsf(uu), binv = dp[e] + 0.5f;
            // beyond u = 1/4 the term amplifies the TF32 error of r by u/(1-u) and more: such a pixel is an
            // extreme outlier (x'G^-1 x > n/4); poison the sum so that the column is searched in FP64.
            // r u / (1 - u) with r = u / beta
            g[e] = (au <= 0.25f) ? (log1pf(-uu) + binv * uu * uu / (1.0f - uu)) + uu : __int_as_float(0x7fc00000);
        }
    }
}

// ------------------------------------------------------------------ shared-memory plan
struct T5Plan {
    int DP, N1, NA, KC;
    uint32_t ph_off, pl_off, wh_off, wl_off, tab_bytes;      // B operand tables (one bulk copy)
    uint32_t ring_off, mu_off, dtab_off, bar_off, total;
};

__host__ __device__ inline T5Plan t5_plan(int NT, int NT16) {
    T5Plan p;
    p.DP = 8 * NT; p.N1 = (p.DP + 15) / 16 * 16; p.NA = 16 * NT16; p.KC = p.DP / 4;
    const uint32_t pbytes = (uint32_t)p.KC * p.N1 * 16, wbytes = (uint32_t)p.KC * p.NA * 16;
    p.ph_off = 0; p.pl_off = pbytes; p.wh_off = 2 * pbytes; p.wl_off = 2 * pbytes + wbytes;
    p.tab_bytes = 2 * pbytes + 2 * wbytes;
    p.ring_off = p.tab_bytes;
    p.mu_off = p.ring_off + (uint32_t)kT5Stages * kT5HalfRows * p.DP * 4;
    p.dtab_off = (p.mu_off + (uint32_t)p.DP * 8 + 15u) & ~15u;         // d2 = 1/beta - 1/2 per alpha
    p.bar_off = p.dtab_off + (uint32_t)p.NA * 4;
    p.total = p.bar_off + 24 * 8;
    return p;
}

// B_ZREADY: the squares are handed over in up to three parts (k-steps of GEMM2), so that GEMM2 starts on the first
// third of zh|zl while the rest is still being squared
enum { B_TAB = 0, B_XFULL = 1, B_XEMPTY = 4, B_XREADY = 7, B_G1 = 8, B_RFULL = 10, B_REMPTY = 12, B_ZREADY = 14,
       B_COUNT = 17 };

// ------------------------------------------------------------------ B operand tables
// tab[s] = Ph | Pl | Wh | Wl, each [K chunk c][row][4]: P rows are eigen-directions j (K = band b),
// W rows are alphas i (K = eigen-direction j), scaled by beta_i so that GEMM2 delivers u = beta r directly;
// hi = tf32(v), lo = tf32(v - hi).
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
            w = be / (dn * be * lam[j] + al);                  // beta folded in: GEMM2 delivers u = beta r
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
    constexpr int ZP = (NT >= 6) ? CMF_S5_ZP : 1, ZC = (NT + ZP - 1) / ZP;    // hand-over parts of the squares, k-steps per part
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
    S5_CTA(tid == 0, 24);
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
    float* muh = reinterpret_cast<float*>(smem_raw + p.mu_off);        // column mean, FP32 head and tail, [DP] each
    float* mul = muh + DP;
    float* dtab = reinterpret_cast<float*>(smem_raw + p.dtab_off);     // d2 per alpha
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + p.bar_off);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    if (tid == 0) {
        mbar_init(&bars[B_TAB], 1);
        for (int i = 0; i < kT5Stages; ++i) { mbar_init(&bars[B_XFULL + i], 1); mbar_init(&bars[B_XEMPTY + i], 2); }
        mbar_init(&bars[B_XREADY], 128);
        mbar_init(&bars[B_G1], 1);
        for (int i = 0; i < 3; ++i) mbar_init(&bars[B_ZREADY + i], 128);
        mbar_init(&bars[B_RFULL], 1); mbar_init(&bars[B_RFULL + 1], 1);
        mbar_init(&bars[B_REMPTY], 128); mbar_init(&bars[B_REMPTY + 1], 128);
        fence_mbar_init();
    }
    if (warp == kT5WarpIssuer) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = tid; i < DP; i += blockDim.x) {
        const double m = mu_g[(long long)s * DP + i];
        const float mh = (float)m;
        muh[i] = mh;
        mul[i] = (float)(m - (double)mh);
    }
    for (int i = tid; i < NA; i += blockDim.x) {
        const float b = (i < na_out) ? betaf_g[(long long)s * AP16 + a_off + i] : 0.f;
        const float binv = 1.0f / b;
        // d_k = 1/beta - 1/k, k = 2..5: d2 comes from the table, d3..d5 = d2 + (1/2 - 1/k) are formed in the epilogue.
        // beta == 0 (alpha == 1, padding): u == 0 and every term vanishes whatever d2 is; keep it finite
        dtab[i] = (b > 0.f) ? binv - 0.5f : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    S5_CTA(tid == 0, 25);

    const int wg = warp >> 2;
    if (wg == 3) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_CTL) ";");
        if (warp == kT5WarpProducer && lane == 0) {
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
        } else if (warp == kT5WarpIssuer) {
            // ---------------- MMA issuer: the whole warp runs the loop, one elected lane issues (see elect_one()).
            // The bases go through a lane-0 shuffle so that the compiler knows them warp-uniform: descriptors and TMEM
            // addresses are then formed on the uniform datapath, one add per operand with the k-steps unrolled
            // (the rolled loop with per-thread bases spent ~12 instructions and ~65 cycles per MMA on R2UR and
            // descriptor encoding, more than the tensor pipe needs for it).
            const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem_raw), 0);
            const uint32_t lbo1 = N1 * 16, lbo2 = (uint32_t)NA * 16;
            const uint32_t id1 = idesc_tf32(N1), id2a = idesc_tf32(NA_a), id2b = idesc_tf32(NA_b > 0 ? NA_b : 16);
            // descriptor of k-step ks = descriptor of k-step 0 + ks * (2 K chunks of LBO bytes, in 16-byte units), low word
            const uint32_t dhi = (uint32_t)(smem_desc(0, 0, 128) >> 32);
            const uint32_t st1 = (2 * lbo1) >> 4, st2 = (2 * lbo2) >> 4;
            mbar_wait_guard(&bars[B_TAB], 0);
            S5_CTA(lane == 0, 26);
            for (int t = 0; t < ntiles; ++t) {
                const uint32_t ph = (uint32_t)(t & 1);
                // formed per tile (a handful of uniform adds) rather than kept live across the loop
                uint32_t sb = sbase;
                asm volatile("" : "+r"(sb));
                sb = __shfl_sync(0xffffffffu, sb, 0);
                uint32_t tm = tmem;
                asm volatile("" : "+r"(tm));
                tm = __shfl_sync(0xffffffffu, tm, 0);
                const uint32_t d1h = (uint32_t)smem_desc(sb + p.ph_off, lbo1, 128), d1l = (uint32_t)smem_desc(sb + p.pl_off, lbo1, 128);
                const uint32_t d2h = (uint32_t)smem_desc(sb + p.wh_off, lbo2, 128), d2l = (uint32_t)smem_desc(sb + p.wl_off, lbo2, 128);
                const uint32_t d2hb = d2h + (uint32_t)NA_a, d2lb = d2l + (uint32_t)NA_a;
                S5_TL(true, t, 8);
                mbar_wait_guard(&bars[B_XREADY], ph);
                S5_TL(true, t, 9);
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < NT; ++ks) {
                        mma_ts_tf32_w(tm + C_Y, tm + C_XH + 8 * ks, d1h + ks * st1, dhi, id1, ks > 0);
                        mma_ts_tf32_w(tm + C_Y, tm + C_XL + 8 * ks, d1h + ks * st1, dhi, id1, 1);
                        mma_ts_tf32_w(tm + C_Y, tm + C_XH + 8 * ks, d1l + ks * st1, dhi, id1, 1);
                    }
                    tc_commit(&bars[B_G1]);
                }
                __syncwarp();
                S5_TL(true, t, 10);
                if (t > 0) mbar_wait_guard(&bars[B_REMPTY], ph ^ 1u);
                S5_TL(true, t, 12);
#pragma unroll
                for (int zp = 0; zp < ZP; ++zp) {
                    if (zp * ZC < NT) {
                        mbar_wait_guard(&bars[B_ZREADY + zp], ph);
                        tc_fence_after();
                        if (zp == 0) S5_TL(true, t, 11);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = zp * ZC; ks < (zp + 1) * ZC && ks < NT; ++ks) {
                                mma_ts_tf32_w(tm + C_R, tm + C_Y + 8 * ks, d2h + ks * st2, dhi, id2a, ks > 0);
                                mma_ts_tf32_w(tm + C_R, tm + C_Y + 8 * ks, d2l + ks * st2, dhi, id2a, 1);
                                mma_ts_tf32_w(tm + C_R, tm + C_ZL + 8 * ks, d2h + ks * st2, dhi, id2a, 1);
                            }
                            if ((zp + 1) * ZC >= NT) tc_commit(&bars[B_RFULL]);
                        }
                        __syncwarp();
                    }
                }
                S5_TL(true, t, 13);
                if (NA_b > 0) {
                    if (t > 0) { mbar_wait_guard(&bars[B_REMPTY + 1], ph ^ 1u); tc_fence_after(); }
                    S5_TL(true, t, 14);
                    if (elect_one()) {
#pragma unroll
                        for (int ks = 0; ks < NT; ++ks) {
                            mma_ts_tf32_w(tm + C_R + NA_a, tm + C_Y + 8 * ks, d2hb + ks * st2, dhi, id2b, ks > 0);
                            mma_ts_tf32_w(tm + C_R + NA_a, tm + C_Y + 8 * ks, d2lb + ks * st2, dhi, id2b, 1);
                            mma_ts_tf32_w(tm + C_R + NA_a, tm + C_ZL + 8 * ks, d2hb + ks * st2, dhi, id2b, 1);
                        }
                        tc_commit(&bars[B_RFULL + 1]);
                    }
                    __syncwarp();
                    S5_TL(true, t, 15);
                }
            }
        }
        __syncwarp();
    } else if (wg == 2) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_CVT) ";");
        // ---------------- convert (x -> xh|xl) and square (y -> zh|zl): thread = pixel = TMEM lane
        const int q = warp & 3, row = 32 * q + lane, myhalf = q >> 1, rin = row & (kT5HalfRows - 1);
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        // loops stay rolled: one warp per scheduler runs this code, so its footprint has to sit in the
        // instruction cache (the unrolled first version spent most of its time in instruction fetch)
        auto square8 = [&](uint32_t (&y)[8], int c) {
            // packed pairs for the FP32 part; zl goes to TMEM unrounded (the tensor core reads the TF32 bits only)
            uint32_t lo[8];
#pragma unroll
            for (int e = 0; e < 8; e += 2) {
                const uint64_t yy = f2_pack_u(y[e], y[e + 1]);
                const uint64_t z = f2_mul(yy, yy);
                float z0, z1, l0, l1;
                f2_unpack(z, z0, z1);
                y[e] = tf32_round(z0);
                y[e + 1] = tf32_round(z1);
                f2_unpack(f2_sub(z, f2_pack_u(y[e], y[e + 1])), l0, l1);
                lo[e] = __float_as_uint(l0);
                lo[e + 1] = __float_as_uint(l1);
            }
            tmem_st8(tl + C_Y + 8 * c, y);
            tmem_st8(tl + C_ZL + 8 * c, lo);
        };
#pragma unroll 1
        for (int t = 0; t <= ntiles; ++t) {
            if (t > 0) {
                // ---- y -> zh | zl of tile t-1
                S5_TL(q == 0, t - 1, 0);
                mbar_wait_guard(&bars[B_G1], (uint32_t)((t - 1) & 1));
                S5_TL(q == 0, t - 1, 1);
                tc_fence_after();
                // one hand-over part (ZC k-steps) per stage, double-buffered: the loads of the next part are in flight
                // while this part is squared, so the TMEM load latency shows once per tile, not once per k-step
                uint32_t yb[2][ZC][8];
#pragma unroll
                for (int i = 0; i < ZC; ++i)
                    if (i < NT) tmem_ld8(tl + C_Y + 8 * i, yb[0][i]);
#pragma unroll
                for (int zp = 0; zp < ZP; ++zp) {
                    if (zp * ZC < NT) {
                        tc_wait_ld();
#pragma unroll
                        for (int i = 0; i < ZC; ++i)
                            if ((zp + 1) * ZC + i < NT) tmem_ld8(tl + C_Y + 8 * ((zp + 1) * ZC + i), yb[(zp + 1) & 1][i]);
#pragma unroll
                        for (int i = 0; i < ZC; ++i)
                            if (zp * ZC + i < NT) square8(yb[zp & 1][i], zp * ZC + i);
                        tc_wait_st();
                        tc_fence_before();
                        mbar_arrive(&bars[B_ZREADY + zp]);
                    }
                }
                S5_TL(q == 0, t - 1, 2);
            }
            if (t < ntiles) {
                // ---- x -> xh | xl of tile t
                const int h = 2 * t + myhalf, slot = h % kT5Stages, use = h / kT5Stages;
                const bool have = h < nhalf;
                S5_TL(q == 0, t, 3);
                if (have) mbar_wait_guard(&bars[B_XFULL + slot], (uint32_t)(use & 1));
                S5_TL(q == 0, t, 4);
                const float* xrow = ring + (slot * kT5HalfRows + rin) * DP;
                // a dropped pixel is NaN in every band of xt (repack), so band 0 tells; rows past the chunk end and
                // dropped pixels become exact zeros through the masks of the TF32 split (NaN & 0 == 0)
                const float x0 = xrow[0];                                   // stale but in-bounds when !have
                const bool row_ok = have && (128 * t + row < nrows) && x0 == x0;
                const ulonglong2* src = reinterpret_cast<const ulonglong2*>(xrow);
                const ulonglong2* mh2 = reinterpret_cast<const ulonglong2*>(muh);
                const ulonglong2* ml2 = reinterpret_cast<const ulonglong2*>(mul);
                auto convert = [&](auto masked) {
                    constexpr bool MASKED = decltype(masked)::value;
                    const uint32_t mask = (!MASKED || row_ok) ? 0xffffe000u : 0u, keep = row_ok ? 0xffffffffu : 0u;
#pragma unroll 1
                    for (int c = 0; c < NT; ++c) {
                        uint32_t hi[8], lo[8];
#pragma unroll
                        for (int h2 = 0; h2 < 2; ++h2) {
                            const ulonglong2 x = src[2 * c + h2], mh = mh2[2 * c + h2], ml = ml2[2 * c + h2];
                            const uint64_t va = f2_sub(f2_sub(x.x, mh.x), ml.x), vb = f2_sub(f2_sub(x.y, mh.y), ml.y);
                            float v0, v1, v2, v3, l0, l1, l2, l3;
                            f2_unpack(va, v0, v1);
                            f2_unpack(vb, v2, v3);
                            hi[4 * h2 + 0] = (__float_as_uint(v0) + 0x1000u) & mask;
                            hi[4 * h2 + 1] = (__float_as_uint(v1) + 0x1000u) & mask;
                            hi[4 * h2 + 2] = (__float_as_uint(v2) + 0x1000u) & mask;
                            hi[4 * h2 + 3] = (__float_as_uint(v3) + 0x1000u) & mask;
                            // xl goes to TMEM unrounded: the tensor core reads the TF32 bits only
                            f2_unpack(f2_sub(va, f2_pack_u(hi[4 * h2 + 0], hi[4 * h2 + 1])), l0, l1);
                            f2_unpack(f2_sub(vb, f2_pack_u(hi[4 * h2 + 2], hi[4 * h2 + 3])), l2, l3);
                            lo[4 * h2 + 0] = MASKED ? (__float_as_uint(l0) & keep) : __float_as_uint(l0);
                            lo[4 * h2 + 1] = MASKED ? (__float_as_uint(l1) & keep) : __float_as_uint(l1);
                            lo[4 * h2 + 2] = MASKED ? (__float_as_uint(l2) & keep) : __float_as_uint(l2);
                            lo[4 * h2 + 3] = MASKED ? (__float_as_uint(l3) & keep) : __float_as_uint(l3);
                        }
                        tmem_st8(tl + C_XH + 8 * c, hi);
                        tmem_st8(tl + C_XL + 8 * c, lo);
                    }
                };
                // the masks cost one operation per element: only warps that hold a dropped pixel or a row past the end pay
                if (__all_sync(0xffffffffu, row_ok)) convert(std::false_type{});
                else convert(std::true_type{});
                tc_wait_st();
                if (have) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bars[B_XEMPTY + slot]);
                }
                tc_fence_before();
                mbar_arrive(&bars[B_XREADY]);
                S5_TL(q == 0, t, 5);
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 " CMF_STR(CMF_S5_REGS_EPI) ";");
        // ---------------- epilogue: thread = pixel, one half of the alphas, per-alpha sums in registers
        const int q = warp & 3, half = warp >> 2;
        const int ntile = half ? nt_b : nt_a;
        const int cb = half ? NA_a : 0;
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16) + C_R + (uint32_t)cb;
        uint64_t acc[NH16][8];                                    // per-alpha sums, two alphas per 64-bit register
#pragma unroll
        for (int k = 0; k < NH16; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[k][e] = 0ull;
        if (ntile > 0) {
            for (int t = 0; t < ntiles; ++t) {
                S5_TL(q == 0, t, 16 + 4 * half);
                mbar_wait_guard(&bars[B_RFULL + half], (uint32_t)(t & 1));
                S5_TL(q == 0, t, 17 + 4 * half);
                tc_fence_after();
                uint32_t rbuf[2][16];
                tmem_ld16(tl, rbuf[0]);
#pragma unroll
                for (int k = 0; k < NH16; ++k) {
                    if (k < ntile) {
                        tc_wait_ld();
                        if (k + 1 < NH16 && k + 1 < ntile) tmem_ld16(tl + 16 * (k + 1), rbuf[(k + 1) & 1]);
                        // GEMM2 delivers u = beta r (beta is folded into W).  Packed FP32 pairs throughout:
                        //   g = u^2 (d2 + d3 u + d4 u^2 + d5 u^3),  d_k = 1/beta - 1/k = d2 + (1/2 - 1/k)
                        //     = u^2 (d2 (1 + u)(1 + u^2) + u/6)     [+ u^4/4 + ...: u^2 / (4 d2) of g, < 2e-8 for |u| <= 2^-6 at
                        //     flightline size (d2 > 3e3), < 6e-7 for a 100-pixel column]
                        // one broadcast 16-byte read of d2 per four alphas, after the range check
                        uint64_t u[8];
                        float umax = 0.f;
                        const float* dp = dtab + cb + 16 * k;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            u[e] = f2_pack_u(rbuf[k & 1][2 * e], rbuf[k & 1][2 * e + 1]);
                            umax = fmaxf(umax, fmaxf(fabsf(__uint_as_float(rbuf[k & 1][2 * e])),
                                                     fabsf(__uint_as_float(rbuf[k & 1][2 * e + 1]))));
                        }
                        if (!(umax == umax)) umax = 1.0f;                   // NaN -> slow path -> poison
                        const unsigned big = __ballot_sync(0xffffffffu, umax > 0x1p-6f);
                        if (big == 0u) {
                            const uint64_t one = f2_pack(1.0f, 1.0f), c3 = f2_pack(1.0f / 6.0f, 1.0f / 6.0f);
                            const ulonglong2* d2p = reinterpret_cast<const ulonglong2*>(dp);
#pragma unroll
                            for (int v4 = 0; v4 < 4; ++v4) {
                                const ulonglong2 dd = d2p[v4];
#pragma unroll
                                for (int h2 = 0; h2 < 2; ++h2) {
                                    const int e = 2 * v4 + h2;
                                    const uint64_t d2 = h2 ? dd.y : dd.x;
                                    const uint64_t uu = f2_mul(u[e], u[e]);
                                    const uint64_t a1 = f2_add(u[e], one);
                                    const uint64_t s4 = f2_fma(uu, a1, a1);                          // (1 + u)(1 + u^2)
                                    acc[k][e] = f2_fma(uu, f2_fma(d2, s4, f2_mul(c3, u[e])), acc[k][e]);
                                }
                            }
                        } else {
                            float g[16], uc[16];
#pragma unroll
                            for (int e = 0; e < 16; ++e) uc[e] = __uint_as_float(rbuf[k & 1][e]);
                            g_cold16(uc, dp, g);
#pragma unroll
                            for (int e = 0; e < 8; ++e) acc[k][e] = f2_add(acc[k][e], f2_pack(g[2 * e], g[2 * e + 1]));
                        }
                    }
                }
                tc_fence_before();
                mbar_arrive(&bars[B_REMPTY + half]);
                S5_TL(q == 0, t, 18 + 4 * half);
            }
        }
        S5_CTA(warp == 0 && lane == 0, 28);
        // every x tile has been consumed by now: the ring doubles as the reduction scratch [8 warps][NH16*16].
        // Butterfly with halving: after the five exchanges lane l holds the 32-lane total of value l.
        double* red = reinterpret_cast<double*>(ring) + warp * (NH16 * 16);
#pragma unroll
        for (int g0 = 0; g0 < NH16 * 16; g0 += 32) {
            constexpr int kTot = NH16 * 16;
            double v[32];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                float a0 = 0.f, a1 = 0.f;
                if (g0 + i < kTot) f2_unpack(acc[(g0 + i) >> 4][((g0 + i) & 15) >> 1], a0, a1);
                v[i] = (double)a0; v[i + 1] = (double)a1;
            }
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
    S5_CTA(warp == 0 && lane == 0, 29);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    S5_CTA(tid == 0, 30);
    if (warp == kT5WarpIssuer) {
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
    S5_CTA(tid == 0, 31);
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

// Issue rate of the TS-form TF32 MMA in isolation (cmf_microbench kind 40 + N/16): one CTA, `reps` rounds of 27 MMAs
// (9 k-steps x {A0.B0, A1.B0, A0.B1}, the pattern of GEMM1 / GEMM2) into one 128 x N accumulator; cycles per MMA.
__global__ void __launch_bounds__(128, 1) tc5_rate_kernel(int N, int reps, int same_d, long long* out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 18 * N * 8; i += blockDim.x) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
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
    {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const uint32_t tl = tmem + ((uint32_t)(32 * warp) << 16);
        for (int c = 0; c < 18; ++c) tmem_st8(tl + 8 * c, z);
        tc_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t sbase = __shfl_sync(0xffffffffu, smem_u32(smem_raw), 0);
        const uint32_t tm = __shfl_sync(0xffffffffu, tmem, 0);
        const uint32_t lbo = (uint32_t)N * 16;
        const uint64_t d0 = smem_desc(sbase, lbo, 128), d1 = smem_desc(sbase + 9 * 2 * lbo, lbo, 128);
        const uint64_t st = (2 * lbo) >> 4;
        const uint32_t id = idesc_tf32(N);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < 9; ++ks) {
                    const uint32_t dd = tm + 256 + (same_d ? 0 : (ks & 1) * 0);
                    mma_ts_tf32(dd, tm + 8 * ks, d0 + ks * st, id, (r | ks) > 0);
                    mma_ts_tf32(dd, tm + 72 + 8 * ks, d0 + ks * st, id, 1);
                    mma_ts_tf32(dd, tm + 8 * ks, d1 + ks * st, id, 1);
                }
            }
            __syncwarp();
        }
        if (elect_one()) tc_commit(&bar);
        __syncwarp();
        mbar_wait_guard(&bar, 0);
        const long long t1 = clock64();
        if (tid == 0) out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

// Issue interval of FFMA and of the packed FFMA2 (cmf_microbench kinds 60..63): `warps` warps per scheduler, 16
// independent accumulator chains per thread; cycles per warp instruction and scheduler.
template <int PACKED>
__global__ void __launch_bounds__(512, 1) fma_rate_kernel(int iters, float seed, long long* out, float* sink) {
    float a[16];
    uint64_t b[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = seed + i; b[i] = f2_pack(seed + i, seed - i); }
    const float m = seed * 0.5f, c = seed * 0.25f;
    const uint64_t m2 = f2_pack(m, m), c2 = f2_pack(c, c);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (PACKED) b[i] = f2_fma(b[i], m2, c2);
            else a[i] = fmaf(a[i], m, c);
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) { float x, y; f2_unpack(b[i], x, y); r += a[i] + x + y; }
    if (r == 123.456f) sink[0] = r;
    if (threadIdx.x == 0) out[0] = t1 - t0;
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

// Line chunks per column for the screening pass: a CTA costs its 128-pixel tiles plus a fixed start / end (tables,
// pipeline fill, final reduction: about 2.5 tiles, measured with tools/s5_timeline.py), CTAs run in waves of one per SM.
int screen5_pick_chunks(const Dims& d, int sm_count) {
    int best = 1;
    double best_cost = 0.0;
    for (int n = 1; n <= 16; ++n) {
        int lpc = (d.L + n - 1) / n;
        lpc = (lpc + 127) / 128 * 128;
        const int n_eff = (d.L + lpc - 1) / lpc;
        if (n_eff != n) continue;
        const long long ctas = (long long)d.S * n;
        const long long waves = (ctas + sm_count - 1) / sm_count;
        const double cost = (double)waves * (lpc / 128 + 2.5);
        if (n == 1 || cost < best_cost * 0.995) { best_cost = cost; best = n; }
    }
    return best;
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
// the timeline the last screening launch left behind: [4 tiles][32 events] SM clocks (0 = event not reached)
int screen5_timeline(long long* out) {
    return cudaMemcpyFromSymbol(out, g_s5_tl, sizeof(long long) * 4 * 32) == cudaSuccess ? 0 : -1;
}

// cycles per FFMA (packed = 0) or FFMA2 (packed = 1) warp instruction and scheduler with `warps` warps per scheduler
double fma_issue_rate(int packed, int warps) {
    long long* d = nullptr;
    float* sink = nullptr;
    long long h = 0;
    cudaMalloc(&d, sizeof(long long));
    cudaMalloc(&sink, sizeof(float));
    const int iters = 2000;
    if (packed) fma_rate_kernel<1><<<1, 128 * warps>>>(iters, 1.0f, d, sink);
    else fma_rate_kernel<0><<<1, 128 * warps>>>(iters, 1.0f, d, sink);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d); cudaFree(sink);
    return e == cudaSuccess ? (double)h / (16.0 * iters * warps) : -1.0;
}

// cycles per TS-form TF32 MMA (M = 128, K = 8) with N columns, issued back to back by one CTA
double screen5_mma_rate(int N, int reps) {
    long long* d = nullptr;
    long long h = 0;
    cudaMalloc(&d, sizeof(long long));
    const size_t smem = (size_t)18 * N * 32;
    cudaFuncSetAttribute(tc5_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc5_rate_kernel<<<1, 128, smem>>>(N, reps, 1, d);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e == cudaSuccess ? (double)h / (27.0 * reps) : -1.0;
}

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
