// K3a: tensor-core screening of the leave-one-out alpha search, K3b: candidate selection.
//
// The alpha search of looshrinkage (cmf/robust_mf.py:105-127) only needs argmin_i nll_i.  Evaluating
// every (pixel, alpha) term in FP64 costs 72 x 208 FP64 MACs per pixel and on B200 the FP64 tensor rate
// equals the plain FP64 rate (37 vs 33 TFLOP/s), so that contraction is the whole run time.  This pass
// evaluates all alphas with the contraction r = (y*y) . W on the TF32 tensor path instead:
//     GEMM1  Y = Xc . P                 FP64 DMMA (as the exact pass: y carries the cancellation)
//     GEMM2  R^T = W^T . (Y*Y)^T        3 x TF32 mma.m16n8k8: Wh.zh + Wl.zh + Wh.zl, FP32 accumulate;
//                                       every term is positive, so r keeps ~2^-21 relative accuracy
//     h(r)   = log(1-u) + r u/(1-u), u = beta r, in FP32; the dominant part sum_k r_k is known in closed
//            form from the eigenvalues ((n-1) sum_j lam_j W_ji) and is added in FP64 by K3b.
// K3b turns the sums into approximate nll values, takes their minimum and marks every alpha within
// `tol` of it (or any non-finite value) as a candidate; only columns with more than one candidate are
// re-evaluated by the exact FP64 pass (K3), restricted to the 8-alpha tiles that hold candidates.  The
// selected index is therefore the exact FP64 argmin as long as the screening error is below tol / 2.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

constexpr int kScreenWarps = 8;
constexpr int kScreenMT = 2;    // 8-pixel m-tiles per warp step

// h(r) = log(1 - u) + r u / (1 - u), u = beta r  (so that log q + r/q = r + h), as the series
//     h = u (r s1 - s2),  s1 = 1 + u + ... + u^M,  s2 = 1 + u/2 + ... + u^M/(M+1)
// The number of terms is chosen per warp from max|u| so that the 16 evaluations of a tile pair are
// branch-free, independent FMA chains: M = 4 for |u| <= 2^-6 (truncation u^5 < 2^-30), M = 9 for
// |u| <= 2^-3 (u^10 = 2^-30); anything larger takes log1pf and a division.  u >= 1 gives NaN / -inf like
// the reference's log(q) and sends the column to the exact pass.
template <int M>
__device__ __forceinline__ float screen_series(float r, float u) {
    float s1 = 1.0f, s2 = 1.0f / (float)(M + 1);
#pragma unroll
    for (int m = M; m >= 1; --m) {
        s1 = fmaf(u, s1, 1.0f);
        s2 = fmaf(u, s2, 1.0f / (float)m);
    }
    return u * fmaf(r, s1, -s2);
}

__device__ __forceinline__ float screen_exact32(float r, float u) {
    return log1pf(-u) + r * u / (1.0f - u);
}

// sum over the 4 x MT accumulator values of one 16-alpha tile: lanes get (alpha g: e = 0,1), (alpha g+8: e = 2,3)
template <int MT>
__device__ __forceinline__ void screen_terms(const float (&acc)[MT][4], float b_lo, float b_hi, float& f_lo,
                                             float& f_hi) {
    float u[MT][4], umax = 0.f;
#pragma unroll
    for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            u[m][e] = ((e < 2) ? b_lo : b_hi) * acc[m][e];
            umax = fmaxf(umax, fabsf(u[m][e]));
        }
    if (!(umax == umax)) umax = 1.0f;                       // NaN -> poison
    const unsigned big = __ballot_sync(0xffffffffu, umax > 0x1p-6f);
    const unsigned huge = __ballot_sync(0xffffffffu, umax > 0x1p-3f);
    f_lo = 0.f; f_hi = 0.f;
    if (big == 0u) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            f_lo += screen_series<4>(acc[m][0], u[m][0]) + screen_series<4>(acc[m][1], u[m][1]);
            f_hi += screen_series<4>(acc[m][2], u[m][2]) + screen_series<4>(acc[m][3], u[m][3]);
        }
    } else if (huge == 0u) {
#pragma unroll
        for (int m = 0; m < MT; ++m) {
            f_lo += screen_series<9>(acc[m][0], u[m][0]) + screen_series<9>(acc[m][1], u[m][1]);
            f_hi += screen_series<9>(acc[m][2], u[m][2]) + screen_series<9>(acc[m][3], u[m][3]);
        }
    } else {
#pragma unroll
        for (int m = 0; m < MT; ++m)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                // beyond u = 1/4 the term amplifies the TF32 error of r by u/(1-u) and more: such a pixel is an
                // extreme outlier (x'G^-1 x > n/4); poison the sum so that the column is searched in FP64
                const float au = fabsf(u[m][e]);
                const float h = (au <= 0x1p-3f) ? screen_series<9>(acc[m][e], u[m][e])
                              : (au <= 0.25f)   ? screen_exact32(acc[m][e], u[m][e])
                                                : __int_as_float(0x7fc00000);
                if (e < 2) f_lo += h; else f_hi += h;
            }
    }
}

// G1T: GEMM1 also on the TF32 path (3 products, FP32 accumulate) instead of FP64 DMMA; NPROD: 3 = Wh.zh +
// Wl.zh + Wh.zl, 2 = drop the zl term (z kept to 11 bits).
template <int NT, bool G1T, int NPROD>
__global__ void __launch_bounds__(kScreenWarps * 32, 1)
    loo_screen_kernel(const float* __restrict__ xt, const double* __restrict__ mu_g,
                      const double* __restrict__ Pf_g, const float* __restrict__ Ps_g,
                      const float* __restrict__ Ws_g, const float* __restrict__ betaf_g,
                      const int* __restrict__ n_g, int L, int NT16, int lines_per_chunk,
                      double* __restrict__ fscreen, const int* __restrict__ nrows) {
    constexpr int DP = 8 * NT, KS = DP / 4, MT = kScreenMT, TL = 8 * MT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int AP16 = NT16 * 16;
    float4* Wh = reinterpret_cast<float4*>(smem_raw);            // [NT16][NT][32]
    float4* Wl = Wh + NT16 * NT * 32;                            // [NT16][NT][32]
    double* Pf = reinterpret_cast<double*>(Wl + NT16 * NT * 32); // [KS][NT][32]  (G1T: float2 Ph/Pl [NT][NT][32] each)
    double* mu_s = Pf + KS * NT * 32;                            // [DP]
    double* fsm = mu_s + DP;                                     // [warps][AP16]
    float* beta_s = reinterpret_cast<float*>(fsm + kScreenWarps * AP16);   // [AP16]
    float* ring = beta_s + AP16;                                 // [warps][TL*DP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kScreenWarps * TL * DP);   // [warps] + 1

    const int s = blockIdx.x, chunk = blockIdx.y;
    if (n_g[s] < 2) return;                                      // nothing to search (K4 handles n < 2)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q4 = lane & 3;
    const int c_begin = chunk * lines_per_chunk;
    const int c_end = max(c_begin, min(nrows ? min(L, nrows[blockIdx.x]) : L, c_begin + lines_per_chunk));   // compacted mode pass
    const int ntiles = (c_end - c_begin + TL - 1) / TL;
    const float* col_base = xt + (long long)s * L * DP;
    float* mytile = ring + warp * TL * DP;
    uint64_t* mybar = bars + warp;
    uint64_t* tabbar = bars + kScreenWarps;

    if (lane == 0) mbar_init(mybar, 1);
    if (threadIdx.x == 0) mbar_init(tabbar, 1);
    if (lane == 0) fence_mbar_init();
    __syncthreads();

    const uint32_t wbytes = (uint32_t)(2 * NT16 * NT * 32 * sizeof(float4));
    const uint32_t pbytes = (uint32_t)(KS * NT * 32 * sizeof(double));
    if (threadIdx.x == 0) {
        mbar_expect_tx(tabbar, wbytes + pbytes);
        bulk_g2s(Wh, Ws_g + (long long)s * 2 * NT16 * NT * 32 * 4, wbytes, tabbar);
        if (G1T) bulk_g2s(Pf, Ps_g + (long long)s * 2 * NT * NT * 32 * 2, pbytes, tabbar);   // same byte count
        else bulk_g2s(Pf, Pf_g + (long long)s * KS * NT * 32, pbytes, tabbar);
    }
    auto issue = [&](int it) {
        const int t = warp + kScreenWarps * it;
        if (t < ntiles) {
            const int l0 = c_begin + t * TL;
            const int nl = min(TL, c_end - l0);
            const uint32_t bytes = (uint32_t)(nl * DP * sizeof(float));
            mbar_expect_tx(mybar, bytes);
            bulk_g2s(mytile, col_base + (long long)l0 * DP, bytes, mybar);
        }
    };
    if (lane == 0) issue(0);
    for (int i = threadIdx.x; i < DP; i += blockDim.x) mu_s[i] = mu_g[(long long)s * DP + i];
    for (int i = threadIdx.x; i < AP16; i += blockDim.x) beta_s[i] = betaf_g[(long long)s * AP16 + i];
    for (int i = threadIdx.x; i < kScreenWarps * AP16; i += blockDim.x) fsm[i] = 0.0;
    __syncthreads();
    mbar_wait(tabbar, 0);

    double* myf = fsm + warp * AP16;
    for (int it = 0; warp + kScreenWarps * it < ntiles; ++it) {
        const int t = warp + kScreenWarps * it;
        const int nl = min(TL, c_end - (c_begin + t * TL));
        mbar_wait(mybar, (uint32_t)(it & 1));
        // ---- GEMM1, then z = y*y split into two TF32 terms
        uint32_t zh[MT][2 * NT], zl[MT][2 * NT];
        if constexpr (!G1T) {
            // FP64 DMMA (same fragments as the exact pass)
            double a[MT][KS];
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                const int row = 8 * m + g;
                const bool rowok = row < nl;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const float x = mytile[row * DP + 4 * ks + q4];
                    a[m][ks] = (rowok && x == x) ? (double)x - mu_s[4 * ks + q4] : 0.0;
                }
            }
            __syncwarp();
            if (lane == 0) issue(it + 1);
            double c[MT][NT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int nt1 = 0; nt1 < NT; ++nt1) { c[m][nt1][0] = 0.0; c[m][nt1][1] = 0.0; }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int nt1 = 0; nt1 < NT; ++nt1) {
                    const double b = Pf[(ks * NT + nt1) * 32 + lane];
#pragma unroll
                    for (int m = 0; m < MT; ++m) mma884(c[m][nt1][0], c[m][nt1][1], a[m][ks], b);
                }
            }
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int nt1 = 0; nt1 < NT; ++nt1)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float zf = (float)(c[m][nt1][e] * c[m][nt1][e]);
                        const float hi = to_tf32(zf);
                        zh[m][2 * nt1 + e] = __float_as_uint(hi);
                        zl[m][2 * nt1 + e] = __float_as_uint(to_tf32(zf - hi));
                    }
        } else {
            // 3 x TF32: A = xc (16 pixels x 8 bands per k-step), B = P (8 bands x 8 j), FP32 accumulate.
            // a0 (pixel g, band q) a1 (pixel g+8, q) a2 (g, q+4) a3 (g+8, q+4); the two pixel halves of the
            // accumulator are exactly the two 8-pixel m-tiles of GEMM2.
            static_assert(MT == 2, "TF32 GEMM1 maps one m16 tile onto two 8-pixel tiles");
            const float2* Ph = reinterpret_cast<const float2*>(Pf);        // [NT][NT][32]: b0, b1
            const float2* Pl = Ph + NT * NT * 32;
            float4 xh[NT], xl[NT];
#pragma unroll
            for (int ks = 0; ks < NT; ++ks) {
                float v[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int row = g + ((e & 1) ? 8 : 0), b = 8 * ks + q4 + ((e & 2) ? 4 : 0);
                    const float x = mytile[row * DP + b];
                    const double m = mu_s[b];
                    const float mh = (float)m, ml = (float)(m - (double)mh);
                    v[e] = (row < nl && x == x) ? (x - mh) - ml : 0.f;
                }
                xh[ks] = make_float4(to_tf32(v[0]), to_tf32(v[1]), to_tf32(v[2]), to_tf32(v[3]));
                xl[ks] = make_float4(to_tf32(v[0] - xh[ks].x), to_tf32(v[1] - xh[ks].y), to_tf32(v[2] - xh[ks].z),
                                     to_tf32(v[3] - xh[ks].w));
            }
            __syncwarp();
            if (lane == 0) issue(it + 1);
#pragma unroll
            for (int nt1 = 0; nt1 < NT; ++nt1) {
                float c0[4] = {0.f, 0.f, 0.f, 0.f}, c1[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int ks = 0; ks < NT; ++ks) {
                    const float2 ph = Ph[(ks * NT + nt1) * 32 + lane], pl = Pl[(ks * NT + nt1) * 32 + lane];
                    mma_tf32(c0, xh[ks], __float_as_uint(ph.x), __float_as_uint(ph.y));
                    mma_tf32(c1, xl[ks], __float_as_uint(ph.x), __float_as_uint(ph.y));
                    mma_tf32(c2, xh[ks], __float_as_uint(pl.x), __float_as_uint(pl.y));
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float y = c0[e] + (c1[e] + c2[e]);
                    const float zf = y * y;
                    const float hi = to_tf32(zf);
                    zh[e >> 1][2 * nt1 + (e & 1)] = __float_as_uint(hi);
                    zl[e >> 1][2 * nt1 + (e & 1)] = __float_as_uint(to_tf32(zf - hi));
                }
            }
        }
        // ---- GEMM2 on the TF32 tensor path, two 16-alpha tiles at a time, + FP32 epilogue
        for (int at = 0; at < NT16; at += 2) {
            const int at1 = (at + 1 < NT16) ? at + 1 : at;
            float acc0[MT][4], acc1[MT][4];          // Wh.zh
            float lcc0[MT][4], lcc1[MT][4];          // Wl.zh
            float zcc0[MT][4], zcc1[MT][4];          // Wh.zl
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc0[m][e] = 0.f; acc1[m][e] = 0.f; lcc0[m][e] = 0.f; lcc1[m][e] = 0.f;
                    zcc0[m][e] = 0.f; zcc1[m][e] = 0.f;
                }
            const float4* wh0 = Wh + (at * NT) * 32 + lane;
            const float4* wl0 = Wl + (at * NT) * 32 + lane;
            const float4* wh1 = Wh + (at1 * NT) * 32 + lane;
            const float4* wl1 = Wl + (at1 * NT) * 32 + lane;
#pragma unroll
            for (int ks = 0; ks < NT; ++ks) {
                const float4 h0 = wh0[ks * 32], l0 = wl0[ks * 32], h1 = wh1[ks * 32], l1 = wl1[ks * 32];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    mma_tf32(acc0[m], h0, zh[m][2 * ks], zh[m][2 * ks + 1]);
                    mma_tf32(acc1[m], h1, zh[m][2 * ks], zh[m][2 * ks + 1]);
                    mma_tf32(lcc0[m], l0, zh[m][2 * ks], zh[m][2 * ks + 1]);
                    mma_tf32(lcc1[m], l1, zh[m][2 * ks], zh[m][2 * ks + 1]);
                    if (NPROD == 3) {
                        mma_tf32(zcc0[m], h0, zl[m][2 * ks], zl[m][2 * ks + 1]);
                        mma_tf32(zcc1[m], h1, zl[m][2 * ks], zl[m][2 * ks + 1]);
                    }
                }
            }
#pragma unroll
            for (int m = 0; m < MT; ++m)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    acc0[m][e] += lcc0[m][e] + zcc0[m][e];
                    acc1[m][e] += lcc1[m][e] + zcc1[m][e];
                }
            // accumulator: r^T[alpha = 16 at + g (+8)][pixel = 2 q4 + {0,1}]
            const float b00 = beta_s[16 * at + g], b01 = beta_s[16 * at + g + 8];
            const float b10 = beta_s[16 * at1 + g], b11 = beta_s[16 * at1 + g + 8];
            float f00, f01, f10, f11;
            screen_terms<MT>(acc0, b00, b01, f00, f01);
            screen_terms<MT>(acc1, b10, b11, f10, f11);
            f00 += __shfl_xor_sync(0xffffffffu, f00, 1); f00 += __shfl_xor_sync(0xffffffffu, f00, 2);
            f01 += __shfl_xor_sync(0xffffffffu, f01, 1); f01 += __shfl_xor_sync(0xffffffffu, f01, 2);
            f10 += __shfl_xor_sync(0xffffffffu, f10, 1); f10 += __shfl_xor_sync(0xffffffffu, f10, 2);
            f11 += __shfl_xor_sync(0xffffffffu, f11, 1); f11 += __shfl_xor_sync(0xffffffffu, f11, 2);
            if (q4 == 0) {
                myf[16 * at + g] += (double)f00;
                myf[16 * at + g + 8] += (double)f01;
                if (at1 != at) {
                    myf[16 * at1 + g] += (double)f10;
                    myf[16 * at1 + g + 8] += (double)f11;
                }
            }
        }
    }
    __syncthreads();
    double* out = fscreen + ((long long)s * gridDim.y + chunk) * AP16;
    for (int i = threadIdx.x; i < AP16; i += blockDim.x) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kScreenWarps; ++w) a += fsm[w * AP16 + i];
        out[i] = a;
    }
}

// ---------------------------------------------------------------------------------------- K3b
// One warp per column: approximate nll, its minimum, the candidate set and the refinement tile mask.
__global__ void __launch_bounds__(128)
    select_kernel(const double* __restrict__ fscreen, int nchunk, const double* __restrict__ logdet_g,
                  const double* __restrict__ rsum_g, const int* __restrict__ n_g,
                  const int* __restrict__ nloo_g, int A, int AP, int AP16,
                  int D, int S, double tol, double* __restrict__ nll_g, int* __restrict__ sel_index,
                  unsigned long long* __restrict__ tile_mask, int* __restrict__ ncand_g,
                  double* __restrict__ tol_g, const float* __restrict__ betaf_fold, int* __restrict__ probe_g) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (s >= S) return;
    const int n = n_g[s];
    if (n < 2) {                               // K4 writes the degenerate results
        if (lane == 0) { sel_index[s] = -3; tile_mask[s] = 0ull; ncand_g[s] = 0; tol_g[s] = 0.0; probe_g[s] = -1; }
        return;
    }
    const double inf = __longlong_as_double(0x7ff0000000000000LL);
    const double const_term = (double)D * log(2.0 * M_PI);
    const double nl = (double)(nloo_g ? nloo_g[s] : n);       // the n of looshrinkage (:355-356)
    double vmin = inf, hmax = 0.0;
    bool bad = false;                          // a non-finite sum: the exact pass decides everything
    for (int i = lane; i < A; i += 32) {
        double fs = 0.0;
        for (int c = 0; c < nchunk; ++c) fs += fscreen[((long long)s * nchunk + c) * AP16 + i];
        // the tcgen05 pass sums h + u; sum_k u_k = beta sum_k r_k is the closed form (k_screen5.cu)
        if (betaf_fold) fs -= (double)betaf_fold[(long long)s * AP16 + i] * rsum_g[(long long)s * AP + i];
        const double ld = det_roundtrip(logdet_g[(long long)s * AP + i]);
        double v;
        if (!(fabs(ld) < inf)) v = inf;                    // det under/overflow (:112-113)
        else {
            v = 0.5 * (const_term + ld) + (rsum_g[(long long)s * AP + i] + fs) / (2.0 * nl);
            if (!(fabs(v) < inf)) bad = true;  // NaN or +-inf out of the data
            hmax = fmax(hmax, fabs(fs) / (2.0 * nl));
        }
        nll_g[(long long)s * A + i] = v;
        if (v < vmin) vmin = v;
    }
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        hmax = fmax(hmax, __shfl_xor_sync(0xffffffffu, hmax, o));
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    unsigned long long mask = 0ull;
    int cnt = 0, first = 0x7fffffff;
    if (!bad && vmin < inf) {
        // the screening error is relative to the part of nll that is screened (the h sums): tol = rel * max|h term|
        const double lim = vmin + (tol * hmax + 1.0e-10);
        for (int i = lane; i < A; i += 32) {
            const double v = nll_g[(long long)s * A + i];
            if (v <= lim) { ++cnt; if (i < first) first = i; mask |= 1ull << (i >> 3); }
        }
        for (int o = 16; o > 0; o >>= 1) {
            cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            first = min(first, __shfl_xor_sync(0xffffffffu, first, o));
            mask |= __shfl_xor_sync(0xffffffffu, mask, o);
        }
    }
    // Runtime certificate, part 1: a column that is refined anyway also gets its best EXCLUDED tile evaluated
    // exactly (the probe).  K4 then checks that the exact minimum does not sit in the probe tile and measures the
    // screening error on every evaluated alpha against the margin.
    // Every 32nd column is a sentinel: refined (best tile + probe) even when the screen separates its alphas, so that
    // every flightline carries measurements of the screening error (part 3 extends them to the other columns).
    const bool sentinel = (s & 31) == 0;
    if (!bad && vmin < inf && cnt == 1 && sentinel) { cnt = 2; mask = 1ull << (first >> 3); }
    int probe = -1;
    if (!bad && vmin < inf && cnt > 1) {
        double pv = inf;
        int pi = 0x7fffffff;
        for (int i = lane; i < A; i += 32) {
            if ((mask >> (i >> 3)) & 1ull) continue;
            const double v = nll_g[(long long)s * A + i];
            if (v < pv) { pv = v; pi = i; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, pv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, pi, o);
            if (ov < pv || (ov == pv && oi < pi)) { pv = ov; pi = oi; }
        }
        if (pv < inf) probe = pi >> 3;
    }
    if (lane == 0) {
        tol_g[s] = tol * hmax + 1.0e-10;
        probe_g[s] = -1;
        if (bad) { sel_index[s] = -2; tile_mask[s] = ~0ull; ncand_g[s] = A; }          // refine everything
        else if (!(vmin < inf)) { sel_index[s] = -1; tile_mask[s] = 0ull; ncand_g[s] = 0; }   // all inf (:123-127)
        else if (cnt == 1) { sel_index[s] = first; tile_mask[s] = 0ull; ncand_g[s] = 1; }
        else {
            sel_index[s] = -2; ncand_g[s] = cnt;
            if (probe >= 0) { mask |= 1ull << probe; probe_g[s] = probe; }
            tile_mask[s] = mask;
        }
    }
}

// ---------------------------------------------------------------------------------------- launchers
size_t screen_smem_bytes(const Dims& d) {
    const int DP = d.DP, KS = DP / 4, TL = 8 * kScreenMT;
    return (size_t)2 * d.NT16 * d.NT * 32 * sizeof(float4) + (size_t)(KS * d.NT * 32 + DP) * sizeof(double) +
           (size_t)kScreenWarps * d.AP16 * sizeof(double) + (size_t)d.AP16 * sizeof(float) +
           (size_t)kScreenWarps * TL * DP * sizeof(float) + (kScreenWarps + 1) * sizeof(uint64_t);
}

// tuning hook (tools/ only): CMF_SCREEN_VARIANT = "<g1t>,<nprod>"
static void screen_variant(int* g1t, int* nprod) {
    static int v[2] = {-1, -1};
    if (v[0] < 0) {
        v[0] = 1; v[1] = 3;   // measured: TF32 projection keeps the screening error at 0.07 of the margin
        if (const char* e = cmf_hook("CMF_SCREEN_VARIANT")) sscanf(e, "%d,%d", &v[0], &v[1]);
    }
    *g1t = v[0]; *nprod = v[1];
}

template <int NT, bool G1T, int NPROD>
static void launch_screen_v(const Dims& d, const float* xt, const double* mu, const double* Pf, const float* Ps,
                            const float* Ws, const float* betaf, const int* n, int nchunk, double* fscreen,
                            cudaStream_t st) {
    constexpr int TL = 8 * kScreenMT;
    const size_t smem = screen_smem_bytes(d);
    cudaFuncSetAttribute(loo_screen_kernel<NT, G1T, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int lpc = (d.L + nchunk - 1) / nchunk;
    lpc = (lpc + TL - 1) / TL * TL;
    dim3 grid(d.S, nchunk);
    loo_screen_kernel<NT, G1T, NPROD><<<grid, kScreenWarps * 32, smem, st>>>(xt, mu, Pf, Ps, Ws, betaf, n, d.L,
                                                                             d.NT16, lpc, fscreen, d.nrows);
}

template <int NT>
static void launch_screen_t(const Dims& d, const float* xt, const double* mu, const double* Pf, const float* Ps,
                            const float* Ws, const float* betaf, const int* n, int nchunk, double* fscreen,
                            cudaStream_t st) {
    int g1t, nprod;
    screen_variant(&g1t, &nprod);
#ifdef CMF_TUNING_HOOKS
    if (g1t && nprod == 3) launch_screen_v<NT, true, 3>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st);
    else if (g1t) launch_screen_v<NT, true, 2>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st);
    else if (nprod == 3) launch_screen_v<NT, false, 3>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st);
    else launch_screen_v<NT, false, 2>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st);
#else
    (void)g1t; (void)nprod;
    launch_screen_v<NT, true, 3>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st);   // the measured-best variant
#endif
}

void launch_screen(const Dims& d, const float* xt, const double* mu, const double* Pf, const float* Ps,
                   const float* Ws, const float* betaf, const int* n, int nchunk, double* fscreen,
                   cudaStream_t st) {
    switch (d.NT) {
#define CMF_CASE(k) case k: launch_screen_t<k>(d, xt, mu, Pf, Ps, Ws, betaf, n, nchunk, fscreen, st); break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6)
        CMF_CASE(7) CMF_CASE(8) CMF_CASE(9) CMF_CASE(10) CMF_CASE(11) CMF_CASE(12)
#undef CMF_CASE
        default: break;
    }
}

void launch_select(const Dims& d, const double* fscreen, int nchunk, const double* logdet, const double* rsum,
                   const int* n, const int* nloo, double tol, double* nll, int* sel_index,
                   unsigned long long* tile_mask, int* ncand, double* tol_out, const float* betaf_fold,
                   int* probe, cudaStream_t st) {
    select_kernel<<<(d.S + 3) / 4, 128, 0, st>>>(fscreen, nchunk, logdet, rsum, n, nloo, d.A, d.AP, d.AP16, d.D, d.S, tol,
                                                 nll, sel_index, tile_mask, ncand, tol_out, betaf_fold, probe);
}

// Runtime certificate, part 3 (one CTA): the largest measured screening error of the flightline, as a fraction of the
// margin, also vouches for the columns the screen decided alone.  If it reaches kCertFactor^-1 of the margin every
// such column is sent to the exact pass as well; columns whose own check failed in K4 already are.
__global__ void __launch_bounds__(256)
    certify_kernel(const double* __restrict__ check, const int* __restrict__ sel_index, int S, int enabled,
                   unsigned long long* __restrict__ redo, double* __restrict__ worst_out) {
    __shared__ double red[8];
    double w = 0.0;
    for (int s = threadIdx.x; s < S; s += blockDim.x) if (check[s] > w) w = check[s];
    for (int o = 16; o > 0; o >>= 1) w = fmax(w, __shfl_xor_sync(0xffffffffu, w, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = w;
    __syncthreads();
    w = red[0];
    for (int i = 1; i < 8; ++i) w = fmax(w, red[i]);
    if (threadIdx.x == 0 && worst_out) *worst_out = w;
    const bool all = enabled && (w * kCertFactor > 1.0);
    for (int s = threadIdx.x; s < S; s += blockDim.x) {
        unsigned long long r = redo[s];                 // set by K4 for a column whose own check failed
        if (all && sel_index[s] >= 0) r = ~0ull;
        redo[s] = enabled ? r : 0ull;
    }
}

void launch_certify(const Dims& d, const double* check, const int* sel_index, int enabled, unsigned long long* redo,
                    double* worst, cudaStream_t st) {
    certify_kernel<<<1, 256, 0, st>>>(check, sel_index, d.S, enabled, redo, worst);
}

}  // namespace cmf
