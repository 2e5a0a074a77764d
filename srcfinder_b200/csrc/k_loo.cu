// K3: leave-one-out likelihood sums for all alphas in one pass over the column (FP64 tensor path).
//
// Replaces the 201-iteration loop of looshrinkage (cmf/robust_mf.py:105-117; Theiler 2012 eq. 29):
// per pixel k and alpha i the reference needs r_ki = x_k^T G_i^-1 x_k, q = 1 - beta_i r and
// log(q) + r/q.  With the spectral form prepared by K2 this becomes two chained FP64 GEMMs per 8-pixel
// tile, both on DMMA.8x8x4:
//     GEMM1   Y = Xc . P            (8 x DP) . (DP x DP)      y_kj
//     GEMM2   R^T = W^T . (Y*Y)^T   (A x DP) . (DP x 8)       r_ki = sum_j y_kj^2 W_ji
// The accumulator registers of GEMM1 (squared in place) are exactly the B fragments of GEMM2 because
// K2 stores W with its j axis permuted to the accumulator layout, so nothing moves between the GEMMs.
// The epilogue forms log(q) + r/q per (pixel, alpha) in FP64 and reduces over pixels (in-lane, then two
// shuffles); partial sums per warp live in shared memory and leave the CTA in a fixed order.
//
// Mapping: CTA = (column, chunk of lines).  P (fragment order), W (fragment order), beta stay resident
// in shared memory (162 KB for D = 72, A = 201); each warp streams its own 16-line tiles with 1-D bulk
// async copies into a single private stage that is refilled as soon as the tile sits in registers.
#include <math.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

constexpr int kLooWarps = 8;

// log(q) + r/q with q = 1 - u, u = beta r  (cmf/robust_mf.py:115-117), evaluated as
//     r + sum_{m>=1} u^m (r - 1/m)
// The series is exact to below one ulp of r once u^(M+1) < 2^-56:  M = 8 for |u| <= 2^-7 (the common case,
// u ~ D/n), M = 13 for |u| <= 2^-4; anything larger takes the library log and division.  The measured cost
// of log + division is ~65 FP64 issue slots against 18 for the short series, and on this part FP64 FMA and
// DMMA share one pipe (33 vs 37 TFLOP/s), so the epilogue is a first-order term of the kernel's run time.
__device__ __forceinline__ double loo_term(double r, double beta) {
    const double u = beta * r;
    const double au = fabs(u);
    if (au <= 0x1p-7) {
        double s1 = fma(u, 1.0, 1.0);                   // 1 + u
        double s2 = fma(u, 1.0 / 8.0, 1.0 / 7.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 6.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 5.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 4.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 3.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 2.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0);
        // s1 = 1 + u + ... + u^7, s2 = 1 + u/2 + ... + u^7/8 ;  result r + u (r s1 - s2)
        return fma(u, fma(r, s1, -s2), r);
    }
    if (au <= 0x1p-4) {
        double s1 = 1.0, s2 = 1.0 / 13.0;
#pragma unroll
        for (int m = 12; m >= 1; --m) {
            s1 = fma(u, s1, 1.0);
            s2 = fma(u, s2, 1.0 / (double)m);
        }
        return fma(u, fma(r, s1, -s2), r);
    }
    const double q = 1.0 - u;
    return log(q) + r / q;
}

template <int NT, bool WSMEM>
__global__ void __launch_bounds__(kLooWarps * 32, 1)
    loo_kernel(const float* __restrict__ xt, const double* __restrict__ mu_g,
               const double* __restrict__ Pf_g, const double* __restrict__ Wf_g,
               const double* __restrict__ beta_g, int L, int NT2, int lines_per_chunk,
               double* __restrict__ fpart, const unsigned long long* __restrict__ tile_mask,
               const int* __restrict__ nrows) {
    constexpr int DP = 8 * NT, KS = DP / 4, MT = kLooMT, TL = 8 * MT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int AP = NT2 * 8;
    // W table: resident in shared memory when it fits (WSMEM), else read through L1/L2 (large D)
    double* Wf_s = reinterpret_cast<double*>(smem_raw);        // [KS][NT2][32]
    double* Pf = Wf_s + (WSMEM ? KS * NT2 * 32 : 0);           // [KS][NT][32]
    double* mu_s = Pf + KS * NT * 32;                          // [DP]
    double* beta_s = mu_s + DP;                                // [AP]
    double* fsm = beta_s + AP;                                 // [warps][AP]
    float* ring = reinterpret_cast<float*>(fsm + kLooWarps * AP);   // [warps][TL*DP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + kLooWarps * TL * DP);   // [warps] + 1

    const int s = blockIdx.x, chunk = blockIdx.y;
    // refinement mode: bit t of the mask = "8-alpha tile t holds a candidate" (K3b); 0 = column decided
    const unsigned long long tmask = tile_mask ? tile_mask[s] : ~0ull;
    if (tmask == 0ull) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, q4 = lane & 3;
    const int c_begin = chunk * lines_per_chunk;
    const int c_end = max(c_begin, min(nrows ? min(L, nrows[s]) : L, c_begin + lines_per_chunk));   // compacted mode pass
    const int ntiles = (c_end - c_begin + TL - 1) / TL;
    const float* col_base = xt + (long long)s * L * DP;
    float* mytile = ring + warp * TL * DP;
    uint64_t* mybar = bars + warp;
    uint64_t* tabbar = bars + kLooWarps;

    if (lane == 0) mbar_init(mybar, 1);
    if (threadIdx.x == 0) mbar_init(tabbar, 1);
    if (lane == 0) fence_mbar_init();
    __syncthreads();

    // resident tables: three bulk copies on one barrier (sizes are multiples of 256 B)
    const uint32_t wbytes = (uint32_t)(KS * NT2 * 32 * sizeof(double));
    const uint32_t pbytes = (uint32_t)(KS * NT * 32 * sizeof(double));
    if (threadIdx.x == 0) {
        mbar_expect_tx(tabbar, (WSMEM ? wbytes : 0u) + pbytes);
        if (WSMEM) bulk_g2s(Wf_s, Wf_g + (long long)s * KS * NT2 * 32, wbytes, tabbar);
        bulk_g2s(Pf, Pf_g + (long long)s * KS * NT * 32, pbytes, tabbar);
    }
    auto issue = [&](int it) {
        const int t = warp + kLooWarps * it;
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
    for (int i = threadIdx.x; i < AP; i += blockDim.x) beta_s[i] = beta_g[(long long)s * AP + i];
    for (int i = threadIdx.x; i < kLooWarps * AP; i += blockDim.x) fsm[i] = 0.0;
    __syncthreads();
    mbar_wait(tabbar, 0);

    const double* Wf = WSMEM ? Wf_s : (Wf_g + (long long)s * KS * NT2 * 32);
    double* myf = fsm + warp * AP;
    for (int it = 0; warp + kLooWarps * it < ntiles; ++it) {
        const int t = warp + kLooWarps * it;
        const int nl = min(TL, c_end - (c_begin + t * TL));
        mbar_wait(mybar, (uint32_t)(it & 1));
        // ---- A fragments of GEMM1: a[m][ks] = xc[pixel = 8 m + lane/4][b = 4 ks + lane%4]
        double z[MT][KS];
        {
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
            if (lane == 0) issue(it + 1);   // the tile now lives in registers: refill the stage
            // ---- GEMM1 (+ squaring), k-outer: NT*MT independent accumulator chains
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
                for (int nt1 = 0; nt1 < NT; ++nt1) {
                    z[m][2 * nt1] = c[m][nt1][0] * c[m][nt1][0];
                    z[m][2 * nt1 + 1] = c[m][nt1][1] * c[m][nt1][1];
                }
        }
        // ---- GEMM2 + epilogue, two alpha tiles at a time for ILP on the tensor pipe
        for (int at = 0; at < NT2; at += 2) {
            if (tile_mask && !((tmask >> (at & 63)) & 3ull)) continue;
            const bool two = (at + 1 < NT2);
            double c0[MT][2], c1[MT][2];
#pragma unroll
            for (int m = 0; m < MT; ++m) { c0[m][0] = c0[m][1] = 0.0; c1[m][0] = c1[m][1] = 0.0; }
            const double* w0 = Wf + at * 32 + lane;
            const double* w1 = Wf + (two ? at + 1 : at) * 32 + lane;
#pragma unroll
            for (int ks2 = 0; ks2 < KS; ++ks2) {
                const double wa = w0[ks2 * NT2 * 32];
                const double wb = w1[ks2 * NT2 * 32];
#pragma unroll
                for (int m = 0; m < MT; ++m) {
                    mma884(c0[m][0], c0[m][1], wa, z[m][ks2]);
                    mma884(c1[m][0], c1[m][1], wb, z[m][ks2]);
                }
            }
            // accumulator: r^T[alpha = 8 at + lane/4][pixel = 2 (lane%4) + e]
            const double be0 = beta_s[8 * at + g];
            const double be1 = beta_s[8 * (two ? at + 1 : at) + g];
            double f0 = 0.0, f1 = 0.0;
#pragma unroll
            for (int m = 0; m < MT; ++m) {
                f0 += loo_term(c0[m][0], be0) + loo_term(c0[m][1], be0);
                f1 += loo_term(c1[m][0], be1) + loo_term(c1[m][1], be1);
            }
            f0 += shfl_xor_f64(f0, 1); f0 += shfl_xor_f64(f0, 2);
            f1 += shfl_xor_f64(f1, 1); f1 += shfl_xor_f64(f1, 2);
            if (q4 == 0) {
                myf[8 * at + g] += f0;
                if (two) myf[8 * (at + 1) + g] += f1;
            }
        }
    }
    __syncthreads();
    double* out = fpart + ((long long)s * gridDim.y + chunk) * AP;
    for (int i = threadIdx.x; i < AP; i += blockDim.x) {
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < kLooWarps; ++w) a += fsm[w * AP + i];
        out[i] = a;
    }
}

template <int NT>
static void launch_loo_t(const Dims& d, const float* xt, const double* mu, const double* Pf, const double* Wf,
                         const double* beta, int nchunk, double* fpart, const unsigned long long* tile_mask,
                         cudaStream_t st) {
    constexpr int DP = 8 * NT, KS = DP / 4, TL = 8 * kLooMT;
    const size_t base = (size_t)(KS * NT * 32 + DP + d.AP + kLooWarps * d.AP) * sizeof(double) +
                        (size_t)kLooWarps * TL * DP * sizeof(float) + (kLooWarps + 1) * sizeof(uint64_t);
    const size_t wtab = (size_t)KS * d.NT2 * 32 * sizeof(double);
    int lpc = (d.L + nchunk - 1) / nchunk;
    lpc = (lpc + TL - 1) / TL * TL;
    dim3 grid(d.S, nchunk);
    if (base + wtab <= 227 * 1024) {
        cudaFuncSetAttribute(loo_kernel<NT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(base + wtab));
        loo_kernel<NT, true><<<grid, kLooWarps * 32, base + wtab, st>>>(xt, mu, Pf, Wf, beta, d.L, d.NT2, lpc, fpart, tile_mask, d.nrows);
    } else {
        cudaFuncSetAttribute(loo_kernel<NT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)base);
        loo_kernel<NT, false><<<grid, kLooWarps * 32, base, st>>>(xt, mu, Pf, Wf, beta, d.L, d.NT2, lpc, fpart, tile_mask, d.nrows);
    }
}

void launch_loo(const Dims& d, const float* xt, const double* mu, const double* Pf, const double* Wf,
                const double* beta, int nchunk, double* fpart, const unsigned long long* tile_mask,
                cudaStream_t st) {
    switch (d.NT) {
#define CMF_CASE(k) case k: launch_loo_t<k>(d, xt, mu, Pf, Wf, beta, nchunk, fpart, tile_mask, st); break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6)
        CMF_CASE(7) CMF_CASE(8) CMF_CASE(9) CMF_CASE(10) CMF_CASE(11) CMF_CASE(12)
#undef CMF_CASE
        default: break;
    }
}

}  // namespace cmf
