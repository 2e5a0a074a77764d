// Wide-window kernel set: active windows wider than 96 bands, where the per-column matrices no longer fit the
// shared-memory kernels of k_gram / k_eigen / k_loo.  The reference selects such a window for `-R` with the CH4
// library (cmf/robust_mf.py:186-187: bands 5..420, D = 416); the arithmetic per column is the same
// (:92-136, :297-397), only the shapes change: 416 x 416 covariance, 416 x 416 eigenproblem, n x 416 x 416
// projection.  Everything here is generic in D (any width, also used for small D by the one-column
// looshrinkage entry point) and blocked:
//   W0  wide_valid_kernel / wide_sums_kernel / wide_mean_kernel   mask (:282), column sums, range, mean (:347)
//   W1  wide_pack_kernel      BIL slab -> column-major copy xt [S][L][DP] (+ the int8 digit images of k_gram8.cu)
//   W2  wide_gram64_kernel    centred Gram in FP64 (DMMA), 64 x 64 blocks (cross-check / FP64-input path)
//   W3  wide_cov_kernel, wide_tred_kernel, wide_ql_kernel, wide_rot_kernel   correlation matrix, Householder
//       tridiagonalisation and Q in global memory, QL recurrences of all columns in parallel (one warp per
//       column; the rotation sequences go to global memory), rotations applied to row blocks of Q in shared memory
//   W4  wide_tables_kernel    log det G_alpha, beta, sum_k r_k, W[j][alpha]
//   W5  wide_proj_kernel      Z = (Xc P)^2 in FP64 (DMMA)            (x^T G^-1 x in the spectral form, :114)
//   W6  wide_loo_kernel       R = Z W, log q + r/q, sums per alpha    (:115-117)
// K4 (finalize_kernel), the scalar scoring kernel and K6 are shared with the narrow path.
#include <math.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

// ---------------------------------------------------------------------------------------- W0
// Validity of every pixel (cmf/robust_mf.py:282): thread <-> (line, column), all D active bands; the 32 lanes of
// a warp read 32 neighbouring columns of one (line, band) row.
__global__ void __launch_bounds__(256)
    wide_valid_kernel(const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S, int D,
                      uint8_t* __restrict__ mask) {
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int line = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (col >= S || line >= L) return;
    const float* p = slab + (long long)line * line_pitch + col;
    bool ok = true;
    int b = 0;
    for (; b + 8 <= D; b += 8) {
        float v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldg_nc_f1(p + (long long)(b + k) * band_pitch);
#pragma unroll
        for (int k = 0; k < 8; ++k) ok = ok && pixel_value_ok(v[k]);
    }
    for (; b < D; ++b) ok = ok && pixel_value_ok(ldg_nc_f1(p + (long long)b * band_pitch));
    mask[(long long)line * S + col] = ok ? 1 : 0;
}

// Column sums, counts and value ranges over the pixels that enter the statistics (valid, and selected when a
// selection is given): thread <-> (column, band), a fixed range of lines in order -> deterministic partials.
__global__ void __launch_bounds__(256)
    wide_sums_kernel(const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S, int D, int DP,
                     const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sel, int lines_per_split,
                     double* __restrict__ colsum_part, int* __restrict__ colcnt_part, float* __restrict__ lo_part,
                     float* __restrict__ hi_part) {
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int b = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int split = blockIdx.z;
    if (col >= S || b >= DP) return;
    const int l0 = split * lines_per_split, l1 = min(L, l0 + lines_per_split);
    double sum = 0.0;
    float lo = 3.402823466e38f, hi = -3.402823466e38f;
    int cnt = 0;
    if (b < D) {
        const float* p = slab + (long long)b * band_pitch + col;
        for (int l = l0; l < l1; ++l) {
            const long long o = (long long)l * S + col;
            if (mask[o] && (sel == nullptr || sel[o])) {
                const float v = ldg_nc_f1(p + (long long)l * line_pitch);
                sum += (double)v;
                lo = fminf(lo, v);
                hi = fmaxf(hi, v);
                ++cnt;
            }
        }
    }
    const long long o = ((long long)split * S + col) * DP + b;
    colsum_part[o] = sum;
    lo_part[o] = lo;
    hi_part[o] = hi;
    if (b == 0) colcnt_part[split * S + col] = cnt;
}

// mean (:347), count, the centre the Gram products are taken about (the mean rounded to float, so that
// (double)x - ctr is exact) and the binary exponent of the largest |x - ctr| of every (column, band)
__global__ void __launch_bounds__(128)
    wide_mean_kernel(const double* __restrict__ colsum_part, const int* __restrict__ colcnt_part,
                     const float* __restrict__ lo_part, const float* __restrict__ hi_part, int nsplit, int S, int D,
                     int DP, double* __restrict__ mu, int* __restrict__ n, double* __restrict__ ctr,
                     int* __restrict__ qexp) {
    const int s = blockIdx.x;
    int cnt = 0;
    for (int k = 0; k < nsplit; ++k) cnt += colcnt_part[k * S + s];
    for (int b = threadIdx.x; b < DP; b += blockDim.x) {
        double a = 0.0;
        float lo = 3.402823466e38f, hi = -3.402823466e38f;
        for (int k = 0; k < nsplit; ++k) {
            const long long o = ((long long)k * S + s) * DP + b;
            a += colsum_part[o];
            lo = fminf(lo, lo_part[o]);
            hi = fmaxf(hi, hi_part[o]);
        }
        const double m = cnt > 0 ? a / (double)cnt : 0.0;       // numpy mean: sum / n
        const double c = (b < D && cnt > 0) ? (double)(float)m : 0.0;
        mu[(long long)s * DP + b] = (b < D) ? m : 0.0;
        ctr[(long long)s * DP + b] = c;
        int e = 0;
        if (b < D && cnt > 0) {
            const double r = fmax(fabs((double)hi - c), fabs((double)lo - c));
            if (r > 0.0) (void)frexp(r, &e);                    // r = f * 2^e, f in [0.5, 1)  =>  r < 2^e
        }
        qexp[(long long)s * DP + b] = e;
    }
    if (threadIdx.x == 0) n[s] = cnt;
}

// ---------------------------------------------------------------------------------------- W1
// BIL slab -> (a) column-major copy xt [S][L][DP] float (NaN rows for pixels outside the statistics) and
// (b) optionally the balanced base-256 digit images of the centred radiances that the tcgen05 integer Gram pass
// reads (k_gram8.cu):  q = rint((x - ctr) 2^(30 - e)) = d0 2^24 + d1 2^16 + d2 2^8 + d3, d in [-128, 127],
// stored as ready-made shared-memory operand tiles  img[s][64-line block][32-band block][16-line chunk 0..3]
// [row = digit * 32 + band][16 lines]  (the canonical K-major no-swizzle core-matrix layout, K = line).
// CTA = 32 columns x 16 lines (one chunk); per 32-band block the tile goes through shared memory so that both
// outputs are written in full 128 / 512-byte runs.
__global__ void __launch_bounds__(256)
    wide_pack_kernel(const float* __restrict__ slab, long long line_pitch, int band_pitch, int L, int S, int D, int DP,
                     const uint8_t* __restrict__ mask, const uint8_t* __restrict__ sel, const double* __restrict__ ctr,
                     const int* __restrict__ qexp, float* __restrict__ xt, int8_t* __restrict__ img, int nrb,
                     int nkb) {
    extern __shared__ __align__(16) unsigned char pack_smem[];
    float (*tile)[32][33] = reinterpret_cast<float (*)[32][33]>(pack_smem);                    // [16][32][33]
    uint8_t (*use)[32] = reinterpret_cast<uint8_t (*)[32]>(pack_smem + 16 * 32 * 33 * sizeof(float));   // [16][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s0 = blockIdx.x * 32, l0 = blockIdx.y * 16;
    const int kb = blockIdx.y >> 2, c4 = blockIdx.y & 3;
    for (int i = tid; i < 512; i += 256) {
        const int l = i >> 5, c = i & 31;
        bool u = false;
        if (l0 + l < L && s0 + c < S) {
            const long long o = (long long)(l0 + l) * S + s0 + c;
            u = mask[o] && (sel == nullptr || sel[o]);
        }
        use[l][c] = u ? 1 : 0;
    }
    for (int rb = 0; rb < nrb; ++rb) {
        __syncthreads();
        // ---- load [16 lines][32 bands][32 columns]: a warp reads 32 neighbouring columns of one (line, band) row
        for (int i = warp; i < 512; i += 8) {
            const int l = i >> 5, bi = i & 31, b = rb * 32 + bi;
            float v = 0.f;
            if (l0 + l < L && b < D && s0 + lane < S)
                v = ldg_nc_f1(slab + (long long)(l0 + l) * line_pitch + (long long)b * band_pitch + s0 + lane);
            tile[l][bi][lane] = v;
        }
        __syncthreads();
        // ---- (a) xt rows: a warp writes the 32 bands of one (column, line)
        if (xt != nullptr && rb * 32 < DP) {
            const float qnan = __int_as_float(0x7fc00000);
            for (int i = warp; i < 512; i += 8) {
                const int c = i >> 4, l = i & 15;
                if (l0 + l < L && s0 + c < S && rb * 32 + lane < DP)
                    xt[((long long)(s0 + c) * L + l0 + l) * DP + rb * 32 + lane] = use[l][c] ? tile[l][lane][c] : qnan;
            }
        }
        // ---- (b) digit images: thread <-> (column, band), 16 lines -> four 16-byte units
        if (img != nullptr) {
#pragma unroll 1
            for (int k = 0; k < 4; ++k) {
                const int c = warp * 4 + k, b = rb * 32 + lane;
                if (s0 + c >= S) continue;
                double cc = 0.0, sc = 0.0;
                if (b < D) {
                    cc = ctr[(long long)(s0 + c) * DP + b];
                    sc = scalbn(1.0, 30 - qexp[(long long)(s0 + c) * DP + b]);
                }
                uint32_t w[4][4];
#pragma unroll
                for (int dgt = 0; dgt < 4; ++dgt)
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[dgt][j] = 0u;
#pragma unroll
                for (int l = 0; l < 16; ++l) {
                    int q = 0;
                    if (b < D && use[l][c]) q = __double2int_rn(((double)tile[l][lane][c] - cc) * sc);
                    // balanced digits, least significant first: d = ((q + 128) & 255) - 128, q = (q - d) >> 8
                    const int d3 = ((q + 128) & 255) - 128; q = (q - d3) >> 8;
                    const int d2 = ((q + 128) & 255) - 128; q = (q - d2) >> 8;
                    const int d1 = ((q + 128) & 255) - 128; q = (q - d1) >> 8;
                    const int d0 = q;
                    const int sh = 8 * (l & 3);
                    w[0][l >> 2] |= (uint32_t)(d0 & 255) << sh;
                    w[1][l >> 2] |= (uint32_t)(d1 & 255) << sh;
                    w[2][l >> 2] |= (uint32_t)(d2 & 255) << sh;
                    w[3][l >> 2] |= (uint32_t)(d3 & 255) << sh;
                }
                int8_t* base = img + ((((long long)(s0 + c) * nkb + kb) * nrb + rb) * 4 + c4) * (128 * 16);
#pragma unroll
                for (int dgt = 0; dgt < 4; ++dgt)
                    *reinterpret_cast<uint4*>(base + (dgt * 32 + lane) * 16) =
                        make_uint4(w[dgt][0], w[dgt][1], w[dgt][2], w[dgt][3]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------- FP64 tensor blocks
// C(64 x 64) += sum_k A[k][m] B[k][n] over a 16-deep stage held in shared memory (k-major rows of kWP doubles;
// kWP == 4 mod 16 keeps the fragment reads of a half-warp on 16 different 8-byte banks).  4 warps, 2 x 2, each a
// 32 x 32 block of 4 x 4 DMMA.8x8x4 tiles.
constexpr int kWP = 68;

__device__ __forceinline__ void dmma_stage(const double* __restrict__ As, const double* __restrict__ Bs,
                                           double (&acc)[4][4][2], int wm, int wn, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[t] = As[(4 * kk + q) * kWP + wm * 32 + t * 8 + g];
            b[t] = Bs[(4 * kk + q) * kWP + wn * 32 + t * 8 + g];
        }
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj) mma884(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
    }
}

template <typename T>
__device__ __forceinline__ double centred(T x, double c) {
    return (x == x) ? (double)x - c : 0.0;       // NaN rows mark pixels outside the statistics
}

// ---------------------------------------------------------------------------------------- W2
// G = sum_l (x_l - c)(x_l - c)^T for one 64 x 64 block pair (bi >= bj) of one column, all lines in order.
template <typename T>
__global__ void __launch_bounds__(128)
    wide_gram64_kernel(const T* __restrict__ xt, const double* __restrict__ ctr, int L, int DP,
                       double* __restrict__ gram) {
    __shared__ double As[16 * kWP], Bs[16 * kWP];
    const int s = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int bi = 0, bj = blockIdx.x;
    while (bj > bi) { bj -= bi + 1; ++bi; }          // pair index -> (bi, bj), bj <= bi
    const int wm = warp >> 1, wn = warp & 1;
    const T* col = xt + (long long)s * L * DP;
    const double* cs = ctr + (long long)s * DP;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    const bool diag = bi == bj;
    // element e of a stage: line = e / 64, band = e % 64; thread t owns e = t + 128 r
    const int eb = tid & 63, el = tid >> 6;
    const int ba = bi * 64 + eb, bb = bj * 64 + eb;
    const double ca = ba < DP ? cs[ba] : 0.0, cb = bb < DP ? cs[bb] : 0.0;
    double ra[8], rb[8];
    auto fetch = [&](int l0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int l = l0 + el + 2 * r;
            ra[r] = (l < L && ba < DP) ? centred(col[(long long)l * DP + ba], ca) : 0.0;
            rb[r] = (!diag && l < L && bb < DP) ? centred(col[(long long)l * DP + bb], cb) : 0.0;
        }
    };
    fetch(0);
    for (int l0 = 0; l0 < L; l0 += 16) {
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            As[(el + 2 * r) * kWP + eb] = ra[r];
            if (!diag) Bs[(el + 2 * r) * kWP + eb] = rb[r];
        }
        __syncthreads();
        if (l0 + 16 < L) fetch(l0 + 16);
        dmma_stage(As, diag ? As : Bs, acc, wm, wn, lane);
    }
    const int g = lane >> 2, q = lane & 3;
    double* G = gram + (long long)s * DP * DP;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int r = bi * 64 + wm * 32 + ti * 8 + g, c = bj * 64 + wn * 32 + tj * 8 + 2 * q + e;
                if (r < DP && c < DP) {
                    if (!diag || c <= r) G[(long long)r * DP + c] = acc[ti][tj][e];
                    if (!diag || c < r) G[(long long)c * DP + r] = acc[ti][tj][e];
                }
            }
}

// ---------------------------------------------------------------------------------------- W3: covariance
// Gram about ctr -> covariance (pilot term removed as in K2), shrinkage target diag(S) (:100), correlation matrix
// R = T^-1/2 S T^-1/2 (full, symmetric) into the work matrix; dinv, sum log(1e4 T_jj) (:94-99).
// scale_rc: optional per-band binary exponents of the integer Gram (G_real = G * 2^(e_r + e_c - 60)).
__global__ void __launch_bounds__(256)
    wide_cov_kernel(const double* __restrict__ gram, const int* __restrict__ n_g, int D, int DP,
                    const double* __restrict__ mu_g, const double* __restrict__ ctr_g, const int* __restrict__ qexp,
                    int mode, double* __restrict__ work, double* __restrict__ dinv_g, double* __restrict__ slogT_g,
                    int* __restrict__ status_g) {
    extern __shared__ double sm[];
    double* dinv = sm;          // [DP]
    double* dm = dinv + DP;     // [DP] mu - ctr
    double* sc = dm + DP;       // [DP] 2^(e - 30) for the integer Gram, else 1
    const int s = blockIdx.x, tid = threadIdx.x;
    const int n = n_g[s];
    double* A = work + (long long)s * DP * DP;
    if (n < 2) {
        for (int i = tid; i < DP; i += blockDim.x) dinv_g[(long long)s * DP + i] = 0.0;
        if (tid == 0) { status_g[s] = (n == 0) ? kStatusEmpty : kStatusDegenerate; slogT_g[s] = 0.0; }
        return;
    }
    const double* G = gram + (long long)s * DP * DP;
    const double inv_nm1 = 1.0 / (double)(n - 1);
    for (int b = tid; b < DP; b += blockDim.x) {
        dm[b] = (b < D) ? mu_g[(long long)s * DP + b] - ctr_g[(long long)s * DP + b] : 0.0;
        sc[b] = qexp ? scalbn(1.0, qexp[(long long)s * DP + b] - 30) : 1.0;
    }
    __syncthreads();
    for (int b = tid; b < DP; b += blockDim.x) {
        double t0 = 0.0;
        if (b < D) t0 = (G[(long long)b * DP + b] * sc[b] * sc[b] - (double)n * dm[b] * dm[b]) * inv_nm1;
        dinv[b] = (mode == 1) ? ((b < D) ? 1.0 : 0.0) : ((t0 > 0.0) ? 1.0 / sqrt(t0) : 0.0);
        dinv_g[(long long)s * DP + b] = dinv[b];
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;                                   // log det of the scaled T (:94-99), bands in order
        if (mode == 0)
            for (int b = 0; b < D; ++b) {
                const double t0 = (G[(long long)b * DP + b] * sc[b] * sc[b] - (double)n * dm[b] * dm[b]) * inv_nm1;
                a += log(1.0e4 * t0);
            }
        slogT_g[s] = a;
        status_g[s] = kStatusOk;
    }
    for (long long idx = tid; idx < (long long)DP * DP; idx += blockDim.x) {
        const int r = (int)(idx / DP), c = (int)(idx % DP);
        double v = 0.0;
        if (r < D && c < D) {
            // the lower triangle is the value used for both halves, so the matrix is exactly symmetric
            const int rr = max(r, c), cc2 = min(r, c);
            v = (G[(long long)rr * DP + cc2] * sc[rr] * sc[cc2] - (double)n * dm[rr] * dm[cc2]) * inv_nm1;
            v *= dinv[rr] * dinv[cc2];
            if (r == c && mode == 0) v = (dinv[r] > 0.0) ? 1.0 : 0.0;
        }
        A[idx] = v;
    }
}

// ---------------------------------------------------------------------------------------- W3: tridiagonalisation
// Householder reduction (EISPACK tred2 scheme, the one k_eigen.cu runs in shared memory) of the D x D matrix of
// one column, in global memory / L2: one CTA per column, the leading block is kept fully symmetric so that
// every pass reads whole rows (coalesced), then the accumulation of the transformations leaves Q in the matrix.
constexpr int kWtThreads = 1024;      // largest shape (shared-memory plan of wide_eigen_fits)

template <int NTHR>
__device__ __forceinline__ double wt_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NTHR / 32; ++w) t += red[w];
    return t;
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR, 1024 / NTHR)
    wide_tred_kernel(double* __restrict__ work, const int* __restrict__ n_g, int D, int DP, double* __restrict__ d_g,
                     double* __restrict__ e_g) {
    constexpr int NW = NTHR / 32;
    extern __shared__ double sm[];
    double* u = sm;              // [DP] Householder vector (row i)
    double* ev = u + DP;         // [DP] e / p / q vector
    double* dv = ev + DP;        // [DP] d
    double* cv = dv + DP;        // [DP] column i (u / h) in the accumulation phase
    double* gw = cv + DP;        // [NW][DP] per-warp partial sums of the accumulation phase
    __shared__ double red[NW];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n_g[s] < 2) {
        for (int i = tid; i < DP; i += blockDim.x) { d_g[(long long)s * DP + i] = 0.0; e_g[(long long)s * DP + i] = 0.0; }
        return;
    }
    double* a = work + (long long)s * DP * DP;
    const int LD = DP;
    for (int i = tid; i < DP; i += blockDim.x) { dv[i] = 0.0; ev[i] = 0.0; }
    __syncthreads();
    for (int i = D - 1; i >= 1; --i) {
        const int l = i - 1;
        double h = 0.0;
        if (l > 0) {
            double part = 0.0;
            for (int k = tid; k <= l; k += blockDim.x) { const double v = a[(long long)i * LD + k]; u[k] = v; part += fabs(v); }
            const double scale = wt_block_sum<NTHR>(part, red);
            if (scale == 0.0) {
                if (tid == 0) ev[i] = u[l];
            } else {
                part = 0.0;
                for (int k = tid; k <= l; k += blockDim.x) { const double v = u[k] / scale; u[k] = v; part += v * v; }
                h = wt_block_sum<NTHR>(part, red);
                const double f = u[l];
                const double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
                h -= f * g;
                __syncthreads();
                if (tid == 0) { ev[i] = scale * g; u[l] = f - g; }
                __syncthreads();
                // row i keeps u, column i keeps u / h (read again by the accumulation phase)
                for (int k = tid; k <= l; k += blockDim.x) {
                    a[(long long)i * LD + k] = u[k];
                    a[(long long)k * LD + i] = u[k] / h;
                }
                // p = A u / h: one warp per row of the (full, symmetric) leading block; 4 loads in flight per lane
                for (int j = warp; j <= l; j += NW) {
                    const double* row = a + (long long)j * LD;
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                    int k = lane;
                    for (; k + 96 <= l; k += 128) {
                        const double x0 = row[k], x1 = row[k + 32], x2 = row[k + 64], x3 = row[k + 96];
                        g0 += x0 * u[k]; g1 += x1 * u[k + 32]; g2 += x2 * u[k + 64]; g3 += x3 * u[k + 96];
                    }
                    for (; k <= l; k += 32) g0 += row[k] * u[k];
                    double gj = (g0 + g1) + (g2 + g3);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) gj += __shfl_xor_sync(0xffffffffu, gj, o);
                    if (lane == 0) ev[j] = gj / h;
                }
                __syncthreads();
                part = 0.0;
                for (int j = tid; j <= l; j += blockDim.x) part += ev[j] * u[j];
                const double ff = wt_block_sum<NTHR>(part, red);
                const double hh = ff / (h + h);
                __syncthreads();
                for (int j = tid; j <= l; j += blockDim.x) ev[j] -= hh * u[j];
                __syncthreads();
                // A <- A - u q^T - q u^T on the whole leading block
                for (int j = warp; j <= l; j += NW) {
                    double* row = a + (long long)j * LD;
                    const double fj = u[j], gj = ev[j];
                    int k = lane;
                    for (; k + 96 <= l; k += 128) {
                        const double x0 = row[k], x1 = row[k + 32], x2 = row[k + 64], x3 = row[k + 96];
                        row[k] = x0 - (fj * ev[k] + gj * u[k]);
                        row[k + 32] = x1 - (fj * ev[k + 32] + gj * u[k + 32]);
                        row[k + 64] = x2 - (fj * ev[k + 64] + gj * u[k + 64]);
                        row[k + 96] = x3 - (fj * ev[k + 96] + gj * u[k + 96]);
                    }
                    for (; k <= l; k += 32) row[k] -= fj * ev[k] + gj * u[k];
                }
            }
        } else {
            if (tid == 0) ev[i] = a[(long long)i * LD + l];
        }
        if (tid == 0) dv[i] = h;
        __syncthreads();
    }
    if (tid == 0) { dv[0] = 0.0; ev[0] = 0.0; }
    __syncthreads();
    // ---- accumulate the transformations: the matrix becomes Q
    for (int i = 0; i < D; ++i) {
        const int l = i - 1;
        if (dv[i] != 0.0) {
            for (int k = tid; k <= l; k += blockDim.x) { u[k] = a[(long long)i * LD + k]; cv[k] = a[(long long)k * LD + i]; }
            __syncthreads();
            // g_j = sum_k u_k a[k][j]: warp w takes rows k = w, w + NW, ... (coalesced along j) into its own partial
            // vector; the partials are then added in warp order -> deterministic
            for (int j = lane; j <= l; j += 32) gw[warp * DP + j] = 0.0;
            for (int k = warp; k <= l; k += NW) {
                const double* row = a + (long long)k * LD;
                const double uk = u[k];
                int j = lane;
                for (; j + 96 <= l; j += 128) {
                    const double x0 = row[j], x1 = row[j + 32], x2 = row[j + 64], x3 = row[j + 96];
                    gw[warp * DP + j] += uk * x0; gw[warp * DP + j + 32] += uk * x1;
                    gw[warp * DP + j + 64] += uk * x2; gw[warp * DP + j + 96] += uk * x3;
                }
                for (; j <= l; j += 32) gw[warp * DP + j] += uk * row[j];
            }
            __syncthreads();
            for (int j = tid; j <= l; j += blockDim.x) {
                double g = 0.0;
                for (int w = 0; w < NW; ++w) g += gw[w * DP + j];
                u[j] = g;                                   // u (row i) is no longer needed: reuse it for g
            }
            __syncthreads();
            for (int k = warp; k <= l; k += NW) {
                double* row = a + (long long)k * LD;
                const double aki = cv[k];
                int j = lane;
                for (; j + 96 <= l; j += 128) {
                    const double x0 = row[j], x1 = row[j + 32], x2 = row[j + 64], x3 = row[j + 96];
                    row[j] = x0 - u[j] * aki; row[j + 32] = x1 - u[j + 32] * aki;
                    row[j + 64] = x2 - u[j + 64] * aki; row[j + 96] = x3 - u[j + 96] * aki;
                }
                for (; j <= l; j += 32) row[j] -= u[j] * aki;
            }
        }
        __syncthreads();
        if (tid == 0) { dv[i] = a[(long long)i * LD + i]; a[(long long)i * LD + i] = 1.0; }
        for (int j = tid; j <= l; j += blockDim.x) { a[(long long)j * LD + i] = 0.0; a[(long long)i * LD + j] = 0.0; }
        __syncthreads();
    }
    for (int i = tid; i < DP; i += blockDim.x) {
        d_g[(long long)s * DP + i] = (i < D) ? dv[i] : 0.0;
        e_g[(long long)s * DP + i] = (i < D) ? ev[i] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------- W3: blocked tridiagonalisation
// The same reduction in the blocked (LAPACK dsytrd / dlatrd / dorgtr) form, lower variant, panels of 8 columns.
// wide_tred_kernel above touches the whole trailing matrix three times per column (matvec, rank-2 update read and
// write) and ncu shows it bound by DRAM (800 MB of traffic per 416 x 416 matrix: 148 matrices do not fit the L2).
// Here the trailing matrix is only READ once per column (w = A v with the corrections -V (W^T v) - W (V^T v) of the
// panel's earlier reflectors) and updated once per panel (A -= V W^T + W V^T); Q is formed panel by panel with the
// compact WY form (I - V T V^T), two passes per 16 reflectors instead of per reflector: ~240 MB per matrix.
// A = Q T Q^T, Q = H_0 H_1 ... H_(n-2), H_c = I - tau_c v_c v_c^T, v_c zero above row c+1; v_c is kept in ROW c of
// the (full, symmetric) matrix.  The panels V, W live in shared memory t-major ([16][DP]) so that every access is
// either a broadcast or conflict-free.  Output: d, e in the EISPACK convention of wide_ql_kernel, Q in `work`.
constexpr int kTbNB = 8;       // 1024 threads leave 64 registers each: 2 x 8 panel values per thread in the update
constexpr int kTbThreads = 1024;

__device__ __forceinline__ double tb_block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kTbThreads / 32; ++w) t += red[w];
    return t;
}

__global__ void __launch_bounds__(kTbThreads, 1)
    wide_tredb_kernel(double* __restrict__ work, const int* __restrict__ n_g, int D, int DP, double* __restrict__ d_g,
                      double* __restrict__ e_g) {
    constexpr int NB = kTbNB, NW = kTbThreads / 32;
    extern __shared__ double sm[];
    double* Vt = sm;                    // [NB][DP]   (t-major)
    double* Wt = Vt + NB * DP;          // [NB][DP]
    double* vv = Wt + NB * DP;          // [DP] current reflector
    double* ww = vv + DP;               // [DP] column / w vector
    double* dv = ww + DP;               // [DP]
    double* ev = dv + DP;               // [DP]
    double* tauv = ev + DP;             // [DP]
    double* tmp = tauv + DP;            // [2 NB]
    double* Gm = tmp + 2 * NB;          // [NB][NB]
    double* Tm = Gm + NB * NB;          // [NB][NB]
    double* pbuf = Tm + NB * NB;        // [2][NB][512] partial sums of phase 2
    __shared__ double red[NW];
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n_g[s] < 2) {
        for (int i = tid; i < DP; i += blockDim.x) { d_g[(long long)s * DP + i] = 0.0; e_g[(long long)s * DP + i] = 0.0; }
        return;
    }
    double* A = work + (long long)s * DP * DP;
    const int LD = DP, n = D;
    for (int i = tid; i < DP; i += blockDim.x) { dv[i] = 0.0; ev[i] = 0.0; tauv[i] = 0.0; }
    // ---------------------------------------------------------------- phase 1: A -> tridiagonal, reflectors in rows
    for (int i0 = 0; i0 < n - 1; i0 += NB) {
        const int pb = min(NB, n - 1 - i0);
        for (int i = tid; i < 2 * NB * DP; i += blockDim.x) Vt[i] = 0.0;      // Vt and Wt are contiguous
        __syncthreads();
        for (int i = 0; i < pb; ++i) {
            const int c = i0 + i;
            // (1) column c of the matrix as updated by the panel's earlier reflectors (row c read: symmetric)
            for (int r = c + tid; r < n; r += blockDim.x) {
                double a = A[(long long)c * LD + r];
                for (int t = 0; t < i; ++t)
                    a -= __dadd_rn(__dmul_rn(Vt[t * DP + r], Wt[t * DP + c]), __dmul_rn(Wt[t * DP + r], Vt[t * DP + c]));
                ww[r] = a;
            }
            __syncthreads();
            // (2) reflector annihilating rows c+2.. of the column
            double part = 0.0;
            for (int r = c + 2 + tid; r < n; r += blockDim.x) part += ww[r] * ww[r];
            const double xn2 = tb_block_sum(part, red);
            const double alpha = ww[c + 1];
            double beta = alpha, tau_c = 0.0, scal = 0.0;
            if (xn2 > 0.0) {
                beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
                tau_c = (beta - alpha) / beta;
                scal = 1.0 / (alpha - beta);
            }
            if (tid == 0) { dv[c] = ww[c]; ev[c + 1] = beta; tauv[c] = tau_c; }
            for (int r = c + 1 + tid; r < n; r += blockDim.x) {
                const double v = (r == c + 1) ? 1.0 : ww[r] * scal;
                vv[r] = v;
                Vt[i * DP + r] = v;
            }
            __syncthreads();
            // (3) w = A[c+1:, c+1:] v on the matrix as it was when the panel started: one warp per row
            for (int r = c + 1 + warp; r < n; r += NW) {
                const double* row = A + (long long)r * LD;
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                int k = c + 1 + lane;
                for (; k + 96 < n; k += 128) {
                    const double x0 = row[k], x1 = row[k + 32], x2 = row[k + 64], x3 = row[k + 96];
                    g0 += x0 * vv[k]; g1 += x1 * vv[k + 32]; g2 += x2 * vv[k + 64]; g3 += x3 * vv[k + 96];
                }
                for (; k < n; k += 32) g0 += row[k] * vv[k];
                double gj = (g0 + g1) + (g2 + g3);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) gj += __shfl_xor_sync(0xffffffffu, gj, o);
                if (lane == 0) ww[r] = gj;
            }
            // (4) tmp[2t] = W_t . v, tmp[2t+1] = V_t . v over rows c+1.. : one warp per product (2 i <= 32)
            if (warp < 2 * i) {
                const double* src = ((warp & 1) ? Vt : Wt) + (warp >> 1) * DP;
                double g = 0.0;
                for (int r = c + 1 + lane; r < n; r += 32) g += src[r] * vv[r];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
                if (lane == 0) tmp[warp] = g;
            }
            __syncthreads();
            // (5) w <- tau (w - V (W^T v) - W (V^T v)); w <- w - (tau/2) (w . v) v
            part = 0.0;
            for (int r = c + 1 + tid; r < n; r += blockDim.x) {
                double w = ww[r];
                for (int t = 0; t < i; ++t)
                    w -= __dadd_rn(__dmul_rn(Vt[t * DP + r], tmp[2 * t]), __dmul_rn(Wt[t * DP + r], tmp[2 * t + 1]));
                w *= tau_c;
                ww[r] = w;
                part += w * vv[r];
            }
            const double wv = tb_block_sum(part, red);
            const double alpha2 = -0.5 * tau_c * wv;
            for (int r = c + 1 + tid; r < n; r += blockDim.x) Wt[i * DP + r] = ww[r] + alpha2 * vv[r];
            __syncthreads();
        }
        // the panel's reflectors go into rows c of the matrix (dead from now on), v_c[c+1] = 1 stored explicitly
        for (int i = warp; i < pb; i += NW) {
            const int c = i0 + i;
            for (int r = c + 1 + lane; r < n; r += 32) A[(long long)c * LD + r] = Vt[i * DP + r];
        }
        // trailing update A[r0:, r0:] -= V W^T + W V^T (every element, so the matrix stays exactly symmetric)
        const int r0 = i0 + pb;
        for (int r = r0 + warp; r < n; r += NW) {
            double vr[NB], wr[NB];
#pragma unroll
            for (int t = 0; t < NB; ++t) { vr[t] = Vt[t * DP + r]; wr[t] = Wt[t * DP + r]; }
            double* row = A + (long long)r * LD;
            for (int k = r0 + lane; k < n; k += 32) {
                double sacc = 0.0;
#pragma unroll
                for (int t = 0; t < NB; ++t)
                    sacc += __dadd_rn(__dmul_rn(vr[t], Wt[t * DP + k]), __dmul_rn(wr[t], Vt[t * DP + k]));
                row[k] -= sacc;
            }
        }
        __syncthreads();
    }
    if (tid == 0) dv[n - 1] = A[(long long)(n - 1) * LD + (n - 1)];
    __syncthreads();
    for (int i = tid; i < DP; i += blockDim.x) {
        d_g[(long long)s * DP + i] = (i < n) ? dv[i] : 0.0;
        e_g[(long long)s * DP + i] = (i < n) ? ev[i] : 0.0;
    }
    // ---------------------------------------------------------------- phase 2: Q = H_0 ... H_(n-2), panels in reverse
    int prev_lo = n;                                   // Q occupies [prev_lo:, prev_lo:] so far
    for (int i0 = ((n - 2) / NB) * NB; i0 >= 0; i0 -= NB) {
        const int pb = min(NB, n - 1 - i0);
        const int lo = i0 + 1;
        // the panel's reflectors, from rows c of the matrix
        for (int i = tid; i < NB * DP; i += blockDim.x) Vt[i] = 0.0;
        __syncthreads();
        for (int i = warp; i < pb; i += NW) {
            const int c = i0 + i;
            for (int r = c + 1 + lane; r < n; r += 32) Vt[i * DP + r] = A[(long long)c * LD + r];
        }
        __syncthreads();
        // extend Q to [lo:, lo:] with the identity: rows lo..prev_lo-1 entirely, columns lo..prev_lo-1 of the rows below
        for (int r = lo + warp; r < n; r += NW) {
            double* row = A + (long long)r * LD;
            const int kend = (r < prev_lo) ? n : prev_lo;
            for (int k = lo + lane; k < kend; k += 32) row[k] = (k == r) ? 1.0 : 0.0;
        }
        // Gram of the reflectors G[t][u] = v_t . v_u (u < t), one warp per pair
        for (int pidx = warp; pidx < pb * pb; pidx += NW) {
            const int t = pidx / pb, u = pidx % pb;
            if (u >= t) continue;
            double g = 0.0;
            for (int r = lo + lane; r < n; r += 32) g += Vt[t * DP + r] * Vt[u * DP + r];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
            if (lane == 0) { Gm[t * NB + u] = g; Gm[u * NB + t] = g; }
        }
        __syncthreads();
        // T of the compact WY form (forward, columnwise): T[i][i] = tau_i, T[0:i, i] = -tau_i T[0:i, 0:i] (V^T v_i)
        if (tid == 0) {
            for (int i = 0; i < pb; ++i) {
                const double ti = tauv[i0 + i];
                for (int u = 0; u < i; ++u) {
                    double a = 0.0;
                    for (int k = u; k < i; ++k) a += Tm[u * NB + k] * Gm[k * NB + i];
                    Tm[u * NB + i] = -ti * a;
                }
                Tm[i * NB + i] = ti;
            }
        }
        __syncthreads();
        // Q_sub <- Q_sub - V (T (V^T Q_sub)): thread <-> (column j, row parity), coalesced across j; the two partial
        // sums of V^T Q are added in a fixed order
        {
            const int part = tid >> 9, jj = tid & 511;
            for (int jb = lo; jb < n; jb += 512) {
                const int j = jb + jj;
                double acc[NB];
#pragma unroll
                for (int t = 0; t < NB; ++t) acc[t] = 0.0;
                if (j < n) {
                    int r = lo + part;
                    for (; r + 6 < n; r += 8) {
                        const double q0 = A[(long long)r * LD + j], q1 = A[(long long)(r + 2) * LD + j],
                                     q2 = A[(long long)(r + 4) * LD + j], q3 = A[(long long)(r + 6) * LD + j];
#pragma unroll
                        for (int t = 0; t < NB; ++t)
                            acc[t] += (Vt[t * DP + r] * q0 + Vt[t * DP + r + 2] * q1) +
                                      (Vt[t * DP + r + 4] * q2 + Vt[t * DP + r + 6] * q3);
                    }
                    for (; r < n; r += 2) {
                        const double qv = A[(long long)r * LD + j];
#pragma unroll
                        for (int t = 0; t < NB; ++t) acc[t] += Vt[t * DP + r] * qv;
                    }
                }
#pragma unroll
                for (int t = 0; t < NB; ++t) pbuf[(part * NB + t) * 512 + jj] = acc[t];
                __syncthreads();
                if (j < n) {
#pragma unroll
                    for (int t = 0; t < NB; ++t) acc[t] = pbuf[t * 512 + jj] + pbuf[(NB + t) * 512 + jj];
                    // x = T acc, in place (x[t] needs acc[u >= t] only, T is upper triangular)
#pragma unroll
                    for (int t = 0; t < NB; ++t) {
                        double a = 0.0;
#pragma unroll
                        for (int u = 0; u < NB; ++u)
                            if (u >= t && u < pb && t < pb) a += Tm[t * NB + u] * acc[u];
                        acc[t] = a;
                    }
                    int r = lo + part;
                    for (; r + 6 < n; r += 8) {
                        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
#pragma unroll
                        for (int t = 0; t < NB; ++t) {
                            a0 += Vt[t * DP + r] * acc[t]; a1 += Vt[t * DP + r + 2] * acc[t];
                            a2 += Vt[t * DP + r + 4] * acc[t]; a3 += Vt[t * DP + r + 6] * acc[t];
                        }
                        A[(long long)r * LD + j] -= a0; A[(long long)(r + 2) * LD + j] -= a1;
                        A[(long long)(r + 4) * LD + j] -= a2; A[(long long)(r + 6) * LD + j] -= a3;
                    }
                    for (; r < n; r += 2) {
                        double a = 0.0;
#pragma unroll
                        for (int t = 0; t < NB; ++t) a += Vt[t * DP + r] * acc[t];
                        A[(long long)r * LD + j] -= a;
                    }
                }
                __syncthreads();
            }
        }
        prev_lo = lo;
        __syncthreads();
    }
    // row and column 0 of Q
    for (int k = tid; k < n; k += blockDim.x) {
        A[k] = (k == 0) ? 1.0 : 0.0;
        A[(long long)k * LD] = (k == 0) ? 1.0 : 0.0;
    }
    if (n == 1 && tid == 0) A[0] = 1.0;
}

// ---------------------------------------------------------------------------------------- W3: QL recurrences
// Implicit-shift QL (tql2) on the tridiagonal (d, e) of every column, WITHOUT the eigenvectors: the plane-rotation
// recurrence is a serial FP64 chain that only needs (d, e), so one warp per column runs it (lane 0; all lanes
// scan for the deflation point) and writes each iteration's rotation sequence to global memory; W3c applies
// them to Q with full parallelism.  Same arithmetic as the producer warp of eigen_ql_kernel.
constexpr int kWqMaxIter = 60;

__global__ void __launch_bounds__(32)
    wide_ql_kernel(const double* __restrict__ d_g, const double* __restrict__ e_g, const int* __restrict__ n_g, int D,
                   int DP, double2* __restrict__ rot, long long rot_cap, int2* __restrict__ iters, int iter_cap,
                   int* __restrict__ niter_g, double* __restrict__ lam_g, int* __restrict__ status_g) {
    extern __shared__ double sm[];
    double* d = sm;            // [DP]
    double* e = d + DP;        // [DP]
    const int s = blockIdx.x, lane = threadIdx.x;
    if (n_g[s] < 2) {
        for (int i = lane; i < DP; i += 32) lam_g[(long long)s * DP + i] = 0.0;
        if (lane == 0) niter_g[s] = 0;
        return;
    }
    for (int i = lane; i < DP; i += 32) {
        d[i] = d_g[(long long)s * DP + i];
        e[i] = (i + 1 < D) ? e_g[(long long)s * DP + i + 1] : 0.0;      // e[i-1] = e[i]
    }
    __syncwarp();
    double2* myrot = rot + (long long)s * rot_cap;
    int2* myit = iters + (long long)s * iter_cap;
    long long nrot = 0;
    int nit = 0;
    bool failed = false;
    for (int l = 0; l < D && !failed; ++l) {
        for (int iter = 0;; ++iter) {
            int m = D - 1;
            for (int base = l; base < D - 1; base += 32) {
                const int mm = base + lane;
                bool small = false;
                if (mm < D - 1) {
                    const double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                    small = fabs(e[mm]) <= 1.1102230246251565e-16 * dd;
                }
                const unsigned hit = __ballot_sync(0xffffffffu, small);
                if (hit) { m = base + __ffs(hit) - 1; break; }
            }
            if (m == l) break;
            if (iter >= kWqMaxIter || nit >= iter_cap || nrot + (m - l) > rot_cap) { failed = true; break; }
            int cnt = 0;
            if (lane == 0) {
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = sqrt(g * g + 1.0);
                g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                double sn = 1.0, c = 1.0, p = 0.0;
                int i = m - 1;
                bool under = false;
                double e_i = e[i], d_i = d[i], d_i1 = d[i + 1];
                for (; i >= l && !under; --i) {
                    double e_n = 0.0, d_n = 0.0;
                    if (i > l) { e_n = e[i - 1]; d_n = d[i - 1]; }
                    const double f = sn * e_i;
                    const double b = c * e_i;
                    const double h = f * f + g * g;
                    under = !(h > 0.0);
                    double x0;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(h));
                    const double err = fma(-h * x0, x0, 1.0);
                    const double rinv = under ? 0.0 : fma(fma(err, 0.375, 0.5), x0 * err, x0);
                    const double g2 = d_i1 - p;
                    const double w = fma(d_i - g2, f, 2.0 * g * b);
                    e[i + 1] = h * rinv;
                    sn = f * rinv;
                    c = g * rinv;
                    const double r2 = w * rinv;
                    if (!under) {
                        p = sn * r2;
                        d[i + 1] = g2 + p;
                        g = c * r2 - b;
                        myrot[nrot + cnt] = make_double2(c, sn);        // rotation of columns (i, i+1)
                        ++cnt;
                    } else {
                        d[i + 1] = g2;
                        e[m] = 0.0;
                    }
                    e_i = e_n; d_i1 = d_i; d_i = d_n;
                }
                if (!under) { d[l] -= p; e[l] = g; e[m] = 0.0; }
                myit[nit] = make_int2(m, cnt);
            }
            cnt = __shfl_sync(0xffffffffu, cnt, 0);
            nrot += cnt;
            ++nit;
            __syncwarp();
        }
    }
    __syncwarp();
    for (int i = lane; i < DP; i += 32) lam_g[(long long)s * DP + i] = (i < D) ? d[i] : 0.0;
    if (lane == 0) {
        niter_g[s] = nit;
        if (failed) status_g[s] |= kStatusNoConverge;
    }
}

// ---------------------------------------------------------------------------------------- W3: apply rotations
// Q <- Q J_1 J_2 ...: the rows of Q are independent, so CTA = (column, block of 64 rows) keeps its rows in shared
// memory ([column of Q][row], 64 threads, thread <-> row) and streams the rotation sequences of W3b through a
// double-buffered shared-memory window.  Then P = T^-1/2 V leaves in row-major order.
constexpr int kWrThreads = 64;

__global__ void __launch_bounds__(kWrThreads, 1)
    wide_rot_kernel(const double* __restrict__ work, const int* __restrict__ n_g, int D, int DP, int rows,
                    const double2* __restrict__ rot, long long rot_cap, const int2* __restrict__ iters, int iter_cap,
                    const int* __restrict__ niter_g, const double* __restrict__ dinv_g, double* __restrict__ P_g) {
    extern __shared__ double sm[];
    double* q = sm;                                                     // [DP][rows]
    double2* rbuf = reinterpret_cast<double2*>(q + (size_t)DP * rows);  // [2][DP]
    const int s = blockIdx.y, r0 = blockIdx.x * rows, tid = threadIdx.x;
    const int row = r0 + tid;
    const bool mine = tid < rows && row < DP;
    double* Pout = P_g + (long long)s * DP * DP;
    if (n_g[s] < 2) {
        if (mine) for (int j = 0; j < DP; ++j) Pout[(long long)row * DP + j] = 0.0;
        return;
    }
    const double* a = work + (long long)s * DP * DP;
    // rows of Q, read along the rows, kept transposed as q[column][row]
    for (int idx = tid; idx < rows * DP; idx += kWrThreads) {
        const int rr = idx / DP, c = idx - rr * DP;
        q[c * rows + rr] = (r0 + rr < D && c < D) ? a[(long long)(r0 + rr) * DP + c] : 0.0;
    }
    const int nit = niter_g[s];
    const double2* myrot = rot + (long long)s * rot_cap;
    const int2* myit = iters + (long long)s * iter_cap;
    long long off = 0;
    int2 cur = nit > 0 ? myit[0] : make_int2(0, 0);
    for (int t = tid; t < cur.y; t += kWrThreads) rbuf[t] = myrot[t];
    __syncthreads();
    for (int it = 0; it < nit; ++it) {
        const int buf = it & 1;
        const int2 nxt = (it + 1 < nit) ? myit[it + 1] : make_int2(0, 0);
        const long long noff = off + cur.y;
        // the next iteration's rotations are fetched before this one's are applied
        double2 pre[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = tid + k * kWrThreads;
            pre[k] = (t < nxt.y) ? myrot[noff + t] : make_double2(1.0, 0.0);
        }
        const int m = cur.x, cnt = cur.y;
        if (cnt > 0 && tid < rows) {
            const double2* cr = rbuf + buf * DP;
            double f = q[m * rows + tid];
            // the chain through f is two dependent FP64 operations per rotation; everything else (the rotation, the
            // element it meets, s * z for the stored column) is loaded / formed eight rotations ahead of it
            int t = 0;
            for (; t + 8 <= cnt; t += 8) {
                double2 cs[8];
                double z[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) { cs[k] = cr[t + k]; z[k] = q[(m - 1 - t - k) * rows + tid]; }
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int i = m - 1 - t - k;             // rotation t + k acts on columns (i, i + 1)
                    q[(i + 1) * rows + tid] = fma(cs[k].x, f, cs[k].y * z[k]);
                    f = fma(-cs[k].y, f, cs[k].x * z[k]);       // one FMA on the chain: c * z does not wait for f
                }
            }
            for (; t < cnt; ++t) {
                const int i = m - 1 - t;
                const double2 c1 = cr[t];
                const double zi = q[i * rows + tid];
                q[(i + 1) * rows + tid] = fma(c1.x, f, c1.y * zi);
                f = fma(-c1.y, f, c1.x * zi);
            }
            q[(m - cnt) * rows + tid] = f;
        }
        double2* nb = rbuf + (buf ^ 1) * DP;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int t = tid + k * kWrThreads;
            if (t < nxt.y) nb[t] = pre[k];
        }
        for (int t = tid + 8 * kWrThreads; t < nxt.y; t += kWrThreads) nb[t] = myrot[noff + t];
        __syncthreads();
        off = noff;
        cur = nxt;
    }
    // P[b][j] = dinv[b] V[b][j]: thread <-> row
    if (mine) {
        const double di = dinv_g[(long long)s * DP + row];
        for (int j = 0; j < DP; ++j) Pout[(long long)row * DP + j] = (row < D && j < D) ? di * q[j * rows + tid] : 0.0;
    }
}

// ---------------------------------------------------------------------------------------- W4
// log det G_alpha, beta, the closed form of sum_k r_k (as tables_kernel) and W[j][alpha] = 1/(n beta lam_j + alpha)
// as a row-major [DP][APW] table.
__global__ void __launch_bounds__(256)
    wide_tables_kernel(const int* __restrict__ n_g, const int* __restrict__ nloo_g, const double* __restrict__ alphas,
                       int A, int AP, int APW, int D, int DP, int model, const double* __restrict__ lam_g,
                       const double* __restrict__ slogT_g, double* __restrict__ logdet_g, double* __restrict__ beta_g,
                       double* __restrict__ rsum_g, double* __restrict__ W_g) {
    extern __shared__ double lam[];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int n = n_g[s];
    double* W = W_g + (long long)s * DP * APW;
    if (model != 0) return;
    if (n < 2) {
        for (int i = tid; i < AP; i += blockDim.x) {
            logdet_g[(long long)s * AP + i] = 0.0; beta_g[(long long)s * AP + i] = 0.0; rsum_g[(long long)s * AP + i] = 0.0;
        }
        for (long long i = tid; i < (long long)DP * APW; i += blockDim.x) W[i] = 0.0;
        return;
    }
    for (int j = tid; j < DP; j += blockDim.x) lam[j] = lam_g[(long long)s * DP + j];
    __syncthreads();
    const double dn = (double)(nloo_g ? nloo_g[s] : n);
    const double sumlogT = slogT_g[s];
    for (int i = tid; i < AP; i += blockDim.x) {
        double ld = 0.0, be = 0.0, rs = 0.0;
        if (i < A) {
            const double al = alphas[i];
            be = (1.0 - al) / (dn - 1.0);
            ld = sumlogT;
            for (int j = 0; j < D; ++j) {
                const double den = dn * be * lam[j] + al;
                ld += log(den);
                rs += lam[j] / den;
            }
            rs *= ((double)n - 1.0);
        }
        logdet_g[(long long)s * AP + i] = ld;
        beta_g[(long long)s * AP + i] = be;
        rsum_g[(long long)s * AP + i] = rs;
    }
    for (long long idx = tid; idx < (long long)DP * APW; idx += blockDim.x) {
        const int j = (int)(idx / APW), i = (int)(idx % APW);
        double w = 0.0;
        if (j < D && i < A) {
            const double al = alphas[i];
            const double be = (1.0 - al) / (dn - 1.0);
            w = 1.0 / (dn * be * lam[j] + al);
        }
        W[idx] = w;
    }
}

// ---------------------------------------------------------------------------------------- W5 / W6
// C[m][n] = sum_k A[m][k] B[k][n] for a 128 x 64 block with A row-major in global memory (pitch lda, rows are pixels)
// and B row-major (pitch ldb).  8 warps (4 in m x 2 in n), each a 32 x 32 block of 4 x 4 DMMA tiles; 16-deep stages:
// A is kept as it arrives, [m][k] with a pitch of 20 doubles (== 4 mod 16: the fragment reads of a half-warp and the
// stores of a half-warp both fall on 16 different 8-byte banks), B as [k][n] with pitch kWP.  The next stage is
// fetched into registers while the current one is multiplied.
constexpr int kRP = 20;
constexpr int kRowsThreads = 256;

__device__ __forceinline__ void dmma_stage_rows(const double* __restrict__ As, const double* __restrict__ Bs,
                                                double (&acc)[4][4][2], int wm, int wn, int lane) {
    const int g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        double a[4], b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            a[t] = As[(wm * 32 + t * 8 + g) * kRP + 4 * kk + q];
            b[t] = Bs[(4 * kk + q) * kWP + wn * 32 + t * 8 + g];
        }
#pragma unroll
        for (int ti = 0; ti < 4; ++ti)
#pragma unroll
            for (int tj = 0; tj < 4; ++tj) mma884(acc[ti][tj][0], acc[ti][tj][1], a[ti], b[tj]);
    }
}

template <typename TA, bool CENTRE>
__device__ __forceinline__ void gemm_rows_block(const TA* __restrict__ A, long long lda, int m_valid, int K,
                                                const double* __restrict__ ctr_s, const double* __restrict__ B,
                                                long long ldb, int n_valid, double* As, double* Bs,
                                                double (&acc)[4][4][2], int tid) {
    const int lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
    // A stage: element e = tid + 256 r: k = e % 16, row m = e / 16;  B stage: n = e % 64, k = e / 64
    const int ak = tid & 15, am = tid >> 4, bn = tid & 63, bk = tid >> 6;
    TA ra[8];
    double rb[4];
    auto fetch = [&](int k0) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const int m = am + 16 * r, k = k0 + ak;
            ra[r] = (m < m_valid && k < K) ? A[(long long)m * lda + k] : (TA)0;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int kb = k0 + bk + 4 * r;
            rb[r] = (kb < K && bn < n_valid) ? B[(long long)kb * ldb + bn] : 0.0;
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < K; k0 += 16) {
        __syncthreads();
        {
            const int k = k0 + ak;
            const double c = (CENTRE && k < K) ? ctr_s[k] : 0.0;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const TA x = ra[r];
                As[(am + 16 * r) * kRP + ak] = (k < K && am + 16 * r < m_valid) ? (CENTRE ? centred(x, c) : (double)x) : 0.0;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) Bs[(bk + 4 * r) * kWP + bn] = rb[r];
        __syncthreads();
        if (k0 + 16 < K) fetch(k0 + 16);
        dmma_stage_rows(As, Bs, acc, wm, wn, lane);
    }
}

// Z[px][j] = (sum_b (x[px][b] - mu_b) P[b][j])^2 for the columns s0 .. s0 + gridDim.z - 1 (Z is a per-batch buffer)
template <typename T>
__global__ void __launch_bounds__(kRowsThreads, 2)
    wide_proj_kernel(const T* __restrict__ xt, const double* __restrict__ mu_g, const double* __restrict__ P_g,
                     const int* __restrict__ n_g, int L, int D, int DP, int s0, double* __restrict__ Z) {
    extern __shared__ double rows_sm[];
    double* As = rows_sm;                    // [128][kRP]
    double* Bs = As + 128 * kRP;             // [16][kWP]
    double* mu_s = Bs + 16 * kWP;            // [DP]
    const int s = s0 + blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (n_g[s] < 2) return;
    const int j0 = blockIdx.x * 64, p0 = blockIdx.y * 128;
    for (int i = tid; i < DP; i += blockDim.x) mu_s[i] = mu_g[(long long)s * DP + i];
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm_rows_block<T, true>(xt + ((long long)s * L + p0) * DP, DP, min(128, L - p0), D, mu_s,
                             P_g + (long long)s * DP * DP + j0, DP, min(64, DP - j0), As, Bs, acc, tid);
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double* Zc = Z + (long long)blockIdx.z * L * DP;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
            const int px = p0 + wm * 32 + ti * 8 + g, j = j0 + wn * 32 + tj * 8 + 2 * q;
            if (px < L && j < DP) {      // DP is a multiple of 8: j + 1 < DP too
                const double y0 = acc[ti][tj][0], y1 = acc[ti][tj][1];
                *reinterpret_cast<double2*>(Zc + (long long)px * DP + j) = make_double2(y0 * y0, y1 * y1);
            }
        }
}

// log(q) + r/q, the same evaluation as k_loo.cu (series for small u = beta r)
__device__ __forceinline__ double wide_loo_term(double r, double beta) {
    const double u = beta * r;
    const double au = fabs(u);
    if (au <= 0x1p-7) {
        double s1 = fma(u, 1.0, 1.0);
        double s2 = fma(u, 1.0 / 8.0, 1.0 / 7.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 6.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 5.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 4.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 3.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0 / 2.0);
        s1 = fma(u, s1, 1.0); s2 = fma(u, s2, 1.0);
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

// fpart[s][chunk][alpha] = sum over the chunk's pixels of log q + r/q with r = sum_j Z[px][j] W[j][alpha]
__global__ void __launch_bounds__(kRowsThreads, 2)
    wide_loo_kernel(const double* __restrict__ Z, const double* __restrict__ W_g, const double* __restrict__ beta_g,
                    const int* __restrict__ n_g, int L, int D, int DP, int AP, int APW, int s0, int lines_per_chunk,
                    int nchunk, double* __restrict__ fpart) {
    extern __shared__ double rows_sm[];
    double* As = rows_sm;                    // [128][kRP]
    double* Bs = As + 128 * kRP;             // [16][kWP]
    double* xch = Bs + 16 * kWP;             // [4][64]
    const int s = s0 + blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = blockIdx.x * 64, chunk = blockIdx.y;
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
    double* out = fpart + ((long long)s * nchunk + chunk) * AP;
    if (n_g[s] < 2) return;
    const int c_begin = chunk * lines_per_chunk, c_end = min(L, c_begin + lines_per_chunk);
    double be[4][2], fsum[4][2];
#pragma unroll
    for (int tj = 0; tj < 4; ++tj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int al = a0 + wn * 32 + tj * 8 + 2 * q + e;
            be[tj][e] = al < AP ? beta_g[(long long)s * AP + al] : 0.0;
            fsum[tj][e] = 0.0;
        }
    const double* Zc = Z + (long long)blockIdx.z * L * DP;
    for (int p0 = c_begin; p0 < c_end; p0 += 128) {
        double acc[4][4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        gemm_rows_block<double, false>(Zc + (long long)p0 * DP, DP, min(128, c_end - p0), D, nullptr,
                                       W_g + (long long)s * DP * APW + a0, APW, min(64, APW - a0), As, Bs, acc, tid);
        // rows past the chunk end were loaded as zero: r = 0 and the term vanishes
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double f = 0.0;
#pragma unroll
                for (int ti = 0; ti < 4; ++ti) f += wide_loo_term(acc[ti][tj][e], be[tj][e]);
                fsum[tj][e] += f;
            }
    }
    // sum over the 8 pixel rows of a fragment (lanes with equal q), then over the four warps in m, fixed order
#pragma unroll
    for (int tj = 0; tj < 4; ++tj)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double f = fsum[tj][e];
            f += shfl_xor_f64(f, 4); f += shfl_xor_f64(f, 8); f += shfl_xor_f64(f, 16);
            fsum[tj][e] = f;
        }
    __syncthreads();
    if (g == 0) {
#pragma unroll
        for (int tj = 0; tj < 4; ++tj)
#pragma unroll
            for (int e = 0; e < 2; ++e) xch[wm * 64 + wn * 32 + tj * 8 + 2 * q + e] = fsum[tj][e];
    }
    __syncthreads();
    if (tid < 64 && a0 + tid < AP) out[a0 + tid] = ((xch[tid] + xch[64 + tid]) + xch[128 + tid]) + xch[192 + tid];
}

// ---------------------------------------------------------------------------------------- -f on a wide window
// The full-column covariance as the shrinkage target (cmf/robust_mf.py:99, :353-356): T = cov(I_reg) is whitened
// with its own spectral factor W = T0^-1/2 U M^-1/2 (T0 = diag T, T0^-1/2 T T0^-1/2 = U M U^T from one solve of the
// blocked eigen-solver per run, W^T T W = I), the mode's covariance becomes R = W^T S W, and with R = V Lambda V^T
// the search sees P = W V and log det T = sum log T0 + sum log M exactly as with the diagonal target.
// C = A B for the DP x DP row-major matrices of every column (FP64 DMMA, the 128 x 64 blocks of W5)
__global__ void __launch_bounds__(kRowsThreads, 2)
    wide_dgemm_kernel(const double* __restrict__ A_g, const double* __restrict__ B_g, int D, int DP,
                      double* __restrict__ C_g) {
    extern __shared__ double rows_sm[];
    double* As = rows_sm;                    // [128][kRP]
    double* Bs = As + 128 * kRP;             // [16][kWP]
    const int s = blockIdx.z, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int j0 = blockIdx.x * 64, r0 = blockIdx.y * 128;
    const long long base = (long long)s * DP * DP;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    gemm_rows_block<double, false>(A_g + base + (long long)r0 * DP, DP, min(128, DP - r0), D, nullptr, B_g + base + j0, DP,
                                   min(64, DP - j0), As, Bs, acc, tid);
    const int wm = warp >> 1, wn = warp & 1, g = lane >> 2, q = lane & 3;
#pragma unroll
    for (int ti = 0; ti < 4; ++ti)
#pragma unroll
        for (int tj = 0; tj < 4; ++tj) {
            const int r = r0 + wm * 32 + ti * 8 + g, j = j0 + wn * 32 + tj * 8 + 2 * q;
            if (r < DP && j < DP)
                *reinterpret_cast<double2*>(C_g + base + (long long)r * DP + j) = make_double2(acc[ti][tj][0], acc[ti][tj][1]);
        }
}

// W = P_f diag(M^-1/2) and its transpose, log det of the scaled target, status of the target (singular: some M <= 0)
__global__ void __launch_bounds__(256)
    wide_target_kernel(const double* __restrict__ P_g, const double* __restrict__ lam_g,
                       const double* __restrict__ slogT_g, const int* __restrict__ status_g, int D, int DP,
                       double* __restrict__ W_g, double* __restrict__ Wt_g, double* __restrict__ slogT_full,
                       int* __restrict__ fstatus) {
    extern __shared__ double isq[];          // [DP]
    const int s = blockIdx.x, tid = threadIdx.x;
    const double* lam = lam_g + (long long)s * DP;
    for (int j = tid; j < DP; j += blockDim.x) isq[j] = (j < D && lam[j] > 0.0) ? 1.0 / sqrt(lam[j]) : 0.0;
    __syncthreads();
    if (tid == 0) {
        double a = slogT_g[s];
        int bad = status_g[s] & (kStatusNoConverge | kStatusEmpty | kStatusDegenerate);
        for (int j = 0; j < D; ++j) {
            if (lam[j] > 0.0) a += log(lam[j]);
            else bad |= kStatusSingular;
        }
        slogT_full[s] = a;
        fstatus[s] = bad;
    }
    const long long base = (long long)s * DP * DP;
    for (long long idx = tid; idx < (long long)DP * DP; idx += blockDim.x) {
        const int b = (int)(idx / DP), j = (int)(idx % DP);
        const double v = P_g[base + idx] * isq[j];
        W_g[base + idx] = v;
        Wt_g[base + (long long)j * DP + b] = v;
    }
}

// the whitened covariance is symmetric up to rounding: the lower triangle is the value both halves use
__global__ void __launch_bounds__(256) wide_mirror_kernel(double* __restrict__ work, int DP) {
    double* A = work + (long long)blockIdx.x * DP * DP;
    for (long long idx = threadIdx.x; idx < (long long)DP * DP; idx += blockDim.x) {
        const int r = (int)(idx / DP), c = (int)(idx % DP);
        if (c > r) A[idx] = A[(long long)c * DP + r];
    }
}

__global__ void __launch_bounds__(256)
    wide_target_finish_kernel(const double* __restrict__ slogT_full, const int* __restrict__ fstatus, int S,
                              double* __restrict__ slogT, int* __restrict__ status) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    slogT[s] = slogT_full[s];
    if ((status[s] & (kStatusEmpty | kStatusDegenerate)) == 0) status[s] |= fstatus[s];
}

// ---------------------------------------------------------------------------------------- launchers
void launch_wide_stats(const Dims& d, const float* slab, uint8_t* mask, const uint8_t* sel, int write_mask,
                       int nsplit, int lps, double* colsum_part, int* colcnt_part, float* lo_part, float* hi_part,
                       double* mu, int* n, double* ctr, int* qexp, cudaStream_t st) {
    if (write_mask) {
        dim3 grid((d.S + 31) / 32, (d.L + 7) / 8);
        wide_valid_kernel<<<grid, 256, 0, st>>>(slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, mask);
    }
    dim3 grid2((d.S + 31) / 32, (d.DP + 7) / 8, nsplit);
    wide_sums_kernel<<<grid2, 256, 0, st>>>(slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, d.DP, mask, sel, lps,
                                            colsum_part, colcnt_part, lo_part, hi_part);
    wide_mean_kernel<<<d.S, 128, 0, st>>>(colsum_part, colcnt_part, lo_part, hi_part, nsplit, d.S, d.D, d.DP, mu, n,
                                          ctr, qexp);
}

void launch_wide_pack(const Dims& d, const float* slab, const uint8_t* mask, const uint8_t* sel, const double* ctr,
                      const int* qexp, float* xt, int8_t* img, cudaStream_t st) {
    const int nrb = (d.D + 31) / 32, nkb = (d.L + 63) / 64;
    dim3 grid((d.S + 31) / 32, nkb * 4);
    const size_t smem = 16 * 32 * 33 * sizeof(float) + 16 * 32;
    cudaFuncSetAttribute(wide_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    wide_pack_kernel<<<grid, 256, smem, st>>>(slab, d.line_pitch, d.band_pitch, d.L, d.S, d.D, d.DP, mask, sel, ctr, qexp,
                                           xt, img, nrb, nkb);
}

void launch_wide_gram64(const Dims& d, const float* xt, const double* ctr, double* gram, cudaStream_t st) {
    const int nb = (d.DP + 63) / 64;
    dim3 grid(nb * (nb + 1) / 2, d.S);
    wide_gram64_kernel<float><<<grid, 128, 0, st>>>(xt, ctr, d.L, d.DP, gram);
}

void launch_wide_gram64_f64(int L, int DP, int S, const double* x, const double* ctr, double* gram, cudaStream_t st) {
    const int nb = (DP + 63) / 64;
    dim3 grid(nb * (nb + 1) / 2, S);
    wide_gram64_kernel<double><<<grid, 128, 0, st>>>(x, ctr, L, DP, gram);
}

size_t wide_rot_cap(const Dims& d) { return (size_t)3 * d.D * d.D + 1024; }
int wide_iter_cap(const Dims& d) { return 8 * d.D + 64; }

// rows of Q one CTA of wide_rot_kernel keeps in shared memory
static int wide_rot_rows(int DP) {
    const long long avail = 227ll * 1024 - (long long)2 * DP * sizeof(double2);
    int rows = (int)(avail / ((long long)DP * sizeof(double)));
    if (rows > kWrThreads) rows = kWrThreads;
    return rows & ~7;
}

bool wide_eigen_fits(const Dims& d) { return wide_rot_rows(d.DP) >= 8 && (size_t)(4 + kWtThreads / 32) * d.DP * 8 <= 200u * 1024u; }

static void launch_wide_dgemm(const Dims& d, const double* A, const double* B, double* C, cudaStream_t st) {
    const size_t sm = (size_t)(128 * kRP + 16 * kWP) * sizeof(double);
    dim3 grid((d.DP + 63) / 64, (d.DP + 127) / 128, d.S);
    wide_dgemm_kernel<<<grid, kRowsThreads, sm, st>>>(A, B, d.D, d.DP, C);
}

void launch_wide_target(const Dims& d, const double* P, const double* lam, const double* slogT, const int* status,
                        const WideTarget& t, cudaStream_t st) {
    wide_target_kernel<<<d.S, 256, (size_t)d.DP * sizeof(double), st>>>(P, lam, slogT, status, d.D, d.DP, t.W, t.Wt,
                                                                        t.slogT, t.status);
}

// mode: 0 = diag(S) target (correlation matrix), 1 = no scaling (plain eigenvectors of the covariance),
//       2 = the full-column target held in `tgt` (-f)
void launch_wide_eigen(const Dims& d, const double* gram, const int* n, const double* mu, const double* ctr,
                       const int* qexp, int mode, double* work, double* dinv, double* dvec, double* evec,
                       double2* rot, int2* iters, int* niter, double* P, double* lam, double* slogT, int* status,
                       cudaStream_t st, const WideTarget* tgt) {
    const size_t rot_cap = wide_rot_cap(d);
    const int iter_cap = wide_iter_cap(d);
    wide_cov_kernel<<<d.S, 256, (size_t)3 * d.DP * sizeof(double), st>>>(gram, n, d.D, d.DP, mu, ctr, qexp,
                                                                         mode == 2 ? 1 : mode, work, dinv, slogT, status);
    if (mode == 2) {
        launch_wide_dgemm(d, work, tgt->W, tgt->tmp, st);            // S W
        launch_wide_dgemm(d, tgt->Wt, tgt->tmp, work, st);           // W^T (S W)
        wide_mirror_kernel<<<d.S, 256, 0, st>>>(work, d.DP);
    }
    int nthr = 0;                                                       // 0: the blocked kernel
    if (const char* e = cmf_hook("CMF_WIDE_TRED")) nthr = atoi(e);      // tuning hook (tools build): 1024 / 512 = unblocked
    if (nthr == 0) {
        const size_t tsm = (size_t)((2 * kTbNB + 5) * d.DP + 2 * kTbNB + 2 * kTbNB * kTbNB + 2 * kTbNB * 512) * sizeof(double);
        cudaFuncSetAttribute(wide_tredb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        wide_tredb_kernel<<<d.S, kTbThreads, tsm, st>>>(work, n, d.D, d.DP, dvec, evec);
    } else if (nthr == 512) {
        const size_t tsm = (size_t)(4 + 512 / 32) * d.DP * sizeof(double);
        cudaFuncSetAttribute(wide_tred_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        wide_tred_kernel<512><<<d.S, 512, tsm, st>>>(work, n, d.D, d.DP, dvec, evec);
    } else {
        const size_t tsm = (size_t)(4 + 1024 / 32) * d.DP * sizeof(double);
        cudaFuncSetAttribute(wide_tred_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsm);
        wide_tred_kernel<1024><<<d.S, 1024, tsm, st>>>(work, n, d.D, d.DP, dvec, evec);
    }
    wide_ql_kernel<<<d.S, 32, (size_t)2 * d.DP * sizeof(double), st>>>(dvec, evec, n, d.D, d.DP, rot, (long long)rot_cap,
                                                                       iters, iter_cap, niter, lam, status);
    const int rows = wide_rot_rows(d.DP);
    const size_t smem = (size_t)d.DP * rows * sizeof(double) + (size_t)2 * d.DP * sizeof(double2);
    cudaFuncSetAttribute(wide_rot_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid((d.DP + rows - 1) / rows, d.S);
    wide_rot_kernel<<<grid, kWrThreads, smem, st>>>(work, n, d.D, d.DP, rows, rot, (long long)rot_cap, iters, iter_cap,
                                                    niter, dinv, P);
    if (mode == 2) {
        launch_wide_dgemm(d, tgt->W, P, tgt->tmp, st);               // P = W V
        cudaMemcpyAsync(P, tgt->tmp, (size_t)d.S * d.DP * d.DP * sizeof(double), cudaMemcpyDeviceToDevice, st);
        wide_target_finish_kernel<<<(d.S + 255) / 256, 256, 0, st>>>(tgt->slogT, tgt->status, d.S, slogT, status);
    }
}

void launch_wide_tables(const Dims& d, int APW, const int* n, const int* nloo, const double* alphas, int model,
                        const double* lam, const double* slogT, double* logdet, double* beta, double* rsum, double* W,
                        cudaStream_t st) {
    wide_tables_kernel<<<d.S, 256, (size_t)d.DP * sizeof(double), st>>>(n, nloo, alphas, d.A, d.AP, APW, d.D, d.DP,
                                                                        model, lam, slogT, logdet, beta, rsum, W);
}

// exact LOO sums of columns [s0, s0 + ns): projection + squares into the batch buffer Z, then the alpha contraction
void launch_wide_loo(const Dims& d, int APW, const float* xt, const double* mu, const double* P, const double* W,
                     const double* beta, const int* n, int s0, int ns, int nchunk, double* Z, double* fpart,
                     cudaStream_t st) {
    const size_t sm1 = (size_t)(128 * kRP + 16 * kWP + d.DP) * sizeof(double);
    const size_t sm2 = (size_t)(128 * kRP + 16 * kWP + 256) * sizeof(double);
    dim3 g1((d.DP + 63) / 64, (d.L + 127) / 128, ns);
    wide_proj_kernel<float><<<g1, kRowsThreads, sm1, st>>>(xt, mu, P, n, d.L, d.D, d.DP, s0, Z);
    int lpc = (d.L + nchunk - 1) / nchunk;
    lpc = (lpc + 127) / 128 * 128;
    dim3 g2(APW / 64, nchunk, ns);
    wide_loo_kernel<<<g2, kRowsThreads, sm2, st>>>(Z, W, beta, n, d.L, d.D, d.DP, d.AP, APW, s0, lpc, nchunk, fpart);
}

void launch_wide_loo_f64(int L, int D, int DP, int AP, int APW, const double* x, const double* zero_mu,
                         const double* P, const double* W, const double* beta, const int* n, double* Z,
                         double* fpart, cudaStream_t st) {
    const size_t sm1 = (size_t)(128 * kRP + 16 * kWP + DP) * sizeof(double);
    const size_t sm2 = (size_t)(128 * kRP + 16 * kWP + 256) * sizeof(double);
    dim3 g1((DP + 63) / 64, (L + 127) / 128, 1);
    wide_proj_kernel<double><<<g1, kRowsThreads, sm1, st>>>(x, zero_mu, P, n, L, D, DP, 0, Z);
    const int lpc = (L + 127) / 128 * 128;
    dim3 g2(APW / 64, 1, 1);
    wide_loo_kernel<<<g2, kRowsThreads, sm2, st>>>(Z, W, beta, n, L, D, DP, AP, APW, 0, lpc, 1, fpart);
}


// ---------------------------------------------------------------------------------------- one-column helpers
// (cmf_looshrinkage: the importable looshrinkage(I_zm, alphas, nll, n, I_reg) of the reference, :92-136)
// column means of an FP64 sample matrix [rows][DP] (numpy.cov re-centres its input, :68): thread <-> band, rows in order
__global__ void __launch_bounds__(128) wide_mean64_kernel(const double* __restrict__ x, int rows, int D, int DP,
                                                          double* __restrict__ mean) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= DP) return;
    double a = 0.0;
    if (b < D) for (int r = 0; r < rows; ++r) a += x[(long long)r * DP + b];
    mean[b] = (b < D && rows > 0) ? a / (double)rows : 0.0;
}

// C = (1 - alpha) S + alpha T, S = cov = G / (m - 1), T = diag(diag(S))  (:130-134); alpha = 0 when mindex = -1
__global__ void __launch_bounds__(256) wide_cmat_kernel(const double* __restrict__ gram, int m, int D, int DP,
                                                        const int* __restrict__ mindex, const double* __restrict__ alphas,
                                                        double* __restrict__ C, const double* __restrict__ gram_reg,
                                                        int m_reg) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= D * D) return;
    const int r = idx / D, c = idx % D;
    const int mi = mindex[0];
    const double al = mi >= 0 ? alphas[mi] : 0.0;
    const double inv = 1.0 / (double)(m - 1);
    const double sv = gram[(long long)max(r, c) * DP + min(r, c)] * inv;
    // the target: diag(S), or cov(I_reg) when the caller gave one (:131-133)
    const double tv = gram_reg ? gram_reg[(long long)max(r, c) * DP + min(r, c)] / (double)(m_reg - 1)
                               : ((r == c) ? sv : 0.0);
    C[idx] = (1.0 - al) * sv + al * tv;
}

void launch_wide_mean64(const double* x, int rows, int D, int DP, double* mean, cudaStream_t st) {
    wide_mean64_kernel<<<(DP + 127) / 128, 128, 0, st>>>(x, rows, D, DP, mean);
}
void launch_wide_cmat(const double* gram, int m, int D, int DP, const int* mindex, const double* alphas, double* C,
                      cudaStream_t st, const double* gram_reg, int m_reg) {
    wide_cmat_kernel<<<(D * D + 255) / 256, 256, 0, st>>>(gram, m, D, DP, mindex, alphas, C, gram_reg, m_reg);
}

}  // namespace cmf
