// K1: per-column centred Gram matrix  G = sum_l (x_l - c)(x_l - c)^T  on the FP64 tensor path (c = column mean,
// or a pilot centre near it whose offset K2 removes exactly: sum (x-c)(x-c)^T - n (mu-c)(mu-c)^T).
//
// Replaces numpy.cov in the reference (cmf/robust_mf.py:52-70, called at :98 and :130).  Precision
// scheme: FP32 radiances converted to FP64, mean removed in FP64, products and accumulation in
// DMMA.8x8x4 (FP64) -- SURVEY.md 7.3(1): FP32-class accumulation (3xTF32) misses the 1e-3 sigma
// tolerance through the alpha search, FP64 accumulation holds it with ~7 orders of margin.
//
// Mapping: one CTA = one cross-track column x one chunk of lines; every warp is an independent
// pipeline: lane 0 streams 16-line tiles of the column-major copy (contiguous 16*DP*4 bytes) into the
// warp's private shared-memory ring with 1-D bulk async copies (TMA engine) tracked by mbarriers, the
// warp turns each 4-line slice into NT fragments (one LDS + cvt + FP64 subtract each; the same register
// serves as A and B operand because G is X^T X) and issues the NT(NT+1)/2 lower-triangle 8x8 tiles.
// Warps are summed through shared memory in a fixed order at the end -> deterministic results.
#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

constexpr int kGramNS = 3;      // ring stages per warp
constexpr int kGramWarps = 8;

template <int NT, int R0, int R1>
struct GramTiles {
    // tiles (i, j), R0 <= i < R1, j <= i
    static constexpr int count = R1 * (R1 + 1) / 2 - R0 * (R0 + 1) / 2;
};

template <int NT, int R0, int R1>
__device__ __forceinline__ void gram_warp(const float* __restrict__ col_base, const double* mu_s,
                                          float* myring, uint64_t* mybars, int c_begin, int c_end,
                                          int first_tile, int tile_stride, double* red, int order,
                                          int lane) {
    constexpr int DP = 8 * NT, TL = kGramTL, NS = kGramNS;
    constexpr int NTILE = GramTiles<NT, R0, R1>::count;
    const int ntiles = (c_end - c_begin + TL - 1) / TL;
    const int g = lane >> 2, q = lane & 3;

    double acc[NTILE][2];
#pragma unroll
    for (int t = 0; t < NTILE; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; }
    double mu_r[R1];
#pragma unroll
    for (int i = 0; i < R1; ++i) mu_r[i] = mu_s[8 * i + g];

    auto issue = [&](int it) {
        const int t = first_tile + tile_stride * it;
        if (t < ntiles) {
            const int l0 = c_begin + t * TL;
            const int nl = min(TL, c_end - l0);
            const uint32_t bytes = (uint32_t)(nl * DP * sizeof(float));
            uint64_t* bar = mybars + (it % NS);
            mbar_expect_tx(bar, bytes);
            bulk_g2s(myring + (it % NS) * TL * DP, col_base + (long long)l0 * DP, bytes, bar);
        }
    };
    if (lane == 0) {
#pragma unroll
        for (int it = 0; it < NS; ++it) issue(it);
    }
    for (int it = 0; first_tile + tile_stride * it < ntiles; ++it) {
        const int t = first_tile + tile_stride * it;
        const int nl = min(TL, c_end - (c_begin + t * TL));
        mbar_wait(mybars + (it % NS), (uint32_t)((it / NS) & 1));
        const float* tile = myring + (it % NS) * TL * DP;
#pragma unroll
        for (int kk = 0; kk < TL / 4; ++kk) {
            const int row = kk * 4 + q;
            const bool rowok = row < nl;
            double f[R1];
#pragma unroll
            for (int i = 0; i < R1; ++i) {
                const float x = tile[row * DP + 8 * i + g];
                f[i] = (rowok && x == x) ? (double)x - mu_r[i] : 0.0;
            }
            int tt = 0;
#pragma unroll
            for (int i = R0; i < R1; ++i) {
#pragma unroll
                for (int j = 0; j <= i; ++j) {
                    mma884(acc[tt][0], acc[tt][1], f[i], f[j]);
                    ++tt;
                }
            }
        }
        __syncwarp();
        if (lane == 0) issue(it + NS);
    }
    // fixed-order accumulation of this warp's tiles into the CTA result
    constexpr int T0 = R0 * (R0 + 1) / 2;
    for (int w = 0; w < kGramWarps; ++w) {
        __syncthreads();
        if (w == order) {
#pragma unroll
            for (int t = 0; t < NTILE; ++t) {
                double2* p = reinterpret_cast<double2*>(red + ((T0 + t) * 32 + lane) * 2);
                double2 v = *p;
                v.x += acc[t][0];
                v.y += acc[t][1];
                *p = v;
            }
        }
    }
}

template <int NT>
__global__ void __launch_bounds__(kGramWarps * 32, 1)
    gram_kernel(const float* __restrict__ xt, const double* __restrict__ mu_g, int L, int lines_per_chunk,
                int chunk_lo, int nchunk, double* __restrict__ gram_part, const int* __restrict__ nrows) {
    constexpr int DP = 8 * NT, TL = kGramTL, NS = kGramNS, NTRI = NT * (NT + 1) / 2;
    constexpr int NROLE = (NT > 9) ? 2 : 1;
    constexpr int RSPLIT = (NT > 9) ? 8 : NT;   // role 0: tile rows [0,RSPLIT), role 1: [RSPLIT,NT)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* ring = reinterpret_cast<float*>(smem_raw);                     // [warps][NS][TL*DP]
    double* red = reinterpret_cast<double*>(ring + kGramWarps * NS * TL * DP);   // [NTRI*64]
    double* mu_s = red + NTRI * 64;                                       // [DP]
    uint64_t* bars = reinterpret_cast<uint64_t*>(mu_s + DP);              // [warps][NS]

    const int s = blockIdx.x, chunk = chunk_lo + blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c_begin = chunk * lines_per_chunk;
    // compacted background-mode pass: only the first nrows[s] rows of the column hold pixels
    const int c_end = max(c_begin, min(nrows ? min(L, nrows[s]) : L, c_begin + lines_per_chunk));

    for (int i = threadIdx.x; i < DP; i += blockDim.x) mu_s[i] = mu_g[(long long)s * DP + i];
    for (int i = threadIdx.x; i < NTRI * 64; i += blockDim.x) red[i] = 0.0;
    if (lane == 0) {
        for (int st = 0; st < NS; ++st) mbar_init(&bars[warp * NS + st], 1);
        fence_mbar_init();
    }
    __syncthreads();

    const float* col_base = xt + (long long)s * L * DP;
    float* myring = ring + warp * NS * TL * DP;
    uint64_t* mybars = bars + warp * NS;
    const int role = warp % NROLE;
    const int first_tile = warp / NROLE;
    const int stride = kGramWarps / NROLE;
    if constexpr (NROLE == 1) {
        gram_warp<NT, 0, NT>(col_base, mu_s, myring, mybars, c_begin, c_end, first_tile, stride, red, warp,
                             lane);
    } else {
        if (role == 0)
            gram_warp<NT, 0, RSPLIT>(col_base, mu_s, myring, mybars, c_begin, c_end, first_tile, stride, red,
                                     warp, lane);
        else
            gram_warp<NT, RSPLIT, NT>(col_base, mu_s, myring, mybars, c_begin, c_end, first_tile, stride, red,
                                      warp, lane);
    }
    __syncthreads();
    double* out = gram_part + ((long long)s * nchunk + chunk) * NTRI * 64;
    for (int i = threadIdx.x; i < NTRI * 64; i += blockDim.x) out[i] = red[i];
}

size_t gram_part_elems(const Dims& d, int nchunk) { return (size_t)d.S * nchunk * ntri(d.NT) * 64; }

template <int NT>
static void launch_gram_t(const Dims& d, const float* xt, const double* mu, int nchunk, int lpc, int chunk_lo,
                          int chunk_hi, double* gram_part, cudaStream_t st) {
    constexpr int DP = 8 * NT, NTRI = NT * (NT + 1) / 2;
    const size_t smem = (size_t)kGramWarps * kGramNS * kGramTL * DP * sizeof(float) +
                        (size_t)(NTRI * 64 + DP) * sizeof(double) + kGramWarps * kGramNS * sizeof(uint64_t);
    cudaFuncSetAttribute(gram_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (chunk_hi <= chunk_lo) return;
    dim3 grid(d.S, chunk_hi - chunk_lo);
    gram_kernel<NT><<<grid, kGramWarps * 32, smem, st>>>(xt, mu, d.L, lpc, chunk_lo, nchunk, gram_part, d.nrows);
}

// Chunks [chunk_lo, chunk_hi) of `nchunk` chunks of `lpc` lines each (lpc a multiple of the Gram tile);
// `ctr` is the point the products are centred on (the column mean, or the pilot centre of the chased pass).
void launch_gram(const Dims& d, const float* xt, const double* ctr, int nchunk, int lpc, int chunk_lo,
                 int chunk_hi, double* gram_part, cudaStream_t st) {
    const double* mu = ctr;
    switch (d.NT) {
#define CMF_CASE(n) case n: launch_gram_t<n>(d, xt, mu, nchunk, lpc, chunk_lo, chunk_hi, gram_part, st); break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6)
        CMF_CASE(7) CMF_CASE(8) CMF_CASE(9) CMF_CASE(10) CMF_CASE(11) CMF_CASE(12)
#undef CMF_CASE
        default: break;
    }
}

}  // namespace cmf
