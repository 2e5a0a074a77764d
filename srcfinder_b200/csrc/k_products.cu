// Kernels for the two steps either side of the matched filter, sm_100a:
//
//   pixel_flags_kernel       per-pixel spectrometer flags of a radiance cube: saturated / specular / dark / cloud
//                            (spectrometer_masks/masks_sds.py:133-233, the per-pixel tests only; region growing,
//                            buffers and morphology are image operations and stay outside)
//   profile_* kernels        column profiles of a score image (triage/cmf_profile.py:110-140): per cross-track
//                            column npix/avg/std/min/max, or robust npix/med/mad/p05/p95, in float32 exactly as
//                            numpy evaluates them (sequential float32 sums down the lines, nearest-rank
//                            percentiles, mean of the two middle values)
//
// Both are HBM/L2-bound integer-and-compare work; nothing here touches tensor cores.
#include <stdint.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

// ---------------------------------------------------------------------------------------- flags
// thread <-> (line, sample); every band row is read as a coalesced run of samples.
__global__ void __launch_bounds__(256) pixel_flags_kernel(const float* __restrict__ cube, long long line_pitch,
                                                          int band_pitch, int L, int S, FlagSpec f,
                                                          uint8_t* __restrict__ flags) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y;
    if (s >= S || l >= L) return;
    const float* p = cube + (long long)l * line_pitch + s;
    // saturated: ANY value above the threshold inside the window (masks_sds.py:149); NaN compares false
    bool sat = false;
    for (int b = f.sat_lo; b <= f.sat_hi; ++b) sat |= ldg_nc_f1(p + (long long)b * band_pitch) > f.sat_thresh;
    uint8_t out = sat ? kFlagSaturated : 0;
    // specular: saturated and bright in the visible band (:152-163)
    if (f.spec_band >= 0 && sat && ldg_nc_f1(p + (long long)f.spec_band * band_pitch) > f.spec_thresh)
        out |= kFlagSpecular;
    // dark: low radiance at 2139 nm that is not the no-data value (:165-180)
    if (f.dark_band >= 0) {
        const float v = ldg_nc_f1(p + (long long)f.dark_band * band_pitch);
        if (v < f.dark_thresh && !(v <= -9999.0f)) out |= kFlagDark;
    }
    // cloud: bright at band a and a negative a->b slope.  The reference's third argument of np.logical_and
    // (the b->c slope) is numpy's `out` parameter, so it does not take part in the result (:228).
    if (f.cloud_a >= 0 && f.cloud_b >= 0) {
        const float ra = ldg_nc_f1(p + (long long)f.cloud_a * band_pitch);
        const float rb = ldg_nc_f1(p + (long long)f.cloud_b * band_pitch);
        const float diff = rb - ra;                         // np.diff of the float32 pair (:213)
        const bool neg_slope = f.cloud_dwl > 0.0f ? diff < 0.0f : (f.cloud_dwl < 0.0f ? diff > 0.0f : false);
        if (ra > f.cloud_thresh && neg_slope) out |= kFlagCloud;
    }
    flags[(long long)l * S + s] = out;
}

void launch_pixel_flags(const float* cube, long long line_pitch, int band_pitch, int L, int S, const FlagSpec& f,
                        uint8_t* flags, cudaStream_t st) {
    dim3 grid((S + 255) / 256, L);
    pixel_flags_kernel<<<grid, 256, 0, st>>>(cube, line_pitch, band_pitch, L, S, f, flags);
}

// ---------------------------------------------------------------------------------------- exclusion helpers
__global__ void invert_u8_kernel(uint8_t* p, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = p[i] ? 0 : 1;
}

void launch_invert_u8(uint8_t* p, long long n, cudaStream_t st) {
    invert_u8_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n);
}

// valid pixels per column (nuse, cmf/robust_mf.py:302) from the mask image; one CTA per column, fixed order
__global__ void __launch_bounds__(256) count_mask_kernel(const uint8_t* __restrict__ mask, int L, int S,
                                                         int* __restrict__ count) {
    __shared__ int red[8];
    const int s = blockIdx.x;
    int c = 0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) c += mask[(long long)l * S + s] ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < 8; ++i) t += red[i];
        count[s] = t;
    }
}

void launch_count_mask(const Dims& d, const uint8_t* mask, int* count, cudaStream_t st) {
    count_mask_kernel<<<d.S, 256, 0, st>>>(mask, d.L, d.S, count);
}

// ---------------------------------------------------------------------------------------- profiles
// K-P0: score image f64 [L][S] -> float32 [S][L] (the reference converts to float32 first, cmf_profile.py:112),
// NaN where the pixel is no-data, NaN or not positive (:113-114, :122).  32x32 transposing tiles.
__global__ void __launch_bounds__(256) profile_gather_kernel(const double* __restrict__ mf, int L, int S,
                                                             float nodata32, float* __restrict__ colv) {
    __shared__ float t[32][33];
    const int l0 = blockIdx.y * 32, s0 = blockIdx.x * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int l = l0 + r, s = s0 + tx;
        float v = __int_as_float(0x7fc00000);
        if (l < L && s < S) {
            const float x = (float)mf[(long long)l * S + s];
            if (!(x == nodata32) && x == x && x > 0.0f) v = x;
        }
        t[r][tx] = v;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int s = s0 + r, l = l0 + tx;
        if (l < L && s < S) colv[(long long)s * L + l] = t[tx][r];
    }
}

__device__ __forceinline__ float block_reduce_minmax(float v, bool is_max, float* red) {
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : fminf(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
    for (int i = 1; i < nw; ++i) r = is_max ? fmaxf(r, red[i]) : fminf(r, red[i]);
    return r;
}

__device__ __forceinline__ int block_reduce_sum_int(int v, int* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    int r = 0;
    for (int i = 0; i < nw; ++i) r += red[i];
    return r;
}

// K-P1 (plain): one CTA per column.  numpy reduces axis 0 of a C-ordered (L, S) float32 array by adding whole
// rows into the accumulator, i.e. each column is a SEQUENTIAL float32 sum down the lines (no pairwise blocks);
// the masked pixels are zeros in that sum (np.nanmean / np.nanstd, :127-128).  One thread replays exactly that;
// the rest of the block does the count and the extrema.
__global__ void __launch_bounds__(256) profile_plain_kernel(const float* __restrict__ colv, int L, int S,
                                                            double* __restrict__ out) {
    __shared__ float redf[8];
    __shared__ int redi[8];
    __shared__ float s_avg;
    const int s = blockIdx.x;
    const float* x = colv + (long long)s * L;
    int cnt = 0;
    float vmin = __int_as_float(0x7f800000), vmax = -__int_as_float(0x7f800000);
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        const float v = x[l];
        if (v == v) { ++cnt; vmin = fminf(vmin, v); vmax = fmaxf(vmax, v); }
    }
    const int n = block_reduce_sum_int(cnt, redi);
    vmin = block_reduce_minmax(vmin, false, redf);
    vmax = block_reduce_minmax(vmax, true, redf);
    const float qnan = __int_as_float(0x7fc00000);
    if (threadIdx.x == 0) {
        float tot = 0.0f;
#pragma unroll 8
        for (int l = 0; l < L; ++l) {
            const float v = x[l];
            tot = __fadd_rn(tot, v == v ? v : 0.0f);
        }
        // np.true_divide(float32 total, int count) runs in double and is stored back as float32
        s_avg = n > 0 ? (float)((double)tot / (double)n) : qnan;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const float avg = s_avg;
        float sq = 0.0f;
#pragma unroll 8
        for (int l = 0; l < L; ++l) {
            const float v = x[l];
            const float dlt = v == v ? __fsub_rn(v, avg) : 0.0f;     // masked entries stay 0 (_nanvar)
            sq = __fadd_rn(sq, __fmul_rn(dlt, dlt));
        }
        const float var = n > 0 ? (float)((double)sq / (double)n) : qnan;
        out[0 * S + s] = (double)n;
        out[1 * S + s] = (double)avg;
        out[2 * S + s] = (double)__fsqrt_rn(var);
        out[3 * S + s] = n > 0 ? (double)vmin : (double)qnan;
        out[4 * S + s] = n > 0 ? (double)vmax : (double)qnan;
    }
}

// k-th smallest (0-based) of the non-NaN keys of a column by 4 rounds of 8-bit radix selection.  Keys are
// non-negative floats, whose order is that of their bit patterns.  MAD = true selects on |x - centre| (float32).
template <bool MAD>
__device__ float radix_select(const float* __restrict__ x, int L, int k, float centre, int* hist, int* pick) {
    uint32_t prefix = 0, known = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int l = threadIdx.x; l < L; l += blockDim.x) {
            const float v = x[l];
            if (v == v) {
                const uint32_t key = __float_as_uint(MAD ? fabsf(__fsub_rn(v, centre)) : v);
                if ((key & known) == prefix) atomicAdd(&hist[(key >> shift) & 255], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            int acc = 0, b = 0;
            for (; b < 256; ++b) {
                if (acc + hist[b] > k) break;
                acc += hist[b];
            }
            pick[0] = b;
            pick[1] = k - acc;
        }
        __syncthreads();
        prefix |= (uint32_t)pick[0] << shift;
        known |= 255u << shift;
        k = pick[1];
        __syncthreads();
    }
    return __uint_as_float(prefix);
}

// K-P1 (robust): median, MAD about the median, nearest-rank percentiles (np.nanmedian / extrema(p=0.95), :123-125).
__global__ void __launch_bounds__(256) profile_robust_kernel(const float* __restrict__ colv, int L, int S,
                                                             double qlo, double qhi, double* __restrict__ out) {
    __shared__ int hist[256];
    __shared__ int pick[2];
    __shared__ int redi[8];
    const int s = blockIdx.x;
    const float* x = colv + (long long)s * L;
    int cnt = 0;
    for (int l = threadIdx.x; l < L; l += blockDim.x) cnt += (x[l] == x[l]) ? 1 : 0;
    const int n = block_reduce_sum_int(cnt, redi);
    const float qnan = __int_as_float(0x7fc00000);
    if (n == 0) {
        if (threadIdx.x == 0) {
            out[0 * S + s] = 0.0;
            for (int r = 1; r < 5; ++r) out[r * S + s] = (double)qnan;
        }
        return;
    }
    // np.median: mean of the two middle order statistics for even n, as float32
    const float m_lo = radix_select<false>(x, L, (n - 1) / 2, 0.0f, hist, pick);
    const float m_hi = (n & 1) ? m_lo : radix_select<false>(x, L, n / 2, 0.0f, hist, pick);
    const float med = (n & 1) ? m_lo : __fmul_rn(__fadd_rn(m_lo, m_hi), 0.5f);
    const float d_lo = radix_select<true>(x, L, (n - 1) / 2, med, hist, pick);
    const float d_hi = (n & 1) ? d_lo : radix_select<true>(x, L, n / 2, med, hist, pick);
    const float mad = (n & 1) ? d_lo : __fmul_rn(__fadd_rn(d_lo, d_hi), 0.5f);
    // method 'nearest': index = around((n - 1) * q), half to even
    const int i_lo = (int)rint((double)(n - 1) * qlo), i_hi = (int)rint((double)(n - 1) * qhi);
    const float p_lo = radix_select<false>(x, L, min(max(i_lo, 0), n - 1), 0.0f, hist, pick);
    const float p_hi = radix_select<false>(x, L, min(max(i_hi, 0), n - 1), 0.0f, hist, pick);
    if (threadIdx.x == 0) {
        out[0 * S + s] = (double)n;
        out[1 * S + s] = (double)med;
        out[2 * S + s] = (double)mad;
        out[3 * S + s] = (double)p_lo;
        out[4 * S + s] = (double)p_hi;
    }
}

void launch_column_profile(const double* mf, int L, int S, double nodata, int robust, double qlo, double qhi,
                           float* colv, double* out, cudaStream_t st) {
    dim3 grid((S + 31) / 32, (L + 31) / 32);
    profile_gather_kernel<<<grid, 256, 0, st>>>(mf, L, S, (float)nodata, colv);
    if (robust) profile_robust_kernel<<<S, 256, 0, st>>>(colv, L, S, qlo, qhi, out);
    else profile_plain_kernel<<<S, 256, 0, st>>>(colv, L, S, out);
}


// ---------------------------------------------------------------------------------------- detection pre-filter
// The per-pixel head of filtdet (srcfinder_util.py:1428-1436) with kde (:1383-1387):
//     imgkde = gaussian_filter(mf, sigma = k, truncate = 1)        separable, radius int(k + 0.5), 'reflect' borders,
//                                                                   axis 0 (lines) first, then axis 1 (samples)
//     imgkde = (imgkde - min) / (max - min);  detkde = mf * imgkde
//     detkde = clip((detkde - mfmin) / (mfmax - mfmin), 0, 1);  ch4min = mf >= mfmin;  detmask = detkde > 0
// Same FP64 evaluation order as scipy's correlate1d for a symmetric filter (centre tap first, then the pairs from
// the outermost inwards, no fused multiply-add), so the result equals the reference's to the last bit when the
// C library is compiled without contraction.  The connected-component steps that follow (:1437-1470) are image
// morphology and not part of this call.
__device__ __forceinline__ int reflect_index(int i, int n) {
    // scipy 'reflect' (d c b a | a b c d | d c b a): period 2n
    const int p = 2 * n;
    int m = i % p;
    if (m < 0) m += p;
    return m < n ? m : p - 1 - m;
}

__global__ void __launch_bounds__(256)
    kde_blur_kernel(const double* __restrict__ in, int L, int S, int axis, int radius, const double* __restrict__ w,
                    double* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (long long)L * S) return;
    const int l = (int)(idx / S), s = (int)(idx % S);
    const int n = axis == 0 ? L : S, pos = axis == 0 ? l : s;
    const long long stride = axis == 0 ? S : 1;
    const double* line = in + (axis == 0 ? (long long)s : (long long)l * S);
    double acc = __dmul_rn(line[(long long)pos * stride], w[radius]);
    const bool interior = pos - radius >= 0 && pos + radius < n;
    for (int jj = -radius; jj < 0; ++jj) {
        const int a = interior ? pos + jj : reflect_index(pos + jj, n);
        const int b = interior ? pos - jj : reflect_index(pos - jj, n);
        const double pair = __dadd_rn(line[(long long)a * stride], line[(long long)b * stride]);
        acc = __dadd_rn(acc, __dmul_rn(pair, w[radius + jj]));
    }
    out[idx] = acc;
}

// block partials of min / max with numpy's NaN propagation (any NaN -> NaN)
__global__ void __launch_bounds__(256)
    kde_minmax_kernel(const double* __restrict__ v, long long n, double* __restrict__ part) {
    __shared__ double smin[8], smax[8];
    __shared__ int snan[8];
    double lo = __longlong_as_double(0x7ff0000000000000LL), hi = -lo;
    int nan = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = v[i];
        if (x != x) nan = 1;
        lo = fmin(lo, x); hi = fmax(hi, x);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        nan |= __shfl_xor_sync(0xffffffffu, nan, o);
    }
    if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; snan[threadIdx.x >> 5] = nan; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { lo = fmin(lo, smin[i]); hi = fmax(hi, smax[i]); nan |= snan[i]; }
        lo = fmin(lo, smin[0]); hi = fmax(hi, smax[0]); nan |= snan[0];
        part[3 * blockIdx.x] = lo; part[3 * blockIdx.x + 1] = hi; part[3 * blockIdx.x + 2] = nan ? 1.0 : 0.0;
    }
}

__global__ void __launch_bounds__(256)
    kde_finish_kernel(const double* __restrict__ mf, const double* __restrict__ blur, long long n,
                      const double* __restrict__ part, int nparts, double mfmin, double mfmax,
                      double* __restrict__ detkde, uint8_t* __restrict__ ch4min, uint8_t* __restrict__ detmask) {
    double lo = part[0], hi = part[1];
    bool nan = part[2] != 0.0;
    for (int i = 1; i < nparts; ++i) { lo = fmin(lo, part[3 * i]); hi = fmax(hi, part[3 * i + 1]); nan = nan || part[3 * i + 2] != 0.0; }
    if (nan) { lo = __longlong_as_double(0x7ff8000000000000LL); hi = lo; }
    const double range = __dsub_rn(hi, lo), den = __dsub_rn(mfmax, mfmin);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const double x = mf[i];
        const double k01 = __ddiv_rn(__dsub_rn(blur[i], lo), range);
        double d = __ddiv_rn(__dsub_rn(__dmul_rn(x, k01), mfmin), den);
        d = (d != d) ? d : fmin(fmax(d, 0.0), 1.0);          // np.clip keeps NaN
        detkde[i] = d;
        if (ch4min) ch4min[i] = x >= mfmin ? 1 : 0;
        if (detmask) detmask[i] = d > 0.0 ? 1 : 0;
    }
}

void launch_detection_prefilter(const double* mf, int L, int S, int radius, const double* w_dev, double mfmin,
                                double mfmax, double* tmp, double* blur, double* part, double* detkde,
                                uint8_t* ch4min, uint8_t* detmask, cudaStream_t st) {
    const long long n = (long long)L * S;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    kde_blur_kernel<<<blocks, 256, 0, st>>>(mf, L, S, 0, radius, w_dev, tmp);
    kde_blur_kernel<<<blocks, 256, 0, st>>>(tmp, L, S, 1, radius, w_dev, blur);
    const int nparts = 296;
    kde_minmax_kernel<<<nparts, 256, 0, st>>>(blur, n, part);
    kde_finish_kernel<<<1184, 256, 0, st>>>(mf, blur, n, part, nparts, mfmin, mfmax, detkde, ch4min, detmask);
}

// ---------------------------------------------------------------------------------------- CNN input
// cnn/cnn_pred_pipeline.py:19-30, 126-157: ClampCH4(vmin, vmax) then transforms.Normalize(mean, std) on the float32
// score image (torch.clamp, then (x - mean) / std in float32; no-data pixels are clamped like any other value).
__global__ void __launch_bounds__(256)
    cnn_input_kernel(const double* __restrict__ mf64, const float* __restrict__ mf32, long long n, float vmin,
                     float vmax, float mean, float stdv, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float x = mf32 ? mf32[i] : (float)mf64[i];
        float c = x;
        if (x == x) c = fminf(fmaxf(x, vmin), vmax);          // torch.clamp propagates NaN
        out[i] = __fdiv_rn(__fsub_rn(c, mean), stdv);
    }
}

void launch_cnn_input(const double* mf64, const float* mf32, long long n, float vmin, float vmax, float mean,
                      float stdv, float* out, cudaStream_t st) {
    cnn_input_kernel<<<1184, 256, 0, st>>>(mf64, mf32, n, vmin, vmax, mean, stdv, out);
}

}  // namespace cmf
