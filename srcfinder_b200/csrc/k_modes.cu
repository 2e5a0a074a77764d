// Background modes (-k > 1 / -r): per-column mode lists from per-pixel labels, the per-pass member
// masks, and the final column statistics over the inlier pixels.
//
// Reference (cmf/robust_mf.py:306-344, 388-391): labels come from a k-means on the column's leading
// principal components; clusters with fewer than bgminsamp samples are relabelled -l when -r is given
// (label 0 can never flip: -0 == 0, :323-324); if every cluster was flagged the flips are undone
// (:330-332).  The modes are then fitted in the order of the (partly negated) unique-label list: an entry
// ki >= 0 fits and scores the pixels of that cluster, an entry ki < 0 fits and scores ALL inlier pixels
// (labels >= 0) with the pooled model (:341), later entries overwriting earlier ones (:386).
// Here the labels are an input (the reference's MiniBatchKMeans is unseeded, :312); everything that
// follows from them is computed on the device in this file, one pass of the pipeline per list entry.
#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

// One CTA per column: histogram of the labels of the valid pixels, rejection flags, entry list.
//   entries[s][t]  t-th entry of the mode list (kModeNone when the list is shorter)
//   rejmask[s]     bit l set = cluster l is rejected (its pixels are outliers)
__global__ void __launch_bounds__(256)
    modes_kernel(const int32_t* __restrict__ labels, const uint8_t* __restrict__ mask, int L, int S,
                 int reject_min, int8_t* __restrict__ entries, uint32_t* __restrict__ rejmask,
                 int* __restrict__ nentries, uint32_t* __restrict__ flagmask) {
    __shared__ int cnt[kMaxLabels];
    const int s = blockIdx.x, tid = threadIdx.x;
    if (tid < kMaxLabels) cnt[tid] = 0;
    __syncthreads();
    int local[4] = {0, 0, 0, 0};          // most columns use a handful of labels: count 0..3 in registers
    for (int l = tid; l < L; l += blockDim.x) {
        const long long o = (long long)l * S + s;
        if (mask[o]) {
            const int lab = labels[o];
            if (lab >= 0 && lab < 4) ++local[lab];
            else if (lab >= 4 && lab < kMaxLabels) atomicAdd(&cnt[lab], 1);
        }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (local[k]) atomicAdd(&cnt[k], local[k]);
    __syncthreads();
    if (tid == 0) {
        uint32_t rej = 0u;
        int ne = 0, nneg = 0;
        int8_t list[kMaxLabels];
        for (int l = 0; l < kMaxLabels; ++l) {
            if (cnt[l] == 0) continue;
            const bool flag = reject_min > 0 && cnt[l] < reject_min && l != 0;   // -0 == 0 never flips
            if (flag) { rej |= 1u << l; ++nneg; }
            list[ne++] = (int8_t)(flag ? -l : l);
        }
        // _bgmeta band 0 is written inside the counting loop (:326-327), i.e. with the negated ids even when the
        // rejection is undone afterwards: keep the as-flagged set for the cluster image
        flagmask[s] = rej;
        if (ne > 0 && nneg == ne) {            // all clusters rejected: proceed without rejection (:330-332)
            rej = 0u;
            for (int t = 0; t < ne; ++t) list[t] = (int8_t)(-list[t]);
        }
        for (int t = 0; t < kMaxLabels; ++t) entries[(long long)s * kMaxLabels + t] = (t < ne) ? list[t] : kModeNone;
        rejmask[s] = rej;
        nentries[s] = ne;
    }
}

// Member mask of pass t, cluster-id image (first pass only) and inlier mask.
__global__ void __launch_bounds__(256)
    members_kernel(const int32_t* __restrict__ labels, const uint8_t* __restrict__ mask, long long LS, int S,
                   int t, const int8_t* __restrict__ entries, const uint32_t* __restrict__ rejmask,
                   const uint32_t* __restrict__ flagmask, uint8_t* __restrict__ sel, int16_t* __restrict__ cluster_img,
                   uint8_t* __restrict__ inlier) {
    const long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (o >= LS) return;
    const int s = (int)(o % S);
    const bool valid = mask[o] != 0;
    const int lab = labels[o];
    const bool known = valid && lab >= 0 && lab < kMaxLabels;
    const uint32_t rej = rejmask[s];
    const bool rejected = known && ((rej >> lab) & 1u);
    const int e = entries[(long long)s * kMaxLabels + t];
    bool member = false;
    if (known && e != kModeNone) member = (e >= 0) ? (lab == e && !rejected) : !rejected;
    sel[o] = member ? 1 : 0;
    if (cluster_img != nullptr) {
        const bool flagged = known && ((flagmask[s] >> lab) & 1u);
        cluster_img[o] = known ? (int16_t)(flagged ? -lab : lab) : (int16_t)0;
        inlier[o] = (known && !rejected) ? 1 : 0;
    }
}

__global__ void __launch_bounds__(256) fill_f64_kernel(double* __restrict__ p, long long n, double v) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// colnum = nuse, colavg / colstd (ddof = 0) of the final scores of the inlier pixels (:388-391).
// One CTA per column, fixed reduction order.
__global__ void __launch_bounds__(256)
    colstats_modes_kernel(const double* __restrict__ mf, const uint8_t* __restrict__ inlier,
                          const int* __restrict__ nuse, int L, int S, double nodata,
                          double* __restrict__ colstats) {
    __shared__ double ssum[256], ssq[256];
    __shared__ int scnt[256];
    const int s = blockIdx.x, tid = threadIdx.x;
    double sum = 0.0, sq = 0.0;
    int cnt = 0;
    for (int l = tid; l < L; l += blockDim.x) {
        const long long o = (long long)l * S + s;
        if (inlier[o]) { const double v = mf[o]; sum += v; sq += v * v; ++cnt; }
    }
    ssum[tid] = sum; ssq[tid] = sq; scnt[tid] = cnt;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (tid < w) { ssum[tid] += ssum[tid + w]; ssq[tid] += ssq[tid + w]; scnt[tid] += scnt[tid + w]; }
        __syncthreads();
    }
    if (tid == 0) {
        const int n = nuse[s];
        if (n == 0) {
            colstats[s] = nodata; colstats[S + s] = nodata; colstats[2 * S + s] = nodata;
        } else {
            // np.mean / np.std of an empty selection are NaN; otherwise population statistics
            const double c = (double)scnt[0];
            const double mean = ssum[0] / c;
            double var = ssq[0] / c - mean * mean;
            if (var < 0.0) var = 0.0;
            colstats[s] = (double)n;
            colstats[S + s] = mean;
            colstats[2 * S + s] = sqrt(var);
        }
    }
}

void launch_modes(const Dims& d, const int32_t* labels, const uint8_t* mask, int reject_min, int8_t* entries,
                  uint32_t* rejmask, int* nentries, uint32_t* flagmask, cudaStream_t st) {
    modes_kernel<<<d.S, 256, 0, st>>>(labels, mask, d.L, d.S, reject_min, entries, rejmask, nentries, flagmask);
}

void launch_members(const Dims& d, const int32_t* labels, const uint8_t* mask, int t, const int8_t* entries,
                    const uint32_t* rejmask, const uint32_t* flagmask, uint8_t* sel, int16_t* cluster_img, uint8_t* inlier,
                    cudaStream_t st) {
    const long long LS = (long long)d.L * d.S;
    members_kernel<<<(unsigned)((LS + 255) / 256), 256, 0, st>>>(labels, mask, LS, d.S, t, entries, rejmask, flagmask, sel,
                                                                 cluster_img, inlier);
}

void launch_fill_f64(double* p, long long n, double v, cudaStream_t st) {
    fill_f64_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, n, v);
}

void launch_colstats_modes(const Dims& d, const double* mf, const uint8_t* inlier, const int* nuse,
                           double nodata, double* colstats, cudaStream_t st) {
    colstats_modes_kernel<<<d.S, 256, 0, st>>>(mf, inlier, nuse, d.L, d.S, nodata, colstats);
}


// One warp per column: exclusive count of the selected pixels above every line (ballot + popc, 32 lines at a time).
__global__ void __launch_bounds__(128)
    rank_kernel(const uint8_t* __restrict__ sel, int L, int S, int32_t* __restrict__ rowidx) {
    const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (s >= S) return;
    int base = 0;
    for (int l0 = 0; l0 < L; l0 += 32) {
        const int l = l0 + lane;
        const bool on = l < L && sel[(long long)l * S + s] != 0;
        const unsigned b = __ballot_sync(0xffffffffu, on);
        if (l < L) rowidx[(long long)l * S + s] = base + __popc(b & ((1u << lane) - 1u));
        base += __popc(b);
    }
}

void launch_rank(const Dims& d, const uint8_t* sel, int32_t* rowidx, cudaStream_t st) {
    rank_kernel<<<(d.S + 3) / 4, 128, 0, st>>>(sel, d.L, d.S, rowidx);
}


// Member rows of the full column-major copy -> compacted copy of a background-mode pass, with the column sums and
// counts of the members (the K0 partials of the pass).  CTA = (column, line split); thread <-> (row lane, 4 bands);
// every thread adds its rows in line order and the row lanes are added in a fixed order -> deterministic.
constexpr int kCompactRows = 16;

__global__ void __launch_bounds__(512)
    compact_kernel(const float* __restrict__ xt_full, const uint8_t* __restrict__ sel,
                   const int32_t* __restrict__ rowidx, int L, int S, int DP, int lines_per_split,
                   float* __restrict__ xt_mode, double* __restrict__ colsum_part, int* __restrict__ colcnt_part) {
    extern __shared__ double csm[];                       // [kCompactRows][DP] partial sums
    __shared__ int cnt_sh[kCompactRows];
    const int s = blockIdx.x, split = blockIdx.y;
    const int Q = DP / 4;                                 // float4 lanes per row
    const int rr = threadIdx.x / Q, q = threadIdx.x % Q;
    const int l_begin = split * lines_per_split, l_end = min(L, l_begin + lines_per_split);
    const bool active = rr < kCompactRows;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int cnt = 0;
    if (active) {
        const float* src = xt_full + (long long)s * L * DP + 4 * q;
        float* dst = xt_mode + (long long)s * L * DP + 4 * q;
        for (int l = l_begin + rr; l < l_end; l += kCompactRows) {
            const long long o = (long long)l * S + s;
            if (sel[o]) {
                const float4 x = *reinterpret_cast<const float4*>(src + (long long)l * DP);
                *reinterpret_cast<float4*>(dst + (long long)rowidx[o] * DP) = x;
                a0 += (double)x.x; a1 += (double)x.y; a2 += (double)x.z; a3 += (double)x.w;
                ++cnt;
            }
        }
        double* p = csm + rr * DP + 4 * q;
        p[0] = a0; p[1] = a1; p[2] = a2; p[3] = a3;
        if (q == 0) cnt_sh[rr] = cnt;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < DP; b += blockDim.x) {
        double a = 0.0;
        for (int r = 0; r < kCompactRows; ++r) a += csm[r * DP + b];
        colsum_part[((long long)split * S + s) * DP + b] = a;
    }
    if (threadIdx.x == 0) {
        int c = 0;
        for (int r = 0; r < kCompactRows; ++r) c += cnt_sh[r];
        colcnt_part[split * S + s] = c;
    }
}

// returns false when the window is too wide for the thread plan (the caller then repacks from the slab)
bool launch_compact(const Dims& d, const float* xt_full, const uint8_t* sel, const int32_t* rowidx, int nsplit,
                    int lines_per_split, float* xt_mode, double* colsum_part, int* colcnt_part, cudaStream_t st) {
    const int Q = d.DP / 4;
    if (Q * kCompactRows > 512) return false;
    dim3 grid(d.S, nsplit);
    compact_kernel<<<grid, 512, (size_t)kCompactRows * d.DP * sizeof(double), st>>>(
        xt_full, sel, rowidx, d.L, d.S, d.DP, lines_per_split, xt_mode, colsum_part, colcnt_part);
    return true;
}

}  // namespace cmf
