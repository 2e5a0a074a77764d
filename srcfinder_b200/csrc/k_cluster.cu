// Background-mode partition on the device (-k > 1): PCA projection of every column's valid pixels onto the
// leading `pcadim` eigenvectors of the column covariance, then a k-means on the projections.
//
// Reference (cmf/robust_mf.py:306-313): Icol_pca = Icol_zm . eig(cov(Icol_zm))[1][:, :pcadim];
// labels = MiniBatchKMeans(n_clusters=k).fit(Icol_pca).labels_.  That k-means is unseeded (:312), so two
// runs of the reference give different partitions; a drop-in can only be deterministic by
// choosing its own rule, which is stated here and restated in oracle/cluster_oracle.py:
//   * components: the `pcadim` largest eigenvalues, descending; each eigenvector signed so that v . mu >= 0
//     (component 1 then orders pixels dark -> bright and label numbering is reproducible);
//   * the projections are quantised to a 2^-24 grid of the column's largest |projection| and the Lloyd
//     iterations accumulate the cluster sums in 64-bit integers, so the centroids do not depend on the
//     order of summation and the numpy restatement reproduces the labels bit for bit;
//   * initial partition: k approximately equal-count slices of component 1 (4096-bin histogram between its
//     extremes); empty clusters keep their centroid; ties go to the lowest label; at most `max_iter` (100,
//     the sklearn default) iterations.
#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

// ---- per column: indices of the pcadim largest eigenvalues (descending, ties -> lower index) and signs
__global__ void __launch_bounds__(128)
    pca_pick_kernel(const double* __restrict__ lam_g, const double* __restrict__ P_g, const double* __restrict__ mu_g,
                    const int* __restrict__ n_g, int D, int DP, int pd, int* __restrict__ pick_g,
                    double* __restrict__ vtop_g) {
    __shared__ int pick[kMaxPcaDim];
    __shared__ double sign[kMaxPcaDim];
    const int s = blockIdx.x, tid = threadIdx.x;
    const double* lam = lam_g + (long long)s * DP;
    const double* P = P_g + (long long)s * DP * DP;
    const double* mu = mu_g + (long long)s * DP;
    if (tid == 0) {
        for (int p = 0; p < pd; ++p) {
            int best = -1;
            for (int j = 0; j < D; ++j) {
                bool used = false;
                for (int q = 0; q < p; ++q) used |= (pick[q] == j);
                if (used) continue;
                if (best < 0 || lam[j] > lam[best]) best = j;
            }
            pick[p] = best;
            pick_g[(long long)s * kMaxPcaDim + p] = best;
        }
    }
    __syncthreads();
    if (tid < pd) {
        const int j = pick[tid];
        double dot = 0.0;
        for (int b = 0; b < D; ++b) dot += P[b * DP + j] * mu[b];
        sign[tid] = (dot < 0.0) ? -1.0 : 1.0;
    }
    __syncthreads();
    // vtop[s][b][p]
    for (int idx = tid; idx < D * pd; idx += blockDim.x) {
        const int b = idx / pd, p = idx % pd;
        vtop_g[((long long)s * DP + b) * kMaxPcaDim + p] = (n_g[s] >= 2) ? sign[p] * P[b * DP + pick[p]] : 0.0;
    }
}

// ---- y[s][l][p] = sum_b (x[l][b] - mu[b]) v[b][p]  (FP64, b ascending), 0 for invalid pixels
// CTA = (256 lines, column): the tile of xt rows is staged in shared memory with coalesced 16-byte loads (row pitch
// DP + 1 floats: the row-wise reads of the threads fall on different banks); the wide-window layout (rows of up to
// 432 floats) does not fit and takes the direct path.
__global__ void __launch_bounds__(256)
    pca_project_kernel(const float* __restrict__ xt, const uint8_t* __restrict__ mask, const double* __restrict__ mu_g,
                       const double* __restrict__ vtop_g, int L, int S, int D, int DP, int pd, int staged,
                       double* __restrict__ y_g) {
    extern __shared__ double sm[];
    double* v = sm;                 // [D][pd]
    double* mu = v + D * pd;        // [D]
    float* tile = reinterpret_cast<float*>(mu + D);     // [256][DP + 1] when staged
    const int s = blockIdx.y, tid = threadIdx.x;
    for (int idx = tid; idx < D * pd; idx += blockDim.x)
        v[idx] = vtop_g[((long long)s * DP + idx / pd) * kMaxPcaDim + idx % pd];
    for (int b = tid; b < D; b += blockDim.x) mu[b] = mu_g[(long long)s * DP + b];
    const int l0 = blockIdx.x * blockDim.x;
    const int nl = min((int)blockDim.x, L - l0);
    const int TP = DP + 1;
    if (staged) {
        const float4* src = reinterpret_cast<const float4*>(xt + ((long long)s * L + l0) * DP);
        const int q4 = DP / 4;
        for (int i = tid; i < nl * q4; i += blockDim.x) {
            const float4 x = src[i];
            float* dst = tile + (i / q4) * TP + 4 * (i % q4);
            dst[0] = x.x; dst[1] = x.y; dst[2] = x.z; dst[3] = x.w;
        }
    }
    __syncthreads();
    const int l = l0 + tid;
    if (l >= L) return;
    double acc[kMaxPcaDim];
#pragma unroll
    for (int p = 0; p < kMaxPcaDim; ++p) acc[p] = 0.0;
    if (mask[(long long)l * S + s]) {
        const float* row = staged ? tile + tid * TP : xt + ((long long)s * L + l) * DP;
        for (int b = 0; b < D; ++b) {
            const double xc = (double)row[b] - mu[b];
#pragma unroll
            for (int p = 0; p < kMaxPcaDim; ++p)
                if (p < pd) acc[p] = __dadd_rn(acc[p], __dmul_rn(xc, v[b * pd + p]));
        }
    }
    double* out = y_g + ((long long)s * L + l) * pd;
#pragma unroll
    for (int p = 0; p < kMaxPcaDim; ++p)
        if (p < pd) out[p] = acc[p];
}

// ---- k-means, one CTA per column
constexpr int kKmThreads = 256;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmInitBins = 4096;
constexpr int kSmallK = 4;          // up to this many clusters the sums are kept per thread

__device__ __forceinline__ double block_max(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[0];
#pragma unroll
    for (int w = 1; w < kKmWarps; ++w) t = fmax(t, red[w]);
    return t;
}

__device__ __forceinline__ long long block_minmax_i64(long long v, bool want_max, long long* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long u = __shfl_xor_sync(0xffffffffu, v, o);
        v = want_max ? (u > v ? u : v) : (u < v ? u : v);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    long long t = red[0];
#pragma unroll
    for (int w = 1; w < kKmWarps; ++w) t = want_max ? (red[w] > t ? red[w] : t) : (red[w] < t ? red[w] : t);
    return t;
}

// The Lloyd iterations work on the quantised projections only: they are written once as int32 [pd][L] per column
// (component-major, so every load is a coalesced run of lines; INT_MIN in component 0 marks an invalid pixel, so
// the strided mask image is not read again) and every iteration is ONE pass: a pixel is reassigned and added to the
// integer sums of its new cluster in the same sweep.  The per-warp sums are built from warp-wide integer
// reductions (REDUX: |q| <= 2^24, 32 lanes fit int32) instead of shared-memory atomics that all lanes of a
// spatially coherent batch would aim at the same address.  Arithmetic and tie rules are unchanged (exact integer
// sums, double distances summed over the components in order), so labels are bit-identical to the numpy
// restatement in oracle/cluster_oracle.py.
template <int PDM>
__global__ void __launch_bounds__(kKmThreads, PDM <= 8 ? 3 : 2)
    kmeans_kernel(const double* __restrict__ y_g, const uint8_t* __restrict__ mask, const int* __restrict__ n_g,
                  int L, int S, int pd, int k, int max_iter, int32_t* __restrict__ q_g, uint8_t* __restrict__ lab8_g,
                  int32_t* __restrict__ labels, int* __restrict__ iters_g) {
    extern __shared__ unsigned long long smu[];
    // per-warp integer accumulators [warp][k][pd + 1] (sums, then the count), centroids [k][pd]
    const int stride = pd + 1;
    long long* acc = reinterpret_cast<long long*>(smu);
    double* cen = reinterpret_cast<double*>(acc + kKmWarps * k * stride);
    int* priv = reinterpret_cast<int*>(cen + k * pd);   // [k][pd + 1][threads], only when k <= kSmallK
    const bool small = k <= kSmallK;
    __shared__ double redd[kKmWarps];
    __shared__ long long redi[kKmWarps];
    __shared__ int changed;
    __shared__ int hist[kKmInitBins], part[kKmThreads];
    constexpr int32_t kInvalid = INT32_MIN;

    const int s = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const double* y = y_g + (long long)s * L * pd;
    int32_t* q = q_g + (long long)s * L * pd;              // [pd][L]
    uint8_t* lab8 = lab8_g + (long long)s * L;
    const int n = n_g[s];
    if (n == 0) {
        for (int l = tid; l < L; l += blockDim.x) labels[(long long)l * S + s] = 0;
        if (tid == 0) iters_g[s] = 0;
        return;
    }
    // quantisation scale: 2^(24 - e) with 2^(e-1) <= max|y| < 2^e
    double mx = 0.0;
    for (int l = tid; l < L; l += blockDim.x)
        if (mask[(long long)l * S + s])
            for (int p = 0; p < pd; ++p) mx = fmax(mx, fabs(y[(long long)l * pd + p]));
    mx = block_max(mx, redd);
    int e = 0;
    if (mx > 0.0) frexp(mx, &e);
    const double scale = ldexp(1.0, 24 - e);
    // quantise once; extremes of component 1
    long long lo = (1LL << 62), hi = -(1LL << 62);
    for (int l = tid; l < L; l += blockDim.x) {
        if (mask[(long long)l * S + s]) {
            for (int p = 0; p < pd; ++p) q[(long long)p * L + l] = (int32_t)llrint(y[(long long)l * pd + p] * scale);
            const long long q0 = q[l];
            lo = q0 < lo ? q0 : lo;
            hi = q0 > hi ? q0 : hi;
        } else {
            q[l] = kInvalid;
        }
    }
    lo = block_minmax_i64(lo, false, redi);
    hi = block_minmax_i64(hi, true, redi);
    const long long span = hi - lo + 1;
    // initial partition: approximately equal-count slices of component 1 -- histogram, exclusive prefix,
    // label(bin) = min(k-1, below(bin) * k / n).  (Equal-width slices would hand whole clusters to single
    // saturated outliers.)
    for (int i = tid; i < kKmInitBins; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    for (int l = tid; l < L; l += blockDim.x) {
        const int32_t q0 = q[l];
        if (q0 != kInvalid) atomicAdd(&hist[(int)((((long long)q0 - lo) * kKmInitBins) / span)], 1);
    }
    __syncthreads();
    {
        constexpr int per = kKmInitBins / kKmThreads;
        int local = 0;
        for (int i = 0; i < per; ++i) local += hist[tid * per + i];
        part[tid] = local;
        __syncthreads();
        if (tid == 0) {
            int run = 0;
            for (int i = 0; i < kKmThreads; ++i) { const int v = part[i]; part[i] = run; run += v; }
        }
        __syncthreads();
        long long below = part[tid];
        for (int i = 0; i < per; ++i) {
            const int cnt = hist[tid * per + i];
            const long long lab = (below * k) / n;
            hist[tid * per + i] = (int)(lab < k - 1 ? lab : k - 1);
            below += cnt;
        }
    }
    for (int i = tid; i < k * pd; i += blockDim.x) cen[i] = 0.0;
    for (int i = tid; i < kKmWarps * k * stride; i += blockDim.x) acc[i] = 0;
    if (small) for (int i = tid; i < k * stride * kKmThreads; i += blockDim.x) priv[i] = 0;
    if (tid == 0) changed = 0;
    __syncthreads();

    long long* mine = acc + warp * k * stride;
    const int Lpad = (L + 31) & ~31;                       // whole warps walk the column together
    // one sweep: (assign) the label of every valid pixel, then its quantised components into the sums of that label
    auto sweep = [&](bool assign) {
        int any = 0;                                       // pixels of this thread that changed cluster
        int since = 0;
        // thread-private 32-bit sums in shared memory, element (c, p) of thread t at priv[(c * stride + p) * T + t]
        auto flush = [&]() {
            for (int i = 0; i < k * stride; ++i) {
                const int v = priv[i * kKmThreads + tid];
                if (v != 0) {
                    atomicAdd(reinterpret_cast<unsigned long long*>(mine + i), (unsigned long long)(long long)v);
                    priv[i * kKmThreads + tid] = 0;
                }
            }
        };
        for (int l = tid; l < Lpad; l += blockDim.x) {
            int32_t qi[PDM];
#pragma unroll
            for (int p = 0; p < PDM; ++p) qi[p] = (p < pd && l < L) ? q[(long long)p * L + l] : kInvalid;
            const bool valid = l < L && qi[0] != kInvalid;
            int lab = -1;
            if (valid) {
                if (assign) {
                    // nearest centroid, squared distance summed over components in order, first minimum
                    int best = 0;
                    double bd = 0.0;
                    for (int c = 0; c < k; ++c) {
                        double dsq = 0.0;
#pragma unroll
                        for (int p = 0; p < PDM; ++p)
                            if (p < pd) {
                                const double df = __dsub_rn((double)qi[p], cen[c * pd + p]);
                                dsq = __dadd_rn(dsq, __dmul_rn(df, df));
                            }
                        if (c == 0 || dsq < bd) { bd = dsq; best = c; }
                    }
                    if (best != lab8[l]) { lab8[l] = (uint8_t)best; ++any; }
                    lab = best;
                } else {
                    lab = hist[(int)((((long long)qi[0] - lo) * kKmInitBins) / span)];
                    lab8[l] = (uint8_t)lab;
                }
            } else if (!assign && l < L) {
                lab8[l] = 0;
            }
            if (small) {
                // few clusters: private 32-bit sums per thread (|q| <= 2^24, flushed before 64 additions could
                // overflow) -- plain read-modify-write, no atomics, no warp reductions
                if (lab >= 0) {
                    int* dst = priv + (lab * stride) * kKmThreads + tid;
#pragma unroll
                    for (int p = 0; p < PDM; ++p)
                        if (p < pd) dst[p * kKmThreads] += qi[p];
                    dst[pd * kKmThreads] += 1;
                }
                if (++since == 64) { flush(); since = 0; }
                continue;
            }
            // integer sums of the batch, one cluster at a time (a batch of 32 neighbouring lines rarely holds
            // more than two)
            unsigned todo = __ballot_sync(0xffffffffu, lab >= 0);
            while (todo) {
                const int c = __shfl_sync(0xffffffffu, lab, __ffs(todo) - 1);
                const bool in = lab == c;
                const unsigned members = __ballot_sync(0xffffffffu, in);
#pragma unroll
                for (int p = 0; p < PDM; ++p)
                    if (p < pd) {
                        const int v = __reduce_add_sync(0xffffffffu, in ? qi[p] : 0);
                        if (lane == 0) mine[c * stride + p] += v;
                    }
                if (lane == 0) mine[c * stride + pd] += __popc(members);
                todo &= ~members;
            }
        }
        if (small) flush();
        return any;
    };
    sweep(false);
    __syncthreads();

    int iter = 0;
    for (;; ++iter) {
        // ---- centroids of the current partition (integer sums: order-free, exact)
        for (int i = tid; i < k * pd; i += blockDim.x) {
            const int c = i / pd, p = i % pd;
            long long sum = 0, cnt = 0;
            for (int w = 0; w < kKmWarps; ++w) {
                sum += acc[(w * k + c) * stride + p];
                cnt += acc[(w * k + c) * stride + pd];
            }
            if (cnt > 0) cen[i] = __ddiv_rn((double)sum, (double)cnt);
        }
        __syncthreads();
        if (iter >= max_iter) break;
        for (int i = tid; i < kKmWarps * k * stride; i += blockDim.x) acc[i] = 0;
        __syncthreads();
        // ---- reassign and rebuild the sums in the same pass
        {
            const int mychg = sweep(true);
            if (mychg) atomicAdd(&changed, mychg);         // integer count: order-free
        }
        __syncthreads();
        const int ch = changed;
        __syncthreads();
        if (tid == 0) changed = 0;
        if (!ch) break;
        // converged to within 2^-8 of the column: the reassignment is kept and the iteration stops (the last
        // per-mille of boundary pixels otherwise flip for dozens of sweeps; the reference's MiniBatchKMeans stops
        // on a tolerance as well)
        if ((long long)ch * 256 <= (long long)n) { ++iter; break; }
    }
    for (int l = tid; l < L; l += blockDim.x)
        labels[(long long)l * S + s] = (q[l] != kInvalid) ? (int32_t)lab8[l] : 0;
    if (tid == 0) iters_g[s] = iter;
}

void launch_pca_kmeans(const Dims& d, const float* xt, const uint8_t* mask, const double* mu, const int* n,
                       const double* P, const double* lam, int pcadim, int k, int max_iter, int* pick,
                       double* vtop, double* y, int32_t* q, uint8_t* lab8, int32_t* labels, int* iters,
                       cudaStream_t st) {
    pca_pick_kernel<<<d.S, 128, 0, st>>>(lam, P, mu, n, d.D, d.DP, pcadim, pick, vtop);
    const dim3 grid((d.L + 255) / 256, d.S);
    size_t smem = (size_t)(d.D * pcadim + d.D) * sizeof(double);
    const size_t tile = (size_t)256 * (d.DP + 1) * sizeof(float);
    const int staged = (smem + tile <= 160 * 1024) ? 1 : 0;
    if (staged) smem += tile;
    cudaFuncSetAttribute(pca_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    pca_project_kernel<<<grid, 256, smem, st>>>(xt, mask, mu, vtop, d.L, d.S, d.D, d.DP, pcadim, staged, y);
    const size_t smem2 = (size_t)(kKmWarps * k * (pcadim + 1)) * sizeof(long long) + (size_t)k * pcadim * sizeof(double) +
                         (k <= kSmallK ? (size_t)k * (pcadim + 1) * kKmThreads * sizeof(int) : 0);
    if (pcadim <= 8) {
        cudaFuncSetAttribute(kmeans_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        kmeans_kernel<8><<<d.S, kKmThreads, smem2, st>>>(y, mask, n, d.L, d.S, pcadim, k, max_iter, q, lab8, labels, iters);
    } else {
        cudaFuncSetAttribute(kmeans_kernel<kMaxPcaDim>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        kmeans_kernel<kMaxPcaDim><<<d.S, kKmThreads, smem2, st>>>(y, mask, n, d.L, d.S, pcadim, k, max_iter, q, lab8, labels,
                                                              iters);
    }
}

}  // namespace cmf
