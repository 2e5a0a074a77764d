// K2: per-column spectral factorisation of the shrinkage family, K4: alpha selection + filter weights.
//
// The reference evaluates, for each of 201 alphas, det(G), inv(G) and an n x D x D product with
// G = n*beta*S + alpha*T, T = diag(S)  (cmf/robust_mf.py:105-117).  With R = T^-1/2 S T^-1/2 = V Lam V^T
// (one symmetric eigendecomposition per column, done here by cyclic Jacobi in shared memory):
//     G_alpha^-1       = T^-1/2 V (n*beta*Lam + alpha I)^-1 V^T T^-1/2
//     r_k(alpha)       = sum_j y_kj^2 / (n*beta*lam_j + alpha),          y_k = (T^-1/2 V)^T x_k
//     log det G_alpha  = sum_j log(1e4 T_jj) + sum_j log(n*beta*lam_j + alpha)   (1e4: the x100 scaling, :94)
//     C_alpha^-1 t     = T^-1/2 V diag(1 / ((1-alpha) lam_j + alpha)) V^T T^-1/2 t          (:130-136, :363)
// K2 emits P = T^-1/2 V and the table W[j][i] = 1/(n beta_i lam_j + alpha_i) in DMMA fragment order for
// the LOO pass (K3); K4 reduces the LOO sums, reproduces the reference's argmin rules and forms the
// matched-filter weights  w = C^-1 t / (t^T C^-1 t) * scale  (:376-384).
#include <math.h>

#include "cmf_common.cuh"
#include "cmf_internal.h"

namespace cmf {

constexpr int kJacobiMaxSweeps = 30;

__device__ __forceinline__ void tri_unrank(int t, int& i, int& j) {
    // t = i(i+1)/2 + j, j <= i
    int ii = (int)((sqrtf(8.0f * t + 1.0f) - 1.0f) * 0.5f);
    while ((ii + 1) * (ii + 2) / 2 <= t) ++ii;
    while (ii * (ii + 1) / 2 > t) --ii;
    i = ii;
    j = t - ii * (ii + 1) / 2;
}

// sum_l (x-c)(x-c)^T = sum_l (x-mu)(x-mu)^T + n (mu-c)(mu-c)^T: the term the pilot-centred Gram pass leaves in
__device__ __forceinline__ double pilot_term(const double* __restrict__ mu, const double* __restrict__ ctr,
                                             long long base, int row, int col, int n) {
    const double dr = mu[base + row] - ctr[base + row], dc = mu[base + col] - ctr[base + col];
    return (double)n * dr * dc;
}

#ifdef CMF_TUNING_HOOKS     // cyclic Jacobi: cross-check of the QL solver, tools build only
template <int NT>
__global__ void __launch_bounds__(512)
    eigen_kernel(const double* __restrict__ gram_part, int nchunk, const int* __restrict__ n_g, int D,
        const double* __restrict__ mu_g, const double* __restrict__ ctr_g,
                 double* __restrict__ P_g, double* __restrict__ lam_g, double* __restrict__ slogT_g,
                 int* __restrict__ status_g, int* __restrict__ sweeps_g) {
    constexpr int DP = 8 * NT, LD = DP + 1, NTRI = NT * (NT + 1) / 2;
    extern __shared__ double sm[];
    double* Am = sm;                 // [DP][LD]  correlation matrix -> diagonal
    double* Vm = Am + DP * LD;       // [DP][LD]  eigenvectors (columns)
    double* dinv = Vm + DP * LD;     // [DP]      T^-1/2
    double* lam = dinv + DP;         // [DP]
    double* rc = lam + DP;           // [DP/2+1] rotation cosines
    double* rs = rc + (DP / 2 + 1);  // [DP/2+1] rotation sines
    __shared__ int rp[DP / 2 + 1], rq[DP / 2 + 1];
    __shared__ int rotated;
    __shared__ double sumlogT;

    const int s = blockIdx.x, tid = threadIdx.x;
    const int n = n_g[s];
    double* Pout = P_g + (long long)s * DP * DP;

    if (n < 2) {  // empty or single-pixel column: nothing to factorise (handled in K4)
        for (int i = tid; i < DP * DP; i += blockDim.x) Pout[i] = 0.0;
        for (int i = tid; i < DP; i += blockDim.x) lam_g[(long long)s * DP + i] = 0.0;
        if (tid == 0) {
            status_g[s] = (n == 0) ? kStatusEmpty : kStatusDegenerate;
            sweeps_g[s] = 0;
            slogT_g[s] = 0.0;
        }
        return;
    }

    // ---- assemble the symmetric Gram from the chunk partials (fixed order) and scale to a covariance
    const double inv_nm1 = 1.0 / (double)(n - 1);
    for (int idx = tid; idx < NTRI * 64; idx += blockDim.x) {
        const int t = idx >> 6, within = idx & 63;
        const int lane = within >> 1, e = within & 1;
        int ti, tj;
        tri_unrank(t, ti, tj);
        double v = 0.0;
        for (int c = 0; c < nchunk; ++c) v += gram_part[((long long)s * nchunk + c) * NTRI * 64 + idx];
        const int row = 8 * ti + (lane >> 2), col = 8 * tj + 2 * (lane & 3) + e;
        if (ctr_g != nullptr) v -= pilot_term(mu_g, ctr_g, (long long)s * DP, row, col, n);
        v *= inv_nm1;
        Am[row * LD + col] = v;
        if (ti != tj) Am[col * LD + row] = v;
    }
    __syncthreads();
    if (tid < DP) {
        const double t0 = (tid < D) ? Am[tid * LD + tid] : 0.0;
        dinv[tid] = (t0 > 0.0) ? 1.0 / sqrt(t0) : 0.0;
    }
    __syncthreads();
    if (tid == 0) {
        double a = 0.0;
        for (int b = 0; b < D; ++b) a += log(1.0e4 * Am[b * LD + b]);   // log det of the scaled T (:94-99)
        sumlogT = a;
    }
    __syncthreads();
    for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
        const int r = idx / DP, c = idx % DP;
        double v = Am[r * LD + c] * dinv[r] * dinv[c];
        if (r == c) v = (r < D && dinv[r] > 0.0) ? 1.0 : 0.0;
        if (r >= D || c >= D) v = 0.0;
        Am[r * LD + c] = v;
        Vm[r * LD + c] = (r == c) ? 1.0 : 0.0;
    }
    __syncthreads();

    // ---- cyclic Jacobi, round-robin ordering: m/2 disjoint rotations per step, m-1 steps per sweep.
    // Each step is two barriers: (1) one thread per pair picks its rotation, (2) A <- J^T A J is applied
    // as independent 2x2 blocks (pair k rows x pair l columns, updated in place by one thread) together
    // with V <- V J.  For odd D the bye slot is the zero padding row/column D (identity rotation).
    const int m = (D & 1) ? D + 1 : D;
    const int half = m / 2;
    int sweep = 0;
    for (; sweep < kJacobiMaxSweeps; ++sweep) {
        if (tid == 0) rotated = 0;
        __syncthreads();
        for (int r = 0; r < m - 1; ++r) {
            if (tid < half) {
                int p, q;
                if (tid == 0) { p = m - 1; q = r; }
                else { p = (r + tid) % (m - 1); q = (r - tid + (m - 1)) % (m - 1); }
                if (p > q) { const int t = p; p = q; q = t; }
                double c = 1.0, sn = 0.0;
                if (q < D) {
                    const double apq = Am[p * LD + q];
                    const double app = Am[p * LD + p], aqq = Am[q * LD + q];
                    if (fabs(apq) > 1.0e-300 && fabs(apq) > 1.1e-16 * sqrt(fabs(app * aqq))) {
                        const double theta = (aqq - app) / (2.0 * apq);
                        const double t = copysign(1.0, theta) / (fabs(theta) + sqrt(theta * theta + 1.0));
                        c = 1.0 / sqrt(t * t + 1.0);
                        sn = t * c;
                        rotated = 1;
                    }
                }
                rc[tid] = c; rs[tid] = sn; rp[tid] = p; rq[tid] = q;
            }
            __syncthreads();
            const int nblk = half * half;
            for (int idx = tid; idx < nblk + half * D; idx += blockDim.x) {
                if (idx < nblk) {
                    const int k = idx / half, l = idx - k * half;
                    const double ck = rc[k], sk = rs[k], cl = rc[l], sl = rs[l];
                    if (sk != 0.0 || sl != 0.0) {
                        const int pk = rp[k], qk = rq[k], pl = rp[l], ql = rq[l];
                        const double b00 = Am[pk * LD + pl], b01 = Am[pk * LD + ql];
                        const double b10 = Am[qk * LD + pl], b11 = Am[qk * LD + ql];
                        const double t00 = ck * b00 - sk * b10, t01 = ck * b01 - sk * b11;
                        const double t10 = sk * b00 + ck * b10, t11 = sk * b01 + ck * b11;
                        Am[pk * LD + pl] = cl * t00 - sl * t01;
                        Am[pk * LD + ql] = sl * t00 + cl * t01;
                        Am[qk * LD + pl] = cl * t10 - sl * t11;
                        Am[qk * LD + ql] = sl * t10 + cl * t11;
                    }
                } else {
                    const int j = idx - nblk;
                    const int l = j / D, row = j - l * D;
                    const double cl = rc[l], sl = rs[l];
                    if (sl != 0.0) {
                        const int pl = rp[l], ql = rq[l];
                        const double vp = Vm[row * LD + pl], vq = Vm[row * LD + ql];
                        Vm[row * LD + pl] = cl * vp - sl * vq;
                        Vm[row * LD + ql] = sl * vp + cl * vq;
                    }
                }
            }
            __syncthreads();
        }
        if (!rotated) break;
        __syncthreads();
    }
    if (tid < DP) {
        lam[tid] = (tid < D) ? Am[tid * LD + tid] : 0.0;
        lam_g[(long long)s * DP + tid] = lam[tid];
    }
    if (tid == 0) {
        status_g[s] = (sweep >= kJacobiMaxSweeps) ? kStatusNoConverge : kStatusOk;
        sweeps_g[s] = sweep;
    }
    __syncthreads();

    // ---- P = T^-1/2 V (row-major; K2b derives the fragment-ordered tables)
    for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
        const int b = idx / DP, j = idx % DP;
        Pout[idx] = (b < D && j < D) ? dinv[b] * Vm[b * LD + j] : 0.0;
    }
    if (tid == 0) slogT_g[s] = sumlogT;
}

#endif  // CMF_TUNING_HOOKS

// ---------------------------------------------------------------------------------------- K2 (QL)
// Householder tridiagonalisation + implicit-shift QL with accumulated transformations (the EISPACK
// tred2 / tql2 pair; the same reduction LAPACK's dsteqr path uses) of the D x D correlation matrix, one
// 128-thread CTA per column with the matrix in shared memory.  Against the cyclic Jacobi above it needs
// ~10x fewer shared-memory passes (Jacobi: ~10 sweeps x 71 steps x the whole of A and V; QL: ~1.7
// iterations per eigenvalue, each a chain of plane rotations on two columns), so all columns of a
// flightline are resident at once (5 CTAs per SM) and the pass is bound by the serial rotation recurrence.
constexpr int kQlThreads = 128;
constexpr int kQlMaxIter = 60;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum of one double per thread (all threads get the result); red has one slot per warp
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kQlThreads / 32; ++w) t += red[w];
    return t;
}

// MODE selects the shrinkage target T the column covariance S is whitened with before the eigen-solve:
//   kEigDiag  T = diag(S) (the default target, :100): R = T^-1/2 S T^-1/2, P = T^-1/2 V
//   kEigNone  T = I: plain eigenvectors of the covariance (the PCA of the multimodal partition, :310-311)
//   kEigFull  T = covariance of the whole column (-f, :100 with I_reg, :358): T = L L^T (Cholesky in shared
//             memory), R = L^-1 S L^-T, P = L^-T V, log det T = 2 sum log L_jj.  Every downstream formula
//             (G_alpha^-1 = P (n beta Lam + alpha)^-1 P^T, C^-1 = P ((1-alpha) Lam + alpha)^-1 P^T) is unchanged.
enum { kEigDiag = 0, kEigNone = 1, kEigFull = 2 };

template <int NT, int MODE>
__global__ void __launch_bounds__(kQlThreads, 5)
    eigen_ql_kernel(const double* __restrict__ gram_part, int nchunk, const int* __restrict__ n_g, int D,
        const double* __restrict__ mu_g, const double* __restrict__ ctr_g,
                    double* __restrict__ P_g, double* __restrict__ lam_g, double* __restrict__ slogT_g,
                    int* __restrict__ status_g, int* __restrict__ sweeps_g,
                    const double* __restrict__ gramT_part, const int* __restrict__ nT_g) {
    constexpr int DP = 8 * NT, LD = DP + 1, NTRI = NT * (NT + 1) / 2;
    extern __shared__ double sm[];
    double* a = sm;                  // [DP][LD]  correlation matrix -> Householder vectors -> eigenvectors
    double* dinv = a + DP * LD;      // [DP]      T^-1/2
    double* d = dinv + DP;           // [DP]      diagonal -> eigenvalues
    double* e = d + DP;              // [DP]      off-diagonal
    double* cs = e + DP;             // [2*DP]    rotation (c, s) pairs of one QL iteration
    double* red = cs + 2 * DP;       // [8]       reduction scratch / broadcast
    double* tl = red + 8;            // [DP][LD]  kEigFull only: T -> its Cholesky factor L (lower)
    __shared__ int sh_m2[2], sh_cnt2[2], sh_flag, sh_bad, sh_iter;

#ifdef CMF_EIGEN_PROF
    const long long pstart = clock64();
#endif
    const int s = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = kQlThreads / 32;
    const int n = n_g[s];
    double* Pout = P_g + (long long)s * DP * DP;
    if (n < 2) {
        for (int i = tid; i < DP * DP; i += blockDim.x) Pout[i] = 0.0;
        for (int i = tid; i < DP; i += blockDim.x) lam_g[(long long)s * DP + i] = 0.0;
        if (tid == 0) {
            status_g[s] = (n == 0) ? kStatusEmpty : kStatusDegenerate;
            sweeps_g[s] = 0;
            slogT_g[s] = 0.0;
        }
        return;
    }
    // ---- covariance from the Gram partials (fixed order), correlation scaling
    const double inv_nm1 = 1.0 / (double)(n - 1);
    for (int idx = tid; idx < NTRI * 64; idx += blockDim.x) {
        const int t = idx >> 6, within = idx & 63;
        const int ln = within >> 1, ee = within & 1;
        int ti, tj;
        tri_unrank(t, ti, tj);
        // same left-to-right order as a plain loop, but four chunk loads are in flight at a time
        double v = 0.0;
        const double* gp = gram_part + (long long)s * nchunk * NTRI * 64 + idx;
        int c = 0;
        for (; c + 4 <= nchunk; c += 4) {
            const double x0 = gp[(long long)c * NTRI * 64], x1 = gp[(long long)(c + 1) * NTRI * 64],
                         x2 = gp[(long long)(c + 2) * NTRI * 64], x3 = gp[(long long)(c + 3) * NTRI * 64];
            v = (((v + x0) + x1) + x2) + x3;
        }
        for (; c < nchunk; ++c) v += gp[(long long)c * NTRI * 64];
        const int row = 8 * ti + (ln >> 2), col = 8 * tj + 2 * (ln & 3) + ee;
        if (ctr_g != nullptr) v -= pilot_term(mu_g, ctr_g, (long long)s * DP, row, col, n);
        v *= inv_nm1;
        a[row * LD + col] = v;
        if (ti != tj) a[col * LD + row] = v;
    }
    __syncthreads();
    if (MODE == kEigDiag) {
        if (tid < DP) {
            const double t0 = (tid < D) ? a[tid * LD + tid] : 0.0;
            dinv[tid] = (t0 > 0.0) ? 1.0 / sqrt(t0) : 0.0;
        }
        __syncthreads();
        if (tid == 0) {
            double acc = 0.0;
            for (int b = 0; b < D; ++b) acc += log(1.0e4 * a[b * LD + b]);   // log det of the scaled T (:94-99)
            slogT_g[s] = acc;
        }
        __syncthreads();
        for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
            const int r = idx / DP, c = idx % DP;
            double v = a[r * LD + c] * dinv[r] * dinv[c];
            if (r == c) v = (r < D && dinv[r] > 0.0) ? 1.0 : 0.0;
            if (r >= D || c >= D) v = 0.0;
            a[r * LD + c] = v;
        }
        __syncthreads();
    } else if (MODE == kEigNone) {
        if (tid < DP) dinv[tid] = (tid < D) ? 1.0 : 0.0;
        if (tid == 0) slogT_g[s] = 0.0;
        for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
            const int r = idx / DP, c = idx % DP;
            if (r >= D || c >= D) a[r * LD + c] = 0.0;
        }
        __syncthreads();
    } else {
        // ---- T = covariance of the whole column from its own Gram partials (centred on the column mean)
        const int nT = nT_g[s];
        const double inv_nT = 1.0 / (double)(nT - 1);
        for (int idx = tid; idx < NTRI * 64; idx += blockDim.x) {
            const int t = idx >> 6, within = idx & 63;
            const int ln = within >> 1, ee = within & 1;
            int ti, tj;
            tri_unrank(t, ti, tj);
            double v = 0.0;
            for (int c = 0; c < nchunk; ++c) v += gramT_part[((long long)s * nchunk + c) * NTRI * 64 + idx];
            const int row = 8 * ti + (ln >> 2), col = 8 * tj + 2 * (ln & 3) + ee;
            v *= inv_nT;
            tl[row * LD + col] = v;
            if (ti != tj) tl[col * LD + row] = v;
        }
        if (tid == 0) sh_bad = 0;
        __syncthreads();
        // right-looking Cholesky, lower triangle in place
        for (int j = 0; j < D; ++j) {
            const double piv = tl[j * LD + j];
            __syncthreads();
            if (!(piv > 0.0)) { if (tid == 0) sh_bad = 1; break; }
            const double ljj = sqrt(piv);
            for (int i = j + tid; i < D; i += blockDim.x) tl[i * LD + j] = (i == j) ? ljj : tl[i * LD + j] / ljj;
            __syncthreads();
            for (int i = j + 1 + warp; i < D; i += NW) {
                const double lij = tl[i * LD + j];
                for (int k = j + 1 + lane; k <= i; k += 32) tl[i * LD + k] -= lij * tl[k * LD + j];
            }
            __syncthreads();
        }
        __syncthreads();
        if (sh_bad) {   // T not positive definite: no usable factor, report the mode as singular (:371-374)
            for (int i = tid; i < DP * DP; i += blockDim.x) Pout[i] = 0.0;
            for (int i = tid; i < DP; i += blockDim.x) lam_g[(long long)s * DP + i] = 0.0;
            if (tid == 0) { status_g[s] = kStatusSingular; sweeps_g[s] = 0; slogT_g[s] = 0.0; }
            return;
        }
        if (tid == 0) {
            double acc = (double)D * log(1.0e4);                            // the x100 scaling of I_reg (:100)
            for (int b = 0; b < D; ++b) acc += 2.0 * log(tl[b * LD + b]);
            slogT_g[s] = acc;
        }
        if (tid < DP) dinv[tid] = (tid < D) ? 1.0 : 0.0;
        // X = L^-1 S (forward substitution, one column of S per thread), then R = X L^-T = (L^-1 X^T)^T
        for (int c = tid; c < D; c += blockDim.x) {
            for (int i = 0; i < D; ++i) {
                double v = a[i * LD + c];
                for (int k = 0; k < i; ++k) v -= tl[i * LD + k] * a[k * LD + c];
                a[i * LD + c] = v / tl[i * LD + i];
            }
        }
        __syncthreads();
        // rows of X are independent: thread r owns row r; strided by LD so the accesses do not conflict
        for (int r = tid; r < D; r += blockDim.x) {
            for (int i = 0; i < D; ++i) {
                double v = a[r * LD + i];
                for (int k = 0; k < i; ++k) v -= tl[i * LD + k] * a[r * LD + k];
                a[r * LD + i] = v / tl[i * LD + i];
            }
        }
        __syncthreads();
        for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
            const int r = idx / DP, c = idx % DP;
            if (r >= D || c >= D) a[r * LD + c] = 0.0;
            else if (c < r) a[r * LD + c] = 0.5 * (a[r * LD + c] + a[c * LD + r]);   // tred2 reads the lower triangle
        }
        __syncthreads();
    }

#ifdef CMF_EIGEN_PROF
    long long pt0 = clock64(), pt1 = 0, pt2 = 0, pt3 = 0, pser = 0;
#endif
    // ---- tred2: reduce to tridiagonal form, lower triangle, rows n-1 .. 1
    for (int i = D - 1; i >= 1; --i) {
        const int l = i - 1;
        double h = 0.0;
        if (l > 0) {
            double part = 0.0;
            for (int k = tid; k <= l; k += blockDim.x) part += fabs(a[i * LD + k]);
            const double scale = block_sum(part, red);
            if (scale == 0.0) {
                if (tid == 0) e[i] = a[i * LD + l];
            } else {
                part = 0.0;
                for (int k = tid; k <= l; k += blockDim.x) {
                    const double v = a[i * LD + k] / scale;
                    a[i * LD + k] = v;
                    part += v * v;
                }
                h = block_sum(part, red);
                double f = a[i * LD + l];
                const double g = (f >= 0.0) ? -sqrt(h) : sqrt(h);
                h -= f * g;
                __syncthreads();                       // everyone has read a[i][l]
                if (tid == 0) { e[i] = scale * g; a[i * LD + l] = f - g; }
                __syncthreads();
                // p = A u / h  ->  e[0..l]
                part = 0.0;
                for (int j = tid; j <= l; j += blockDim.x) {
                    a[j * LD + i] = a[i * LD + j] / h;
                    // four independent partial sums: a dependent FP64 add costs far more than its issue slot
                    double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                    int k = 0;
                    for (; k + 3 <= j; k += 4) {
                        g0 += a[j * LD + k] * a[i * LD + k];
                        g1 += a[j * LD + k + 1] * a[i * LD + k + 1];
                        g2 += a[j * LD + k + 2] * a[i * LD + k + 2];
                        g3 += a[j * LD + k + 3] * a[i * LD + k + 3];
                    }
                    for (; k <= j; ++k) g0 += a[j * LD + k] * a[i * LD + k];
                    k = j + 1;
                    for (; k + 3 <= l; k += 4) {
                        g0 += a[k * LD + j] * a[i * LD + k];
                        g1 += a[(k + 1) * LD + j] * a[i * LD + k + 1];
                        g2 += a[(k + 2) * LD + j] * a[i * LD + k + 2];
                        g3 += a[(k + 3) * LD + j] * a[i * LD + k + 3];
                    }
                    for (; k <= l; ++k) g0 += a[k * LD + j] * a[i * LD + k];
                    double gj = ((g0 + g1) + (g2 + g3)) / h;
                    e[j] = gj;
                    part += gj * a[i * LD + j];
                }
                f = block_sum(part, red);
                const double hh = f / (h + h);
                for (int j = tid; j <= l; j += blockDim.x) e[j] -= hh * a[i * LD + j];
                __syncthreads();
                // A <- A - u q^T - q u^T on the lower triangle
                for (int j = warp; j <= l; j += NW) {
                    const double fj = a[i * LD + j], gj = e[j];
                    for (int k = lane; k <= j; k += 32) a[j * LD + k] -= fj * e[k] + gj * a[i * LD + k];
                }
            }
        } else {
            if (tid == 0) e[i] = a[i * LD + l];
        }
        if (tid == 0) d[i] = h;
        __syncthreads();
    }
    if (tid == 0) { d[0] = 0.0; e[0] = 0.0; }
    __syncthreads();
#ifdef CMF_EIGEN_PROF
    pt1 = clock64();
#endif
    // ---- accumulate the transformations: a becomes the orthogonal matrix Q
    for (int i = 0; i < D; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            // g_j = sum_k a[i][k] a[k][j]; kept in cs[] (free until the QL phase)
            for (int j = tid; j <= l; j += blockDim.x) {
                double g0 = 0.0, g1 = 0.0, g2 = 0.0, g3 = 0.0;
                int k = 0;
                for (; k + 3 <= l; k += 4) {
                    g0 += a[i * LD + k] * a[k * LD + j];
                    g1 += a[i * LD + k + 1] * a[(k + 1) * LD + j];
                    g2 += a[i * LD + k + 2] * a[(k + 2) * LD + j];
                    g3 += a[i * LD + k + 3] * a[(k + 3) * LD + j];
                }
                for (; k <= l; ++k) g0 += a[i * LD + k] * a[k * LD + j];
                cs[j] = (g0 + g1) + (g2 + g3);
            }
            __syncthreads();
            for (int k = warp; k <= l; k += NW) {
                const double aki = a[k * LD + i];
                for (int j = lane; j <= l; j += 32) a[k * LD + j] -= cs[j] * aki;
            }
        }
        __syncthreads();
        if (tid == 0) { d[i] = a[i * LD + i]; a[i * LD + i] = 1.0; }
        for (int j = tid; j <= l; j += blockDim.x) { a[j * LD + i] = 0.0; a[i * LD + j] = 0.0; }
        __syncthreads();
    }

#ifdef CMF_EIGEN_PROF
    pt2 = clock64();
#endif
    // ---- tql2: implicit-shift QL on (d, e).  The rotation recurrence of an iteration is a serial FP64 chain
    // (it was 60 % of this kernel), and it only needs (d, e) -- not the eigenvectors.  So warp 0 (lane 0) runs
    // the recurrences of ALL iterations back to back and hands each iteration's (c, s) sequence to warps 1-3,
    // which apply it to their rows of Q while the next recurrence is already running: two message buffers,
    // named barriers 1/2 = "buffer full", 3/4 = "buffer free".  The chain itself is kept short: one rsqrt instead
    // of sqrt + divide, and the next (d, e) entries are loaded before they are needed.
    // Buffer 0 is cs[]; buffer 1 lives in dinv[] (saved in a register meanwhile) and the unused pad column of a.
    const double dinv_keep = (tid < DP) ? dinv[tid] : 0.0;
    if (tid == 0) {
        for (int i = 1; i < D; ++i) e[i - 1] = e[i];
        e[D - 1] = 0.0;
        sh_flag = 0;
        sh_iter = 0;
    }
    __syncthreads();
    // (giving the CTAs of one SM different producer warps through per-SM tickets was measured: 7 % slower)
    constexpr int pw = 0;                               // the producer warp
    int total_iter = 0;
    // barrier ids are immediates (a register id makes ptxas reserve all 16 barriers, which costs a resident CTA)
    auto bar_sync = [](int id) {
        switch (id) {
            case 1: asm volatile("bar.sync 1, %0;" ::"n"(kQlThreads) : "memory"); break;
            case 2: asm volatile("bar.sync 2, %0;" ::"n"(kQlThreads) : "memory"); break;
            case 3: asm volatile("bar.sync 3, %0;" ::"n"(kQlThreads) : "memory"); break;
            default: asm volatile("bar.sync 4, %0;" ::"n"(kQlThreads) : "memory"); break;
        }
    };
    auto bar_arrive = [](int id) {
        switch (id) {
            case 1: asm volatile("bar.arrive 1, %0;" ::"n"(kQlThreads) : "memory"); break;
            case 2: asm volatile("bar.arrive 2, %0;" ::"n"(kQlThreads) : "memory"); break;
            case 3: asm volatile("bar.arrive 3, %0;" ::"n"(kQlThreads) : "memory"); break;
            default: asm volatile("bar.arrive 4, %0;" ::"n"(kQlThreads) : "memory"); break;
        }
    };
    // element t of buffer `buf`: c at cptr[t * cstride], s at sptr[t * sstride]
    auto cbase = [&](int buf) { return buf ? dinv : cs; };
    auto sbase = [&](int buf) { return buf ? a + DP : cs + 1; };
    auto cstr = [](int buf) { return buf ? 1 : 2; };
    auto sstr = [](int buf) { return buf ? LD : 2; };
    if (warp == pw) {
        int k = 0;                                     // message counter
        for (int l = 0; l < D; ++l) {
            for (int iter = 0;; ++iter) {
                const int buf = k & 1;
                if (k >= 2) bar_sync(3 + buf);          // the consumers are done with message k-2
                int m = l, cnt = 0;
#ifdef CMF_EIGEN_PROF
                const long long ps0 = clock64();
#endif
                // first m >= l with a negligible e[m] (m = D-1 if none): the 32 lanes test 32 entries at a time
                {
                    m = D - 1;
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
                }
                if (lane == 0) {
                    if (m != l && iter < kQlMaxIter) {
                        double* cw = cbase(buf);
                        double* sw = sbase(buf);
                        const int cst = cstr(buf), sst = sstr(buf);
                        double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                        double r = sqrt(g * g + 1.0);
                        g = d[m] - d[l] + e[l] / (g + copysign(r, g));
                        double sn = 1.0, c = 1.0, p = 0.0;
                        int i = m - 1;
                        bool under = false;
                        double e_i = e[i], d_i = d[i], d_i1 = d[i + 1];
                        for (; i >= l && !under; --i) {
                            double e_n = 0.0, d_n = 0.0;
                            if (i > l) { e_n = e[i - 1]; d_n = d[i - 1]; }      // next step's operands, early
                            const double f = sn * e_i;
                            const double b = c * e_i;
                            const double h = f * f + g * g;
                            // 1/sqrt(h) without a branch on the chain: hardware seed (2^-22) and one third-order
                            // correction, exact to rounding for the normal, positive h that occur here
                            under = !(h > 0.0);
                            double x0;
                            asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x0) : "d"(h));
                            const double err = fma(-h * x0, x0, 1.0);
                            const double rinv = under ? 0.0 : fma(fma(err, 0.375, 0.5), x0 * err, x0);
                            // r2 = (d_i - g2) sn' + 2 c' b with sn' = f rinv, c' = g rinv: everything that does not
                            // need rinv is formed beside the rsqrt, so only two levels follow it on the chain
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
                                cw[cnt * cst] = c; sw[cnt * sst] = sn;          // rotation of columns (i, i+1)
                                ++cnt;
                            } else {
                                d[i + 1] = g2;                                  // r == 0: d[i+1] -= p, e[m] = 0
                                e[m] = 0.0;
                            }
                            e_i = e_n; d_i1 = d_i; d_i = d_n;
                        }
                        if (!under) { d[l] -= p; e[l] = g; e[m] = 0.0; }
                    } else if (m != l) {
                        sh_flag = 1;       // iteration cap: give up on this eigenvalue (status NoConverge)
                        m = l;
                    }
                    sh_m2[buf] = m; sh_cnt2[buf] = cnt;
                }
#ifdef CMF_EIGEN_PROF
                if (lane == 0) pser += clock64() - ps0;
#endif
                m = __shfl_sync(0xffffffffu, m, 0);
                __threadfence_block();
                __syncwarp();
                bar_arrive(1 + buf);                    // message k is ready
                ++k;
                if (m == l) break;
                ++total_iter;
            }
        }
        // end-of-stream message
        const int buf = k & 1;
        if (k >= 2) bar_sync(3 + buf);
        if (lane == 0) { sh_m2[buf] = 0; sh_cnt2[buf] = -1; sh_iter = total_iter; }
        __threadfence_block();
        __syncwarp();
        bar_arrive(1 + buf);
    } else {
        const int row_id = ((warp - pw - 1) & (NW - 1)) * 32 + lane;   // 3 consumer warps = 96 >= D rows of Q
        for (int k = 0;; ++k) {
            const int buf = k & 1;
            bar_sync(1 + buf);
            const int m = sh_m2[buf], cnt = sh_cnt2[buf];
            if (cnt < 0) break;
            if (cnt > 0 && row_id < D) {
                // rotation t acts on columns (i, i+1), i = m-1-t
                const double* cr = cbase(buf);
                const double* sr = sbase(buf);
                const int cst = cstr(buf), sst = sstr(buf);
                double* row = a + row_id * LD;
                double f = row[m];
                for (int t = 0; t < cnt; ++t) {
                    const int i = m - 1 - t;
                    const double c = cr[t * cst], sn = sr[t * sst];
                    const double zi = row[i];
                    row[i + 1] = sn * zi + c * f;
                    f = c * zi - sn * f;
                }
                row[m - cnt] = f;
            }
            bar_arrive(3 + buf);                        // buffer free again
        }
    }
    __syncthreads();
    if (tid < DP) dinv[tid] = dinv_keep;
    __syncthreads();
    if (tid < DP) lam_g[(long long)s * DP + tid] = (tid < D) ? d[tid] : 0.0;
    if (tid == 0) {
        status_g[s] = sh_flag ? kStatusNoConverge : kStatusOk;
#ifdef CMF_EIGEN_PROF
        pt3 = clock64();
#else
        sweeps_g[s] = sh_iter;
#endif
    }
    if (MODE == kEigFull) {
        // P = L^-T V: back substitution, one eigenvector (column of a) per thread
        __syncthreads();
        for (int j = tid; j < D; j += blockDim.x) {
            for (int b = D - 1; b >= 0; --b) {
                double v = a[b * LD + j];
                for (int k = b + 1; k < D; ++k) v -= tl[k * LD + b] * a[k * LD + j];
                a[b * LD + j] = v / tl[b * LD + b];
            }
        }
        __syncthreads();
    }
    for (int idx = tid; idx < DP * DP; idx += blockDim.x) {
        const int b = idx / DP, j = idx % DP;
        Pout[idx] = (b < D && j < D) ? dinv[b] * a[b * LD + j] : 0.0;
    }
#ifdef CMF_EIGEN_PROF
    __syncthreads();
    if (tid == 0) {   // profiling build: phase cycles packed into the sweeps word instead of the iteration count
        const long long pend = clock64();
        auto q = [](long long v, int sh) { long long t = v >> sh; return (int)(t > 255 ? 255 : t); };
        (void)pt1; (void)pser;
        sweeps_g[s] = q(pt0 - pstart, 13) | (q(pt2 - pt0, 14) << 8) | (q(pt3 - pt2, 15) << 16) | (q(pend - pt3, 13) << 24);
    }
#endif
}

// ---------------------------------------------------------------------------------------- K2b
// Fragment-ordered tables for the LOO passes from (lam, P): Pf (FP64 DMMA B operand of y = xc . P),
// Wf (FP64 DMMA A operand of r = W^T z), log det G_alpha, beta, the closed-form sum_k r_k, and for the
// screening pass the TF32 hi/lo splits Ws (W) and Ps (P) in mma.m16n8k8 fragment order.
template <int NT>
__global__ void __launch_bounds__(512)
    tables_kernel(const int* __restrict__ n_g, const int* __restrict__ nloo_g, const double* __restrict__ alphas,
                  int A, int NT2, int NT16,
                  int D, int model, const double* __restrict__ P_g, const double* __restrict__ lam_g,
                  const double* __restrict__ slogT_g, double* __restrict__ Pf_g, double* __restrict__ Wf_g,
                  double* __restrict__ logdet_g, double* __restrict__ beta_g, float* __restrict__ Ws_g,
                  float* __restrict__ betaf_g, double* __restrict__ rsum_g, float* __restrict__ Ps_g) {
    constexpr int DP = 8 * NT;
    __shared__ double lam[DP];
    const int s = blockIdx.x, tid = threadIdx.x;
    const int n = n_g[s];
    const int AP = NT2 * 8, AP16 = NT16 * 16;
    const double* P = P_g + (long long)s * DP * DP;
    double* Pfout = Pf_g + (long long)s * (DP / 4) * NT * 32;
    double* Wfout = Wf_g + (long long)s * (DP / 4) * NT2 * 32;
    const long long ws_half = (long long)NT16 * NT * 32 * 4;      // floats per (column, hi|lo) table
    float* Wsout = Ws_g ? Ws_g + (long long)s * 2 * ws_half : nullptr;
    if (n < 2) {
        for (int i = tid; i < (DP / 4) * NT * 32; i += blockDim.x) Pfout[i] = 0.0;
        if (model == 0) {
            for (int i = tid; i < (DP / 4) * NT2 * 32; i += blockDim.x) Wfout[i] = 0.0;
            for (int i = tid; i < AP; i += blockDim.x) {
                logdet_g[(long long)s * AP + i] = 0.0;
                beta_g[(long long)s * AP + i] = 0.0;
                rsum_g[(long long)s * AP + i] = 0.0;
            }
        }
        return;                         // the screening and exact passes skip columns with n < 2
    }
    if (tid < DP) lam[tid] = lam_g[(long long)s * DP + tid];
    __syncthreads();
    // Pf[ks][nt][lane] = P[b = 4 ks + lane%4][j = 8 nt + lane/4]
    for (int idx = tid; idx < (DP / 4) * NT * 32; idx += blockDim.x) {
        const int lane = idx & 31, nt = (idx >> 5) % NT, ks = (idx >> 5) / NT;
        const int b = 4 * ks + (lane & 3), j = 8 * nt + (lane >> 2);
        Pfout[idx] = P[b * DP + j];
    }
    // Ps[hl][ks][nt][lane] = float2{ P[b0][j], P[b0+4][j] }, b0 = 8 ks + lane%4, j = 8 nt + lane/4
    if (Ps_g) {
        float* Psout = Ps_g + (long long)s * 2 * NT * NT * 32 * 2;
        for (int idx = tid; idx < NT * NT * 32 * 2; idx += blockDim.x) {
            const int e = idx & 1, lane = (idx >> 1) & 31, nt = (idx >> 6) % NT, ks = (idx >> 6) / NT;
            const int b = 8 * ks + (lane & 3) + 4 * e, j = 8 * nt + (lane >> 2);
            const double pv = P[b * DP + j];
            const float hi = to_tf32((float)pv);
            Psout[idx] = hi;
            Psout[NT * NT * 32 * 2 + idx] = to_tf32((float)(pv - (double)hi));
        }
    }
    if (model != 0) return;  // empirical model: no alpha search

    // ---- LOO tables.  beta_i = (1-alpha_i)/(n-1);  den_ji = n beta_i lam_j + alpha_i, where n is the count the
    // reference hands to looshrinkage: the column's valid pixels even when a cluster subset is fitted
    // (cmf/robust_mf.py:355-356), while the covariance itself is over the m = n_g pixels of the subset
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
                rs += lam[j] / den;          // sum_k r_k(alpha) = (n-1) sum_j lam_j / den_j  (exact identity)
            }
            rs *= ((double)n - 1.0);
        }
        logdet_g[(long long)s * AP + i] = ld;
        beta_g[(long long)s * AP + i] = be;
        rsum_g[(long long)s * AP + i] = rs;
    }
    // Ws[hl][at][ks][lane] = float4{ W[j0][i0], W[j0][i0+8], W[j0+1][i0], W[j0+1][i0+8] },
    //   i0 = 16 at + lane/4, j0 = 8 ks + 2 (lane%4)   (same j permutation as Wf: GEMM1 accumulators feed GEMM2)
    if (Wsout) {
        for (int i = tid; i < AP16; i += blockDim.x) {
            float bf = 0.f;
            if (i < A) bf = (float)((1.0 - alphas[i]) / (dn - 1.0));
            betaf_g[(long long)s * AP16 + i] = bf;
        }
        for (int idx = tid; idx < NT16 * NT * 32 * 4; idx += blockDim.x) {
            const int e = idx & 3, lane = (idx >> 2) & 31, ks = (idx >> 7) % NT, at = (idx >> 7) / NT;
            const int i = 16 * at + (lane >> 2) + ((e & 1) ? 8 : 0);
            const int j = 8 * ks + 2 * (lane & 3) + ((e & 2) ? 1 : 0);
            double w = 0.0;
            if (j < D && i < A) {
                const double al = alphas[i];
                const double be = (1.0 - al) / (dn - 1.0);
                w = 1.0 / (dn * be * lam[j] + al);
            }
            const float hi = to_tf32((float)w);
            const float lo = to_tf32((float)(w - (double)hi));
            Wsout[idx] = hi;
            Wsout[ws_half + idx] = lo;
        }
    }
    // Wf[ks2][at][lane] = W[j = 8 (ks2/2) + 2 (lane%4) + ks2%2][i = 8 at + lane/4]
    // (the j permutation lets the y^2 accumulator registers of GEMM1 feed GEMM2 without any shuffle)
    for (int idx = tid; idx < (DP / 4) * NT2 * 32; idx += blockDim.x) {
        const int lane = idx & 31, at = (idx >> 5) % NT2, ks2 = (idx >> 5) / NT2;
        const int j = 8 * (ks2 >> 1) + 2 * (lane & 3) + (ks2 & 1);
        const int i = 8 * at + (lane >> 2);
        double w = 0.0;
        if (j < D && i < A) {
            const double al = alphas[i];
            const double be = (1.0 - al) / (dn - 1.0);
            w = 1.0 / (dn * be * lam[j] + al);
        }
        Wfout[idx] = w;
    }
}

// ---------------------------------------------------------------------------------------- K4
__global__ void __launch_bounds__(256)
    finalize_kernel(const double* __restrict__ fpart, int nchunk, const double* __restrict__ logdet_g,
                    const int* __restrict__ n_g, const double* __restrict__ alphas, int A, int AP, int D,
                    int DP, int S, const double* __restrict__ P_g, const double* __restrict__ lam_g,
                    const double* __restrict__ mu_g, const double* __restrict__ abscf, int model,
                    int reflectance, double scale, double* __restrict__ nll_g, int* __restrict__ mindex_g,
                    double* __restrict__ w_g, double* __restrict__ wT_g, double* __restrict__ c0_g,
                    int* __restrict__ status_g, const int* __restrict__ sel_index,
                    const unsigned long long* __restrict__ tile_mask, const int* __restrict__ nloo_g,
                    const int* __restrict__ probe_g, const double* __restrict__ tol_g, double* __restrict__ check_g,
                    unsigned long long* __restrict__ redo_g, const unsigned long long* __restrict__ only_g) {
    extern __shared__ double sm[];
    double* nll = sm;           // [AP]
    double* tvec = nll + AP;    // [DP]
    double* uvec = tvec + DP;   // [DP]
    double* vvec = uvec + DP;   // [DP]
    __shared__ double alpha_sel, norm_sh, c0_sh;
    __shared__ int singular;

    const int s = blockIdx.x, tid = threadIdx.x;
    const int n = n_g[s];
    const int Sp = (S + 1) & ~1;
    // second pass of the screening certificate: only the columns that are re-evaluated exactly
    if (only_g != nullptr && only_g[s] == 0ull) return;
    __shared__ double err_sh[16];
    if (tid == 0 && check_g) { check_g[s] = 0.0; redo_g[s] = 0ull; }   // overwritten below for a refined column
    // background-mode pass without members in this column (or past the end of its mode list): the column
    // keeps what earlier passes produced; the scoring pass has no member to write either
    if (nloo_g != nullptr && n == 0) return;
    const double nl = (double)(nloo_g ? nloo_g[s] : n);
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    const double inf = __longlong_as_double(0x7ff0000000000000LL);

    if (n < 2) {
        // n == 0: column skipped by the reference (:303-304).  n == 1: cov() is NaN, every nll is NaN,
        // argmin returns 0 and the scores are NaN (pinned by tests/golden/degenerate_300x4).
        const double fill = (n == 0) ? 0.0 : qnan;
        for (int b = tid; b < DP; b += blockDim.x) {
            w_g[(long long)s * DP + b] = fill;
            wT_g[(long long)b * Sp + s] = fill;
        }
        for (int i = tid; i < A; i += blockDim.x) nll_g[(long long)s * A + i] = (n == 0) ? inf : qnan;
        if (tid == 0) { mindex_g[s] = (n == 0) ? -2 : 0; c0_g[s] = fill; }
        return;
    }

    if (tid == 0) singular = (status_g[s] & kStatusSingular) ? 1 : 0;   // set by K2 when -f finds T not positive definite
    if (model == 0) {
        // screened run (K3a/K3b): sel >= 0 index decided by the screen, -1 all inf, -2 exact values of the
        // tiles in `tmask` decide; alphas outside the mask keep the screened nll (never the minimum).
        const bool screened = sel_index != nullptr;
        const int sel = screened ? sel_index[s] : -2;
        const unsigned long long tmask = screened ? tile_mask[s] : ~0ull;
        const double const_term = (double)D * log(2.0 * M_PI);
        // certificate: spread of (exact - screened) over the alphas evaluated exactly.  Only DIFFERENCES between
        // alphas decide the argmin, so a bias of the screen that is common to the alphas of a column cancels; what can
        // misorder two alphas is how much the screening error varies between them: max - min of (exact - screened).
        const double big = __longlong_as_double(0x7ff0000000000000LL);
        double dmax = -big, dmin = big;
        for (int i = tid; i < A; i += blockDim.x) {
            double v;
            if (!screened || ((tmask >> ((i >> 3) & 63)) & 1ull)) {
                double fs = 0.0;
                for (int c = 0; c < nchunk; ++c) fs += fpart[((long long)s * nchunk + c) * AP + i];
                const double ld = det_roundtrip(logdet_g[(long long)s * AP + i]);
                if (!(fabs(ld) < inf)) v = inf;   // det underflows to 0 -> alpha skipped; overflows -> log(inf) (:112-113)
                else v = 0.5 * (const_term + ld) + fs / (2.0 * nl);
                if (screened && check_g) {
                    const double old = nll_g[(long long)s * A + i];
                    if (fabs(v) < inf && fabs(old) < inf) { dmax = fmax(dmax, v - old); dmin = fmin(dmin, v - old); }
                }
                nll_g[(long long)s * A + i] = v;
            } else {
                v = nll_g[(long long)s * A + i];
            }
            nll[i] = v;
        }
        if (check_g) {
            for (int o = 16; o > 0; o >>= 1) {
                dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
                dmin = fmin(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
            }
            if ((tid & 31) == 0) { err_sh[tid >> 5] = dmax; err_sh[8 + (tid >> 5)] = dmin; }
        }
        __syncthreads();
        if (tid == 0) {
            int best;
            double al;
            if (screened && sel >= 0) {
                best = sel; al = alphas[best];
            } else if (screened && sel == -1) {
                best = -1; al = 0.0; status_g[s] |= kStatusAllInf;
            } else {
                // numpy argmin: the first NaN wins, otherwise the first minimum (:121)
                best = -1;
                bool have_nan = false;
                for (int i = 0; i < A; ++i) {
                    if (screened && !((tmask >> ((i >> 3) & 63)) & 1ull)) continue;
                    const double v = nll[i];
                    if (v != v) { best = i; have_nan = true; break; }
                    if (best < 0 || v < nll[best]) best = i;
                }
                if (best >= 0 && (have_nan || nll[best] != inf)) al = alphas[best];
                else { best = -1; al = 0.0; status_g[s] |= kStatusAllInf; }
            }
            mindex_g[s] = best;
            alpha_sel = al;
            if (check_g) {
                // Runtime certificate, part 2: this column is trusted if the exact minimum does not sit in the probe
                // tile (the best tile the screen had EXCLUDED) and the measured screening error is below
                // 1/kCertFactor of the margin; otherwise every alpha of the column is re-evaluated in FP64.
                double hi = -inf, lo = inf;
                for (int w8 = 0; w8 < (int)(blockDim.x >> 5); ++w8) { hi = fmax(hi, err_sh[w8]); lo = fmin(lo, err_sh[8 + w8]); }
                const double e = (hi >= lo) ? hi - lo : 0.0;        // spread of the screening error over the evaluated alphas
                double frac = 0.0;
                unsigned long long redo = 0ull;
                if (screened && sel == -2 && tmask != ~0ull) {
                    const double margin = tol_g[s];
                    frac = margin > 0.0 ? e / margin : 0.0;
                    const int probe = probe_g ? probe_g[s] : -1;
                    if (frac * kCertFactor > 1.0 || (best >= 0 && probe >= 0 && (best >> 3) == probe)) redo = ~0ull;
                }
                check_g[s] = frac;
                redo_g[s] = redo;
            } else if (only_g != nullptr) {
                status_g[s] |= kStatusRechecked;
            }
        }
    } else {
        if (tid == 0) { mindex_g[s] = -2; alpha_sel = 0.0; }
    }
    __syncthreads();
    const double al = alpha_sel;
    const double* P = P_g + (long long)s * DP * DP;
    const double* lam = lam_g + (long long)s * DP;
    const double* mu = mu_g + (long long)s * DP;
    for (int b = tid; b < DP; b += blockDim.x)
        tvec[b] = (b < D) ? (reflectance ? abscf[b] - mu[b] : abscf[b] * mu[b]) : 0.0;
    __syncthreads();
    for (int j = tid; j < DP; j += blockDim.x) {
        double a = 0.0;
        if (j < D) {
            for (int b = 0; b < D; ++b) a += P[b * DP + j] * tvec[b];
            const double den = (1.0 - al) * lam[j] + al;
            if (!(den > 0.0)) singular = 1;
            a /= den;
        }
        uvec[j] = a;
    }
    __syncthreads();
    for (int b = tid; b < DP; b += blockDim.x) {
        double a = 0.0;
        if (b < D)
            for (int j = 0; j < D; ++j) a += P[b * DP + j] * uvec[j];
        vvec[b] = a;
    }
    __syncthreads();
    if (tid == 0) {
        double nrm = 0.0;
        for (int b = 0; b < D; ++b) nrm += tvec[b] * vvec[b];
        norm_sh = nrm;
    }
    __syncthreads();
    const bool sing = singular != 0;
    for (int b = tid; b < DP; b += blockDim.x) {
        const double w = sing ? 0.0 : vvec[b] / norm_sh * scale;
        vvec[b] = w;
        w_g[(long long)s * DP + b] = w;
        wT_g[(long long)b * Sp + s] = w;
    }
    __syncthreads();
    if (tid == 0) {
        double c = 0.0;
        for (int b = 0; b < D; ++b) c += mu[b] * vvec[b];
        c0_g[s] = c;
        if (sing) status_g[s] |= kStatusSingular;
    }
}

// ---------------------------------------------------------------------------------------- launchers
template <int NT>
static void launch_eigen_t(const Dims& d, const double* gram_part, int nchunk, const int* n, const double* mu,
                           const double* ctr, double* P, double* lam, double* slogT, int* status, int* sweeps,
                           int method, int target, const double* gramT_part, const int* nT, cudaStream_t st) {
    constexpr int DP = 8 * NT, LD = DP + 1;
    if (target == kEigNone) {
        const size_t smem = (size_t)(DP * LD + 5 * DP + 8) * sizeof(double);
        cudaFuncSetAttribute(eigen_ql_kernel<NT, kEigNone>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        eigen_ql_kernel<NT, kEigNone><<<d.S, kQlThreads, smem, st>>>(gram_part, nchunk, n, d.D, mu, ctr, P, lam, slogT,
                                                                    status, sweeps, nullptr, nullptr);
    } else if (target == kEigFull) {
        const size_t smem = (size_t)(2 * DP * LD + 5 * DP + 8) * sizeof(double);
        cudaFuncSetAttribute(eigen_ql_kernel<NT, kEigFull>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        eigen_ql_kernel<NT, kEigFull><<<d.S, kQlThreads, smem, st>>>(gram_part, nchunk, n, d.D, mu, ctr, P, lam, slogT,
                                                                    status, sweeps, gramT_part, nT);
#ifdef CMF_TUNING_HOOKS
    } else if (method == 1) {
        const size_t smem = (size_t)(2 * DP * LD + 2 * DP + 2 * (DP / 2 + 1)) * sizeof(double);
        cudaFuncSetAttribute(eigen_kernel<NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        eigen_kernel<NT><<<d.S, 512, smem, st>>>(gram_part, nchunk, n, d.D, mu, ctr, P, lam, slogT, status, sweeps);
#endif
    } else {
        const size_t smem = (size_t)(DP * LD + 5 * DP + 8) * sizeof(double);
        cudaFuncSetAttribute(eigen_ql_kernel<NT, kEigDiag>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        eigen_ql_kernel<NT, kEigDiag><<<d.S, kQlThreads, smem, st>>>(gram_part, nchunk, n, d.D, mu, ctr, P, lam, slogT,
                                                                    status, sweeps, nullptr, nullptr);
    }
}

// method: 0 = Householder + implicit QL (default), 1 = cyclic Jacobi (cross-check)
// target: 0 = diag(S) (default shrinkage target), 1 = none (PCA), 2 = full-column covariance from gramT_part / nT (-f)
void launch_eigen(const Dims& d, const double* gram_part, int nchunk, const int* n, const double* mu,
                  const double* ctr, double* P, double* lam, double* slogT, int* status, int* sweeps, int method,
                  cudaStream_t st, int target, const double* gramT_part, const int* nT) {
    switch (d.NT) {
#define CMF_CASE(k) \
    case k: launch_eigen_t<k>(d, gram_part, nchunk, n, mu, ctr, P, lam, slogT, status, sweeps, method, target, \
                              gramT_part, nT, st); break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6)
        CMF_CASE(7) CMF_CASE(8) CMF_CASE(9) CMF_CASE(10) CMF_CASE(11) CMF_CASE(12)
#undef CMF_CASE
        default: break;
    }
}

void launch_tables(const Dims& d, const int* n, const int* nloo, const double* alphas, int model, const double* P,
                   const double* lam, const double* slogT, double* Pf, double* Wf, double* logdet, double* beta,
                   float* Ws, float* betaf, double* rsum, float* Ps, cudaStream_t st) {
    switch (d.NT) {
#define CMF_CASE(k)                                                                                           \
    case k:                                                                                                   \
        tables_kernel<k><<<d.S, 512, 0, st>>>(n, nloo, alphas, d.A, d.NT2, d.NT16, d.D, model, P, lam, slogT, Pf, Wf, \
                                              logdet, beta, Ws, betaf, rsum, Ps);                             \
        break;
        CMF_CASE(1) CMF_CASE(2) CMF_CASE(3) CMF_CASE(4) CMF_CASE(5) CMF_CASE(6)
        CMF_CASE(7) CMF_CASE(8) CMF_CASE(9) CMF_CASE(10) CMF_CASE(11) CMF_CASE(12)
#undef CMF_CASE
        default: break;
    }
}

void launch_finalize(const Dims& d, const double* fpart, int nchunk, const double* logdet, const int* n,
                     const double* alphas, const double* P, const double* lam, const double* mu,
                     const double* abscf, int model, int reflectance, double scale, double* nll, int* mindex,
                     double* w, double* wT, double* c0, int* status, const int* sel_index,
                     const unsigned long long* tile_mask, const int* nloo, cudaStream_t st, const int* probe,
                     const double* tol_col, double* check, unsigned long long* redo, const unsigned long long* only) {
    const size_t smem = (size_t)(d.AP + 3 * d.DP) * sizeof(double);
    finalize_kernel<<<d.S, 256, smem, st>>>(fpart, nchunk, logdet, n, alphas, d.A, d.AP, d.D, d.DP, d.S, P,
                                            lam, mu, abscf, model, reflectance, scale, nll, mindex, w, wT, c0,
                                            status, sel_index, tile_mask, nloo, probe, tol_col, check, redo, only);
}

}  // namespace cmf
