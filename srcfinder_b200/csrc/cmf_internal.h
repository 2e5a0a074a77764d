// Internal launch interface between the C-ABI host layer (cmf_api.cu) and the kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

namespace cmf {

// Tuning / cross-check hooks read from the environment exist only in the TOOLS build of the library
// (srcfinder_b200/build.py: libcmf_b200_tools.so, -DCMF_TUNING_HOOKS); the shipped library ignores the environment.
#ifdef CMF_TUNING_HOOKS
inline const char* cmf_hook(const char* name) { return getenv(name); }
#else
inline const char* cmf_hook(const char*) { return nullptr; }
#endif

constexpr int kMaxNT = 12;          // active bands padded to 8*NT, NT <= 12 (D <= 96) for the shared-memory kernels
constexpr int kGramTL = 16;         // lines per Gram tile (4 k-steps of DMMA.8x8x4)
constexpr int kLooMT = 2;           // 8-pixel m-tiles per LOO pass
constexpr int kScoreLines = 8;      // lines per thread in the scoring pass (scalar kernel)
constexpr int kMaxLabels = 32;      // background-mode labels are 0 .. kMaxLabels-1
constexpr int kModeNone = 127;      // mode-list slot past the end of a column's list
constexpr int kMaxPcaDim = 16;      // --pcadim upper bound of the on-device partition
constexpr int kScoreTiledMaxD = 192; // the tiled scoring kernel keeps D x 64 weight pairs in shared memory

// column status bits (per cross-track column)
enum : int {
    kStatusOk = 0,
    kStatusEmpty = 1,        // no valid pixel: MF stays nodata (cmf/robust_mf.py:303-304)
    kStatusDegenerate = 2,   // n == 1: the reference produces NaN scores and alpha index 0
    kStatusSingular = 4,     // C not invertible: MF := 0 (cmf/robust_mf.py:371-374)
    kStatusNoConverge = 8,   // Jacobi hit the sweep cap (results still written)
    kStatusAllInf = 16,      // every nll was inf: alpha := 0, index -1 (cmf/robust_mf.py:123-127)
    kStatusRechecked = 32    // the screening certificate failed: every alpha was re-evaluated in FP64
};

struct Dims {
    int L, S, D, NT, DP;     // lines, samples, active bands, tiles, padded bands
    int A, NT2, AP;          // alphas, 8-alpha tiles, padded alphas
    int NT16, AP16;          // 16-alpha tiles of the screening pass, padded alphas
    long long line_pitch;    // elements between consecutive lines of the active slab
    int band_pitch;          // elements between consecutive bands (== samples of the cube)
    int vec2;                // 8-byte loads allowed (S, pitches even and base aligned)
    // background-mode passes (-k > 1): the member pixels of a column are COMPACTED to the first rows of its xt block
    // (rowidx[l*S + s] = row of pixel (l, s), in line order) and the statistics / search kernels stop at the column's
    // member count nrows[s]; NULL = every line is a row (unimodal path)
    const int32_t* rowidx;
    const int* nrows;
};

// per-pixel spectrometer flags (spectrometer_masks/masks_sds.py:133-233); band indices are 0-based positions
// in the buffer handed to the kernel, < 0 disables a test
enum : uint8_t { kFlagSaturated = 1, kFlagSpecular = 2, kFlagDark = 4, kFlagCloud = 8 };
struct FlagSpec {
    int sat_lo, sat_hi, spec_band, dark_band, cloud_a, cloud_b;
    float sat_thresh, spec_thresh, dark_thresh, cloud_thresh, cloud_dwl;
};
// exclusion support: sel = (exclude == 0), per-column count of the valid-pixel mask
void launch_invert_u8(uint8_t* p, long long n, cudaStream_t st);
void launch_count_mask(const Dims& d, const uint8_t* mask, int* count, cudaStream_t st);
void launch_pixel_flags(const float* cube, long long line_pitch, int band_pitch, int L, int S, const FlagSpec& f,
                        uint8_t* flags, cudaStream_t st);
// column profiles of a score image (triage/cmf_profile.py:110-140); colv is [S][L] float scratch, out [5][S]
void launch_column_profile(const double* mf, int L, int S, double nodata, int robust, double qlo, double qhi,
                           float* colv, double* out, cudaStream_t st);

// detection pre-filter (srcfinder_util.py:1383-1387, 1428-1436) and CNN input normalisation (cnn_pred_pipeline.py:19-30)
void launch_detection_prefilter(const double* mf, int L, int S, int radius, const double* w_dev, double mfmin,
                                double mfmax, double* tmp, double* blur, double* part, double* detkde,
                                uint8_t* ch4min, uint8_t* detmask, cudaStream_t st);
void launch_cnn_input(const double* mf64, const float* mf32, long long n, float vmin, float vmax, float mean,
                      float stdv, float* out, cudaStream_t st);

inline int ntri(int nt) { return nt * (nt + 1) / 2; }

void launch_repack(const Dims& d, const float* slab, float* xt, uint8_t* mask, double* colsum_part,
                   int* colcnt_part, int lines_per_split, int line_base, int line_limit, const uint8_t* sel,
                   int write_mask, cudaStream_t st);
int repack_lines_per_split(const Dims& d, int nsplit);
void launch_mean(const Dims& d, const double* colsum_part, const int* colcnt_part, int nsplit,
                 double* mu, int* n, cudaStream_t st);
void launch_gram(const Dims& d, const float* xt, const double* ctr, int nchunk, int lpc, int chunk_lo,
                 int chunk_hi, double* gram_part, cudaStream_t st);
// ctr != NULL: the Gram partials are centred on ctr, the rank-one term n (mu-ctr)(mu-ctr)^T is removed here
void launch_eigen(const Dims& d, const double* gram_part, int nchunk, const int* n, const double* mu,
                  const double* ctr, double* P, double* lam, double* slogT, int* status, int* sweeps, int method,
                  cudaStream_t st, int target = 0, const double* gramT_part = nullptr, const int* nT = nullptr);
void launch_tables(const Dims& d, const int* n, const int* nloo, const double* alphas, int model, const double* P,
                   const double* lam, const double* slogT, double* Pf, double* Wf, double* logdet, double* beta,
                   float* Ws, float* betaf, double* rsum, float* Ps, cudaStream_t st);
void launch_loo(const Dims& d, const float* xt, const double* mu, const double* Pf, const double* Wf,
                const double* beta, int nchunk, double* fpart, const unsigned long long* tile_mask,
                cudaStream_t st);
void launch_screen(const Dims& d, const float* xt, const double* mu, const double* Pf, const float* Ps,
                   const float* Ws, const float* betaf, const int* n, int nchunk, double* fscreen,
                   cudaStream_t st);
size_t screen_smem_bytes(const Dims& d);
// tcgen05 / TMEM form of the screening pass (k_screen5.cu); falls back to launch_screen when unsupported
bool screen5_supported(const Dims& d);
size_t screen5_table_floats(const Dims& d);
int screen5_pick_chunks(const Dims& d, int sm_count);
void launch_screen5(const Dims& d, const float* xt, const double* mu, const int* n, const int* nloo,
                    const double* alphas, const double* P, const double* lam, float* tab, float* betaf,
                    int nchunk, double* fscreen, cudaStream_t st);
#ifdef CMF_TUNING_HOOKS
double screen5_selftest(int N, int K, int row_off, int swap_lbo_sbo);
int screen5_timeline(long long* out);
double screen5_mma_rate(int N, int reps);
double fma_issue_rate(int packed, int warps);
#endif
void launch_select(const Dims& d, const double* fscreen, int nchunk, const double* logdet, const double* rsum,
                   const int* n, const int* nloo, double tol, double* nll, int* sel_index,
                   unsigned long long* tile_mask, int* ncand, double* tol_out, const float* betaf_fold,
                   int* probe, cudaStream_t st);
// runtime certificate of the screened alpha search: measured screening error must stay below 1/kCertFactor of the margin
constexpr double kCertFactor = 4.0;
void launch_certify(const Dims& d, const double* check, const int* sel_index, int enabled, unsigned long long* redo,
                    double* worst, cudaStream_t st);
void launch_finalize(const Dims& d, const double* fpart, int nchunk, const double* logdet, const int* n,
                     const double* alphas, const double* P, const double* lam, const double* mu,
                     const double* abscf, int model, int reflectance, double scale, double* nll,
                     int* mindex, double* w, double* wT, double* c0, int* status, const int* sel_index,
                     const unsigned long long* tile_mask, const int* nloo, cudaStream_t st,
                     const int* probe = nullptr, const double* tol_col = nullptr, double* check = nullptr,
                     unsigned long long* redo = nullptr, const unsigned long long* only = nullptr);
void launch_score(const Dims& d, const float* slab, const uint8_t* mask, const double* wT, const double* c0,
                  const int* status, double nodata, double* mf, double* stat_part, int nlanes,
                  int lines_per_cta, const uint8_t* sel, const int* mindex, int16_t* alpha_img, cudaStream_t st);
void launch_modes(const Dims& d, const int32_t* labels, const uint8_t* mask, int reject_min, int8_t* entries,
                  uint32_t* rejmask, int* nentries, uint32_t* flagmask, cudaStream_t st);
void launch_members(const Dims& d, const int32_t* labels, const uint8_t* mask, int t, const int8_t* entries,
                    const uint32_t* rejmask, const uint32_t* flagmask, uint8_t* sel, int16_t* cluster_img, uint8_t* inlier,
                    cudaStream_t st);
void launch_fill_f64(double* p, long long n, double v, cudaStream_t st);
// rowidx[l*S + s] = number of selected pixels of column s above line l (exclusive scan down the lines)
void launch_rank(const Dims& d, const uint8_t* sel, int32_t* rowidx, cudaStream_t st);
// member rows of the full column-major copy -> compacted copy + the K0 partial sums / counts of the members
bool launch_compact(const Dims& d, const float* xt_full, const uint8_t* sel, const int32_t* rowidx, int nsplit,
                    int lines_per_split, float* xt_mode, double* colsum_part, int* colcnt_part, cudaStream_t st);
// PCA projection (P, lam = eigenvectors / eigenvalues of the column covariance, launch_eigen target 1) + k-means
void launch_pca_kmeans(const Dims& d, const float* xt, const uint8_t* mask, const double* mu, const int* n,
                       const double* P, const double* lam, int pcadim, int k, int max_iter, int* pick,
                       double* vtop, double* y, int32_t* q, uint8_t* lab8, int32_t* labels, int* iters,
                       cudaStream_t st);
void launch_colstats_modes(const Dims& d, const double* mf, const uint8_t* inlier, const int* nuse,
                           double nodata, double* colstats, cudaStream_t st);
int score_plan(const Dims& d, int sm_count, int* lines_per_cta);
void launch_colstats(const Dims& d, const double* stat_part, int nlanes, const int* n, double nodata,
                     double* colstats, cudaStream_t st);

// ---- wide-window kernel set (k_wide.cu, k_gram8.cu): active windows wider than 8 * kMaxNT bands
void launch_wide_stats(const Dims& d, const float* slab, uint8_t* mask, const uint8_t* sel, int write_mask,
                       int nsplit, int lps, double* colsum_part, int* colcnt_part, float* lo_part, float* hi_part,
                       double* mu, int* n, double* ctr, int* qexp, cudaStream_t st);
void launch_wide_pack(const Dims& d, const float* slab, const uint8_t* mask, const uint8_t* sel, const double* ctr,
                      const int* qexp, float* xt, int8_t* img, cudaStream_t st);
void launch_wide_gram64(const Dims& d, const float* xt, const double* ctr, double* gram, cudaStream_t st);
void launch_wide_gram64_f64(int L, int DP, int S, const double* x, const double* ctr, double* gram, cudaStream_t st);
size_t wide_img_bytes(const Dims& d);
void launch_wide_gram8(const Dims& d, const int8_t* img, double* gram, cudaStream_t st);
size_t wide_rot_cap(const Dims& d);
int wide_iter_cap(const Dims& d);
bool wide_eigen_fits(const Dims& d);
// the -f target of a wide window: W = T^-1/2 in spectral form (and its transpose), log det of the scaled target,
// status of the target, one DP x DP scratch matrix per column
struct WideTarget {
    double *W = nullptr, *Wt = nullptr, *tmp = nullptr, *slogT = nullptr;
    int* status = nullptr;
};
void launch_wide_target(const Dims& d, const double* P, const double* lam, const double* slogT, const int* status,
                        const WideTarget& t, cudaStream_t st);
void launch_wide_eigen(const Dims& d, const double* gram, const int* n, const double* mu, const double* ctr,
                       const int* qexp, int mode, double* work, double* dinv, double* dvec, double* evec,
                       double2* rot, int2* iters, int* niter, double* P, double* lam, double* slogT, int* status,
                       cudaStream_t st, const WideTarget* tgt = nullptr);
void launch_wide_tables(const Dims& d, int APW, const int* n, const int* nloo, const double* alphas, int model,
                        const double* lam, const double* slogT, double* logdet, double* beta, double* rsum, double* W,
                        cudaStream_t st);
void launch_wide_loo(const Dims& d, int APW, const float* xt, const double* mu, const double* P, const double* W,
                     const double* beta, const int* n, int s0, int ns, int nchunk, double* Z, double* fpart,
                     cudaStream_t st);
void launch_wide_loo_f64(int L, int D, int DP, int AP, int APW, const double* x, const double* zero_mu,
                         const double* P, const double* W, const double* beta, const int* n, double* Z,
                         double* fpart, cudaStream_t st);

void launch_wide_mean64(const double* x, int rows, int D, int DP, double* mean, cudaStream_t st);
void launch_wide_cmat(const double* gram, int m, int D, int DP, const int* mindex, const double* alphas, double* C,
                      cudaStream_t st, const double* gram_reg = nullptr, int m_reg = 0);

size_t gram_part_elems(const Dims& d, int nchunk);
int repack_nsplit(const Dims& d);
int pick_chunks(int S, int L, int min_lines, int sm_count, int ctas_per_sm);

}  // namespace cmf
