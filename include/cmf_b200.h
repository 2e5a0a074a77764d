/* cmf_b200.h -- C ABI of the B200-native columnwise matched filter (libcmf_b200.so).
 *
 * The reference (dsmbgu8/srcfinder, cmf/robust_mf.py) exposes no FFI: its boundary is the CLI
 * (robust_mf.py:142-166), the ENVI products it writes (:210-279, :383-403) and the importable
 * looshrinkage(I_zm, alphas, nll, n, I_reg) (:92-136).  This header is the boundary a maintainer
 * binds instead of the body of the column loop (:297-397); INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: every function returns 0 on success or a negative CMF_E_* code and never throws;
 * cmf_last_error() gives the message.  The caller owns host buffers, the context owns device
 * buffers.  One context per (GPU, stream); a context is not thread-safe, independent contexts are.
 * All entry points are asynchronous on the context's stream unless stated; cmf_sync() waits.
 * There is no CPU fallback: without a CUDA device cmf_create() fails with CMF_E_CUDA.
 */
#ifndef CMF_B200_H
#define CMF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cmf_ctx cmf_ctx;

enum {
    CMF_OK = 0,
    CMF_E_ARG = -1,      /* bad argument / unsupported configuration */
    CMF_E_CUDA = -2,     /* CUDA runtime error (message has the detail) */
    CMF_E_STATE = -3,    /* call sequence error (e.g. run before set_problem / upload) */
    CMF_E_NOMEM = -4
};

enum { CMF_MODEL_LOOSHRINKAGE = 0, CMF_MODEL_EMPIRICAL = 1 };      /* -M, robust_mf.py:159-160 */
enum { CMF_INTERLEAVE_BIL = 0 };                                  /* the reference requires BIL, :208 */

/* Per-column status bits returned by CMF_OUT_STATUS. */
enum {
    CMF_COL_OK = 0,
    CMF_COL_EMPTY = 1,        /* no valid pixel; column skipped (:303-304) */
    CMF_COL_DEGENERATE = 2,   /* a single valid pixel: NaN scores, alpha index 0 (reference behaviour) */
    CMF_COL_SINGULAR = 4,     /* C not invertible: scores := 0 ("singular matrix", :371-374) */
    CMF_COL_NOCONVERGE = 8,   /* eigen-solver hit its sweep cap */
    CMF_COL_ALLINF = 16,      /* every nll inf: alpha := 0, index -1 (:123-127) */
    CMF_COL_RECHECKED = 32    /* the screening certificate was not met: every alpha re-evaluated in FP64 */
};

/* What cmf_download()/cmf_device_ptr() can return.  Shapes use L lines, S samples, D active bands,
 * A alphas, DP = D rounded up to a multiple of 8. */
enum {
    CMF_OUT_MF = 0,           /* double [L][S]   matched-filter score (ppm*m), nodata where masked (:266,:383-386) */
    CMF_OUT_MASK = 1,         /* uint8  [L][S]   1 = pixel used (finite and >= 0 in every active band, :282) */
    CMF_OUT_COLSTATS = 2,     /* double [3][S]   npix, mean, std(ddof=0) of the written scores (:388-391) */
    CMF_OUT_ALPHA_INDEX = 3,  /* int32  [S]      argmin index, -1 all-inf, -2 not applicable (:121-127) */
    CMF_OUT_NLL = 4,          /* double [S][A]   leave-one-out negative log likelihood (:117): exact FP64 under
                                                  CMF_RUN_EXACT; otherwise exact for the alphas near the minimum and
                                                  the screened value (error ~1e-7) elsewhere */
    CMF_OUT_MU = 5,           /* double [S][DP]  column mean over valid pixels (:347) */
    CMF_OUT_WEIGHTS = 6,      /* double [S][DP]  Cinv t / (t Cinv t) * scale (:380-384) */
    CMF_OUT_STATUS = 7,       /* int32  [S]      CMF_COL_* bits */
    CMF_OUT_NVALID = 8,       /* int32  [S]      valid pixels per column (nuse, :302) */
    CMF_OUT_EIGVALS = 9,      /* double [S][DP]  eigenvalues of the column correlation matrix */
    CMF_OUT_SWEEPS = 10,      /* int32  [S]      Jacobi sweeps used */
    CMF_OUT_NCAND = 11,       /* int32  [S]      alphas the screening pass could not separate (1 = decided by the
                                                  screen, > 1 = decided by exact FP64 re-evaluation); screened runs only */
    CMF_OUT_SCREEN_TOL = 12,  /* double [S]      nll margin within which the screen treats alphas as tied; the
                                                  selection is exact while the screening error stays below half of it */
    CMF_OUT_CLUSTER_ID = 13,  /* int16  [L][S]   _bgmeta band 0 (:327): cluster label, negated when rejected; labelled runs */
    CMF_OUT_ALPHA_IMAGE = 14, /* int16  [L][S]   _bgmeta band 1 (:365): alpha index of the mode that scored the pixel */
    CMF_OUT_MODE_LIST = 15,   /* int8   [S][32]  the column's mode list (bgulab, :313-332); 127 = past the end */
    CMF_OUT_LABELS = 16,      /* int32  [L][S]   cluster labels in use (given by cmf_set_labels or found by cmf_set_clustering) */
    CMF_OUT_PCA = 17,         /* double [S][L][pcadim] projections the on-device k-means partitioned (:311) */
    CMF_OUT_KMEANS_ITERS = 18,/* int32  [S]      reassignment passes the k-means needed */
    CMF_OUT_FLAGS = 19,       /* uint8  [lines][samples] of the last cmf_pixel_flags() call */
    CMF_OUT_SCREEN_CHECK = 20 /* double [S]      screened runs: spread (max - min) of exact - screened nll over the alphas
                                                  that were re-evaluated exactly, as a fraction of the margin (0 = decided by
                                                  the screen alone or fully exact); the run re-evaluates every alpha of a
                                                  column in FP64 when this reaches 1/4, see cmf_set_screen_margin */
};

typedef struct cmf_problem {
    int32_t lines, bands, samples;   /* cube shape (L, B, S), ENVI BIL: element (l,b,s) at (l*B+b)*S+s */
    int32_t interleave;              /* CMF_INTERLEAVE_BIL */
    int32_t band_lo, band_hi;        /* 1-based inclusive active window, e.g. 351..422 (:186-194) */
    int32_t reflectance;             /* -R: target = abscf - mu, no ppm scaling (:379,:383) */
    int32_t model;                   /* CMF_MODEL_* */
    int32_t num_alphas;              /* 201 in the reference (:241-243); ignored for EMPIRICAL */
    int32_t reserved;
    double nodata;                   /* 'data ignore value', must be <= 0 (:232-234) */
    const double* alphas;            /* [num_alphas] host */
    const double* abscf;             /* [band_hi-band_lo+1] host: library column 3 over the window (:237-238) */
} cmf_problem;

/* ---- lifetime ---- */
int cmf_create(cmf_ctx** out, int device);
void cmf_destroy(cmf_ctx* ctx);
const char* cmf_last_error(const cmf_ctx* ctx);   /* ctx may be NULL: error of the last failed cmf_create */
const char* cmf_version(void);

/* Use an existing CUDA stream (cudaStream_t as void*) instead of the context's own. */
int cmf_set_stream(cmf_ctx* ctx, void* cuda_stream);

/* ---- problem + input ---- */
int cmf_set_problem(cmf_ctx* ctx, const cmf_problem* p);
/* Host cube (L,B,S) float32 BIL -> device: only the active band window is transferred
 * (one strided H2D copy; pinned memory makes it asynchronous).  Replaces the memmap gather at :298. */
int cmf_upload_bil(cmf_ctx* ctx, const float* host_cube);
/* The same in blocks of lines, for inputs that are read from disk while they are uploaded:
 * host_block points at element (line line0, band band_first, sample 0) of a BIL block of `nlines` lines with
 * `block_bands` bands per line (the whole band axis: band_first = 1, block_bands = bands; or only the active window:
 * band_first = band_lo, block_bands = band_hi - band_lo + 1).  The copy is asynchronous on the context stream when
 * the block is pinned (cmf_host_alloc): the caller fills the next block while this one is on the PCIe link, and must
 * cmf_sync() before overwriting a block it has handed in.  The input is complete once every line has been given. */
int cmf_upload_lines(cmf_ctx* ctx, const float* host_block, int32_t line0, int32_t nlines, int32_t band_first,
                     int32_t block_bands);
/* Wait until the copy of the most recent cmf_upload_lines() call that read from `host_block` has finished (an event
 * recorded behind that copy), without waiting for copies from other staging blocks that were enqueued later: this is
 * what lets a reader refill block A while block B is still on the PCIe link.  Returns at once for an unknown block. */
int cmf_upload_wait(cmf_ctx* ctx, const float* host_block);

/* Input already on the device: pointer to element (line 0, band band_lo, sample 0); consecutive lines are
 * line_pitch floats apart, consecutive bands band_pitch floats apart (a full BIL cube: B*S and S). */
int cmf_bind_device_slab(cmf_ctx* ctx, const float* dev_slab, int64_t line_pitch, int32_t band_pitch);

/* Background modes (-k > 1, cmf/robust_mf.py:306-344).  labels[l*S + s] in 0..kmodes-1 is the cluster of pixel
 * (l, s) (ignored where the pixel is invalid); the reference obtains them from an unseeded MiniBatchKMeans
 * (:312), here they are an input so that the path is deterministic.  reject_min > 0 is -r with bgminsamp =
 * reject_min (:200, :316-324).  Everything downstream -- cluster counts, rejection, the mode list, the
 * per-mode fits with n = the column's valid count (:355-356), the overwrite order (:339-386), the inlier
 * statistics (:388-391) -- runs on the device.  labels == NULL returns to the unimodal path.  Host pointer. */
int cmf_set_labels(cmf_ctx* ctx, const int32_t* labels, int kmodes, int reject_min);

/* The same, with the partition found on the device (:308-313): the column's valid pixels are projected on the
 * `pcadim` leading eigenvectors of the column covariance and partitioned by a k-means with `kmodes` clusters.
 * The reference's MiniBatchKMeans is unseeded, so its partition is not reproducible; this one is deterministic
 * (rule in csrc/k_cluster.cu, restated in oracle/cluster_oracle.py): components by descending eigenvalue, signed
 * so that v . mu >= 0; projections quantised to 2^-24 of the column's range and cluster sums kept in 64-bit
 * integers; initial partition = kmodes equal-count slices of component 1; Lloyd iterations until no pixel or at
 * most 1/256 of the column's pixels change cluster, at most max_iter (<= 0: 100) of them.  kmodes <= 1 returns to the unimodal path.  CMF_OUT_LABELS returns the labels found. */
int cmf_set_clustering(cmf_ctx* ctx, int kmodes, int pcadim, int reject_min, int max_iter);

/* -f (:161, :358): in a labelled / clustered looshrinkage run the shrinkage target of every mode fit is the
 * covariance of the whole column instead of diag(S) (looshrinkage's I_reg, :100, :131).  On the device this is a
 * whitening ahead of the same eigen-solve (Cholesky factor in shared memory up to 96 bands, the target's own spectral
 * factor in global memory for wider windows).  No effect on unimodal runs (the reference tests bgmodes > 1). */
int cmf_set_regfull(cmf_ctx* ctx, int enable);

/* Opt-in, default off (= the reference's behaviour): keep pixels out of the BACKGROUND STATISTICS.  A pixel with
 * exclude[l*S + s] != 0 -- e.g. the saturated / specular / dark / cloud pixels cmf_pixel_flags() reports
 *  -- does not enter the column mean (:347), the covariance (:52-70) or the alpha search
 * (:105-127, n = the pixels that do).  Every valid pixel (:282) is still scored with the resulting filter and
 * CMF_OUT_COLSTATS covers every scored pixel; CMF_OUT_NVALID then holds the background count.  Unimodal runs
 * only (ignored while labels / clustering are set).  exclude == NULL returns to the default.  Host pointer. */
int cmf_set_exclusion(cmf_ctx* ctx, const uint8_t* exclude);

/* The alpha search is screened on the tensor cores and every alpha whose screened nll lies within
 * rel_margin * max|screened part of nll| of the minimum is re-evaluated exactly (CMF_OUT_SCREEN_TOL holds the absolute
 * margin per column).  Default 1e-5.  With certify != 0 (default) every run carries a runtime certificate: a refined
 * column also gets its best EXCLUDED 8-alpha tile evaluated in FP64; if the exact minimum falls into that tile, or the
 * measured spread (max - min over the evaluated alphas) of exact - screened nll reaches 1/4 of the margin -- only the
 * variation of the screening error between alphas can misorder them, a common bias cancels -- all alphas of the column
 * are re-evaluated in FP64; and if the worst measurement of the flightline reaches 1/4, so are the columns the screen
 * decided alone.
 * certify = 0 or a smaller margin are for diagnostics (tests/test_gpu_parity.py drives both). */
int cmf_set_screen_margin(cmf_ctx* ctx, double rel_margin, int certify);

/* ---- compute: the whole column loop (:297-392) for every column, no host round trip ---- */
enum {
    CMF_RUN_TIMING = 1,           /* bracket every kernel with CUDA events (see cmf_kernel_times) */
    CMF_RUN_EXACT = 2,            /* evaluate every alpha in FP64 (no tensor-core screening of the search) */
    CMF_RUN_ASYNC = 4             /* cmf_run_host: return once everything is enqueued; cmf_sync() completes it */
};
int cmf_run(cmf_ctx* ctx, uint32_t flags);
int cmf_sync(cmf_ctx* ctx);

/* Upload + run + download of scores / column statistics / alpha indices in one call.  The cube is copied in
 * blocks of one Gram chunk and the repack and statistics passes chase the copies, so only the factorisation,
 * alpha search and scoring remain once the last block has landed.  Any output pointer may be NULL.
 * Synchronous unless CMF_RUN_ASYNC is set (pinned buffers; outputs are valid after cmf_sync()): two contexts
 * used alternately keep the PCIe link busy while the previous flightline is still being scored. */
int cmf_run_host(cmf_ctx* ctx, const float* host_cube, double* mf_out, double* colstats_out,
                 int32_t* alpha_index_out, uint32_t flags);

/* ---- results ---- */
int cmf_download(cmf_ctx* ctx, int what, void* host_dst, size_t bytes);   /* synchronous */
void* cmf_device_ptr(cmf_ctx* ctx, int what);                            /* NULL if not available */
size_t cmf_output_bytes(const cmf_ctx* ctx, int what);

/* Drop-in for the importable looshrinkage(I_zm, alphas, nll, n, I_reg=[]) -> (C, mindex) of the reference
 * (cmf/robust_mf.py:92-136), for one sample matrix: I_zm is double [rows][D] row-major (mean-removed samples), n the
 * sample count the caller passes (:355-356).  nll_out[A] receives the leave-one-out negative log likelihood of every
 * alpha (inf where det(G) under/overflows, :111-113), mindex_out the argmin (-1 when every entry is inf, :121-127),
 * C_out double [D][D] the shrunk covariance (1 - alpha) S + alpha diag(S) (:130-134).  Every alpha is evaluated in
 * FP64 on the device (blocked kernels of csrc/k_wide.cu, any D up to 425+).  I_reg (the -f target, :99, :131):
 * double [reg_rows][D] or NULL / reg_rows 0; when given, the target is cov(I_reg) in the search and in C_out.
 * Independent of cmf_set_problem.  Synchronous. */
int cmf_looshrinkage(cmf_ctx* ctx, const double* I_zm, int32_t rows, int32_t D, const double* alphas, int32_t A,
                     int32_t n, const double* I_reg, int32_t reg_rows, double* nll_out, double* C_out,
                     int32_t* mindex_out);

/* ---- the steps either side of the filter ---- */

/* Per-pixel spectrometer flags of a radiance cube: the per-pixel tests of spectrometer_masks/masks_sds.py
 * (get_saturation_mask :133-150, get_spec_mask :152-163, get_dark_mask :165-180, get_cloud_mask :182-230).
 * Region growing, buffers and dilation (:232-330) are image morphology and are not part of this call.
 * Band numbers are 0-based indices into the cube's band axis, as the reference indexes them; < 0 disables a test. */
enum { CMF_FLAG_SATURATED = 1, CMF_FLAG_SPECULAR = 2, CMF_FLAG_DARK = 4, CMF_FLAG_CLOUD = 8 };
typedef struct cmf_flag_spec {
    int32_t sat_lo, sat_hi;      /* saturation window (bands whose wavelength is in 1945..2485 nm, :148) */
    int32_t spec_band;           /* 25 (:159) */
    int32_t dark_band;           /* 352, 2139 nm (:175) */
    int32_t cloud_a, cloud_b;    /* 15 and 60 (450 nm, 670 nm; :194) */
    float sat_thresh;            /* 6.0 (SAT_THRESH_DEFAULT, :50) */
    float spec_thresh;           /* 9.0 (--visible-mask-growing-threshold, :102) */
    float dark_thresh;           /* 0.104 (:78) */
    float cloud_thresh;          /* 15.0 (SAT_THRESH_CLD, :52) */
    float cloud_dwl;             /* wavelength[cloud_b] - wavelength[cloud_a]; the slope test is
                                    (r_b - r_a) / cloud_dwl < 0 (:213-222) */
} cmf_flag_spec;
/* cube: (lines, bands, samples) float32 BIL, on the host (only the bands the tests read are transferred) or,
 * with on_device != 0, already on the device.  flags_host: uint8 [lines][samples] of CMF_FLAG_* bits (may be
 * NULL; CMF_OUT_FLAGS keeps the device copy).  Independent of cmf_set_problem.  Synchronous. */
int cmf_pixel_flags(cmf_ctx* ctx, const float* cube, int on_device, int32_t lines, int32_t bands, int32_t samples,
                    const cmf_flag_spec* spec, uint8_t* flags_host);

/* Column profile of the scores of the last run (triage/cmf_profile.py:110-140): per cross-track column, over
 * the pixels that are not no-data / NaN and are > 0, evaluated in float32 exactly as numpy does there.
 *   robust == 0: npix, avg, std (ddof 0), min, max               (:127-130)
 *   robust != 0: npix, med, mad, p-low, p-high                   (:123-125; nearest-rank percentiles at
 *                q = (1 - p) * 100 and p * 100, the reference uses p = 0.95)
 * out_host: double [5][samples]; columns without such a pixel give npix = 0 and NaN.  Synchronous. */
int cmf_column_profile(cmf_ctx* ctx, int robust, double p, double* out_host);
/* The same for a score image on the host (the last band of a product on disk, :112): double [lines][samples],
 * `nodata` the product's 'data ignore value'.  Independent of cmf_set_problem.  Synchronous. */
int cmf_column_profile_image(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples, double nodata,
                             int robust, double p, double* out_host);

/* Detection pre-filter: the per-pixel head of srcfinder_util.filtdet (:1428-1436) with kde (:1383-1387) on a score
 * image.  weights[2*radius+1] is the normalised, symmetric 1-D kernel the reference builds with
 * scipy.ndimage.gaussian_filter(sigma = k = 50, truncate = 1) (radius = int(k + 0.5); srcfinder_b200/detect.py
 * builds it with the same numpy expression); it is applied along the lines, then along the samples, with 'reflect'
 * borders and scipy's summation order.  Then imgkde is scaled to [0, 1] by its min / max, detkde = mf * imgkde,
 *   detkde_host  double [lines][samples]  clip((detkde - mfmin) / (mfmax - mfmin), 0, 1)    (mfmin, mfmax = 500, 1500)
 *   ch4min_host  uint8  [lines][samples]  mf >= mfmin                                        (may be NULL)
 *   detmask_host uint8  [lines][samples]  detkde > 0, the candidate mask                      (may be NULL)
 * The connected-component filtering that follows in the reference (:1437-1470) is image morphology and stays with
 * the caller.  mf_host == NULL takes the scores of the last run, which are already on the device (lines / samples
 * ignored).  Synchronous. */
int cmf_detection_prefilter(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples,
                            const double* weights, int32_t radius, double mfmin, double mfmax, double* detkde_host,
                            uint8_t* ch4min_host, uint8_t* detmask_host);

/* CNN input normalisation (cnn/cnn_pred_pipeline.py:19-30, 126-157): ClampCH4(vmin, vmax) followed by
 * transforms.Normalize(mean, std) in float32: out = (clamp((float)mf, vmin, vmax) - mean) / std; no-data pixels are
 * clamped like every other value, as in the reference.  out_host: float [lines][samples].  mf_host == NULL takes the
 * scores of the last run.  Synchronous. */
int cmf_cnn_input(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples, float vmin, float vmax,
                  float mean, float stdv, float* out_host);

/* ---- instrumentation ---- */
int cmf_kernel_count(void);
const char* cmf_kernel_name(int i);
/* mean milliseconds of each kernel over every cmf_run(CMF_RUN_TIMING) since the previous call (the
 * events sit on the context stream, inside whatever region the caller is timing); returns the count */
int cmf_kernel_times(cmf_ctx* ctx, float* ms, int n);
/* launches issued by the last cmf_run()/cmf_run_host() */
int cmf_launch_count(const cmf_ctx* ctx);
/* name of the __global__ function the alpha-search screening pass uses for this problem
 * ("loo_screen5_kernel": tcgen05/TMEM, "loo_screen_kernel": mma.sync, "" if the problem is not screened) */
const char* cmf_screen_kernel(const cmf_ctx* ctx);

/* ---- pinned host memory helpers (for asynchronous uploads) ---- */
void* cmf_host_alloc(size_t bytes);
void cmf_host_free(void* p);
int cmf_host_register(void* p, size_t bytes);
int cmf_host_unregister(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CMF_B200_H */
