// Host side of the C ABI (include/cmf_b200.h): context, device buffers, the kernel sequence.
// No arithmetic happens here and nothing falls back to the CPU.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "../../include/cmf_b200.h"
#include "cmf_internal.h"

using namespace cmf;

namespace {

enum { K_REPACK = 0, K_MEAN, K_GRAM, K_EIGEN, K_TABLES, K_SCREEN, K_SELECT, K_LOO, K_FINALIZE, K_SCORE,
       K_COLSTATS, K_COUNT };
const char* kKernelNames[K_COUNT] = {"repack", "mean", "gram", "eigen", "tables", "screen", "select", "loo",
                                     "finalize", "score", "colstats"};

std::string g_create_error;

}  // namespace

struct cmf_ctx {
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    bool own_stream = false;
    std::string err;

    bool have_problem = false;
    Dims d{};
    int B = 0, band_lo = 0, band_hi = 0, reflectance = 0, model = 0;
    double nodata = -9999.0, scale = 1.0e5;

    // input
    const float* slab = nullptr;     // view used by the kernels (element: line 0, band_lo, sample 0)
    float* slab_own = nullptr;       // allocated by cmf_upload_bil / cmf_run_host
    bool have_input = false;

    // work + output buffers
    std::vector<void*> allocs;
    float* xt = nullptr;
    uint8_t* mask = nullptr;
    double *colsum_part = nullptr, *mu = nullptr, *gram_part = nullptr, *P = nullptr, *Pf = nullptr,
           *Wf = nullptr, *lam = nullptr, *logdet = nullptr, *beta = nullptr, *fpart = nullptr,
           *nll = nullptr, *w = nullptr, *wT = nullptr, *c0 = nullptr, *mf = nullptr, *stat_part = nullptr,
           *colstats = nullptr, *alphas_d = nullptr, *abscf_d = nullptr;
    int *colcnt_part = nullptr, *n = nullptr, *status = nullptr, *sweeps = nullptr, *mindex = nullptr;
    // tensor-core screening of the alpha search (K3a/K3b)
    float *Ws = nullptr, *betaf = nullptr, *Ps = nullptr, *tab5 = nullptr;
    bool use_screen5 = false;     // tcgen05/TMEM screening kernel (k_screen5.cu)
    double *rsum = nullptr, *fscreen = nullptr, *tol_col = nullptr, *slogT = nullptr;
    int eigen_method = 0;         // 0 Householder + QL, 1 cyclic Jacobi (CMF_EIGEN=jacobi, cross-checks)
    int *sel_index = nullptr, *ncand = nullptr, *probe = nullptr;
    unsigned long long *tile_mask = nullptr, *redo = nullptr;
    double *check = nullptr, *check_worst = nullptr;   // runtime certificate of the screen (k_screen.cu, K4)
    int certify = 1;
    // background modes (-k > 1, -r): labels are an input, everything derived from them lives on the device
    bool have_labels = false;
    int kmodes = 1, reject_min = 0;
    int32_t* labels_d = nullptr;
    int8_t* entries = nullptr;
    uint32_t *rejmask = nullptr, *flagmask = nullptr;
    int *nentries = nullptr, *nuse = nullptr;
    uint8_t *sel = nullptr, *inlier = nullptr;
    int16_t *cluster_img = nullptr, *alpha_img = nullptr;
    int32_t* rowidx = nullptr;    // row of every member pixel in the compacted xt of a background-mode pass
    float* xt_mode = nullptr;     // the compacted copy (the full copy stays in xt for every mode of the run)
    // on-device partition (cmf_set_clustering) and the -f regulariser (cmf_set_regfull)
    bool auto_cluster = false, regfull = false;
    int pcadim = 6, km_max_iter = 100, y_pd = 0;
    double *gram_full = nullptr, *vtop = nullptr, *ypca = nullptr;
    int32_t* qpca = nullptr;      // quantised projections [S][pcadim][L] of the k-means
    int *pick = nullptr, *km_iters = nullptr;
    uint8_t* lab8 = nullptr;
    bool can_screen = false;
    // relative to the screened part of nll.  What can misorder two alphas is the VARIATION of the screening error
    // between them; measured (profiles/r02g_margin_sweep.json): 0.06 of this margin, the certificate re-evaluates at 0.25
    double screen_tol = 1.0e-5;
    int nchunk_screen = 1;
    int nsplit = 1, lps = 8, nchunk_gram = 1, nchunk_loo = 1, nlanes = 1, score_lpc = 0;
    int lpc_gram = 16, spc = 1;   // lines per Gram chunk = spc repack splits = one upload block of cmf_run_host
    double* ctr = nullptr;        // pilot centre of the Gram pass: column mean over the first Gram chunk

    cudaEvent_t ev[K_COUNT + 1] = {};              // scratch set (ordering events, untimed runs)
    std::vector<std::vector<cudaEvent_t>> ev_sets;  // one set of K_COUNT+1 events per timed run
    int timed_runs = 0;                             // timed runs recorded since the last cmf_kernel_times()
    std::vector<cudaEvent_t> blk_ev;
    std::vector<std::pair<const void*, cudaEvent_t>> stage_ev;   // one event per staging block of cmf_upload_lines
    bool timed = false;
    bool screened = false;                          // the last run used the screening path
    int launches = 0;
    // products either side of the filter (k_products.cu); independent of the problem buffers
    uint8_t* flags_d = nullptr;
    size_t flags_bytes = 0;
    // opt-in exclusion of pixels from the background statistics (cmf_set_exclusion): 1 = pixel takes part
    bool have_excl = false;
    uint8_t* excl_sel = nullptr;
    // wide-window kernel set (k_wide.cu, k_gram8.cu): windows wider than 8 * kMaxNT bands (-R, robust_mf.py:186-187)
    bool wide = false, use_gram8 = false;
    int APW = 0, zbatch = 1, wsplit = 1, wlps = 1;
    float *lo_part = nullptr, *hi_part = nullptr;
    int* qexp = nullptr;
    int8_t* img = nullptr;
    WideTarget wtarget;           // -f on a wide window (cmf_set_regfull)
    double *wgram = nullptr, *wwork = nullptr, *wdinv = nullptr, *wdvec = nullptr, *wevec = nullptr, *Wtab = nullptr,
           *Zbuf = nullptr;
    double2* rot = nullptr;
    int2* iters = nullptr;
};

namespace {

int fail(cmf_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call)                                                                          \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess)                                                            \
            return fail(ctx, CMF_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
    } while (0)

void free_buffers(cmf_ctx* c) {
    for (void* p : c->allocs) cudaFree(p);
    c->allocs.clear();
    if (c->slab_own) { cudaFree(c->slab_own); c->slab_own = nullptr; }
    c->have_input = false;
    c->slab = nullptr;
}

template <typename T>
cudaError_t dalloc(cmf_ctx* c, T** p, size_t count) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e == cudaSuccess) {
        c->allocs.push_back(q);
        *p = reinterpret_cast<T*>(q);
        // testing hook: CMF_POISON=1 fills every work buffer with 0xFF (NaN / -1) so that any read of memory the
        // pipeline has not written shows up (fresh cudaMalloc pages are usually zero, recycled ones are not)
        static const bool poison = cmf_hook("CMF_POISON") != nullptr;
        if (poison) e = cudaMemset(q, 0xFF, bytes);
    }
    return e;
}

// One pass of the per-column model fit and scoring on the context stream: statistics of the selected
// pixels (sel == NULL: every valid pixel), factorisation, alpha search, weights, scores.  `mark` records
// the per-kernel timing events of an unlabelled run.
template <class Mark>
void fit_and_score(cmf_ctx* ctx, bool exact, const uint8_t* sel, const int* nloo, Mark mark, bool chased = false) {
    const Dims& d = ctx->d;
    cudaStream_t st = ctx->stream;
    // Unimodal passes centre the Gram products on a pilot point (the mean of the first Gram chunk) so that
    // cmf_run_host can run the statistics chunk by chunk behind the upload; K2 removes the offset exactly.
    // Both entry points do the same arithmetic, so their results are bitwise identical.
    const bool pilot = sel == nullptr;
    if (!chased) {
        if (pilot) launch_mean(d, ctx->colsum_part, ctx->colcnt_part, ctx->spc, ctx->ctr, nullptr, st);
        launch_mean(d, ctx->colsum_part, ctx->colcnt_part, ctx->nsplit, ctx->mu, ctx->n, st);
        mark(2);
        launch_gram(d, ctx->xt, pilot ? ctx->ctr : ctx->mu, ctx->nchunk_gram, ctx->lpc_gram, 0, ctx->nchunk_gram,
                    ctx->gram_part, st);
        mark(3);
        ctx->launches += pilot ? 3 : 2;
    } else {
        launch_mean(d, ctx->colsum_part, ctx->colcnt_part, ctx->nsplit, ctx->mu, ctx->n, st);
        ++ctx->launches;
    }
    const bool loo = ctx->model == CMF_MODEL_LOOSHRINKAGE;
    const bool screen = loo && ctx->can_screen && !exact;
    // -f: the shrinkage target of a mode fit is the covariance of the whole column (:100, :358)
    const bool full_target = sel != nullptr && ctx->regfull && ctx->model == CMF_MODEL_LOOSHRINKAGE;
    launch_eigen(d, ctx->gram_part, ctx->nchunk_gram, ctx->n, ctx->mu, pilot ? ctx->ctr : nullptr, ctx->P, ctx->lam,
                 ctx->slogT, ctx->status, ctx->sweeps, ctx->eigen_method, st, full_target ? 2 : 0,
                 full_target ? ctx->gram_full : nullptr, full_target ? ctx->nuse : nullptr);
    mark(4);
    launch_tables(d, ctx->n, nloo, ctx->alphas_d, ctx->model, ctx->P, ctx->lam, ctx->slogT, ctx->Pf, ctx->Wf,
                  ctx->logdet, ctx->beta, (screen && !ctx->use_screen5) ? ctx->Ws : nullptr, ctx->betaf, ctx->rsum,
                  (screen && !ctx->use_screen5) ? ctx->Ps : nullptr, st);
    mark(5);
    ctx->launches += 2;
    if (screen && ctx->use_screen5) {
        launch_screen5(d, ctx->xt, ctx->mu, ctx->n, nloo, ctx->alphas_d, ctx->P, ctx->lam, ctx->tab5, ctx->betaf,
                       ctx->nchunk_screen, ctx->fscreen, st);
        ctx->launches += 2;
    } else if (screen) {
        launch_screen(d, ctx->xt, ctx->mu, ctx->Pf, ctx->Ps, ctx->Ws, ctx->betaf, ctx->n, ctx->nchunk_screen,
                      ctx->fscreen, st);
        ++ctx->launches;
    }
    mark(6);
    if (screen) {
        launch_select(d, ctx->fscreen, ctx->nchunk_screen, ctx->logdet, ctx->rsum, ctx->n, nloo, ctx->screen_tol,
                      ctx->nll, ctx->sel_index, ctx->tile_mask, ctx->ncand, ctx->tol_col,
                      ctx->use_screen5 ? ctx->betaf : nullptr, ctx->probe, st);
        ++ctx->launches;
    }
    mark(7);
    if (loo) {
        launch_loo(d, ctx->xt, ctx->mu, ctx->Pf, ctx->Wf, ctx->beta, ctx->nchunk_loo, ctx->fpart,
                   screen ? ctx->tile_mask : nullptr, st);
        ++ctx->launches;
    }
    mark(8);
    launch_finalize(d, ctx->fpart, ctx->nchunk_loo, ctx->logdet, ctx->n, ctx->alphas_d, ctx->P, ctx->lam,
                    ctx->mu, ctx->abscf_d, ctx->model, ctx->reflectance, ctx->scale, ctx->nll, ctx->mindex,
                    ctx->w, ctx->wT, ctx->c0, ctx->status, screen ? ctx->sel_index : nullptr,
                    screen ? ctx->tile_mask : nullptr, nloo, st, screen ? ctx->probe : nullptr,
                    screen ? ctx->tol_col : nullptr, screen ? ctx->check : nullptr, screen ? ctx->redo : nullptr);
    if (screen) {
        // runtime certificate of the screen: columns whose measured screening error is not safely below the margin
        // (or all screen-decided columns, if the flightline's worst measurement is not) are re-evaluated exactly;
        // when nothing is flagged the two extra launches return at once
        launch_certify(d, ctx->check, ctx->sel_index, ctx->certify, ctx->redo, ctx->check_worst, st);
        launch_loo(d, ctx->xt, ctx->mu, ctx->Pf, ctx->Wf, ctx->beta, ctx->nchunk_loo, ctx->fpart, ctx->redo, st);
        launch_finalize(d, ctx->fpart, ctx->nchunk_loo, ctx->logdet, ctx->n, ctx->alphas_d, ctx->P, ctx->lam,
                        ctx->mu, ctx->abscf_d, ctx->model, ctx->reflectance, ctx->scale, ctx->nll, ctx->mindex,
                        ctx->w, ctx->wT, ctx->c0, ctx->status, nullptr, nullptr, nloo, st, nullptr, nullptr, nullptr,
                        nullptr, ctx->redo);
        ctx->launches += 3;
    }
    mark(9);
    launch_score(d, ctx->slab, ctx->mask, ctx->wT, ctx->c0, ctx->status, ctx->nodata, ctx->mf, ctx->stat_part,
                 ctx->nlanes, ctx->score_lpc, sel, ctx->mindex, ctx->alpha_img, st);
    mark(10);
    ctx->launches += 2;
    ctx->screened = screen;
}

// The same for a wide active window: every step is blocked, the per-column matrices live in global memory.
// exact = true takes the FP64 tensor (DMMA) Gram pass instead of the integer tcgen05 pass (cross-check).
// sel / nloo / write_mask as in fit_and_score (a background-mode pass fits the selected pixels, n of the search = nloo).
template <class Mark>
void wide_fit_and_score(cmf_ctx* ctx, bool exact, const uint8_t* sel, const int* nloo, int write_mask, bool modes_pass,
                        Mark mark) {
    const Dims& d = ctx->d;
    cudaStream_t st = ctx->stream;
    const bool g8 = ctx->use_gram8 && !exact;
    launch_wide_stats(d, ctx->slab, ctx->mask, sel, write_mask, ctx->wsplit, ctx->wlps, ctx->colsum_part,
                      ctx->colcnt_part, ctx->lo_part, ctx->hi_part, ctx->mu, ctx->n, ctx->ctr, ctx->qexp, st);
    launch_wide_pack(d, ctx->slab, ctx->mask, sel, ctx->ctr, ctx->qexp, ctx->xt, g8 ? ctx->img : nullptr, st);
    mark(1);
    mark(2);
    if (g8) launch_wide_gram8(d, ctx->img, ctx->wgram, st);
    else launch_wide_gram64(d, ctx->xt, ctx->ctr, ctx->wgram, st);
    mark(3);
    const bool loo = ctx->model == CMF_MODEL_LOOSHRINKAGE;
    const bool full_target = modes_pass && ctx->regfull && loo;       // -f (:353-356)
    launch_wide_eigen(d, ctx->wgram, ctx->n, ctx->mu, ctx->ctr, g8 ? ctx->qexp : nullptr, full_target ? 2 : 0,
                      ctx->wwork, ctx->wdinv, ctx->wdvec, ctx->wevec, ctx->rot, ctx->iters, ctx->sweeps, ctx->P,
                      ctx->lam, ctx->slogT, ctx->status, st, full_target ? &ctx->wtarget : nullptr);
    mark(4);
    ctx->launches += full_target ? 15 : 10;
    if (loo) {
        launch_wide_tables(d, ctx->APW, ctx->n, nloo, ctx->alphas_d, ctx->model, ctx->lam, ctx->slogT, ctx->logdet,
                           ctx->beta, ctx->rsum, ctx->Wtab, st);
        ++ctx->launches;
    }
    mark(5);
    mark(6);
    mark(7);
    if (loo) {
        for (int s0 = 0; s0 < d.S; s0 += ctx->zbatch) {
            const int ns = std::min(ctx->zbatch, d.S - s0);
            launch_wide_loo(d, ctx->APW, ctx->xt, ctx->mu, ctx->P, ctx->Wtab, ctx->beta, ctx->n, s0, ns, ctx->nchunk_loo,
                            ctx->Zbuf, ctx->fpart, st);
            ctx->launches += 2;
        }
    }
    mark(8);
    launch_finalize(d, ctx->fpart, ctx->nchunk_loo, ctx->logdet, ctx->n, ctx->alphas_d, ctx->P, ctx->lam, ctx->mu,
                    ctx->abscf_d, ctx->model, ctx->reflectance, ctx->scale, ctx->nll, ctx->mindex, ctx->w, ctx->wT,
                    ctx->c0, ctx->status, nullptr, nullptr, nloo, st);
    mark(9);
    launch_score(d, ctx->slab, ctx->mask, ctx->wT, ctx->c0, ctx->status, ctx->nodata, ctx->mf, ctx->stat_part,
                 ctx->nlanes, ctx->score_lpc, modes_pass ? sel : nullptr, ctx->mindex, ctx->alpha_img, st);
    mark(10);
    ctx->launches += 2;
    ctx->screened = false;
}

// Launch the whole column loop on the context stream.  When `blocks_ready` is given, the first repack pass
// runs block by block, each block waiting on the event that marks its upload as complete.
int enqueue(cmf_ctx* ctx, bool timing, bool exact, const std::vector<cudaEvent_t>* blocks_ready,
            int lines_per_block) {
    const Dims& d = ctx->d;
    cudaStream_t st = ctx->stream;
    ctx->launches = 0;
    cudaEvent_t* evs = nullptr;
    if (timing) {
        if (ctx->timed_runs >= 256) ctx->timed_runs = 0;      // ring: keep the last 256 timed runs
        if ((int)ctx->ev_sets.size() <= ctx->timed_runs) {
            std::vector<cudaEvent_t> set(K_COUNT + 1);
            for (auto& e : set) cudaEventCreate(&e);
            ctx->ev_sets.push_back(set);
        }
        evs = ctx->ev_sets[ctx->timed_runs].data();
        ++ctx->timed_runs;
    }
    const bool modes = ctx->have_labels || ctx->auto_cluster;
    auto mark = [&](int i) { if (timing && !modes) cudaEventRecord(evs[i], st); };
    auto no_mark = [](int) {};
    if (timing && modes) cudaEventRecord(evs[0], st);
    mark(0);
    // ---- pass over every valid pixel: validity mask, column-major copy, column sums
    const bool chased = blocks_ready != nullptr && !modes;
    const uint8_t* excl = (!modes && ctx->have_excl) ? ctx->excl_sel : nullptr;
    if (ctx->wide && blocks_ready) for (cudaEvent_t e : *blocks_ready) cudaStreamWaitEvent(st, e, 0);
    if (ctx->wide && !modes) {
        wide_fit_and_score(ctx, exact, excl, nullptr, 1, false, mark);
        if (excl) {
            launch_count_mask(d, ctx->mask, ctx->nuse, st);
            launch_colstats_modes(d, ctx->mf, ctx->mask, ctx->nuse, ctx->nodata, ctx->colstats, st);
            ctx->launches += 2;
        } else {
            launch_colstats(d, ctx->stat_part, ctx->nlanes, ctx->n, ctx->nodata, ctx->colstats, st);
            ++ctx->launches;
        }
        mark(11);
        ctx->timed = timing;
        CK(cudaGetLastError());
        return CMF_OK;
    }
    if (blocks_ready) {
        // one block = one Gram chunk: its repack and (unimodal) its Gram partial run as soon as its copy lands
        int line = 0;
        for (size_t b = 0; b < blocks_ready->size(); ++b) {
            const int lim = std::min(d.L, line + lines_per_block);
            cudaStreamWaitEvent(st, (*blocks_ready)[b], 0);
            launch_repack(d, ctx->slab, ctx->xt, ctx->mask, ctx->colsum_part, ctx->colcnt_part, ctx->lps, line,
                          lim, excl, 1, st);
            ++ctx->launches;
            if (chased) {
                if (b == 0) {
                    launch_mean(d, ctx->colsum_part, ctx->colcnt_part, ctx->spc, ctx->ctr, nullptr, st);
                    ++ctx->launches;
                }
                launch_gram(d, ctx->xt, ctx->ctr, ctx->nchunk_gram, ctx->lpc_gram, (int)b, (int)b + 1,
                            ctx->gram_part, st);
                ++ctx->launches;
            }
            line = lim;
        }
    } else {
        launch_repack(d, ctx->slab, ctx->xt, ctx->mask, ctx->colsum_part, ctx->colcnt_part, ctx->lps, 0, d.L,
                      excl, 1, st);
        ++ctx->launches;
    }
    mark(1);
    if (!modes) {
        fit_and_score(ctx, exact, nullptr, nullptr, mark, chased);
        if (excl) {
            // every valid pixel was scored, the statistics count them all (n holds the background pixels only)
            launch_count_mask(d, ctx->mask, ctx->nuse, st);
            launch_colstats_modes(d, ctx->mf, ctx->mask, ctx->nuse, ctx->nodata, ctx->colstats, st);
            ctx->launches += 2;
        } else {
            launch_colstats(d, ctx->stat_part, ctx->nlanes, ctx->n, ctx->nodata, ctx->colstats, st);
            ++ctx->launches;
        }
        mark(11);
    } else {
        if (ctx->wide) {
            // ---- background modes on a wide window: the same plan with the blocked kernels (no compaction)
            const size_t LSw = (size_t)d.L * d.S;
            const bool g8 = ctx->use_gram8 && !exact;
            launch_wide_stats(d, ctx->slab, ctx->mask, nullptr, 1, ctx->wsplit, ctx->wlps, ctx->colsum_part,
                              ctx->colcnt_part, ctx->lo_part, ctx->hi_part, ctx->mu, ctx->nuse, ctx->ctr, ctx->qexp, st);   // nuse (:302)
            ctx->launches += 3;
            const bool full_target = ctx->regfull && ctx->model == CMF_MODEL_LOOSHRINKAGE;
            if (ctx->auto_cluster || full_target) {
                // scatter of the whole column: the PCA basis (:310) and the -f target (:353-356)
                launch_wide_pack(d, ctx->slab, ctx->mask, nullptr, ctx->ctr, ctx->qexp, ctx->xt, g8 ? ctx->img : nullptr, st);
                if (g8) launch_wide_gram8(d, ctx->img, ctx->wgram, st);
                else launch_wide_gram64(d, ctx->xt, ctx->ctr, ctx->wgram, st);
                ctx->launches += 2;
            }
            if (full_target) {
                launch_wide_eigen(d, ctx->wgram, ctx->nuse, ctx->mu, ctx->ctr, g8 ? ctx->qexp : nullptr, 0, ctx->wwork,
                                  ctx->wdinv, ctx->wdvec, ctx->wevec, ctx->rot, ctx->iters, ctx->sweeps, ctx->P, ctx->lam,
                                  ctx->slogT, ctx->status, st);
                launch_wide_target(d, ctx->P, ctx->lam, ctx->slogT, ctx->status, ctx->wtarget, st);
                ctx->launches += 5;
            }
            if (ctx->auto_cluster) {
                // PCA basis of the whole column (:310-311): plain eigenvectors of its covariance
                launch_wide_eigen(d, ctx->wgram, ctx->nuse, ctx->mu, ctx->ctr, g8 ? ctx->qexp : nullptr, 1, ctx->wwork,
                                  ctx->wdinv, ctx->wdvec, ctx->wevec, ctx->rot, ctx->iters, ctx->sweeps, ctx->P, ctx->lam,
                                  ctx->slogT, ctx->status, st);
                launch_pca_kmeans(d, ctx->xt, ctx->mask, ctx->mu, ctx->nuse, ctx->P, ctx->lam, ctx->pcadim, ctx->kmodes,
                                  ctx->km_max_iter, ctx->pick, ctx->vtop, ctx->ypca, ctx->qpca, ctx->lab8, ctx->labels_d,
                                  ctx->km_iters, st);
                ctx->launches += 7;
            }
            launch_modes(d, ctx->labels_d, ctx->mask, ctx->reject_min, ctx->entries, ctx->rejmask, ctx->nentries, ctx->flagmask, st);
            launch_fill_f64(ctx->mf, (long long)LSw, ctx->nodata, st);
            CK(cudaMemsetAsync(ctx->alpha_img, 0, LSw * sizeof(int16_t), st));
            ctx->launches += 3;
            for (int t = 0; t < ctx->kmodes; ++t) {
                launch_members(d, ctx->labels_d, ctx->mask, t, ctx->entries, ctx->rejmask, ctx->flagmask, ctx->sel,
                               t == 0 ? ctx->cluster_img : nullptr, ctx->inlier, st);
                ++ctx->launches;
                wide_fit_and_score(ctx, exact, ctx->sel, ctx->nuse, 0, true, no_mark);
            }
        } else {
        // ---- background modes (cmf/robust_mf.py:306-344): one fit-and-score pass per mode-list entry
        const size_t LS = (size_t)d.L * d.S;
        launch_mean(d, ctx->colsum_part, ctx->colcnt_part, ctx->nsplit, ctx->mu, ctx->nuse, st);   // nuse (:302)
        if (ctx->auto_cluster || ctx->regfull) {
            // scatter of the whole column about its mean: the PCA basis (:310) and the -f target (:358)
            launch_gram(d, ctx->xt, ctx->mu, ctx->nchunk_gram, ctx->lpc_gram, 0, ctx->nchunk_gram, ctx->gram_full, st);
            ++ctx->launches;
        }
        if (ctx->auto_cluster) {
            launch_eigen(d, ctx->gram_full, ctx->nchunk_gram, ctx->nuse, ctx->mu, nullptr, ctx->P, ctx->lam, ctx->slogT,
                         ctx->status, ctx->sweeps, 0, st, 1);
            launch_pca_kmeans(d, ctx->xt, ctx->mask, ctx->mu, ctx->nuse, ctx->P, ctx->lam, ctx->pcadim, ctx->kmodes,
                              ctx->km_max_iter, ctx->pick, ctx->vtop, ctx->ypca, ctx->qpca, ctx->lab8, ctx->labels_d,
                              ctx->km_iters, st);
            ctx->launches += 4;
        }
        launch_modes(d, ctx->labels_d, ctx->mask, ctx->reject_min, ctx->entries, ctx->rejmask, ctx->nentries, ctx->flagmask, st);
        launch_fill_f64(ctx->mf, (long long)LS, ctx->nodata, st);
        CK(cudaMemsetAsync(ctx->alpha_img, 0, LS * sizeof(int16_t), st));
        ctx->launches += 3;
        for (int t = 0; t < ctx->kmodes; ++t) {
            launch_members(d, ctx->labels_d, ctx->mask, t, ctx->entries, ctx->rejmask, ctx->flagmask, ctx->sel,
                           t == 0 ? ctx->cluster_img : nullptr, ctx->inlier, st);
            // the members of this mode, compacted to the first rows of every column of xt (in line order): the
            // statistics and the search then cost what the mode holds, not what the flightline holds
            launch_rank(d, ctx->sel, ctx->rowidx, st);
            float* const xt_full = ctx->xt;
            // the member rows are gathered from the full column-major copy (one read of the members, one write) ...
            const bool gathered = launch_compact(d, xt_full, ctx->sel, ctx->rowidx, ctx->nsplit, ctx->lps, ctx->xt_mode,
                                                 ctx->colsum_part, ctx->colcnt_part, st);
            if (gathered) {
                ctx->xt = ctx->xt_mode;
            } else {      // ... or, for a window that does not fit the gather's thread plan, repacked from the slab
                ctx->d.rowidx = ctx->rowidx;
                launch_repack(ctx->d, ctx->slab, ctx->xt, ctx->mask, ctx->colsum_part, ctx->colcnt_part, ctx->lps, 0,
                              d.L, ctx->sel, 0, st);
                ctx->d.rowidx = nullptr;
            }
            ctx->d.nrows = ctx->n;                 // the member count of the mean kernel (first launch of the pass)
            ctx->launches += 3;
            fit_and_score(ctx, exact, ctx->sel, ctx->nuse, no_mark);
            ctx->xt = xt_full;
            ctx->d.nrows = nullptr;
        }
        }
        launch_colstats_modes(d, ctx->mf, ctx->inlier, ctx->nuse, ctx->nodata, ctx->colstats, st);
        ++ctx->launches;
        if (timing) for (int i = 1; i <= K_COUNT; ++i) cudaEventRecord(evs[i], st);
    }
    ctx->timed = timing;
    CK(cudaGetLastError());
    return CMF_OK;
}

int ensure_own_slab(cmf_ctx* ctx) {
    const Dims& d = ctx->d;
    if (!ctx->slab_own) {
        void* q = nullptr;
        cudaError_t e = cudaMalloc(&q, (size_t)d.L * d.D * d.S * sizeof(float));
        if (e != cudaSuccess) return fail(ctx, CMF_E_NOMEM, std::string("slab alloc: ") + cudaGetErrorString(e));
        ctx->slab_own = reinterpret_cast<float*>(q);
    }
    ctx->slab = ctx->slab_own;
    ctx->d.line_pitch = (long long)d.D * d.S;
    ctx->d.band_pitch = d.S;
    ctx->d.vec2 = (d.S % 2 == 0) ? 1 : 0;
    return CMF_OK;
}

int ensure_mode_buffers(cmf_ctx* ctx) {
    if (ctx->labels_d) return CMF_OK;
    const size_t LS = (size_t)ctx->d.L * ctx->d.S;
    cudaError_t e = cudaSuccess;
    auto A_ = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A_(dalloc(ctx, &ctx->labels_d, LS));
    A_(dalloc(ctx, &ctx->sel, LS));
    A_(dalloc(ctx, &ctx->inlier, LS));
    A_(dalloc(ctx, &ctx->cluster_img, LS));
    A_(dalloc(ctx, &ctx->alpha_img, LS));
    A_(dalloc(ctx, &ctx->rowidx, LS));
    if (!ctx->wide) A_(dalloc(ctx, &ctx->xt_mode, LS * ctx->d.DP));
    if (e != cudaSuccess) return fail(ctx, CMF_E_NOMEM, std::string("label buffers: ") + cudaGetErrorString(e));
    return CMF_OK;
}

int ensure_full_gram(cmf_ctx* ctx) {
    if (ctx->gram_full) return CMF_OK;
    if (dalloc(ctx, &ctx->gram_full, gram_part_elems(ctx->d, ctx->nchunk_gram)) != cudaSuccess)
        return fail(ctx, CMF_E_NOMEM, "full-column Gram buffer");
    return CMF_OK;
}

struct OutDesc { void* ptr; size_t bytes; };

OutDesc out_desc(const cmf_ctx* c, int what) {
    const Dims& d = c->d;
    const size_t LS = (size_t)d.L * d.S;
    switch (what) {
        case CMF_OUT_MF: return {c->mf, LS * sizeof(double)};
        case CMF_OUT_MASK: return {c->mask, LS};
        case CMF_OUT_COLSTATS: return {c->colstats, (size_t)3 * d.S * sizeof(double)};
        case CMF_OUT_ALPHA_INDEX: return {c->mindex, (size_t)d.S * sizeof(int)};
        case CMF_OUT_NLL: return {c->nll, (size_t)d.S * d.A * sizeof(double)};
        case CMF_OUT_MU: return {c->mu, (size_t)d.S * d.DP * sizeof(double)};
        case CMF_OUT_WEIGHTS: return {c->w, (size_t)d.S * d.DP * sizeof(double)};
        case CMF_OUT_STATUS: return {c->status, (size_t)d.S * sizeof(int)};
        case CMF_OUT_NVALID: return {c->n, (size_t)d.S * sizeof(int)};
        case CMF_OUT_EIGVALS: return {c->lam, (size_t)d.S * d.DP * sizeof(double)};
        case CMF_OUT_SWEEPS: return {c->sweeps, (size_t)d.S * sizeof(int)};
        case CMF_OUT_NCAND: return {c->ncand, (size_t)d.S * sizeof(int)};
        case CMF_OUT_SCREEN_TOL: return {c->tol_col, (size_t)d.S * sizeof(double)};
        case CMF_OUT_CLUSTER_ID: return {c->cluster_img, c->cluster_img ? LS * sizeof(int16_t) : 0};
        case CMF_OUT_ALPHA_IMAGE: return {c->alpha_img, c->alpha_img ? LS * sizeof(int16_t) : 0};
        case CMF_OUT_MODE_LIST: return {c->entries, (size_t)d.S * kMaxLabels};
        case CMF_OUT_LABELS: return {c->labels_d, c->labels_d ? LS * sizeof(int32_t) : 0};
        case CMF_OUT_PCA: return {c->ypca, c->auto_cluster ? LS * c->pcadim * sizeof(double) : 0};
        case CMF_OUT_KMEANS_ITERS: return {c->km_iters, c->auto_cluster ? (size_t)d.S * sizeof(int) : 0};
        case CMF_OUT_FLAGS: return {c->flags_d, c->flags_bytes};
        case CMF_OUT_SCREEN_CHECK: return {c->check, (size_t)d.S * sizeof(double)};
        default: return {nullptr, 0};
    }
}

}  // namespace

extern "C" {

const char* cmf_version(void) { return "cmf_b200 0.1 (sm_100a)"; }

int cmf_create(cmf_ctx** out, int device) {
    cmf_ctx* ctx = nullptr;
    if (!out) return fail(nullptr, CMF_E_ARG, "cmf_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, CMF_E_CUDA,
                    std::string("cmf_create: no CUDA device (") + cudaGetErrorString(e) +
                        "); this library has no CPU path");
    if (device < 0 || device >= ndev) return fail(nullptr, CMF_E_ARG, "cmf_create: bad device index");
    e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail(nullptr, CMF_E_CUDA, cudaGetErrorString(e));
    ctx = new cmf_ctx();
    ctx->device = device;
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    ctx->own_stream = true;
    for (int i = 0; i <= K_COUNT; ++i) CK(cudaEventCreate(&ctx->ev[i]));
    *out = ctx;
    return CMF_OK;
}

void cmf_destroy(cmf_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (ctx->flags_d) { cudaFree(ctx->flags_d); ctx->flags_d = nullptr; }
    free_buffers(ctx);
    for (int i = 0; i <= K_COUNT; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (auto& set : ctx->ev_sets) for (cudaEvent_t e : set) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->blk_ev) cudaEventDestroy(e);
    for (auto& pe : ctx->stage_ev) cudaEventDestroy(pe.second);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    delete ctx;
}

const char* cmf_last_error(const cmf_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int cmf_set_stream(cmf_ctx* ctx, void* cuda_stream) {
    if (!ctx) return CMF_E_ARG;
    cudaSetDevice(ctx->device);
    if (ctx->own_stream && ctx->stream) { cudaStreamSynchronize(ctx->stream); cudaStreamDestroy(ctx->stream); }
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    ctx->own_stream = false;
    return CMF_OK;
}

int cmf_set_problem(cmf_ctx* ctx, const cmf_problem* p) {
    if (!ctx || !p) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (p->interleave != CMF_INTERLEAVE_BIL) return fail(ctx, CMF_E_ARG, "only BIL input is supported (robust_mf.py:208)");
    if (p->lines <= 0 || p->bands <= 0 || p->samples <= 0) return fail(ctx, CMF_E_ARG, "bad cube shape");
    if (p->band_lo < 1 || p->band_hi > p->bands || p->band_hi < p->band_lo)
        return fail(ctx, CMF_E_ARG, "active band window outside the cube");
    if (p->nodata > 0) return fail(ctx, CMF_E_ARG, "nodata value > 0, values will not be masked (robust_mf.py:233-234)");
    if (!p->abscf) return fail(ctx, CMF_E_ARG, "abscf is NULL");
    const int D = p->band_hi - p->band_lo + 1;
    const int NT = (D + 7) / 8;
    // CMF_FORCE_WIDE (tools build only): run a narrow window through the wide-window kernel set, an independent
    // implementation of every step (cross-check, tests/test_gpu_wide.py)
    const bool wide = NT > kMaxNT || cmf_hook("CMF_FORCE_WIDE") != nullptr;
    const bool loo = p->model == CMF_MODEL_LOOSHRINKAGE;
    if (!loo && p->model != CMF_MODEL_EMPIRICAL) return fail(ctx, CMF_E_ARG, "unknown model");
    if (loo && (p->num_alphas < 1 || p->num_alphas > 4096 || !p->alphas))
        return fail(ctx, CMF_E_ARG, "looshrinkage needs 1..4096 alphas");

    CK(cudaStreamSynchronize(ctx->stream));
    free_buffers(ctx);
    ctx->have_labels = false;
    ctx->labels_d = nullptr; ctx->sel = nullptr; ctx->inlier = nullptr; ctx->cluster_img = nullptr;
    ctx->alpha_img = nullptr; ctx->rowidx = nullptr; ctx->xt_mode = nullptr;
    ctx->auto_cluster = false; ctx->regfull = false; ctx->y_pd = 0;
    ctx->have_excl = false; ctx->excl_sel = nullptr;
    ctx->wtarget = WideTarget{};
    ctx->gram_full = nullptr; ctx->vtop = nullptr; ctx->ypca = nullptr; ctx->qpca = nullptr; ctx->pick = nullptr;
    ctx->km_iters = nullptr; ctx->lab8 = nullptr;
    Dims& d = ctx->d;
    d.L = p->lines; d.S = p->samples; d.D = D; d.NT = NT; d.DP = 8 * NT;
    d.A = loo ? p->num_alphas : 1;
    d.NT2 = (d.A + 7) / 8; d.AP = d.NT2 * 8;
    d.NT16 = (d.A + 15) / 16; d.AP16 = d.NT16 * 16;
    d.line_pitch = (long long)D * d.S; d.band_pitch = d.S; d.vec2 = (d.S % 2 == 0);
    d.rowidx = nullptr; d.nrows = nullptr;
    ctx->B = p->bands; ctx->band_lo = p->band_lo; ctx->band_hi = p->band_hi;
    ctx->reflectance = p->reflectance; ctx->model = p->model; ctx->nodata = p->nodata;
    ctx->scale = p->reflectance ? 1.0 : 1.0e5;   // ppmscaling, robust_mf.py:38,:383-386

    ctx->wide = wide;
    if (wide) {
        if (!wide_eigen_fits(d)) return fail(ctx, CMF_E_ARG, "active window too wide for the eigen-solver's shared-memory plan");
        ctx->wsplit = std::max(1, std::min(8, d.L / 256));
        ctx->wlps = (d.L + ctx->wsplit - 1) / ctx->wsplit;
        ctx->wsplit = (d.L + ctx->wlps - 1) / ctx->wlps;
        ctx->nsplit = ctx->wsplit; ctx->lps = ctx->wlps; ctx->spc = 1; ctx->lpc_gram = d.L; ctx->nchunk_gram = 1;
        ctx->APW = (d.AP + 63) / 64 * 64;
        ctx->nchunk_loo = std::max(1, std::min(16, d.L / 1024));
        ctx->nchunk_screen = 1;
        ctx->can_screen = false; ctx->use_screen5 = false;
        // the integer Gram accumulates 32-bit sums of products of balanced base-256 digits: fewer than 2^17 lines
        ctx->use_gram8 = d.L < (1 << 17);
        if (const char* e = cmf_hook("CMF_WIDE_GRAM")) if (strcmp(e, "fp64") == 0) ctx->use_gram8 = false;
        const size_t zcol = (size_t)d.L * d.DP * sizeof(double);
        ctx->zbatch = (int)std::max<size_t>(1, std::min<size_t>((size_t)d.S, ((size_t)4 << 30) / zcol));
        ctx->nlanes = score_plan(d, ctx->sm_count, &ctx->score_lpc);
    } else {
    ctx->nsplit = repack_nsplit(d);
    ctx->lps = repack_lines_per_split(d, ctx->nsplit);
    ctx->nsplit = (d.L + ctx->lps - 1) / ctx->lps;
    {   // Gram chunks are whole repack splits (and upload blocks of cmf_run_host are whole Gram chunks)
        const int want = pick_chunks(d.S, d.L, 256, ctx->sm_count, 1);
        const int lpc0 = (d.L + want - 1) / want;
        ctx->spc = std::max(1, (lpc0 + ctx->lps / 2) / ctx->lps);
        ctx->lpc_gram = ctx->spc * ctx->lps;
        ctx->nchunk_gram = (d.L + ctx->lpc_gram - 1) / ctx->lpc_gram;
    }
    ctx->nchunk_loo = pick_chunks(d.S, d.L, 128, ctx->sm_count, 1);
    ctx->nlanes = score_plan(d, ctx->sm_count, &ctx->score_lpc);
    ctx->nchunk_screen = ctx->nchunk_loo;
    // the screening pass needs its tables in shared memory and a 64-bit tile mask (A <= 512)
    ctx->can_screen = loo && d.NT2 <= 64 && (screen5_supported(d) || screen_smem_bytes(d) <= 227 * 1024);
    ctx->use_screen5 = ctx->can_screen && screen5_supported(d);
    if (const char* e = cmf_hook("CMF_SCREEN_IMPL")) if (strcmp(e, "legacy") == 0) ctx->use_screen5 = false;
    if (ctx->use_screen5) ctx->nchunk_screen = screen5_pick_chunks(d, ctx->sm_count);
    if (const char* e = cmf_hook("CMF_SCREEN_CHUNKS")) ctx->nchunk_screen = std::max(1, atoi(e));   // tuning hook (tools build)
    if (const char* e = cmf_hook("CMF_EIGEN")) ctx->eigen_method = (strcmp(e, "jacobi") == 0) ? 1 : 0;
    }

    const size_t LS = (size_t)d.L * d.S;
    const int Sp = (d.S + 1) & ~1;
    const size_t frag = (size_t)(d.DP / 4) * 32;
    cudaError_t e = cudaSuccess;
    auto A_ = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    A_(dalloc(ctx, &ctx->xt, LS * d.DP));
    A_(dalloc(ctx, &ctx->mask, LS));
    A_(dalloc(ctx, &ctx->colsum_part, (size_t)ctx->nsplit * d.S * d.DP));
    A_(dalloc(ctx, &ctx->colcnt_part, (size_t)ctx->nsplit * d.S));
    A_(dalloc(ctx, &ctx->mu, (size_t)d.S * d.DP));
    A_(dalloc(ctx, &ctx->ctr, (size_t)d.S * d.DP));
    A_(dalloc(ctx, &ctx->n, (size_t)d.S));
    A_(dalloc(ctx, &ctx->P, (size_t)d.S * d.DP * d.DP));
    if (!wide) {
        A_(dalloc(ctx, &ctx->gram_part, gram_part_elems(d, ctx->nchunk_gram)));
        A_(dalloc(ctx, &ctx->Pf, (size_t)d.S * frag * d.NT));
        A_(dalloc(ctx, &ctx->Wf, (size_t)d.S * frag * d.NT2));
    } else {
        const size_t SDD = (size_t)d.S * d.DP * d.DP;
        A_(dalloc(ctx, &ctx->lo_part, (size_t)ctx->nsplit * d.S * d.DP));
        A_(dalloc(ctx, &ctx->hi_part, (size_t)ctx->nsplit * d.S * d.DP));
        A_(dalloc(ctx, &ctx->qexp, (size_t)d.S * d.DP));
        if (ctx->use_gram8) A_(dalloc(ctx, &ctx->img, wide_img_bytes(d)));
        A_(dalloc(ctx, &ctx->wgram, SDD));
        A_(dalloc(ctx, &ctx->wwork, SDD));
        A_(dalloc(ctx, &ctx->wdinv, (size_t)d.S * d.DP));
        A_(dalloc(ctx, &ctx->wdvec, (size_t)d.S * d.DP));
        A_(dalloc(ctx, &ctx->wevec, (size_t)d.S * d.DP));
        A_(dalloc(ctx, &ctx->rot, (size_t)d.S * wide_rot_cap(d)));
        A_(dalloc(ctx, &ctx->iters, (size_t)d.S * wide_iter_cap(d)));
        if (loo) {
            A_(dalloc(ctx, &ctx->Wtab, (size_t)d.S * d.DP * ctx->APW));
            A_(dalloc(ctx, &ctx->Zbuf, (size_t)ctx->zbatch * d.L * d.DP));
        }
    }
    A_(dalloc(ctx, &ctx->lam, (size_t)d.S * d.DP));
    A_(dalloc(ctx, &ctx->logdet, (size_t)d.S * d.AP));
    A_(dalloc(ctx, &ctx->beta, (size_t)d.S * d.AP));
    A_(dalloc(ctx, &ctx->status, (size_t)d.S));
    A_(dalloc(ctx, &ctx->sweeps, (size_t)d.S));
    A_(dalloc(ctx, &ctx->fpart, (size_t)d.S * ctx->nchunk_loo * d.AP));
    A_(dalloc(ctx, &ctx->nll, (size_t)d.S * d.A));
    A_(dalloc(ctx, &ctx->mindex, (size_t)d.S));
    A_(dalloc(ctx, &ctx->w, (size_t)d.S * d.DP));
    A_(dalloc(ctx, &ctx->wT, (size_t)d.DP * Sp));
    A_(dalloc(ctx, &ctx->c0, (size_t)Sp));
    A_(dalloc(ctx, &ctx->mf, LS));
    A_(dalloc(ctx, &ctx->stat_part, (size_t)ctx->nlanes * d.S * 2));
    A_(dalloc(ctx, &ctx->colstats, (size_t)3 * d.S));
    A_(dalloc(ctx, &ctx->alphas_d, (size_t)d.A));
    A_(dalloc(ctx, &ctx->abscf_d, (size_t)d.DP));
    A_(dalloc(ctx, &ctx->rsum, (size_t)d.S * d.AP));
    A_(dalloc(ctx, &ctx->betaf, (size_t)d.S * d.AP16));
    A_(dalloc(ctx, &ctx->sel_index, (size_t)d.S));
    A_(dalloc(ctx, &ctx->ncand, (size_t)d.S));
    A_(dalloc(ctx, &ctx->tile_mask, (size_t)d.S));
    A_(dalloc(ctx, &ctx->redo, (size_t)d.S));
    A_(dalloc(ctx, &ctx->probe, (size_t)d.S));
    A_(dalloc(ctx, &ctx->check, (size_t)d.S));
    A_(dalloc(ctx, &ctx->check_worst, (size_t)1));
    A_(dalloc(ctx, &ctx->tol_col, (size_t)d.S));
    A_(dalloc(ctx, &ctx->slogT, (size_t)d.S));
    A_(dalloc(ctx, &ctx->nuse, (size_t)d.S));
    A_(dalloc(ctx, &ctx->nentries, (size_t)d.S));
    A_(dalloc(ctx, &ctx->rejmask, (size_t)d.S));
    A_(dalloc(ctx, &ctx->flagmask, (size_t)d.S));
    A_(dalloc(ctx, &ctx->entries, (size_t)d.S * kMaxLabels));
    if (ctx->can_screen) {
        A_(dalloc(ctx, &ctx->Ws, (size_t)d.S * 2 * d.NT16 * d.NT * 32 * 4));
        A_(dalloc(ctx, &ctx->fscreen, (size_t)d.S * ctx->nchunk_screen * d.AP16));
        A_(dalloc(ctx, &ctx->Ps, (size_t)d.S * 2 * d.NT * d.NT * 32 * 2));
        if (ctx->use_screen5) A_(dalloc(ctx, &ctx->tab5, (size_t)d.S * screen5_table_floats(d)));
    }
    if (e != cudaSuccess) {
        free_buffers(ctx);
        return fail(ctx, CMF_E_NOMEM, std::string("device allocation failed: ") + cudaGetErrorString(e));
    }
    CK(cudaMemsetAsync(ctx->abscf_d, 0, (size_t)d.DP * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->wT, 0, (size_t)d.DP * Sp * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->c0, 0, (size_t)Sp * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(ctx->fpart, 0, (size_t)d.S * ctx->nchunk_loo * d.AP * sizeof(double), ctx->stream));
    CK(cudaMemcpyAsync(ctx->abscf_d, p->abscf, (size_t)D * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    if (loo)
        CK(cudaMemcpyAsync(ctx->alphas_d, p->alphas, (size_t)d.A * sizeof(double), cudaMemcpyHostToDevice,
                           ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_problem = true;
    return CMF_OK;
}

int cmf_upload_bil(cmf_ctx* ctx, const float* host_cube) {
    if (!ctx || !host_cube) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_upload_bil before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_own_slab(ctx);
    if (rc) return rc;
    const Dims& d = ctx->d;
    const size_t width = (size_t)d.D * d.S * sizeof(float);
    const float* src = host_cube + (size_t)(ctx->band_lo - 1) * d.S;
    CK(cudaMemcpy2DAsync(ctx->slab_own, width, src, (size_t)ctx->B * d.S * sizeof(float), width, (size_t)d.L,
                         cudaMemcpyHostToDevice, ctx->stream));
    ctx->have_input = true;
    return CMF_OK;
}

int cmf_upload_lines(cmf_ctx* ctx, const float* host_block, int32_t line0, int32_t nlines, int32_t band_first,
                     int32_t block_bands) {
    if (!ctx || !host_block) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_upload_lines before cmf_set_problem");
    const Dims& d0 = ctx->d;
    if (line0 < 0 || nlines <= 0 || line0 + nlines > d0.L) return fail(ctx, CMF_E_ARG, "line block outside the cube");
    if (band_first < 1 || band_first > ctx->band_lo || band_first + block_bands - 1 < ctx->band_hi)
        return fail(ctx, CMF_E_ARG, "the block does not hold the active window");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_own_slab(ctx);
    if (rc) return rc;
    const Dims& d = ctx->d;
    const size_t width = (size_t)d.D * d.S * sizeof(float);
    const float* src = host_block + (size_t)(ctx->band_lo - band_first) * d.S;
    CK(cudaMemcpy2DAsync(ctx->slab_own + (size_t)line0 * d.D * d.S, width, src,
                         (size_t)block_bands * d.S * sizeof(float), width, (size_t)nlines, cudaMemcpyHostToDevice,
                         ctx->stream));
    ctx->have_input = true;     // the caller is responsible for handing in every line before cmf_run()
    // an event behind this copy, keyed by the staging block, for cmf_upload_wait()
    cudaEvent_t ev = nullptr;
    for (auto& pe : ctx->stage_ev) if (pe.first == host_block) ev = pe.second;
    if (!ev) {
        if (ctx->stage_ev.size() >= 16) { ev = ctx->stage_ev.front().second; ctx->stage_ev.erase(ctx->stage_ev.begin()); }
        else CK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        ctx->stage_ev.emplace_back(host_block, ev);
    }
    CK(cudaEventRecord(ev, ctx->stream));
    return CMF_OK;
}

int cmf_upload_wait(cmf_ctx* ctx, const float* host_block) {
    if (!ctx) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    for (auto& pe : ctx->stage_ev)
        if (pe.first == host_block) CK(cudaEventSynchronize(pe.second));
    return CMF_OK;
}

int cmf_bind_device_slab(cmf_ctx* ctx, const float* dev_slab, int64_t line_pitch, int32_t band_pitch) {
    if (!ctx || !dev_slab) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_bind_device_slab before cmf_set_problem");
    Dims& d = ctx->d;
    if (band_pitch < d.S || line_pitch < (int64_t)band_pitch * (d.D - 1) + d.S)
        return fail(ctx, CMF_E_ARG, "pitches too small for the active slab");
    ctx->slab = dev_slab;
    d.line_pitch = line_pitch;
    d.band_pitch = band_pitch;
    d.vec2 = (d.S % 2 == 0) && (line_pitch % 2 == 0) && (band_pitch % 2 == 0) &&
             ((reinterpret_cast<uintptr_t>(dev_slab) & 7) == 0);
    ctx->have_input = true;
    return CMF_OK;
}

int cmf_set_labels(cmf_ctx* ctx, const int32_t* labels, int kmodes, int reject_min) {
    if (!ctx) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_set_labels before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    if (labels == nullptr) { ctx->have_labels = false; ctx->auto_cluster = false; return CMF_OK; }
    if (kmodes < 1 || kmodes > kMaxLabels) return fail(ctx, CMF_E_ARG, "kmodes must be 1..32");
    const Dims& d = ctx->d;
    const size_t LS = (size_t)d.L * d.S;
    for (size_t i = 0; i < LS; ++i)
        if (labels[i] < 0 || labels[i] >= kmodes)
            return fail(ctx, CMF_E_ARG, "labels must lie in 0..kmodes-1 (rejection is derived on the device)");
    int rc = ensure_mode_buffers(ctx);
    if (rc) return rc;
    ctx->auto_cluster = false;
    CK(cudaMemcpyAsync(ctx->labels_d, labels, LS * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->kmodes = kmodes;
    ctx->reject_min = reject_min > 0 ? reject_min : 0;
    ctx->have_labels = true;
    return CMF_OK;
}

int cmf_set_clustering(cmf_ctx* ctx, int kmodes, int pcadim, int reject_min, int max_iter) {
    if (!ctx) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_set_clustering before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    if (kmodes <= 1) { ctx->auto_cluster = false; ctx->have_labels = false; return CMF_OK; }
    if (kmodes > kMaxLabels) return fail(ctx, CMF_E_ARG, "kmodes must be 1..32");
    const Dims& d = ctx->d;
    if (pcadim < 1 || pcadim > kMaxPcaDim || pcadim > d.D)
        return fail(ctx, CMF_E_ARG, "pcadim must be 1..min(16, active bands)");
    int rc = ensure_mode_buffers(ctx);
    if (rc) return rc;
    if (!ctx->wide) {
        rc = ensure_full_gram(ctx);
        if (rc) return rc;
    }
    const size_t LS = (size_t)d.L * d.S;
    if (ctx->y_pd < pcadim) {      // (re)allocate the projection buffer; smaller requests reuse it
        cudaError_t e = cudaSuccess;
        auto A_ = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
        A_(dalloc(ctx, &ctx->ypca, LS * pcadim));
        A_(dalloc(ctx, &ctx->qpca, LS * pcadim));
        if (!ctx->lab8) {
            A_(dalloc(ctx, &ctx->lab8, LS));
            A_(dalloc(ctx, &ctx->pick, (size_t)d.S * kMaxPcaDim));
            A_(dalloc(ctx, &ctx->vtop, (size_t)d.S * d.DP * kMaxPcaDim));
            A_(dalloc(ctx, &ctx->km_iters, (size_t)d.S));
        }
        if (e != cudaSuccess) return fail(ctx, CMF_E_NOMEM, std::string("clustering buffers: ") + cudaGetErrorString(e));
        ctx->y_pd = pcadim;
    }
    ctx->kmodes = kmodes;
    ctx->pcadim = pcadim;
    ctx->reject_min = reject_min > 0 ? reject_min : 0;
    ctx->km_max_iter = max_iter > 0 ? max_iter : 100;
    ctx->auto_cluster = true;
    ctx->have_labels = false;
    return CMF_OK;
}

int cmf_set_regfull(cmf_ctx* ctx, int enable) {
    if (!ctx) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_set_regfull before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    if (enable && ctx->wide) {
        // the target in spectral form, its transpose and one scratch matrix per column (k_wide.cu, "-f on a wide window")
        const size_t SDD = (size_t)ctx->d.S * ctx->d.DP * ctx->d.DP;
        WideTarget& t = ctx->wtarget;
        if (!t.W && (dalloc(ctx, &t.W, SDD) != cudaSuccess || dalloc(ctx, &t.Wt, SDD) != cudaSuccess ||
                     dalloc(ctx, &t.tmp, SDD) != cudaSuccess || dalloc(ctx, &t.slogT, (size_t)ctx->d.S) != cudaSuccess ||
                     dalloc(ctx, &t.status, (size_t)ctx->d.S) != cudaSuccess))
            return fail(ctx, CMF_E_NOMEM, "full-column target buffers");
    } else if (enable) {
        int rc = ensure_full_gram(ctx);
        if (rc) return rc;
    }
    ctx->regfull = enable != 0;
    return CMF_OK;
}

int cmf_set_exclusion(cmf_ctx* ctx, const uint8_t* exclude) {
    if (!ctx) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_set_exclusion before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    if (!exclude) { ctx->have_excl = false; return CMF_OK; }
    const size_t LS = (size_t)ctx->d.L * ctx->d.S;
    if (!ctx->excl_sel) {
        cudaError_t e = dalloc(ctx, &ctx->excl_sel, LS);
        if (e != cudaSuccess) return fail(ctx, CMF_E_NOMEM, std::string("exclusion buffer: ") + cudaGetErrorString(e));
    }
    CK(cudaMemcpyAsync(ctx->excl_sel, exclude, LS, cudaMemcpyHostToDevice, ctx->stream));
    launch_invert_u8(ctx->excl_sel, (long long)LS, ctx->stream);      // exclude != 0  ->  sel = 0
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->have_excl = true;
    return CMF_OK;
}

int cmf_set_screen_margin(cmf_ctx* ctx, double rel_margin, int certify) {
    if (!ctx) return CMF_E_ARG;
    if (!(rel_margin > 0.0) || !(rel_margin < 1.0)) return fail(ctx, CMF_E_ARG, "screen margin must lie in (0, 1)");
    ctx->screen_tol = rel_margin;
    ctx->certify = certify ? 1 : 0;
    return CMF_OK;
}

int cmf_run(cmf_ctx* ctx, uint32_t flags) {
    if (!ctx) return CMF_E_ARG;
    if (!ctx->have_problem || !ctx->have_input) return fail(ctx, CMF_E_STATE, "cmf_run needs a problem and an input");
    CK(cudaSetDevice(ctx->device));
    return enqueue(ctx, (flags & CMF_RUN_TIMING) != 0, (flags & CMF_RUN_EXACT) != 0, nullptr, 0);
}

int cmf_sync(cmf_ctx* ctx) {
    if (!ctx) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return CMF_OK;
}

int cmf_run_host(cmf_ctx* ctx, const float* host_cube, double* mf_out, double* colstats_out,
                 int32_t* alpha_index_out, uint32_t flags) {
    if (!ctx || !host_cube) return CMF_E_ARG;
    if (!ctx->have_problem) return fail(ctx, CMF_E_STATE, "cmf_run_host before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_own_slab(ctx);
    if (rc) return rc;
    const Dims& d = ctx->d;
    // upload in blocks of one Gram chunk on the copy stream; the repack and Gram passes chase the copies
    const int lines_per_block = ctx->lpc_gram;
    const int nblocks = ctx->nchunk_gram;
    while ((int)ctx->blk_ev.size() < nblocks) {
        cudaEvent_t e;
        CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->blk_ev.push_back(e);
    }
    const size_t width = (size_t)d.D * d.S * sizeof(float);
    const size_t spitch = (size_t)ctx->B * d.S * sizeof(float);
    CK(cudaEventRecord(ctx->ev[K_COUNT], ctx->stream));           // order copies after earlier work
    CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev[K_COUNT], 0));
    std::vector<cudaEvent_t> ready;
    for (int b = 0; b < nblocks; ++b) {
        const int l0 = b * lines_per_block;
        const int nl = std::min(lines_per_block, d.L - l0);
        const float* src = host_cube + (size_t)l0 * ctx->B * d.S + (size_t)(ctx->band_lo - 1) * d.S;
        CK(cudaMemcpy2DAsync(ctx->slab_own + (size_t)l0 * d.D * d.S, width, src, spitch, width, (size_t)nl,
                             cudaMemcpyHostToDevice, ctx->copy_stream));
        CK(cudaEventRecord(ctx->blk_ev[b], ctx->copy_stream));
        ready.push_back(ctx->blk_ev[b]);
    }
    ctx->have_input = true;
    rc = enqueue(ctx, false, (flags & CMF_RUN_EXACT) != 0, &ready, lines_per_block);
    if (rc) return rc;
    if (mf_out)
        CK(cudaMemcpyAsync(mf_out, ctx->mf, (size_t)d.L * d.S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (colstats_out)
        CK(cudaMemcpyAsync(colstats_out, ctx->colstats, (size_t)3 * d.S * sizeof(double), cudaMemcpyDeviceToHost,
                           ctx->stream));
    if (alpha_index_out)
        CK(cudaMemcpyAsync(alpha_index_out, ctx->mindex, (size_t)d.S * sizeof(int), cudaMemcpyDeviceToHost,
                           ctx->stream));
    if (!(flags & CMF_RUN_ASYNC)) CK(cudaStreamSynchronize(ctx->stream));
    return CMF_OK;
}

const char* cmf_screen_kernel(const cmf_ctx* ctx) {
    if (!ctx || !ctx->have_problem || !ctx->can_screen || ctx->model != CMF_MODEL_LOOSHRINKAGE) return "";
    return ctx->use_screen5 ? "loo_screen5_kernel" : "loo_screen_kernel";
}

size_t cmf_output_bytes(const cmf_ctx* ctx, int what) {
    if (ctx && what == CMF_OUT_FLAGS) return ctx->flags_bytes;
    if (!ctx || !ctx->have_problem) return 0;
    return out_desc(ctx, what).bytes;
}

void* cmf_device_ptr(cmf_ctx* ctx, int what) {
    if (ctx && what == CMF_OUT_FLAGS) return ctx->flags_d;
    if (!ctx || !ctx->have_problem) return nullptr;
    return out_desc(ctx, what).ptr;
}

int cmf_download(cmf_ctx* ctx, int what, void* host_dst, size_t bytes) {
    if (!ctx || !host_dst) return CMF_E_ARG;
    if (!ctx->have_problem && what != CMF_OUT_FLAGS)
        return fail(ctx, CMF_E_STATE, "cmf_download before cmf_set_problem");
    CK(cudaSetDevice(ctx->device));
    const OutDesc o = out_desc(ctx, what);
    if (!o.ptr) return fail(ctx, CMF_E_ARG, "unknown output id");
    if (bytes < o.bytes) return fail(ctx, CMF_E_ARG, "destination buffer too small");
    CK(cudaMemcpyAsync(host_dst, o.ptr, o.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return CMF_OK;
}

// ---- products either side of the filter ----
int cmf_pixel_flags(cmf_ctx* ctx, const float* cube, int on_device, int32_t lines, int32_t bands, int32_t samples,
                    const cmf_flag_spec* spec, uint8_t* flags_host) {
    if (!ctx || !cube || !spec) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    if (lines <= 0 || bands <= 0 || samples <= 0) return fail(ctx, CMF_E_ARG, "bad cube shape");
    const int singles[4] = {spec->spec_band, spec->dark_band, spec->cloud_a, spec->cloud_b};
    if (spec->sat_lo < 0 || spec->sat_hi >= bands || spec->sat_hi < spec->sat_lo)
        return fail(ctx, CMF_E_ARG, "saturation window outside the cube");
    for (int b : singles)
        if (b >= bands) return fail(ctx, CMF_E_ARG, "flag band outside the cube");
    if ((spec->cloud_a < 0) != (spec->cloud_b < 0)) return fail(ctx, CMF_E_ARG, "cloud test needs both bands");
    const size_t LS = (size_t)lines * samples;
    if (ctx->flags_bytes != LS) {
        if (ctx->flags_d) { cudaFree(ctx->flags_d); ctx->flags_d = nullptr; ctx->flags_bytes = 0; }
        void* q = nullptr;
        if (cudaMalloc(&q, LS) != cudaSuccess) { cudaGetLastError(); return fail(ctx, CMF_E_NOMEM, "flags allocation failed"); }
        ctx->flags_d = reinterpret_cast<uint8_t*>(q);
        ctx->flags_bytes = LS;
    }
    FlagSpec f{spec->sat_lo, spec->sat_hi, spec->spec_band, spec->dark_band, spec->cloud_a, spec->cloud_b,
               spec->sat_thresh, spec->spec_thresh, spec->dark_thresh, spec->cloud_thresh, spec->cloud_dwl};
    float* packed = nullptr;
    if (on_device) {
        launch_pixel_flags(cube, (long long)bands * samples, samples, lines, samples, f, ctx->flags_d, ctx->stream);
    } else {
        // only the bands the tests read cross PCIe: the window as one strided copy, the single bands one each
        const int nsat = spec->sat_hi - spec->sat_lo + 1;
        const int nb = nsat + 4;
        void* q = nullptr;
        if (cudaMalloc(&q, LS * nb * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            return fail(ctx, CMF_E_NOMEM, "flag band buffer allocation failed");
        }
        packed = reinterpret_cast<float*>(q);
        const size_t row = (size_t)samples * sizeof(float);
        cudaError_t e = cudaMemcpy2DAsync(packed, row * nb, cube + (size_t)spec->sat_lo * samples, row * bands,
                                          row * nsat, lines, cudaMemcpyHostToDevice, ctx->stream);
        int idx[4];
        for (int k = 0; k < 4 && e == cudaSuccess; ++k) {
            idx[k] = singles[k] < 0 ? -1 : nsat + k;
            if (singles[k] >= 0)
                e = cudaMemcpy2DAsync(packed + (size_t)(nsat + k) * samples, row * nb,
                                      cube + (size_t)singles[k] * samples, row * bands, row, lines,
                                      cudaMemcpyHostToDevice, ctx->stream);
        }
        if (e != cudaSuccess) { cudaFree(packed); return fail(ctx, CMF_E_CUDA, std::string("flag band upload: ") + cudaGetErrorString(e)); }
        f.sat_lo = 0; f.sat_hi = nsat - 1;
        f.spec_band = idx[0]; f.dark_band = idx[1]; f.cloud_a = idx[2]; f.cloud_b = idx[3];
        launch_pixel_flags(packed, (long long)nb * samples, samples, lines, samples, f, ctx->flags_d, ctx->stream);
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess && flags_host)
        e = cudaMemcpyAsync(flags_host, ctx->flags_d, LS, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (packed) cudaFree(packed);
    if (e != cudaSuccess) return fail(ctx, CMF_E_CUDA, std::string("cmf_pixel_flags: ") + cudaGetErrorString(e));
    return CMF_OK;
}

namespace {
int column_profile_impl(cmf_ctx* ctx, const double* mf_dev, int L, int S, double nodata, int robust, double p,
                        double* out_host) {
    void *colv = nullptr, *out = nullptr;
    if (cudaMalloc(&colv, (size_t)L * S * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&out, (size_t)5 * S * sizeof(double)) != cudaSuccess) {
        cudaGetLastError();
        if (colv) cudaFree(colv);
        return fail(ctx, CMF_E_NOMEM, "profile scratch allocation failed");
    }
    // the percentiles the reference asks numpy for: q = (1 - p) * 100 and p * 100, then q / 100 inside numpy
    const double qlo = ((1.0 - p) * 100.0) / 100.0, qhi = (p * 100.0) / 100.0;
    launch_column_profile(mf_dev, L, S, nodata, robust, qlo, qhi, reinterpret_cast<float*>(colv),
                          reinterpret_cast<double*>(out), ctx->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(out_host, out, (size_t)5 * S * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(colv);
    cudaFree(out);
    if (e != cudaSuccess) return fail(ctx, CMF_E_CUDA, std::string("cmf_column_profile: ") + cudaGetErrorString(e));
    return CMF_OK;
}
}  // namespace

int cmf_column_profile(cmf_ctx* ctx, int robust, double p, double* out_host) {
    if (!ctx || !out_host) return CMF_E_ARG;
    if (!ctx->have_problem || !ctx->mf) return fail(ctx, CMF_E_STATE, "cmf_column_profile before a run");
    if (robust && !(p > 0.0 && p < 1.0)) return fail(ctx, CMF_E_ARG, "percentile fraction must be inside (0, 1)");
    CK(cudaSetDevice(ctx->device));
    return column_profile_impl(ctx, ctx->mf, ctx->d.L, ctx->d.S, ctx->nodata, robust, p, out_host);
}

int cmf_column_profile_image(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples, double nodata,
                             int robust, double p, double* out_host) {
    if (!ctx || !mf_host || !out_host) return CMF_E_ARG;
    if (lines <= 0 || samples <= 0) return fail(ctx, CMF_E_ARG, "bad image shape");
    if (robust && !(p > 0.0 && p < 1.0)) return fail(ctx, CMF_E_ARG, "percentile fraction must be inside (0, 1)");
    CK(cudaSetDevice(ctx->device));
    void* img = nullptr;
    const size_t bytes = (size_t)lines * samples * sizeof(double);
    if (cudaMalloc(&img, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, CMF_E_NOMEM, "image allocation failed"); }
    cudaError_t e = cudaMemcpyAsync(img, mf_host, bytes, cudaMemcpyHostToDevice, ctx->stream);
    int rc = CMF_OK;
    if (e != cudaSuccess) rc = fail(ctx, CMF_E_CUDA, std::string("image upload: ") + cudaGetErrorString(e));
    else rc = column_profile_impl(ctx, reinterpret_cast<const double*>(img), lines, samples, nodata, robust, p, out_host);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(img);
    return rc;
}


// ---- detection pre-filter and CNN input ----
int cmf_detection_prefilter(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples,
                            const double* weights, int32_t radius, double mfmin, double mfmax, double* detkde_host,
                            uint8_t* ch4min_host, uint8_t* detmask_host) {
    if (!ctx || !weights || !detkde_host) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const double* src = nullptr;
    int L = lines, S = samples;
    if (mf_host == nullptr) {            // the scores of the last run, already on the device
        if (!ctx->have_problem || !ctx->mf) return fail(ctx, CMF_E_STATE, "cmf_detection_prefilter: no scores on the device");
        L = ctx->d.L; S = ctx->d.S; src = ctx->mf;
    }
    if (L <= 0 || S <= 0 || radius < 0 || radius > 4096) return fail(ctx, CMF_E_ARG, "bad image shape or filter radius");
    if (!(mfmax > mfmin)) return fail(ctx, CMF_E_ARG, "mfmax must exceed mfmin");
    const size_t n = (size_t)L * S;
    const int nparts = 296;
    void* blk = nullptr;
    // one allocation: [image (host input only)] tmp, blur, detkde, partials, weights, the two byte masks
    const size_t dbl = (mf_host ? n : 0) + 3 * n + 3 * nparts + (size_t)(2 * radius + 1);
    if (cudaMalloc(&blk, dbl * sizeof(double) + 2 * n) != cudaSuccess) { cudaGetLastError(); return fail(ctx, CMF_E_NOMEM, "pre-filter scratch allocation failed"); }
    double* p = reinterpret_cast<double*>(blk);
    double* img = nullptr;
    if (mf_host) { img = p; p += n; }
    double *tmp = p, *blur = p + n, *det = p + 2 * n, *part = p + 3 * n, *wd = part + 3 * nparts;
    uint8_t* m1 = reinterpret_cast<uint8_t*>(wd + (2 * radius + 1));
    uint8_t* m2 = m1 + n;
    cudaError_t e = cudaMemcpyAsync(wd, weights, (size_t)(2 * radius + 1) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && mf_host) {
        e = cudaMemcpyAsync(img, mf_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        src = img;
    }
    if (e == cudaSuccess) {
        launch_detection_prefilter(src, L, S, radius, wd, mfmin, mfmax, tmp, blur, part, det, m1, m2, ctx->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(detkde_host, det, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && ch4min_host) e = cudaMemcpyAsync(ch4min_host, m1, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess && detmask_host) e = cudaMemcpyAsync(detmask_host, m2, n, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(blk);
    if (e != cudaSuccess) return fail(ctx, CMF_E_CUDA, std::string("cmf_detection_prefilter: ") + cudaGetErrorString(e));
    return CMF_OK;
}

int cmf_cnn_input(cmf_ctx* ctx, const double* mf_host, int32_t lines, int32_t samples, float vmin, float vmax,
                  float mean, float stdv, float* out_host) {
    if (!ctx || !out_host) return CMF_E_ARG;
    CK(cudaSetDevice(ctx->device));
    const double* src = nullptr;
    int L = lines, S = samples;
    if (mf_host == nullptr) {
        if (!ctx->have_problem || !ctx->mf) return fail(ctx, CMF_E_STATE, "cmf_cnn_input: no scores on the device");
        L = ctx->d.L; S = ctx->d.S; src = ctx->mf;
    }
    if (L <= 0 || S <= 0) return fail(ctx, CMF_E_ARG, "bad image shape");
    if (!(vmax > vmin)) return fail(ctx, CMF_E_ARG, "vmax must exceed vmin (ClampCH4)");
    const size_t n = (size_t)L * S;
    const size_t out_bytes = (n * sizeof(float) + 15) & ~(size_t)15;
    void* blk = nullptr;
    if (cudaMalloc(&blk, out_bytes + (mf_host ? n * sizeof(double) : 0)) != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, CMF_E_NOMEM, "CNN input scratch allocation failed");
    }
    float* out = reinterpret_cast<float*>(blk);
    cudaError_t e = cudaSuccess;
    if (mf_host) {
        double* img = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(blk) + out_bytes);
        e = cudaMemcpyAsync(img, mf_host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
        src = img;
    }
    if (e == cudaSuccess) {
        launch_cnn_input(src, nullptr, (long long)n, vmin, vmax, mean, stdv, out, ctx->stream);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, out, n * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(blk);
    if (e != cudaSuccess) return fail(ctx, CMF_E_CUDA, std::string("cmf_cnn_input: ") + cudaGetErrorString(e));
    return CMF_OK;
}

// ---- the importable looshrinkage(I_zm, alphas, nll, n, I_reg = []) of the reference (cmf/robust_mf.py:92-136) ----
int cmf_looshrinkage(cmf_ctx* ctx, const double* I_zm, int32_t rows, int32_t D, const double* alphas, int32_t A,
                     int32_t n, const double* I_reg, int32_t reg_rows, double* nll_out, double* C_out,
                     int32_t* mindex_out) {
    if (!ctx || !I_zm || !alphas || !nll_out || !C_out || !mindex_out) return CMF_E_ARG;
    if (rows < 1 || D < 1 || D > 1024 || A < 1 || A > 4096) return fail(ctx, CMF_E_ARG, "cmf_looshrinkage: bad shape");
    const bool reg = I_reg != nullptr && reg_rows > 0;
    if (reg && reg_rows < 2) return fail(ctx, CMF_E_ARG, "cmf_looshrinkage: I_reg needs at least two rows");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    Dims d{};
    d.L = rows; d.S = 1; d.D = D; d.NT = (D + 7) / 8; d.DP = 8 * d.NT;
    d.A = A; d.NT2 = (A + 7) / 8; d.AP = d.NT2 * 8; d.NT16 = (A + 15) / 16; d.AP16 = d.NT16 * 16;
    if (!wide_eigen_fits(d)) return fail(ctx, CMF_E_ARG, "cmf_looshrinkage: too many bands for the eigen-solver");
    const int DP = d.DP, APW = (d.AP + 63) / 64 * 64;
    const size_t DD = (size_t)DP * DP;
    // one scratch block; doubles first
    size_t nd = (size_t)rows * DP /*x*/ + (size_t)rows * DP /*Z*/ + 3 * DD /*gram, work, P*/ + 8 * (size_t)DP + 5 * (size_t)d.AP +
                (size_t)DP * APW /*W*/ + (size_t)A /*alphas*/ + (size_t)A /*nll*/ + (size_t)D * D /*C*/ + 3 * (size_t)DP /*w, wT*/ + 8;
    if (reg) nd += (size_t)reg_rows * DP /*I_reg*/ + 4 * DD /*its gram, W, W^T, scratch*/ + (size_t)DP + 2;
    const size_t rot_cap = wide_rot_cap(d);
    const int iter_cap = wide_iter_cap(d);
    const size_t bytes = nd * sizeof(double) + rot_cap * sizeof(double2) + (size_t)iter_cap * sizeof(int2) + 16 * sizeof(int);
    void* blk = nullptr;
    if (cudaMalloc(&blk, bytes) != cudaSuccess) { cudaGetLastError(); return fail(ctx, CMF_E_NOMEM, "cmf_looshrinkage: scratch allocation failed"); }
    cudaError_t e = cudaMemsetAsync(blk, 0, bytes, st);
    double* p = reinterpret_cast<double*>(blk);
    auto take = [&](size_t k) { double* q = p; p += k; return q; };
    double *x = take((size_t)rows * DP), *Z = take((size_t)rows * DP), *gram = take(DD), *work = take(DD), *P = take(DD);
    double *mean = take(DP), *zero = take(DP), *dinv = take(DP), *dvec = take(DP), *evec = take(DP), *lam = take(DP),
           *abscf = take(DP), *mu0 = take(DP);
    double *logdet = take(d.AP), *beta = take(d.AP), *rsum = take(d.AP), *fpart = take(d.AP), *spare = take(d.AP);
    double *W = take((size_t)DP * APW), *al_d = take(A), *nll_d = take(A), *C_d = take((size_t)D * D), *w = take(DP),
           *wT = take(2 * (size_t)DP), *slogT = take(1), *c0 = take(3);
    (void)spare; (void)zero;
    double *xr = nullptr, *gramr = nullptr, *meanr = nullptr;
    WideTarget tgt;
    if (reg) {
        xr = take((size_t)reg_rows * DP); gramr = take(DD); tgt.W = take(DD); tgt.Wt = take(DD); tgt.tmp = take(DD);
        meanr = take(DP); tgt.slogT = take(2);
    }
    if (reinterpret_cast<uintptr_t>(p) & 15) ++p;                       // double2 alignment
    double2* rot = reinterpret_cast<double2*>(p);
    int2* iters = reinterpret_cast<int2*>(rot + rot_cap);
    int* ints = reinterpret_cast<int*>(iters + iter_cap);
    int *n_rows = ints, *n_loo = ints + 1, *status = ints + 2, *niter = ints + 3, *mindex = ints + 4;
    int* n_reg = ints + 5;
    tgt.status = ints + 6;
    const int hv[6] = {rows, n, 0, 0, 0, reg_rows};
    if (reg && e == cudaSuccess)
        e = cudaMemcpy2DAsync(xr, (size_t)DP * sizeof(double), I_reg, (size_t)D * sizeof(double), (size_t)D * sizeof(double),
                              (size_t)reg_rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpy2DAsync(x, (size_t)DP * sizeof(double), I_zm, (size_t)D * sizeof(double),
                                                (size_t)D * sizeof(double), (size_t)rows, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(al_d, alphas, (size_t)A * sizeof(double), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(ints, hv, sizeof(hv), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        launch_wide_mean64(x, rows, D, DP, mean, st);                       // numpy.cov re-centres (:68)
        launch_wide_gram64_f64(rows, DP, 1, x, mean, gram, st);             // sum (x - mean)(x - mean)^T
        if (reg) {
            // T = cov(I_reg) (:99): its spectral factor first, then S whitened with it
            launch_wide_mean64(xr, reg_rows, D, DP, meanr, st);
            launch_wide_gram64_f64(reg_rows, DP, 1, xr, meanr, gramr, st);
            launch_wide_eigen(d, gramr, n_reg, meanr, meanr, nullptr, 0, work, dinv, dvec, evec, rot, iters, niter, P, lam,
                              slogT, status, st);
            launch_wide_target(d, P, lam, slogT, status, tgt, st);
        }
        // S = G / (m - 1), T = diag(S) unless given; the x100 stability scaling enters log det only (:94-99)
        launch_wide_eigen(d, gram, n_rows, mean, mean, nullptr, reg ? 2 : 0, work, dinv, dvec, evec, rot, iters, niter, P,
                          lam, slogT, status, st, reg ? &tgt : nullptr);
        launch_wide_tables(d, APW, n_rows, n_loo, al_d, 0, lam, slogT, logdet, beta, rsum, W, st);
        // r_k = x_k^T G^-1 x_k uses the samples as given, not re-centred (:114)
        launch_wide_loo_f64(rows, D, DP, d.AP, APW, x, mu0, P, W, beta, n_rows, Z, fpart, st);
        launch_finalize(d, fpart, 1, logdet, n_rows, al_d, P, lam, mu0, abscf, 0, 1, 1.0, nll_d, mindex, w, wT, c0, status,
                        nullptr, nullptr, n_loo, st);
        launch_wide_cmat(gram, rows, D, DP, mindex, al_d, C_d, st, gramr, reg_rows);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(nll_out, nll_d, (size_t)A * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(C_out, C_d, (size_t)D * D * sizeof(double), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(mindex_out, mindex, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(blk);
    if (e != cudaSuccess) return fail(ctx, CMF_E_CUDA, std::string("cmf_looshrinkage: ") + cudaGetErrorString(e));
    return CMF_OK;
}

int cmf_kernel_count(void) { return K_COUNT; }
const char* cmf_kernel_name(int i) { return (i >= 0 && i < K_COUNT) ? kKernelNames[i] : ""; }

int cmf_kernel_times(cmf_ctx* ctx, float* ms, int n) {
    if (!ctx || !ms) return CMF_E_ARG;
    if (ctx->timed_runs == 0) return fail(ctx, CMF_E_STATE, "no cmf_run(CMF_RUN_TIMING) since the last query");
    CK(cudaSetDevice(ctx->device));
    const int m = std::min(n, (int)K_COUNT);
    for (int i = 0; i < m; ++i) ms[i] = 0.f;
    for (int r = 0; r < ctx->timed_runs; ++r) {
        cudaEvent_t* evs = ctx->ev_sets[r].data();
        CK(cudaEventSynchronize(evs[K_COUNT]));
        for (int i = 0; i < m; ++i) {
            float t = 0.f;
            CK(cudaEventElapsedTime(&t, evs[i], evs[i + 1]));
            ms[i] += t;
        }
    }
    for (int i = 0; i < m; ++i) ms[i] /= (float)ctx->timed_runs;
    ctx->timed_runs = 0;
    return m;
}

int cmf_launch_count(const cmf_ctx* ctx) { return ctx ? ctx->launches : 0; }

void* cmf_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void cmf_host_free(void* p) { if (p) cudaFreeHost(p); }
int cmf_host_register(void* p, size_t bytes) {
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? CMF_OK : CMF_E_CUDA;
}
int cmf_host_unregister(void* p) { return cudaHostUnregister(p) == cudaSuccess ? CMF_OK : CMF_E_CUDA; }

}  // extern "C"
