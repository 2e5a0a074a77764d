"""TEST INFRASTRUCTURE -- CPU restatement of the reference's columnwise matched filter.

This file is the parity oracle for the CUDA path.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may
import it.  The product package ``srcfinder_b200`` never does (and has no CPU fallback).

It restates ``/root/reference/cmf/robust_mf.py`` with the same NumPy/SciPy (LAPACK) calls
the reference makes -- ``numpy.cov(ddof=1)``, ``scipy.linalg.det``, ``scipy.linalg.inv`` --
so that rounding behaviour (and therefore the selected alpha index) is the reference's.
Each function cites the reference lines it follows.

Pinning: the reference has no tests or golden vectors for this path, so the
restatement is pinned against outputs of the UNMODIFIED reference script executed in the
build container (``oracle/make_golden.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks this module against those files.
"""
from __future__ import annotations

import contextlib

import numpy as np
from numpy.linalg import LinAlgError  # scipy.linalg.LinAlgError is this class
from scipy.linalg import det as _lapack_det
from scipy.linalg import inv as _lapack_inv

try:                                   # tiny-matrix LAPACK calls thrash with many BLAS threads
    from threadpoolctl import threadpool_limits as _blas_limits   # (~40x slower)
except Exception:                      # pragma: no cover
    _blas_limits = None


def single_thread_blas():
    """Context manager pinning BLAS/LAPACK to one thread (also makes the last bits reproducible)."""
    return _blas_limits(limits=1) if _blas_limits is not None else contextlib.nullcontext()


PPM_SCALING = 100000.0        # cmf/robust_mf.py:38
STABILITY_SCALING = 100.0     # cmf/robust_mf.py:94


def alpha_grid(aminexp=-10.0, amaxexp=0.0, astep=0.05):
    """201 shrinkage candidates 1e-10 .. 1 (cmf/robust_mf.py:241-243)."""
    return 10.0 ** np.arange(aminexp, amaxexp + astep, astep)


def active_window(library_name, reflectance=False):
    """1-based inclusive band window chosen from the library file name (cmf/robust_mf.py:186-194)."""
    if reflectance and "ch4" in library_name:
        return [5, 420]
    if "ch4" in library_name:
        return [351, 422]
    if "co2" in library_name:
        return [309, 391]
    raise ValueError("could not set active range")


def min_cluster_samples(active):
    """cmf/robust_mf.py:200."""
    return int((active[1] - active[0]) * 1.2)


def valid_rows(col_ld):
    """Indices of pixels whose every active band is finite and not negative (cmf/robust_mf.py:282)."""
    ok = (~(col_ld < 0)) & np.isfinite(col_ld)
    return np.where(ok.all(axis=1))[0]


def cov_ddof1(a_nd):
    """``numpy.cov`` of samples-in-rows data with ddof=1 (cmf/robust_mf.py:52-70)."""
    return np.cov(a_nd.T, ddof=1)


def loo_shrinkage(x_zm, alphas, nll, n, x_reg=None):
    """Leave-one-out shrinkage search, Theiler 2012 eq. 29 (cmf/robust_mf.py:92-136).

    ``x_zm``  (m, D) mean-removed samples; ``n`` is what the caller passes as the sample count
    (the reference passes the column's valid count even for a cluster subset, :355-356).
    Fills ``nll`` in place; returns ``(C, mindex)``.
    """
    d = x_zm.shape[1]
    xs = x_zm * STABILITY_SCALING
    s_mat = cov_ddof1(xs)
    have_reg = x_reg is not None and len(x_reg) != 0
    t_mat = cov_ddof1(x_reg * STABILITY_SCALING) if have_reg else np.diag(np.diag(s_mat))
    const = d * np.log(2.0 * np.pi)
    nll[:] = np.inf
    for i, alpha in enumerate(alphas):
        try:
            beta = (1.0 - alpha) / (n - 1.0)
            g = n * (beta * s_mat) + (alpha * t_mat)
            g_det = _lapack_det(g, overwrite_a=False, check_finite=False)
            if g_det == 0:
                continue
            g_inv = _lapack_inv(g, overwrite_a=False, check_finite=False)
            r = (xs.dot(g_inv) * xs).sum(axis=1)
            q = 1.0 - beta * r
            nll[i] = 0.5 * (const + np.log(g_det)) + 1.0 / (2.0 * n) * (np.log(q) + (r / q)).sum()
        except LinAlgError:
            pass
    mindex = int(np.argmin(nll))       # NaN entries win argmin, first one -- numpy semantics
    if nll[mindex] != np.inf:
        alpha = alphas[mindex]
    else:
        mindex, alpha = -1, 0.0
    s_fin = cov_ddof1(x_zm)
    t_fin = cov_ddof1(x_reg) if have_reg else np.diag(np.diag(s_fin))
    return (1.0 - alpha) * s_fin + alpha * t_fin, mindex


def column_filter(x_nd, abscf, alphas, nll, n_for_loo, model="looshrinkage", reflectance=False,
                  x_reg=None):
    """One background mode: mean, model fit, inverse, matched filter (cmf/robust_mf.py:346-386).

    Returns dict(mf, mu, alpha_index, weights, singular).  ``mf`` is already scaled (x1e5 for
    radiance).  ``weights`` is ``Cinv t / (t Cinv t)`` (times the ppm scaling) for debugging parity.
    """
    mu = x_nd.mean(axis=0)
    xc = x_nd - mu
    alpha_index = -2
    try:
        if model == "empirical":
            cinv = _lapack_inv(cov_ddof1(xc), overwrite_a=False, check_finite=False)
        else:
            c_mat, alpha_index = loo_shrinkage(xc, alphas, nll, n_for_loo, x_reg=x_reg)
            cinv = _lapack_inv(c_mat, overwrite_a=False, check_finite=False)
    except LinAlgError:
        return dict(mf=np.zeros(x_nd.shape[0]), mu=mu, alpha_index=alpha_index,
                    weights=np.zeros_like(mu), singular=True)
    target = abscf - mu if reflectance else abscf * mu
    normalizer = target.dot(cinv).dot(target.T)
    mf = (xc.dot(cinv).dot(target.T)) / normalizer
    scale = 1.0 if reflectance else PPM_SCALING
    w = cinv.dot(target) / normalizer * scale
    return dict(mf=mf * scale, mu=mu, alpha_index=alpha_index, weights=w, singular=False)


def cmf_cube(cube_lbs, abscf, active, alphas=None, model="looshrinkage", reflectance=False,
             nodata=-9999.0, labels=None, reject_min=None, regfull=False, columns=None,
             keep_nll=False):
    with single_thread_blas():
        return _cmf_cube(cube_lbs, abscf, active, alphas, model, reflectance, nodata, labels, reject_min,
                         regfull, columns, keep_nll)


def _cmf_cube(cube_lbs, abscf, active, alphas, model, reflectance, nodata, labels, reject_min, regfull,
              columns, keep_nll):
    """Column loop of the reference (cmf/robust_mf.py:297-397) on an in-memory BIL cube.

    ``cube_lbs``  (L, B, S) float32;  ``active`` 1-based inclusive [lo, hi];
    ``labels``    optional (L, S) int array of background-mode labels for the valid pixels
                  (the reference gets these from an unseeded MiniBatchKMeans, :306-332; here they
                  are an input so the path is deterministic).  Labels < 0 are rejected pixels.
                  When given, modes are processed in ``np.unique`` order like :339-344.
    Returns dict(mf (L,S) f64 with nodata, mask (L,S) bool, colnum/colavg/colstd (S,),
                 alpha_index (S,) of the last fitted mode, mu (S,D), weights (S,D), nll (S,A)).
    """
    L, B, S = cube_lbs.shape
    lo, hi = active
    D = hi - lo + 1
    alphas = alpha_grid() if alphas is None else np.asarray(alphas, dtype=np.float64)
    nll = np.zeros(len(alphas))
    mf = np.full((L, S), float(nodata))
    mask = np.zeros((L, S), dtype=bool)
    colnum = np.full(S, float(nodata))
    colavg = np.full(S, float(nodata))
    colstd = np.full(S, float(nodata))
    aidx = np.full(S, -2, dtype=np.int32)
    mus = np.zeros((S, D))
    wts = np.zeros((S, D))
    nlls = np.full((S, len(alphas)), np.nan) if keep_nll else None
    for col in (range(S) if columns is None else columns):
        full = cube_lbs[:, lo - 1:hi, col]
        use = valid_rows(full)
        x = np.float64(full[use, :])
        nuse = x.shape[0]
        if nuse == 0:
            continue
        mask[use, col] = True
        if labels is None:
            lab = np.ones(nuse, dtype=np.int64)
            ulab = np.array([1])
        else:
            lab = np.asarray(labels[use, col], dtype=np.int64).copy()
            ulab = np.unique(lab)
            if reject_min is not None:
                # relabel small clusters negative (cmf/robust_mf.py:316-324).  The mode list keeps
                # its original order with flipped entries, and label 0 cannot flip (-0 == 0).
                for i, l in enumerate(ulab.copy()):
                    m = lab == l
                    if m.sum() < reject_min:
                        lab[m] = -l
                        ulab[i] = -l
            if (ulab < 0).all():                                      # :330-332
                lab, ulab = np.abs(lab), np.abs(ulab)
        for ki in ulab:
            kmask = (lab == ki) if ki >= 0 else (lab >= 0)
            xk = x if labels is None else x[kmask, :]
            xreg = None
            if regfull and labels is not None and model == "looshrinkage":
                xreg = x - xk.mean(axis=0)
            res = column_filter(xk, abscf, alphas, nll, nuse, model=model,
                                reflectance=reflectance, x_reg=xreg)
            mf[use[kmask], col] = res["mf"]
            aidx[col] = res["alpha_index"]
            mus[col], wts[col] = res["mu"], res["weights"]
            if keep_nll and model == "looshrinkage":
                nlls[col] = nll
        pix = mf[use[lab >= 0], col]
        colnum[col] = nuse
        colavg[col] = np.mean(pix)
        colstd[col] = np.std(pix)
    out = dict(mf=mf, mask=mask, colnum=colnum, colavg=colavg, colstd=colstd,
               alpha_index=aidx, mu=mus, weights=wts)
    if keep_nll:
        out["nll"] = nlls
    return out


def assemble_product(cube_lbs, mf, colnum, rgb_bands=(60, 42, 24), nodata=-9999.0):
    """4-band (L, S, 4) float64 BIP product: RGB radiance copies + MF (cmf/robust_mf.py:266, 394-397).

    RGB is copied for every column that had at least one valid pixel (columns with none are
    skipped by the ``continue`` at :303-304 and stay zero).
    """
    L, B, S = cube_lbs.shape
    nb = 4 if len(rgb_bands) == 3 else 1
    out = np.zeros((L, S, nb))
    out[:, :, -1] = mf
    if nb == 4:
        done = colnum != float(nodata)
        for i, b in enumerate(rgb_bands):
            out[:, done, i] = cube_lbs[:, b, :][:, done]
    return out
