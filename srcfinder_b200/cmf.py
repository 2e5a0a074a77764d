"""Host-side driver of the CUDA columnwise matched filter: one object per (GPU, problem shape).

Mirrors the inputs of the reference's column loop (cmf/robust_mf.py:297-397): a float32 BIL cube, the
1-based active band window, the unit-absorption coefficients over that window, the model name and the
alpha grid.  All arithmetic happens in libcmf_b200.so; this module only moves pointers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

PPM_SCALING = 100000.0


def alpha_grid(aminexp=-10.0, amaxexp=0.0, astep=0.05):
    """The reference's 201 shrinkage candidates (cmf/robust_mf.py:241-243)."""
    return 10.0 ** np.arange(aminexp, amaxexp + astep, astep)


class CmfError(RuntimeError):
    pass


class ColumnwiseMF(object):
    """GPU columnwise matched filter for cubes of one shape."""

    def __init__(self, lines, bands, samples, active, abscf, model="looshrinkage", reflectance=False,
                 alphas=None, nodata=-9999.0, device=0, stream=None):
        self._lib = _lib.load()
        self._ctx = C.c_void_p()
        rc = self._lib.cmf_create(C.byref(self._ctx), int(device))
        if rc != 0:
            raise CmfError("cmf_create failed (%d): %s" % (rc, self._lib.cmf_last_error(None).decode()))
        self.L, self.B, self.S = int(lines), int(bands), int(samples)
        self.active = [int(active[0]), int(active[1])]
        self.D = self.active[1] - self.active[0] + 1
        self.DP = (self.D + 7) // 8 * 8
        self.model = model
        self.reflectance = bool(reflectance)
        self.nodata = float(nodata)
        self.device = int(device)
        if model not in ("looshrinkage", "empirical"):
            raise CmfError("unknown model %r" % (model,))
        self.alphas = np.ascontiguousarray(alpha_grid() if alphas is None else alphas, dtype=np.float64)
        self.A = len(self.alphas) if model == "looshrinkage" else 1
        ab = np.ascontiguousarray(abscf, dtype=np.float64)
        if ab.shape != (self.D,):
            raise CmfError("abscf must have %d entries (active window), got %r" % (self.D, ab.shape))
        p = _lib.Problem()
        p.lines, p.bands, p.samples, p.interleave = self.L, self.B, self.S, 0
        p.band_lo, p.band_hi = self.active
        p.reflectance = int(self.reflectance)
        p.model = _lib.MODEL_LOOSHRINKAGE if model == "looshrinkage" else _lib.MODEL_EMPIRICAL
        p.num_alphas = len(self.alphas)
        p.nodata = self.nodata
        p.alphas = self.alphas.ctypes.data_as(C.POINTER(C.c_double))
        p.abscf = ab.ctypes.data_as(C.POINTER(C.c_double))
        if stream is not None:
            self._check(self._lib.cmf_set_stream(self._ctx, C.c_void_p(int(stream))))
        self._check(self._lib.cmf_set_problem(self._ctx, C.byref(p)))
        self._keep = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        if rc != 0:
            raise CmfError("libcmf_b200 error %d: %s" % (rc, self._lib.cmf_last_error(self._ctx).decode()))

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx:
            self._lib.cmf_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------ input
    def upload(self, cube_lbs):
        """Host (L, B, S) float32 BIL cube -> device (only the active window is copied)."""
        cube = np.ascontiguousarray(cube_lbs, dtype=np.float32)
        if cube.shape != (self.L, self.B, self.S):
            raise CmfError("cube shape %r != %r" % (cube.shape, (self.L, self.B, self.S)))
        self._keep = cube
        self._check(self._lib.cmf_upload_bil(self._ctx, C.c_void_p(cube.ctypes.data)))

    def upload_stream(self, cube_lbs, block_lines=256):
        """Upload a (L, B, S) float32 BIL cube that lives on disk (``numpy.memmap``) block by block: the active window
        of the next ``block_lines`` lines is read into one of two pinned staging buffers while the previous block
        is on the PCIe link, so disk and PCIe overlap.  Only the active window is read."""
        if cube_lbs.shape != (self.L, self.B, self.S):
            raise CmfError("cube shape %r != %r" % (cube_lbs.shape, (self.L, self.B, self.S)))
        lo, hi = self.active
        nb = int(max(1, min(block_lines, self.L)))
        nbytes = nb * self.D * self.S * 4
        ptrs = [self._lib.cmf_host_alloc(nbytes) for _ in range(2)]
        if not all(ptrs):
            for p in ptrs:
                if p:
                    self._lib.cmf_host_free(C.c_void_p(p))
            raise CmfError("pinned staging buffers (%d bytes each) could not be allocated" % nbytes)
        try:
            views = [np.ctypeslib.as_array((C.c_float * (nb * self.D * self.S)).from_address(p)).reshape(nb, self.D, self.S)
                     for p in ptrs]
            for i, l0 in enumerate(range(0, self.L, nb)):
                l1 = min(self.L, l0 + nb)
                k = i % 2
                if i >= 2:
                    # only the copy that last read THIS staging block has to be done (an event per block); the copy
                    # of the other block, enqueued after it, stays in flight while this block is refilled from disk
                    self._check(self._lib.cmf_upload_wait(self._ctx, C.c_void_p(ptrs[k])))
                np.copyto(views[k][:l1 - l0], cube_lbs[l0:l1, lo - 1:hi, :], casting="same_kind")
                self._check(self._lib.cmf_upload_lines(self._ctx, C.c_void_p(ptrs[k]), l0, l1 - l0, lo, self.D))
            self.sync()
        finally:
            for p in ptrs:
                self._lib.cmf_host_free(C.c_void_p(p))

    def bind_device(self, dev_ptr, line_pitch=None, band_pitch=None):
        """Device-resident slab: pointer to (line 0, band active[0], sample 0)."""
        lp = self.D * self.S if line_pitch is None else int(line_pitch)
        bp = self.S if band_pitch is None else int(band_pitch)
        self._check(self._lib.cmf_bind_device_slab(self._ctx, C.c_void_p(int(dev_ptr)), lp, bp))

    def set_labels(self, labels_ls, kmodes=None, reject_min=0):
        """Background-mode labels (L, S) int in 0..kmodes-1 (cmf/robust_mf.py:312-313); ``reject_min`` > 0 is
        ``-r`` with bgminsamp = reject_min.  ``None`` returns to the unimodal path."""
        if labels_ls is None:
            self._check(self._lib.cmf_set_labels(self._ctx, C.c_void_p(None), 1, 0))
            self._labelled = False
            return
        lab = np.ascontiguousarray(labels_ls, dtype=np.int32)
        if lab.shape != (self.L, self.S):
            raise CmfError("labels shape %r != %r" % (lab.shape, (self.L, self.S)))
        k = int(lab.max()) + 1 if kmodes is None else int(kmodes)
        self._check(self._lib.cmf_set_labels(self._ctx, C.c_void_p(lab.ctypes.data), k, int(reject_min)))
        self._labelled = True

    def set_clustering(self, kmodes, pcadim=6, reject_min=0, max_iter=100):
        """Partition every column on the device (PCA + k-means, cmf/robust_mf.py:308-313; deterministic rule in
        csrc/k_cluster.cu).  ``kmodes`` <= 1 returns to the unimodal path."""
        self._check(self._lib.cmf_set_clustering(self._ctx, int(kmodes), int(pcadim), int(reject_min), int(max_iter)))
        self._labelled = int(kmodes) > 1
        self.pcadim = int(pcadim)

    def set_regfull(self, enable=True):
        """``-f``: regularise every mode fit with the covariance of the whole column (cmf/robust_mf.py:358)."""
        self._check(self._lib.cmf_set_regfull(self._ctx, 1 if enable else 0))

    def set_exclusion(self, exclude):
        """Opt-in (default off = reference behaviour): pixels where ``exclude`` (L, S) is true stay out of the
        background statistics (mean, covariance, alpha search) but are still scored; ``None`` clears."""
        if exclude is None:
            self._check(self._lib.cmf_set_exclusion(self._ctx, C.c_void_p(None)))
            return
        ex = np.ascontiguousarray(np.asarray(exclude) != 0, dtype=np.uint8)
        if ex.shape != (self.L, self.S):
            raise CmfError("exclusion mask must be (lines, samples) = %r, got %r" % ((self.L, self.S), ex.shape))
        self._check(self._lib.cmf_set_exclusion(self._ctx, C.c_void_p(ex.ctypes.data)))

    def set_screen_margin(self, rel_margin=1.0e-5, certify=True):
        """Margin of the tensor-core screen of the alpha search and its runtime certificate (diagnostics)."""
        self._check(self._lib.cmf_set_screen_margin(self._ctx, float(rel_margin), 1 if certify else 0))

    def screen_check(self):
        """Per column: measured screening error as a fraction of the margin (0 where nothing was re-evaluated)."""
        return self._get(_lib.OUT_SCREEN_CHECK, np.float64, (self.S,))

    def labels(self):
        return self._get(_lib.OUT_LABELS, np.int32, (self.L, self.S))

    def pca(self):
        """(S, L, pcadim) projections the on-device k-means partitioned (0 for invalid pixels)."""
        return self._get(_lib.OUT_PCA, np.float64, (self.S, self.L, self.pcadim))

    def kmeans_iters(self):
        return self._get(_lib.OUT_KMEANS_ITERS, np.int32, (self.S,))

    def cluster_id(self):
        return self._get(_lib.OUT_CLUSTER_ID, np.int16, (self.L, self.S))

    def alpha_image(self):
        return self._get(_lib.OUT_ALPHA_IMAGE, np.int16, (self.L, self.S))

    def mode_list(self):
        return self._get(_lib.OUT_MODE_LIST, np.int8, (self.S, 32))

    # ------------------------------------------------------------------ compute
    def run(self, timing=False, sync=True, exact=False):
        """All kernels on the context stream.  ``exact=True`` evaluates every alpha of the leave-one-out
        search in FP64 instead of screening it on the tensor cores first (same selected index)."""
        flags = (_lib.RUN_TIMING if timing else 0) | (_lib.RUN_EXACT if exact else 0)
        self._check(self._lib.cmf_run(self._ctx, flags))
        if sync:
            self.sync()

    def sync(self):
        self._check(self._lib.cmf_sync(self._ctx))

    def run_host(self, host_ptr, mf_out=None, colstats_out=None, alpha_out=None, wait=True):
        """End-to-end call on a host cube pointer (int address): H2D + all kernels + D2H.  ``wait=False``
        returns as soon as the work is enqueued (pinned buffers); ``sync()`` then completes it."""
        def addr(a):
            return C.c_void_p(None if a is None else int(a))
        self._check(self._lib.cmf_run_host(self._ctx, C.c_void_p(int(host_ptr)), addr(mf_out),
                                           addr(colstats_out), addr(alpha_out), 0 if wait else _lib.RUN_ASYNC))

    def kernel_times(self):
        n = self._lib.cmf_kernel_count()
        buf = (C.c_float * n)()
        got = self._lib.cmf_kernel_times(self._ctx, buf, n)
        if got < 0:
            self._check(got)
        return {self._lib.cmf_kernel_name(i).decode(): float(buf[i]) for i in range(got)}

    def launch_count(self):
        return int(self._lib.cmf_launch_count(self._ctx))

    def screen_kernel(self):
        """Name of the kernel that screens the alpha search ('' when the problem is not screened)."""
        return self._lib.cmf_screen_kernel(self._ctx).decode()

    def device_ptr(self, what):
        return self._lib.cmf_device_ptr(self._ctx, int(what))

    # ------------------------------------------------------------------ results
    def _get(self, what, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        self._check(self._lib.cmf_download(self._ctx, int(what), C.c_void_p(out.ctypes.data), out.nbytes))
        return out

    def mf(self):
        return self._get(_lib.OUT_MF, np.float64, (self.L, self.S))

    def mask(self):
        return self._get(_lib.OUT_MASK, np.uint8, (self.L, self.S)).astype(bool)

    def colstats(self):
        """(3, S): npix, mean, std of the written scores; nodata for skipped columns."""
        return self._get(_lib.OUT_COLSTATS, np.float64, (3, self.S))

    def alpha_index(self):
        return self._get(_lib.OUT_ALPHA_INDEX, np.int32, (self.S,))

    def nll(self):
        return self._get(_lib.OUT_NLL, np.float64, (self.S, self.A))

    def mu(self):
        return self._get(_lib.OUT_MU, np.float64, (self.S, self.DP))[:, :self.D]

    def weights(self):
        return self._get(_lib.OUT_WEIGHTS, np.float64, (self.S, self.DP))[:, :self.D]

    def status(self):
        return self._get(_lib.OUT_STATUS, np.int32, (self.S,))

    def nvalid(self):
        return self._get(_lib.OUT_NVALID, np.int32, (self.S,))

    def eigvals(self):
        return self._get(_lib.OUT_EIGVALS, np.float64, (self.S, self.DP))[:, :self.D]

    def sweeps(self):
        return self._get(_lib.OUT_SWEEPS, np.int32, (self.S,))

    def ncand(self):
        """Alphas per column the screening pass left to the exact FP64 re-evaluation (1 = none)."""
        return self._get(_lib.OUT_NCAND, np.int32, (self.S,))

    def screen_tol(self):
        """Per-column nll margin used by the screen (selection is exact while its error is below half of it)."""
        return self._get(_lib.OUT_SCREEN_TOL, np.float64, (self.S,))

    def column_profile(self, robust=False, p=0.95):
        """Column profile of the scores of the last run (triage/cmf_profile.py:110-140): dict of (S,) arrays
        ``npix, avg, std, min, max`` or, with ``robust``, ``npix, med, mad, p05, p95`` (float32 arithmetic as
        in the reference, returned as float64)."""
        out = np.empty((5, self.S), dtype=np.float64)
        self._check(self._lib.cmf_column_profile(self._ctx, int(bool(robust)), float(p), C.c_void_p(out.ctypes.data)))
        names = ("npix", "med", "mad", "p05", "p95") if robust else ("npix", "avg", "std", "min", "max")
        return dict(zip(names, out))

    def pixel_flags(self, cube, spec, on_device=False, shape=None):
        """Per-pixel spectrometer flags (spectrometer_masks/masks_sds.py:133-230) of a float32 BIL cube: a host
        array (L, B, S) or, with ``on_device``, a device address plus ``shape``.  ``spec`` is a _lib.FlagSpec."""
        if on_device:
            L, B, S = shape
            ptr = int(cube)
        else:
            if cube.dtype != np.float32 or not cube.flags.c_contiguous or cube.ndim != 3:
                raise CmfError("cube must be a C-contiguous float32 (lines, bands, samples) array")
            L, B, S = cube.shape
            ptr = cube.ctypes.data
        out = np.empty((L, S), dtype=np.uint8)
        self._check(self._lib.cmf_pixel_flags(self._ctx, C.c_void_p(ptr), int(bool(on_device)), L, B, S,
                                              C.byref(spec), C.c_void_p(out.ctypes.data)))
        return out

    def results(self):
        cs = self.colstats()
        return dict(mf=self.mf(), mask=self.mask(), colnum=cs[0], colavg=cs[1], colstd=cs[2],
                    alpha_index=self.alpha_index(), mu=self.mu(), weights=self.weights(),
                    status=self.status())


def cmf_cube(cube_lbs, abscf, active, model="looshrinkage", reflectance=False, alphas=None,
             nodata=-9999.0, device=0, exact=False, labels=None, reject_min=0, regfull=False, kmodes=1,
             pcadim=6):
    """One-shot convenience: same inputs/outputs as the oracle's ``cmf_cube`` (for parity tests)."""
    L, B, S = cube_lbs.shape
    with ColumnwiseMF(L, B, S, active, abscf, model=model, reflectance=reflectance, alphas=alphas,
                      nodata=nodata, device=device) as eng:
        eng.upload(cube_lbs)
        if labels is not None:
            eng.set_labels(labels, reject_min=reject_min)
        elif kmodes > 1:
            eng.set_clustering(kmodes, pcadim=pcadim, reject_min=reject_min)
        if regfull:
            eng.set_regfull(True)
        eng.run(exact=exact)
        res = eng.results()
        if labels is not None or kmodes > 1:
            res["cluster_id"], res["alpha_image"] = eng.cluster_id(), eng.alpha_image()
        if labels is None and kmodes > 1:
            res["labels"], res["pca"], res["kmeans_iters"] = eng.labels(), eng.pca(), eng.kmeans_iters()
        if model == "looshrinkage":
            res["nll"] = eng.nll()
        return res


_LOO_CTX = {}


def looshrinkage(I_zm, alphas, nll, n, I_reg=[], device=0):
    """Drop-in for the reference's importable ``looshrinkage(I_zm, alphas, nll, n, I_reg=[])``
    (cmf/robust_mf.py:92-136): fills ``nll`` in place and returns ``(C, mindex)``.  Every alpha is evaluated in FP64
    on the GPU (``cmf_looshrinkage``); with ``I_reg`` (the ``-f`` target, :99, :131) the target is ``cov(I_reg)``."""
    lib = _lib.load()
    reg = None
    if len(I_reg) != 0:
        reg = np.ascontiguousarray(I_reg, dtype=np.float64)
        if reg.ndim != 2 or reg.shape[1] != np.shape(I_zm)[1]:
            raise CmfError("looshrinkage: I_reg must have the columns of I_zm")
    ctx = _LOO_CTX.get(int(device))
    if ctx is None:
        ctx = C.c_void_p()
        rc = lib.cmf_create(C.byref(ctx), int(device))
        if rc != 0:
            raise CmfError("cmf_create failed (%d): %s" % (rc, lib.cmf_last_error(None).decode()))
        _LOO_CTX[int(device)] = ctx
    x = np.ascontiguousarray(I_zm, dtype=np.float64)
    al = np.ascontiguousarray(alphas, dtype=np.float64)
    rows, D = x.shape
    if nll.dtype != np.float64 or not nll.flags.c_contiguous or nll.shape != al.shape:
        raise CmfError("nll must be a contiguous float64 array of the same length as alphas")
    Cm = np.empty((D, D), dtype=np.float64)
    mi = C.c_int32(0)
    rc = lib.cmf_looshrinkage(ctx, C.c_void_p(x.ctypes.data), rows, D, C.c_void_p(al.ctypes.data), len(al), int(n),
                              C.c_void_p(reg.ctypes.data if reg is not None else None),
                              reg.shape[0] if reg is not None else 0, C.c_void_p(nll.ctypes.data), C.c_void_p(Cm.ctypes.data),
                              C.byref(mi))
    if rc != 0:
        raise CmfError("cmf_looshrinkage failed (%d): %s" % (rc, lib.cmf_last_error(ctx).decode()))
    return Cm, int(mi.value)
