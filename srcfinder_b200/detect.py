"""The two per-pixel steps between the matched-filter product and the plume detector, on the GPU (SURVEY.md 8(f)
row 4): the head of ``srcfinder_util.filtdet`` (:1422-1436: KDE weighting ``kde`` :1383-1387, clip to
(mfmin, mfmax), candidate mask) and the CNN input normalisation of ``cnn/cnn_pred_pipeline.py`` (``ClampCH4``
:19-30 + ``transforms.Normalize`` :126-157).  All arithmetic runs in libcmf_b200.so; there is no CPU path here.

The connected-component steps of ``filtdet`` (``remove_small_objects``, small-detection rescue, relabelling,
:1437-1470) are image morphology over the whole scene and stay with the caller.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

KERNEL = 50                     # srcfinder_util.py:105
MFMIN, MFMAX = 500, 1500        # :106
# (vmin, vmax, mean, std) per model (cnn/cnn_pred_pipeline.py:126-157)
CNN_MODELS = {
    "COVID_QC": (0, 4000, 110.6390, 183.9152),
    "CalCH4_v8": (0, 4000, 140.6399, 237.5434),
    "Permian_QC": (0, 4000, 100.2635, 158.7060),
    "multi": (0, 4000, 115.0, 190.0),
}


def gaussian_weights(sigma, truncate=1.0):
    """The 1-D kernel scipy.ndimage.gaussian_filter uses (scipy/ndimage/_filters.py, _gaussian_kernel1d, order 0)."""
    sigma = float(sigma)
    radius = int(truncate * sigma + 0.5)
    sigma2 = sigma * sigma
    x = np.arange(-radius, radius + 1)
    phi_x = np.exp(-0.5 / sigma2 * x ** 2)
    return phi_x / phi_x.sum()


def _ctx(device):
    from .cmf import CmfError
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.cmf_create(C.byref(ctx), int(device))
    if rc != 0:
        raise CmfError("cmf_create failed (%d): %s" % (rc, lib.cmf_last_error(None).decode()))
    return lib, ctx


def filtdet_prefilter(ch4mf, k=KERNEL, mfmin=MFMIN, mfmax=MFMAX, engine=None, device=0):
    """``(detkde, ch4min, detmask)`` of ``filtdet`` for a score image (lines, samples); with ``engine`` (a
    ``ColumnwiseMF`` that has run) and ``ch4mf=None`` the scores are taken where they already are, in HBM."""
    from .cmf import CmfError
    w = np.ascontiguousarray(gaussian_weights(k), dtype=np.float64)
    radius = (len(w) - 1) // 2
    if engine is not None and ch4mf is None:
        lib, ctx, own = engine._lib, engine._ctx, False
        L, S, src = engine.L, engine.S, None
    else:
        lib, ctx = _ctx(device)
        own = True
        img = np.ascontiguousarray(ch4mf, dtype=np.float64)
        L, S = img.shape
        src = img.ctypes.data
    try:
        det = np.empty((L, S), dtype=np.float64)
        cmin = np.empty((L, S), dtype=np.uint8)
        dmask = np.empty((L, S), dtype=np.uint8)
        rc = lib.cmf_detection_prefilter(ctx, C.c_void_p(src), L, S, C.c_void_p(w.ctypes.data), radius, float(mfmin),
                                         float(mfmax), C.c_void_p(det.ctypes.data), C.c_void_p(cmin.ctypes.data),
                                         C.c_void_p(dmask.ctypes.data))
        if rc != 0:
            raise CmfError("cmf_detection_prefilter failed (%d): %s" % (rc, lib.cmf_last_error(ctx).decode()))
        return det, cmin.astype(bool), dmask.astype(bool)
    finally:
        if own:
            lib.cmf_destroy(ctx)


def cnn_input(ch4mf, model="COVID_QC", engine=None, device=0):
    """float32 (lines, samples) network input: ClampCH4 + Normalize with the constants of ``model``."""
    from .cmf import CmfError
    vmin, vmax, mean, std = CNN_MODELS[model] if isinstance(model, str) else model
    if engine is not None and ch4mf is None:
        lib, ctx, own = engine._lib, engine._ctx, False
        L, S, src = engine.L, engine.S, None
    else:
        lib, ctx = _ctx(device)
        own = True
        img = np.ascontiguousarray(ch4mf, dtype=np.float64)
        L, S = img.shape
        src = img.ctypes.data
    try:
        out = np.empty((L, S), dtype=np.float32)
        rc = lib.cmf_cnn_input(ctx, C.c_void_p(src), L, S, float(vmin), float(vmax), float(mean), float(std),
                               C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise CmfError("cmf_cnn_input failed (%d): %s" % (rc, lib.cmf_last_error(ctx).decode()))
        return out
    finally:
        if own:
            lib.cmf_destroy(ctx)
