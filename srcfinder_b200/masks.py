"""Per-pixel spectrometer flags on the GPU: host mirror of the per-pixel tests of
``spectrometer_masks/masks_sds.py`` (get_saturation_mask :133-150, get_spec_mask :152-163, get_dark_mask
:165-180, get_cloud_mask :182-230) with the same defaults (:50-54, :78, :102, :194).

The region growing, buffering and dilation that follow in the reference (:232-330) are image morphology over
the whole scene, not per-pixel work, and are left to the caller.  All tests run in libcmf_b200.so
(``cmf_pixel_flags``); there is no CPU path here.
"""
from __future__ import annotations

import numpy as np

from . import _lib

SAT_THRESH_DEFAULT = 6.0            # masks_sds.py:50
SAT_THRESH_CLD = [15.0]             # :52
DARK_THRESH_DEFAULT = [0.104]       # :54
SAT_WINDOW_NM = (1945.0, 2485.0)    # :147
SPEC_BAND, DARK_BAND = 25, 352      # :159, :175
CLOUD_BANDS = (15, 60, 175)         # :194
VISIBLE_MASK_GROWING_THRESHOLD = 9.0  # :102

SATURATED, SPECULAR, DARK, CLOUD = _lib.FLAG_SATURATED, _lib.FLAG_SPECULAR, _lib.FLAG_DARK, _lib.FLAG_CLOUD


def flag_spec(wave, threshold=None, waverange=None, dark_threshold=0.104, cldthreshold=None, cldbands=None,
              visible_mask_growing_threshold=VISIBLE_MASK_GROWING_THRESHOLD, spec_band=SPEC_BAND,
              dark_band=DARK_BAND):
    """Translate the reference's arguments (wavelength vector in nm, thresholds, windows) into band indices.

    The saturation window must select one contiguous run of bands (wavelengths ascending), as it does for every
    imaging spectrometer header; anything else raises ValueError.
    """
    wave = np.asarray(wave, dtype=np.float64)
    threshold = SAT_THRESH_DEFAULT if threshold is None else threshold
    waverange = SAT_WINDOW_NM if waverange is None else waverange
    cldthreshold = SAT_THRESH_CLD if cldthreshold is None else cldthreshold
    cldbands = CLOUD_BANDS if cldbands is None else cldbands
    sel = np.flatnonzero(np.logical_and(wave >= waverange[0], wave <= waverange[1]))
    if sel.size == 0 or not np.array_equal(sel, np.arange(sel[0], sel[-1] + 1)):
        raise ValueError("saturation window does not select a contiguous run of bands")
    nb = len(wave)

    def opt(b):
        return int(b) if b is not None and 0 <= int(b) < nb else -1

    a, b = opt(cldbands[0]), opt(cldbands[1])
    if a < 0 or b < 0:
        a = b = -1
    spec = _lib.FlagSpec()
    spec.sat_lo, spec.sat_hi = int(sel[0]), int(sel[-1])
    spec.spec_band, spec.dark_band = opt(spec_band), opt(dark_band)
    spec.cloud_a, spec.cloud_b = a, b
    spec.sat_thresh = float(threshold)
    spec.spec_thresh = float(visible_mask_growing_threshold)
    spec.dark_thresh = float(dark_threshold)
    spec.cloud_thresh = float(cldthreshold[0])
    spec.cloud_dwl = float(wave[b] - wave[a]) if a >= 0 else 0.0
    return spec


def pixel_flags(cube_lbs, wave, device=0, **kwargs):
    """uint8 (L, S) image of SATURATED | SPECULAR | DARK | CLOUD bits for a float32 BIL cube (L, B, S)."""
    from .cmf import CmfError
    import ctypes as C
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.cmf_create(C.byref(ctx), int(device))
    if rc != 0:
        raise CmfError("cmf_create failed (%d): %s" % (rc, lib.cmf_last_error(None).decode()))
    try:
        cube = np.ascontiguousarray(cube_lbs, dtype=np.float32)
        L, B, S = cube.shape
        spec = flag_spec(wave, **kwargs)
        out = np.empty((L, S), dtype=np.uint8)
        rc = lib.cmf_pixel_flags(ctx, C.c_void_p(cube.ctypes.data), 0, L, B, S, C.byref(spec),
                                 C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise CmfError("cmf_pixel_flags failed (%d): %s" % (rc, lib.cmf_last_error(ctx).decode()))
        return out
    finally:
        lib.cmf_destroy(ctx)
