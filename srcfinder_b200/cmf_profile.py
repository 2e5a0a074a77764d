"""Column profiles of CMF products on the GPU: host mirror of ``triage/cmf_profile.py`` (summarize, :92-140).

    python -m srcfinder_b200.cmf_profile [-v] [--robust] [--outdir DIR] cmf_file [cmf_file ...]

Same arguments as the reference's parser (:45-67; ``--plot``, ``--jobs`` and ``--randomize`` are accepted and
ignored: plotting is not part of the path, and one GPU call per image needs no process pool).  For every
product it writes ``<outdir>/<basename>_column_stats.csv`` with the columns ``npix,avg,std,min,max`` or, with
``--robust``, ``npix,med,mad,p05,p95`` -- one row per cross-track column, as :131-135 does through pandas --
and skips a product whose csv already exists (:104-106).  The statistics are computed by ``cmf_column_profile_image``
in libcmf_b200.so in float32, exactly as numpy evaluates them in the reference; there is no CPU path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np

from . import _lib, envi

PLAIN_COLS = ["npix", "avg", "std", "min", "max"]          # :99
ROBUST_COLS = ["npix", "med", "mad", "p05", "p95"]         # :97


def column_profile_image(mf_ls, nodata=-9999.0, robust=False, p=0.95, device=0):
    """dict of (S,) arrays for a (lines, samples) score image (the last band of a product)."""
    from .cmf import CmfError
    lib = _lib.load()
    ctx = C.c_void_p()
    rc = lib.cmf_create(C.byref(ctx), int(device))
    if rc != 0:
        raise CmfError("cmf_create failed (%d): %s" % (rc, lib.cmf_last_error(None).decode()))
    try:
        img = np.ascontiguousarray(mf_ls, dtype=np.float64)
        L, S = img.shape
        out = np.empty((5, S), dtype=np.float64)
        rc = lib.cmf_column_profile_image(ctx, C.c_void_p(img.ctypes.data), L, S, float(nodata),
                                          int(bool(robust)), float(p), C.c_void_p(out.ctypes.data))
        if rc != 0:
            raise CmfError("cmf_column_profile_image failed (%d): %s" % (rc, lib.cmf_last_error(ctx).decode()))
        return dict(zip(ROBUST_COLS if robust else PLAIN_COLS, out))
    finally:
        lib.cmf_destroy(ctx)


def _fmt(v):
    """pandas' to_csv writes float64 with repr precision; npix is a float column there as well (np.c_)."""
    return repr(float(v)) if np.isfinite(v) else ("" if np.isnan(v) else repr(float(v)))


def summarize(cmff, outdir, use_robust_stats=False, device=0, verbose=False):
    """One product -> ``<outdir>/<base>_column_stats.csv`` (triage/cmf_profile.py:92-135)."""
    outbase = os.path.split(os.path.splitext(cmff)[0])[1]
    colcsv = os.path.join(outdir, outbase + "_column_stats.csv")
    if os.path.exists(colcsv):
        print(colcsv, "exists, exiting")
        return False
    print("processing %s" % outbase)
    hdr = cmff + ".hdr" if os.path.exists(cmff + ".hdr") else os.path.splitext(cmff)[0] + ".hdr"
    meta = envi.read_header(hdr)
    mm = envi.open_memmap(cmff, meta)
    il = str(meta.get("interleave", "bip")).lower()
    band = {"bip": lambda a: a[..., -1], "bil": lambda a: a[:, -1, :], "bsq": lambda a: a[-1]}[il](mm)
    nodata = float(meta.get("data ignore value", -9999))
    cols = ROBUST_COLS if use_robust_stats else PLAIN_COLS
    res = column_profile_image(np.array(band, dtype=np.float64), nodata, use_robust_stats, 0.95, device)
    if verbose:
        print("CMF # positive=%d" % int(res["npix"].sum()))
    print("Saving column stats to", colcsv)
    with open(colcsv, "w") as fh:
        fh.write(",".join(cols) + "\n")
        for i in range(len(res["npix"])):
            fh.write(",".join(_fmt(res[c][i]) for c in cols) + "\n")
    return True


def build_parser():
    p = argparse.ArgumentParser("summarize_cmf.py")
    p.add_argument("-v", "--verbose", action="store_true", help="Verbose output")
    p.add_argument("--robust", action="store_true", help="Use robust statistics")
    p.add_argument("-j", "--jobs", type=int, default=1, help="Number of parallel jobs (ignored: one GPU call per image)")
    p.add_argument("--plot", action="store_true", help="Plot column statistics (ignored)")
    p.add_argument("--randomize", action="store_true", help="Randomize cmffiles processing order (ignored)")
    p.add_argument("--outdir", type=str, default=".", help="Output directory")
    p.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    p.add_argument("cmffiles", help="CMF image file", type=str, metavar="cmf_file", nargs="+")
    return p


def main(argv=None):
    args = build_parser().parse_args(argv)
    for f in args.cmffiles:
        summarize(f, args.outdir, args.robust, args.device, args.verbose)
    return 0


if __name__ == "__main__":
    sys.exit(main())
