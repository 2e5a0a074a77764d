"""Drop-in for ``cmf/robust_mf.py``: same command line, same ENVI products, CUDA arithmetic.

    python -m srcfinder_b200.robust_mf [-v] [-k K] [--pcadim P] [-r] [-f] [--rgb_bands r,g,b] [-m] [-R]
                                       [-M looshrinkage|empirical] INPUT LIBRARY OUTPUT

Argument names, defaults and meanings are the reference's (cmf/robust_mf.py:142-166).  Products
(:210-279, :383-403): OUTPUT (+.hdr) 4-band float64 BIP image [R, G, B radiance copies, CH4 ppm x m]
with ``data ignore value`` in masked pixels and the ``model parameters`` header string;
OUTPUT_bgmeta (+.hdr) with ``-m`` (int16 BIP: cluster_id, alpha_index); <INPUT>_column_stats.csv
(rows npix / avg / std, one column per cross-track sample -- the layout the reference's broken
DataFrame call at :401-402 evidently intends).  All arithmetic runs in libcmf_b200.so on the GPU.
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np

from . import envi
from .cmf import ColumnwiseMF, CmfError, alpha_grid

BAND_NAMES_RGB = ["Red Radiance (uW/nm/sr/cm2)", "Green Radiance (uW/nm/sr/cm2)",
                  "Blue Radiance (uW/nm/sr/cm2)", "CH4 Absorption (ppm x m)"]        # :218-221


def active_window(library, reflectance):
    """Active channels from the library file name and -R (cmf/robust_mf.py:186-194)."""
    if reflectance and "ch4" in library:
        return [5, 420]
    if "ch4" in library:
        return [351, 422]
    if "co2" in library:
        return [309, 391]
    return None


def build_parser():
    p = argparse.ArgumentParser(description="Robust MF")
    p.add_argument("-v", "--verbose", action="store_true", help="verbose output")
    p.add_argument("-k", "--kmeans", type=int, default=1, help="number of columnwise modes (k-means clusters)")
    p.add_argument("--pcadim", type=int, default=6, help="number of PCA dims (for k-means clusters>1)")
    p.add_argument("-r", "--reject", action="store_true", help="enable multimodal covariance outlier rejection")
    p.add_argument("-f", "--full", action="store_true",
                   help="regularize multimodal estimates with the full column covariariance")
    p.add_argument("--rgb_bands", default="60,42,24", help="comma-separated list of RGB channels")
    p.add_argument("-m", "--metadata", action="store_true", help="save metadata image")
    p.add_argument("-R", "--reflectance", action="store_true", help="reflectance signature")
    p.add_argument("-M", "--model", type=str, default="looshrinkage",
                   help="model name (looshrinkage (default)|empirical)")
    p.add_argument("--active", type=str, default=None,
                   help="extension: explicit 1-based active window lo,hi (other sensors, e.g. EMIT)")
    p.add_argument("--exclude", type=str, default=None,
                   help="extension, default off: comma-separated spectrometer flags (saturated,specular,dark,cloud; "
                        "spectrometer_masks/masks_sds.py tests) whose pixels stay out of the background statistics "
                        "(they are still scored); needs a 'wavelength' list in the input header")
    p.add_argument("--device", type=int, default=0, help="extension: CUDA device index")
    p.add_argument("input", type=str, metavar="INPUT", help="path to input image")
    p.add_argument("library", type=str, metavar="LIBRARY", help="path to target library file")
    p.add_argument("output", type=str, metavar="OUTPUT", help="path for output image (mf ch4 ppm)")
    return p


def model_parameters_string(modelname, bgmodes, pcadim, reject, regfull, reflectance, active):
    """The ``model parameters`` header value (cmf/robust_mf.py:246-259)."""
    bgmodel = "unimodal" if bgmodes == 1 else "multimodal"
    parms = "modelname=%s, bgmodel=%s" % (modelname, bgmodel)
    if bgmodes > 1:
        parms += ", bgmodes=%d, pcadim=%d, reject=%s" % (bgmodes, pcadim, reject)
        if modelname == "looshrinkage":
            parms += ", regfull=%s" % regfull
    if modelname == "looshrinkage":
        parms += ", aminexp=%s, amaxexp=%s, astep=%s" % (-10.0, 0.0, 0.05)
    parms += ", reflectance=%s, active_bands=%s" % (reflectance, active)
    return "{ %s }" % parms


def write_column_stats(path, colstats):
    """npix / avg / std rows, one column per sample (pandas-style CSV)."""
    S = colstats.shape[1]
    with open(path, "w") as fh:
        fh.write("," + ",".join(str(i) for i in range(S)) + "\n")
        for name, row in zip(("npix", "avg", "std"), colstats):
            fh.write(name + "," + ",".join(repr(float(v)) for v in row) + "\n")


def run(args, log=print):
    infile, outfile = args.input, args.output
    colcsv = os.path.splitext(infile)[0] + "_column_stats.csv"               # :183
    active = active_window(args.library, args.reflectance)
    if args.active:
        active = [int(t) for t in args.active.split(",")]
    if active is None:
        log("could not set active range")
        return 0                                                             # sys.exit(0), :193-194
    if args.kmeans > 32:
        raise CmfError("at most 32 background modes are supported")
    bgminsamp = int((active[1] - active[0]) * 1.2)                           # :200
    if args.model not in ("looshrinkage", "empirical"):
        raise CmfError("unknown model %r" % args.model)
    log('started processing input file: "%s"' % infile)
    meta = envi.read_header(infile + ".hdr")
    if str(meta.get("interleave", "")).lower() != "bil":
        raise CmfError("input must be BIL (the reference unpacks (lines, bands, samples), :208)")
    img = envi.open_memmap(infile, meta)
    L, B, S = img.shape

    outmeta = dict(meta)
    outmeta["lines"] = L
    outmeta["data type"] = envi.NUMPY_TO_ENVI["float64"]
    rgb = [] if args.rgb_bands == "[]" else [int(t) for t in args.rgb_bands.split(",")]
    if len(rgb) == 3:
        outmeta["bands"] = 4
        outmeta["band names"] = list(BAND_NAMES_RGB)
    elif len(rgb) == 0:
        outmeta["bands"] = 1
        outmeta["band names"] = [BAND_NAMES_RGB[3]]
    else:
        raise Exception("invalid value of rgb_bands argument: %s" % args.rgb_bands)   # :226
    outmeta["interleave"] = "bip"
    for key in ("smoothing factors", "wavelength", "wavelength units", "fwhm"):
        outmeta.pop(key, None)
    nodata = float(outmeta.get("data ignore value", -9999))
    if nodata > 0:
        raise Exception("nodata value=%f > 0, values will not be masked" % nodata)    # :234
    lib = np.float64(np.loadtxt(args.library))
    abscf = lib[active[0] - 1:active[1], 2]
    outmeta["model parameters"] = model_parameters_string(args.model, args.kmeans, args.pcadim, args.reject,
                                                          args.full, args.reflectance, active)
    out = envi.create_image(outfile, outmeta)
    assert out.shape[0] == L and out.shape[1] == S
    out[:, :, -1] = nodata

    log("starting columnwise processing (%d columns)" % S)
    t0 = time.time()
    streamed = img.dtype == np.float32 and not args.exclude
    cube = None if streamed else np.ascontiguousarray(img, dtype=np.float32)
    with ColumnwiseMF(L, B, S, active, abscf, model=args.model, reflectance=args.reflectance,
                      alphas=alpha_grid(), nodata=nodata, device=args.device) as eng:
        if streamed:
            eng.upload_stream(img)       # disk -> pinned blocks -> device, active window only (:298)
        else:
            eng.upload(cube)
        if args.exclude:
            if args.kmeans > 1:
                raise CmfError("--exclude applies to unimodal runs only")
            from . import masks
            names = {"saturated": masks.SATURATED, "specular": masks.SPECULAR, "dark": masks.DARK,
                     "cloud": masks.CLOUD}
            bits = 0
            for tok in args.exclude.split(","):
                if tok.strip() not in names:
                    raise CmfError("unknown flag %r in --exclude" % tok)
                bits |= names[tok.strip()]
            if "wavelength" not in meta:
                raise CmfError("--exclude needs the band wavelengths in the input header")
            wave = np.array([float(w) for w in meta["wavelength"]])
            flags = eng.pixel_flags(cube, masks.flag_spec(wave))
            eng.set_exclusion((flags & bits) != 0)
            log("excluding %d flagged pixels from the background statistics" % int(((flags & bits) != 0).sum()))
        if args.kmeans > 1:
            # PCA + k-means partition on the device (:306-313; deterministic, the reference's is unseeded),
            # rejection of small clusters (-r, :316-324) and the full-column regulariser (-f, :358)
            eng.set_clustering(args.kmeans, pcadim=args.pcadim, reject_min=bgminsamp if args.reject else 0)
            if args.full and args.model == "looshrinkage":
                eng.set_regfull(True)
        eng.run()
        mf = eng.mf()
        colstats = eng.colstats()
        aidx = eng.alpha_index()
        mask = eng.mask() if args.metadata else None
        if args.metadata and args.kmeans > 1:
            cluster_img, alpha_img = eng.cluster_id(), eng.alpha_image()
    out[:, :, -1] = mf
    done = colstats[0] != nodata          # columns with no valid pixel are skipped entirely (:303-304)
    if len(rgb) == 3:
        for i, b in enumerate(rgb):       # 0-based band indices, nodata copied as is (:395-397)
            out[:, done, i] = img[:, b, :][:, done]
    out.flush()
    for c in np.where(done)[0]:
        log("Column %i mean: %e, std: %e" % (c, colstats[1, c], colstats[2, c]))
    if args.metadata:
        bgmeta = dict(outmeta)
        bgmeta["bands"] = 2
        bgmeta["data type"] = envi.NUMPY_TO_ENVI["int16"]
        bgmeta["num alphas"] = len(alpha_grid())
        bgmeta["alphas"] = "{%s}" % (str(alpha_grid())[1:-1])
        bgmeta["band names"] = "{cluster_id, alpha_index}"
        bg = envi.create_image(outfile + "_bgmeta", bgmeta)
        if args.kmeans > 1:
            bg[:, :, 0] = np.where(mask, cluster_img, 0)                          # :327
            if args.model == "looshrinkage":
                bg[:, :, 1] = np.where(mask, alpha_img, 0)                        # :365
        elif args.model == "looshrinkage":
            bg[:, :, 1] = np.where(mask, aidx[None, :].astype(np.int16), 0)       # :365
        bg.flush()
    log("Saving column stats to %s" % colcsv)
    write_column_stats(colcsv, colstats)
    log("done (elapsed time=%ds)" % (time.time() - t0))
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    return run(args)


if __name__ == "__main__":
    sys.exit(main())
