"""TEST INFRASTRUCTURE -- generates ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run once in the build container (needs ``/root/reference``):

    OPENBLAS_NUM_THREADS=1 python oracle/make_golden.py

For every case it builds a seeded synthetic BIL cube (``srcfinder_b200.synth``), writes it as
an ENVI file, drives ``/root/reference/cmf/robust_mf.py`` through ``oracle/ref_shim.py`` and
stores the inputs (only the bands the script reads -- everything else is zero and is
re-created as zero by ``tests/golden_util.load_case``) together with the reference's outputs:
the 4-band product, the ``_bgmeta`` image (alpha index), the parsed output header and the
per-column mean/std lines the script prints.

Also (re)creates ``srcfinder_b200/data/ch4_unit_425.npy`` from the reference's library text file.
"""
from __future__ import annotations

import json
import os
import re
import sys
import tempfile

os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
os.environ.setdefault("OMP_NUM_THREADS", "1")

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from srcfinder_b200 import synth  # noqa: E402

REF_LIB = "/root/reference/cmf/ang_ch4_unit_3col_425chan.txt"
GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = [
    # name, lines, samples, seed, synth kwargs, CLI flags, library file name
    dict(name="unimodal_300x6", L=300, S=6, seed=11, kw={}, flags=["-m"], lib="ang_ch4_unit.txt"),
    dict(name="unimodal_2000x4", L=2000, S=4, seed=12, kw={}, flags=["-m"], lib="ang_ch4_unit.txt"),
    dict(name="badpix_400x6", L=400, S=6, seed=13, kw=dict(bad_pixels=True), flags=["-m"],
         lib="ang_ch4_unit.txt", dead_column=2, sparse_column=4),
    dict(name="empirical_300x4", L=300, S=4, seed=14, kw={}, flags=["-M", "empirical"],
         lib="ang_ch4_unit.txt"),
    dict(name="co2window_300x4", L=300, S=4, seed=15, kw={}, flags=["-m"], lib="ang_co2_unit.txt"),
    # note: ``--rgb_bands []`` cannot be pinned -- the reference crashes at :395 (rgb_bands[0])
    dict(name="degenerate_300x4", L=300, S=4, seed=17, kw={}, flags=["-m"], lib="ang_ch4_unit.txt",
         few_valid={1: 50, 2: 1, 3: 74}),
    # background modes: the reference's MiniBatchKMeans draws from numpy's global RandomState, so seeding it
    # before the (unmodified) script runs makes the partition reproducible; the labels the script used are
    # recovered from _bgmeta band 0.  A block of 40 bright lines forms a cluster below bgminsamp = 85.
    dict(name="modes_k3r_900x4", L=900, S=4, seed=21, kw=dict(bad_pixels=True), flags=["-k", "3", "-r", "-m"],
         lib="ang_ch4_unit.txt", np_seed=5, bright_lines=(100, 140)),
    dict(name="modes_k2_600x3", L=600, S=3, seed=22, kw={}, flags=["-k", "2", "-m"], lib="ang_ch4_unit.txt",
         np_seed=7),
    # -f: every mode fit is regularised with the covariance of the whole column (:358, looshrinkage's I_reg)
    dict(name="modes_k3rf_900x4", L=900, S=4, seed=23, kw=dict(bad_pixels=True),
         flags=["-k", "3", "-r", "-f", "-m"], lib="ang_ch4_unit.txt", np_seed=9, bright_lines=(300, 340)),
    dict(name="modes_k2f_500x3", L=500, S=3, seed=24, kw={}, flags=["-k", "2", "-f", "-m"],
         lib="ang_ch4_unit.txt", np_seed=11),
    # -R with the CH4 library: active window [5, 420] = 416 bands (:186-187), target = abscf - mu, no ppm
    # scaling (:379, :383).  At this width det(G) under/overflows for most alphas (:111-113), which the
    # wide-window kernel set has to reproduce.  Slow in the reference (~1 min per column).
    dict(name="reflectance_700x2", L=700, S=2, seed=31, kw={}, flags=["-R", "-m"], lib="ang_ch4_unit.txt"),
    dict(name="reflectance_1500x3", L=1500, S=3, seed=32, kw=dict(bad_pixels=True), flags=["-R", "-m"],
         lib="ang_ch4_unit.txt"),
]


def _parse_colstats(stdout, ncols):
    avg = np.full(ncols, np.nan)
    std = np.full(ncols, np.nan)
    for m in re.finditer(r"Column (\d+) mean: (\S+), std: (\S+)", stdout):
        c = int(m.group(1))
        avg[c], std[c] = float(m.group(2)), float(m.group(3))
    return avg, std


def run_case(case, tmp):
    L, S = case["L"], case["S"]
    cube = synth.make_cube(L, S, seed=case["seed"], **case["kw"])
    if "dead_column" in case:            # a column with no valid pixel at all (:303-304)
        cube[:, :, case["dead_column"]] = synth.NODATA
    if "sparse_column" in case:          # a column with few (but > D+1) valid pixels
        cube[::3, 360, case["sparse_column"]] = np.nan
    for c, nvalid in case.get("few_valid", {}).items():   # columns with n < D, n == 1, n ~ D+2
        cube[nvalid:, 355, c] = np.nan
    if "bright_lines" in case:           # a small, spectrally distinct population (rejected with -r)
        l0, l1 = case["bright_lines"]
        ok = cube[l0:l1] > 0
        cube[l0:l1] = np.where(ok, cube[l0:l1] * 2.5, cube[l0:l1])
    libname = case["lib"]
    refl = "-R" in case["flags"]
    lo, hi = (5, 420) if (refl and "ch4" in libname) else ((351, 422) if "ch4" in libname else (309, 391))
    rgb = [] if "[]" in case["flags"] else [60, 42, 24]
    keep = sorted(set(range(lo - 1, hi)) | set(rgb))
    slim = np.zeros_like(cube)
    slim[:, keep, :] = cube[:, keep, :]
    cube = slim
    inp = os.path.join(tmp, case["name"] + "_rdn")
    out = os.path.join(tmp, case["name"] + "_mf")
    libpath = os.path.join(tmp, libname)
    synth.write_library_txt(libpath)
    ref_shim.write_bil_cube(inp, cube, extra_meta={
        "wavelength units": "Nanometers", "description": "synthetic AVIRIS-NG radiance",
        "wavelength": ["%.2f" % w for w in synth.load_ch4_library()[:, 1]],
        "fwhm": ["5.0"] * cube.shape[1], "smoothing factors": ["0"] * cube.shape[1],
        "bad pixel map": "none"})
    if "np_seed" in case:
        np.random.seed(case["np_seed"])
    stdout = ref_shim.run_reference_cli(case["flags"] + [inp, libpath, out])
    hdr = ref_shim.parse_envi_header(out + ".hdr")
    nb = int(hdr["bands"]) if "-m" not in case["flags"] else (4 if rgb else 1)
    # with -m the reference aliases the two header dicts (:272), so the main header on disk was
    # written before the aliasing and is intact; read the product with its own header
    prod = np.fromfile(out, dtype=np.float64).reshape(L, S, -1)
    rec = dict(
        cube_kept=cube[:, keep, :], kept_bands=np.array(keep), shape=np.array(cube.shape),
        flags=json.dumps(case["flags"]), libname=libname, active=np.array([lo, hi]),
        product=prod, header=json.dumps(hdr), stdout_avg=None, stdout_std=None)
    rec["stdout_avg"], rec["stdout_std"] = _parse_colstats(stdout, S)
    if "-m" in case["flags"]:
        bg = np.fromfile(out + "_bgmeta", dtype=np.int16).reshape(L, S, 2)
        rec["bgmeta"] = bg
        rec["bgmeta_header"] = json.dumps(ref_shim.parse_envi_header(out + "_bgmeta.hdr"))
    assert prod.shape[2] == nb, (prod.shape, nb)
    np.savez_compressed(os.path.join(GOLDEN, case["name"] + ".npz"), **rec)
    print("%-20s product %s  mf std %.2f  stdout cols %d" % (
        case["name"], prod.shape, np.std(prod[..., -1][prod[..., -1] != -9999.0]),
        np.isfinite(rec["stdout_avg"]).sum()))


def looshrinkage_cases():
    """The reference's importable looshrinkage(I_zm, alphas, nll, n, I_reg=[]) (:92-136), called directly."""
    ref = ref_shim.import_reference_functions()
    alphas = 10.0 ** np.arange(-10, 0.05, 0.05)
    for name, L, seed, window, n_extra in (("looshrinkage_600x72", 600, 41, (351, 422), 0),
                                           ("looshrinkage_250x83", 250, 42, (309, 391), 37),
                                           ("looshrinkage_900x160", 900, 43, (200, 359), 0)):
        cube = synth.make_cube(L, 1, seed=seed)
        x = np.float64(cube[:, window[0] - 1:window[1], 0])
        izm = x - x.mean(axis=0)
        nll = np.zeros(len(alphas))
        n = L + n_extra                      # the reference passes the column count for a cluster subset (:355-356)
        C, mindex = ref.looshrinkage(izm, alphas, nll, n)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), I_zm=izm, alphas=alphas, n=n, nll=nll, C=C,
                            mindex=mindex)
        print("%-22s mindex %d  finite nll %d" % (name, mindex, np.isfinite(nll).sum()))
    # with I_reg, as the column loop calls it for -f (:353-356): the samples of one mode, the whole column (less the
    # mode's mean) as the regulariser, n = the column's pixel count
    for name, L, seed, window in (("looshrinkage_reg_700x72", 700, 47, (351, 422)),
                                  ("looshrinkage_reg_900x140", 900, 48, (200, 339))):
        cube = synth.make_cube(L, 1, seed=seed)
        x = np.float64(cube[:, window[0] - 1:window[1], 0])
        member = x[:, -1] > np.median(x[:, -1])
        mu = x[member].mean(axis=0)
        izm = x[member] - mu
        ireg = x - mu
        nll = np.zeros(len(alphas))
        C, mindex = ref.looshrinkage(izm, alphas, nll, L, I_reg=ireg)
        np.savez_compressed(os.path.join(GOLDEN, name + ".npz"), I_zm=izm, I_reg=ireg, alphas=alphas, n=L, nll=nll, C=C,
                            mindex=mindex)
        print("%-22s mindex %d  finite nll %d" % (name, mindex, np.isfinite(nll).sum()))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    if sys.argv[1:] == ["looshrinkage"]:
        return looshrinkage_cases()
    lib = np.loadtxt(REF_LIB)
    np.save(os.path.join(ROOT, "srcfinder_b200", "data", "ch4_unit_425.npy"), lib)
    only = sys.argv[1:]
    with tempfile.TemporaryDirectory() as tmp:
        for case in CASES:
            if only and case["name"] not in only:
                continue
            run_case(case, tmp)


if __name__ == "__main__":
    main()
