"""The steps either side of the filter through the C ABI against their CPU oracles:
per-pixel spectrometer flags (bit-exact) and column profiles (bit-exact in float32: numpy's own evaluation order
is replayed on the device)."""
import numpy as np
import pytest

from oracle import products_oracle as po
from srcfinder_b200 import ColumnwiseMF, cmf_profile, envi, masks, synth

pytestmark = pytest.mark.gpu

ACTIVE = [351, 422]


def _wavelengths(nb=425):
    return 376.86 + 5.0087 * np.arange(nb)          # AVIRIS-NG band centres (nm), ~5 nm sampling


def _flag_cube(L, S, seed):
    rng = np.random.default_rng(seed)
    cube = synth.make_cube(L, S, seed=seed, bad_pixels=True)
    pick = lambda frac: rng.random((L, S)) < frac
    sat = pick(0.02)
    cube[:, 400, :][sat] = 7.5                       # saturated in the SWIR window
    cube[:, 25, :][sat & pick(0.5)] = 9.5            # ... and bright in the visible: specular
    cube[:, 25, :][pick(0.01)] = 11.0                # bright at band 25 alone is not specular
    cube[:, 352, :][pick(0.03)] = 0.05               # dark at 2139 nm
    cube[:, 352, :][pick(0.01)] = -9999.0            # no-data is not dark
    cube[:, 352, :][pick(0.005)] = np.float32(0.104)  # exactly the threshold (float32): not dark
    cld = pick(0.03)
    cube[:, 15, :][cld] = 18.0
    cube[:, 60, :][cld] = np.where(rng.random(cld.sum()) < 0.5, 12.0, 25.0)   # negative / positive slope
    cube[:, 313, :][pick(0.002)] = 6.5               # first band of the window
    cube[:, 312, :][pick(0.002)] = 50.0              # just outside the window
    cube[:, 420, :][pick(0.002)] = np.nan            # NaN never compares greater
    cube[:, 421, :][pick(0.002)] = np.float32(6.0)   # equal to the threshold: not saturated
    return cube


def test_pixel_flags_against_oracle():
    L, S = 300, 50
    cube = _flag_cube(L, S, seed=21)
    wave = _wavelengths()
    ref = po.pixel_flags(cube, wave)
    got = masks.pixel_flags(cube, wave)
    assert got.dtype == np.uint8 and got.shape == (L, S)
    assert np.array_equal(got, ref)
    for bit in (masks.SATURATED, masks.SPECULAR, masks.DARK, masks.CLOUD):
        assert (ref & bit).any(), "fixture must exercise bit %d" % bit
    assert not ((ref & masks.SPECULAR).astype(bool) & ~(ref & masks.SATURATED).astype(bool)).any()


def test_pixel_flags_device_cube_and_options():
    import torch
    L, S = 128, 37                                   # odd sample count: unaligned rows
    cube = _flag_cube(L, S, seed=22)
    wave = _wavelengths()
    kw = dict(threshold=5.0, waverange=(2000, 2400), dark_threshold=0.2, cldthreshold=[10.0],
              visible_mask_growing_threshold=8.0)
    ref = po.pixel_flags(cube, wave, threshold=5.0, waverange=(2000, 2400), dark_threshold=0.2, cldthreshold=(10.0,),
                         visible_mask_growing_threshold=8.0)
    assert np.array_equal(masks.pixel_flags(cube, wave, **kw), ref)
    dev = torch.from_numpy(cube).cuda()
    with ColumnwiseMF(L, 425, S, ACTIVE, synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1], 2]) as eng:
        got = eng.pixel_flags(dev.data_ptr(), masks.flag_spec(wave, **kw), on_device=True, shape=(L, 425, S))
    assert np.array_equal(got, ref)


def _score_image(L, S, seed):
    rng = np.random.default_rng(seed)
    mf = rng.normal(0.0, 420.0, (L, S)) * rng.uniform(0.5, 2.0, S)[None, :]
    mf[rng.random((L, S)) < 0.03] = -9999.0
    mf[rng.random((L, S)) < 0.002] = np.nan
    mf[:, 3] = -9999.0                               # a skipped column
    mf[:, 5] = -np.abs(mf[:, 5])                     # no positive pixel
    mf[: L - 1, 7] = -9999.0
    mf[L - 1, 7] = 12.5                              # a single valid pixel
    mf[2:, 8] = -9999.0
    mf[:2, 8] = (3.0, 8.0)                           # two valid pixels: the even-count median
    return mf


@pytest.mark.parametrize("robust", [False, True])
def test_column_profile_image_bit_exact(robust):
    L, S = 3001, 23
    mf = _score_image(L, S, seed=31)
    ref = po.column_profile(mf, -9999, use_robust_stats=robust)
    got = cmf_profile.column_profile_image(mf, -9999.0, robust=robust)
    assert list(got) == list(ref)
    for k in ref:
        r = np.asarray(ref[k], dtype=np.float64)
        assert np.array_equal(got[k], r, equal_nan=True), (k, np.flatnonzero(got[k] != r)[:5])
    assert got["npix"][3] == 0 and np.isnan(got[list(got)[1]][3])


def test_column_profile_of_a_run_and_cli(tmp_path):
    """Profile of the scores left on the device by a run == profile of the product on disk == oracle."""
    L, S = 1500, 12
    cube = synth.make_cube(L, S, seed=33, bad_pixels=True)
    ab = synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1], 2]
    with ColumnwiseMF(L, 425, S, ACTIVE, ab) as eng:
        eng.upload(cube)
        eng.run()
        mf = eng.mf()
        plain, robust = eng.column_profile(), eng.column_profile(robust=True)
    for got, ref in ((plain, po.column_profile(mf)), (robust, po.column_profile(mf, use_robust_stats=True))):
        for k in ref:
            assert np.array_equal(got[k], np.asarray(ref[k], dtype=np.float64), equal_nan=True), k
    # the CLI on a 4-band BIP product (triage/cmf_profile.py:92-135)
    out = str(tmp_path / "ang_cmf")
    mm = envi.create_image(out, {"samples": S, "lines": L, "bands": 4, "data type": 5, "interleave": "bip",
                                 "byte order": 0, "data ignore value": -9999})
    mm[..., 3] = mf
    mm.flush()
    assert cmf_profile.main(["--robust", "--outdir", str(tmp_path), out]) == 0
    rows = open(str(tmp_path / "ang_cmf_column_stats.csv")).read().splitlines()
    assert rows[0] == "npix,med,mad,p05,p95" and len(rows) == S + 1
    vals = np.array([[float(t) if t else np.nan for t in r.split(",")] for r in rows[1:]])
    for j, k in enumerate(("npix", "med", "mad", "p05", "p95")):
        assert np.array_equal(vals[:, j], robust[k], equal_nan=True)
    assert cmf_profile.summarize(out, str(tmp_path), True) is False      # exists -> skipped (:104-106)


def test_exclusion_from_background_statistics():
    """Opt-in exclusion: flagged pixels stay out of the mean / covariance / alpha search but are
    scored with the resulting filter.  Expected values come from the oracle's own column fit on the kept pixels."""
    from oracle import cmf_oracle as orc
    L, S = 700, 6
    cube = synth.make_cube(L, S, seed=35, bad_pixels=True)
    rng = np.random.default_rng(35)
    ok = cube[:, 400, :] > 0
    hot = (rng.random((L, S)) < 0.03) & ok
    cube[:, 330:423, :] = np.where(hot[:, None, :], cube[:, 330:423, :] * 4.0 + 5.0, cube[:, 330:423, :])   # flares
    wave = _wavelengths()
    ab = synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1], 2]
    with ColumnwiseMF(L, 425, S, ACTIVE, ab) as eng:
        eng.upload(cube)
        flags = eng.pixel_flags(cube, masks.flag_spec(wave))
        excl = (flags & masks.SATURATED) != 0
        assert excl.sum() > 20 and np.array_equal(flags, po.pixel_flags(cube, wave))
        eng.run()
        plain = eng.results()
        eng.set_exclusion(excl)
        eng.run()
        got = eng.results()
        nbg = eng.nvalid()
        eng.set_exclusion(None)
        eng.run()
        back = eng.results()
    assert np.array_equal(back["mf"], plain["mf"], equal_nan=True)          # clearing restores the default
    assert np.array_equal(got["mask"], plain["mask"])                         # validity is untouched (:282)
    assert not np.array_equal(got["mf"], plain["mf"])
    alphas, nll = orc.alpha_grid(), np.zeros(201)
    for c in range(S):
        full = cube[:, ACTIVE[0] - 1:ACTIVE[1], c]
        use = orc.valid_rows(full)
        keep = use[~excl[use, c]]
        assert nbg[c] == len(keep)
        fit = orc.column_filter(np.float64(full[keep]), ab, alphas, nll, len(keep))
        want = (np.float64(full[use]) - fit["mu"]).dot(fit["weights"])
        assert got["alpha_index"][c] == fit["alpha_index"]
        sd = np.std(want)
        assert np.max(np.abs(got["mf"][use, c] - want)) / sd < 1e-6
        assert np.all(got["mf"][~plain["mask"][:, c], c] == -9999.0)
        assert got["colnum"][c] == len(use)
        assert got["colavg"][c] == pytest.approx(np.mean(want), abs=1e-6 * sd)
        assert got["colstd"][c] == pytest.approx(sd, rel=1e-8)


# ---- against the reference's own functions, executed by oracle/make_product_golden.py (tests/golden/*.npz) ----
import glob
import os

from tests.test_products_oracle import GOLDEN, load_flag_case


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "flags_*.npz"))))
def test_pixel_flags_against_reference_fixture(path):
    """masks_sds.py:133-233 executed unmodified -> fixture; the CUDA flags must equal it bit for bit."""
    cube, wave, want, _ = load_flag_case(path)
    assert np.array_equal(masks.pixel_flags(cube, wave), want)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "profile_*.npz"))))
def test_column_profile_against_reference_fixture(path):
    """triage/cmf_profile.py:110-133 + srcfinder_util.extrema executed unmodified -> fixture (float32, bit-exact)."""
    z = np.load(path)
    for robust, tag in ((False, "plain"), (True, "robust")):
        got = cmf_profile.column_profile_image(z["mf"], nodata=-9999.0, robust=robust)
        names = ("npix", "med", "mad", "p05", "p95") if robust else ("npix", "avg", "std", "min", "max")
        for key, name in zip(("colnum", "colavg", "colstd", "colmin", "colmax"), names):
            want = np.asarray(z["%s_%s" % (tag, key)], dtype=np.float64)
            assert np.array_equal(np.asarray(got[name], dtype=np.float64), want, equal_nan=True), (tag, key)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLDEN, "filtdet_*.npz"))))
def test_detection_prefilter_against_reference_fixture(path):
    """srcfinder_util.kde + the head of filtdet executed unmodified -> fixture.  Same summation order as scipy's
    correlate1d without fused multiply-adds: masks identical, detkde to 1e-12 (it is bit-identical unless the
    host C library contracted an operation)."""
    from srcfinder_b200 import detect
    z = np.load(path)
    det, cmin, dmask = detect.filtdet_prefilter(z["mf"], k=int(z["k"]), mfmin=int(z["mfmin"]), mfmax=int(z["mfmax"]))
    assert np.array_equal(cmin, z["ch4min"])
    assert np.array_equal(dmask, z["detmask"])
    assert np.max(np.abs(det - z["detkde"])) <= 1e-12
    assert np.mean(det == z["detkde"]) > 0.99


def test_cnn_input_against_reference_fixture():
    """ClampCH4 + transforms.Normalize executed from cnn/cnn_pred_pipeline.py -> fixture, float32 bit-exact."""
    from srcfinder_b200 import detect
    z = np.load(os.path.join(GOLDEN, "cnnnorm_70x33.npz"))
    for name in z["names"]:
        got = detect.cnn_input(np.float64(z["x"]), model=str(name))
        assert got.dtype == np.float32
        assert np.array_equal(got, z["out_" + str(name)])


def test_prefilter_and_cnn_input_from_device_scores():
    """Both steps can read the scores where the filter left them (no host round trip): same result as via the host."""
    from srcfinder_b200 import detect
    cube = synth.make_cube(300, 40, seed=27)
    ab = synth.load_ch4_library()[ACTIVE[0] - 1:ACTIVE[1], 2]
    L, B, S = cube.shape
    with ColumnwiseMF(L, B, S, ACTIVE, ab) as eng:
        eng.upload(cube)
        eng.run()
        mf = eng.mf()
        a = detect.filtdet_prefilter(None, engine=eng)
        x = detect.cnn_input(None, engine=eng)
    b = detect.filtdet_prefilter(mf)
    assert all(np.array_equal(p, q) for p, q in zip(a, b))
    assert np.array_equal(x, detect.cnn_input(mf))
    ref = po.detection_prefilter(mf)
    assert np.array_equal(a[2], ref[2]) and np.max(np.abs(a[0] - ref[0])) <= 1e-12
