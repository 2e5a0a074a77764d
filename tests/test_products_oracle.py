"""CPU checks of the oracles for SURVEY.md 8(f) rows 2 and 3 (oracle/products_oracle.py) and of the host-side
argument translation; no GPU."""
import numpy as np
import pytest

from oracle import products_oracle as po
from srcfinder_b200 import cmf_profile, masks


def test_flag_spec_translates_wavelength_window_to_bands():
    wave = 376.86 + 5.0087 * np.arange(425)
    spec = masks.flag_spec(wave)
    sel = np.flatnonzero((wave >= 1945) & (wave <= 2485))
    assert (spec.sat_lo, spec.sat_hi) == (sel[0], sel[-1])
    assert (spec.spec_band, spec.dark_band, spec.cloud_a, spec.cloud_b) == (25, 352, 15, 60)
    assert spec.sat_thresh == 6.0 and spec.spec_thresh == 9.0 and spec.cloud_thresh == 15.0
    assert spec.dark_thresh == np.float32(0.104)
    assert spec.cloud_dwl == np.float32(wave[60] - wave[15])
    with pytest.raises(ValueError):
        masks.flag_spec(wave, waverange=(100, 200))
    short = masks.flag_spec(wave[:330], waverange=(1945, 2020))       # bands 352 absent: test disabled
    assert short.dark_band == -1


def test_oracle_flags_semantics():
    L, S = 4, 3
    cube = np.full((L, 425, S), 1.0, dtype=np.float32)
    wave = 376.86 + 5.0087 * np.arange(425)
    cube[0, 400, 0] = 6.5                         # saturated
    cube[1, 400, 1] = 6.5; cube[1, 25, 1] = 9.5   # saturated + specular
    cube[2, 352, 2] = 0.05                        # dark
    cube[3, 352, 0] = -9999.0                     # no-data, not dark
    cube[3, 15, 1] = 20.0; cube[3, 60, 1] = 5.0   # cloud: bright and falling
    cube[3, 15, 2] = 20.0; cube[3, 60, 2] = 30.0  # bright but rising: not a cloud
    f = po.pixel_flags(cube, wave)
    assert f[0, 0] == po.SATURATED and f[1, 1] == po.SATURATED | po.SPECULAR
    assert f[2, 2] == po.DARK and f[3, 0] == 0
    assert f[3, 1] == po.CLOUD and f[3, 2] == 0
    # the b->c slope is numpy's `out` argument in the reference: it must not change the result
    cube[3, 175, 1] = 100.0
    assert po.pixel_flags(cube, wave)[3, 1] == po.CLOUD


def test_oracle_profile_is_numpy_sequential_float32():
    """The device replays numpy's evaluation order; this pins that order on this container's numpy."""
    rng = np.random.default_rng(5)
    L, S = 2500, 4
    mf = rng.normal(0, 400, (L, S))
    mf[rng.random((L, S)) < 0.05] = -9999.0
    r = po.column_profile(mf)
    cmf = np.float32(mf)
    m = (cmf != np.float32(-9999)) & (cmf > 0)
    for c in range(S):
        tot = np.float32(0)
        for v in np.where(m[:, c], cmf[:, c], np.float32(0)):
            tot = np.float32(tot + v)
        assert r["avg"][c] == np.float32(np.float64(tot) / m[:, c].sum())
    rr = po.column_profile(mf, use_robust_stats=True)
    for c in range(S):
        v = np.sort(cmf[m[:, c], c]); n = len(v)
        med = v[n // 2] if n % 2 else np.float32(np.float32(v[n // 2 - 1] + v[n // 2]) * np.float32(0.5))
        assert rr["med"][c] == med
        assert rr["p05"][c] == v[int(np.rint((n - 1) * (((1 - 0.95) * 100) / 100)))]
        assert rr["p95"][c] == v[int(np.rint((n - 1) * ((0.95 * 100) / 100)))]


def test_profile_cli_parser_matches_reference_flags():
    a = cmf_profile.build_parser().parse_args(["--robust", "--outdir", "x", "-j", "4", "a_cmf", "b_cmf"])
    assert a.robust and a.outdir == "x" and a.cmffiles == ["a_cmf", "b_cmf"] and a.jobs == 4
    assert cmf_profile.PLAIN_COLS == ["npix", "avg", "std", "min", "max"]
    assert cmf_profile.ROBUST_COLS == ["npix", "med", "mad", "p05", "p95"]


# ---- the restatements against the reference's own code, executed by oracle/make_product_golden.py ----
import glob
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _fixtures(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "_*.npz")))


def load_flag_case(path):
    z = np.load(path)
    cube = np.zeros(tuple(int(v) for v in z["shape"]), dtype=np.float32)
    cube[:, z["kept_bands"], :] = z["cube_kept"]
    want = (z["saturated"] * po.SATURATED + z["specular"] * po.SPECULAR + z["dark"] * po.DARK +
            z["cloud"] * po.CLOUD).astype(np.uint8)
    return cube, z["wave"], want, z


@pytest.mark.parametrize("path", _fixtures("flags"))
def test_flags_restatement_matches_reference_functions(path):
    cube, wave, want, z = load_flag_case(path)
    assert float(z["sat_thresh"]) == 6.0 and float(z["spec_thresh"]) == 9.0 and float(z["dark_thresh"]) == 0.104
    assert want.any() and len(np.unique(want)) >= 5
    assert np.array_equal(po.pixel_flags(cube, wave), want)


@pytest.mark.parametrize("path", _fixtures("profile"))
def test_profile_restatement_matches_reference_statements(path):
    z = np.load(path)
    for robust, tag in ((False, "plain"), (True, "robust")):
        r = po.column_profile(z["mf"], use_robust_stats=robust)
        names = ("npix", "med", "mad", "p05", "p95") if robust else ("npix", "avg", "std", "min", "max")
        for key, name in zip(("colnum", "colavg", "colstd", "colmin", "colmax"), names):
            assert np.array_equal(r[name], z["%s_%s" % (tag, key)], equal_nan=True), (tag, key)


@pytest.mark.parametrize("path", _fixtures("filtdet"))
def test_prefilter_restatement_matches_reference_statements(path):
    z = np.load(path)
    det, cmin, dmask = po.detection_prefilter(z["mf"], k=int(z["k"]), mfmin=int(z["mfmin"]), mfmax=int(z["mfmax"]))
    assert np.array_equal(det, z["detkde"]) and np.array_equal(cmin, z["ch4min"]) and np.array_equal(dmask, z["detmask"])
    assert dmask.any() and not dmask.all()


def test_cnn_input_restatement_matches_reference_transform():
    from srcfinder_b200 import detect
    z = np.load(os.path.join(GOLDEN, "cnnnorm_70x33.npz"))
    assert sorted(str(n) for n in z["names"]) == sorted(detect.CNN_MODELS)
    for name in z["names"]:
        vmin, vmax, mean, std = z["par_" + str(name)]
        assert tuple(detect.CNN_MODELS[str(name)]) == (vmin, vmax, mean, std)
        assert np.array_equal(po.cnn_input(z["x"], vmin, vmax, mean, std), z["out_" + str(name)])


def test_gaussian_weights_are_scipys():
    from scipy.ndimage import correlate1d, gaussian_filter1d
    from srcfinder_b200 import detect
    w = detect.gaussian_weights(50)
    assert len(w) == 101 and detect.KERNEL == 50 and (detect.MFMIN, detect.MFMAX) == (500, 1500)
    x = np.random.default_rng(3).normal(size=400)
    assert np.array_equal(correlate1d(x, w[::-1], mode="reflect"), gaussian_filter1d(x, 50, truncate=1))
