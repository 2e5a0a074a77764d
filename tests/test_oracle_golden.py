"""The numpy restatement (oracle/cmf_oracle.py) against outputs of the UNMODIFIED reference.

The reference holds no tests or golden vectors for cmf/robust_mf.py; these
fixtures were produced by running the reference script itself (oracle/make_golden.py).
"""
import numpy as np
import pytest

from oracle import cmf_oracle as orc
from srcfinder_b200 import synth
from tests.golden_util import case_names, load_case


def _abscf(active):
    lib = synth.load_ch4_library()
    return lib[active[0] - 1:active[1], 2]


@pytest.mark.parametrize("name", case_names())
def test_restatement_matches_reference_run(name):
    case = load_case(name)
    cube, active = case["cube"], case["active"]
    assert orc.active_window(case["libname"], case["reflectance"]) == active
    res = orc.cmf_cube(cube, _abscf(active), active, model=case["model"],
                       reflectance=case["reflectance"], labels=case.get("labels"),
                       reject_min=case["reject_min"] or None, regfull=case["regfull"])
    ref_mf = case["product"][..., -1]
    # masks: pixels left at nodata must be identical (invalid pixels, plus the rejected clusters with -r)
    if case["kmodes"] > 1:
        rejected = case["bgmeta"][..., 0] < 0
        assert np.array_equal(ref_mf == -9999.0, ~res["mask"] | rejected)
        assert np.array_equal(res["mf"] == -9999.0, ref_mf == -9999.0)
        ok = res["mask"] & ~rejected
    else:
        assert np.array_equal(ref_mf == -9999.0, ~res["mask"])
        ok = res["mask"]
    a, b = res["mf"][ok], ref_mf[ok]
    assert np.array_equal(np.isnan(a), np.isnan(b))
    fin = np.isfinite(b)
    # same LAPACK calls in the same order -> agreement far below the 1e-3 sigma tolerance
    scale = np.nanstd(b[fin]) if fin.any() else 1.0
    assert np.max(np.abs(a[fin] - b[fin])) <= 1e-9 * scale
    if "bgmeta" in case and case["kmodes"] == 1:
        for col in range(cube.shape[2]):
            used = ok[:, col]
            if used.any():
                assert set(np.unique(case["bgmeta"][used, col, 1])) == {res["alpha_index"][col]}
    # the per-column statistics the reference prints (:392), 7 significant digits
    for col in range(cube.shape[2]):
        if np.isfinite(case["stdout_std"][col]):
            assert res["colstd"][col] == pytest.approx(case["stdout_std"][col], rel=2e-6)
            assert res["colavg"][col] == pytest.approx(case["stdout_avg"][col], rel=1e-3, abs=1e-9)
    prod = orc.assemble_product(cube, res["mf"], res["colnum"])
    assert np.array_equal(prod[..., :3], case["product"][..., :3])


def test_alpha_grid_and_window():
    a = orc.alpha_grid()
    assert len(a) == 201 and a[0] == 1e-10 and abs(a[-1] - 1.0) < 1e-9
    assert orc.min_cluster_samples([351, 422]) == 85


def test_known_answer_scaled_target():
    """MF of mu + s*t is s*1e5 for any positive-definite model (SURVEY.md 8c self-check)."""
    rng = np.random.default_rng(5)
    cube = synth.make_cube(400, 2, seed=21, plume=False)
    active = [351, 422]
    ab = _abscf(active)
    res = orc.cmf_cube(cube, ab, active, columns=[0])
    mu, w = res["mu"][0], res["weights"][0]
    s = 3.0e-3
    probe = mu + s * (ab * mu)
    assert (probe - mu).dot(w) == pytest.approx(s * 1e5, rel=1e-9)
    assert abs(res["colavg"][0]) < 1e-6 * res["colstd"][0]


import glob
import os

_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(_GOLDEN, "looshrinkage_*.npz"))))
def test_loo_shrinkage_restatement_matches_reference_function(path):
    """oracle.loo_shrinkage against the reference's importable looshrinkage() called directly (fixtures of
    oracle/make_golden.py looshrinkage), with and without the -f regulariser I_reg (:99, :131): same index, same
    finite set, same nll and C to rounding (same LAPACK calls in the same order)."""
    z = np.load(path)
    nll = np.zeros(len(z["alphas"]))
    reg = z["I_reg"] if "I_reg" in z.files else None
    c_mat, mindex = orc.loo_shrinkage(z["I_zm"], z["alphas"], nll, int(z["n"]), x_reg=reg)
    assert mindex == int(z["mindex"])
    assert np.array_equal(np.isfinite(nll), np.isfinite(z["nll"]))
    fin = np.isfinite(z["nll"])
    assert np.allclose(nll[fin], z["nll"][fin], rtol=1e-12, atol=0.0)
    assert np.allclose(c_mat, z["C"], rtol=1e-13, atol=0.0)
