"""Host logic of the drop-in CLI that needs no GPU: argument surface, active windows, header strings,
ENVI round trip, column-stats CSV -- checked against what the unmodified reference wrote (goldens)."""
import numpy as np
import pytest

from srcfinder_b200 import envi, robust_mf
from tests.golden_util import load_case


def test_argument_surface_matches_reference():
    p = robust_mf.build_parser()
    a = p.parse_args(["in", "lib_ch4.txt", "out"])
    assert (a.kmeans, a.pcadim, a.reject, a.full, a.rgb_bands, a.metadata, a.reflectance, a.model) == \
        (1, 6, False, False, "60,42,24", False, False, "looshrinkage")
    a = p.parse_args(["-k", "3", "-r", "-f", "-m", "-R", "-M", "empirical", "--pcadim", "4", "--rgb_bands",
                      "1,2,3", "in", "lib", "out"])
    assert (a.kmeans, a.pcadim, a.reject, a.full, a.metadata, a.reflectance, a.model) == \
        (3, 4, True, True, True, True, "empirical")


def test_active_windows():
    assert robust_mf.active_window("ang_ch4_unit.txt", False) == [351, 422]
    assert robust_mf.active_window("ang_ch4_unit.txt", True) == [5, 420]
    assert robust_mf.active_window("ang_co2_unit.txt", False) == [309, 391]
    assert robust_mf.active_window("something.txt", False) is None


@pytest.mark.parametrize("name", ["unimodal_300x6", "empirical_300x4", "co2window_300x4"])
def test_model_parameters_string_matches_reference_header(name, tmp_path):
    case = load_case(name)
    s = robust_mf.model_parameters_string(case["model"], 1, 6, False, False, case["reflectance"], case["active"])
    hdr = str(tmp_path / "x.hdr")
    envi.write_header(hdr, {"samples": 1, "lines": 1, "bands": 1, "data type": 5, "interleave": "bip",
                            "model parameters": s})
    assert envi.read_header(hdr)["model parameters"] == case["header"]["model parameters"]


def test_envi_roundtrip(tmp_path):
    meta = {"samples": 5, "lines": 4, "bands": 3, "data type": 4, "interleave": "bil", "byte order": 0,
            "data ignore value": -9999, "band names": ["a", "b", "c"], "description": "x y z"}
    mm = envi.create_image(str(tmp_path / "img"), meta)
    assert mm.shape == (4, 3, 5) and mm.dtype == np.float32
    mm[:] = np.arange(60, dtype=np.float32).reshape(4, 3, 5)
    mm.flush()
    back = envi.read_header(str(tmp_path / "img.hdr"))
    assert back["band names"] == ["a", "b", "c"] and back["interleave"] == "bil"
    assert back["description"] == "x y z"
    again = envi.open_memmap(str(tmp_path / "img"))
    assert np.array_equal(again, mm)


def test_column_stats_csv(tmp_path):
    cs = np.array([[300.0, -9999.0], [1.5e-12, -9999.0], [774.8, -9999.0]])
    path = str(tmp_path / "c.csv")
    robust_mf.write_column_stats(path, cs)
    rows = open(path).read().splitlines()
    assert rows[0] == ",0,1" and rows[1].startswith("npix,300.0,") and rows[3].split(",")[0] == "std"
