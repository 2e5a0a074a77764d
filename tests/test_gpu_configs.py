"""BASELINE.json configs[2..4] on the GPU: the bad-pixel flightline (C3), the flightline batch (C4) and the
EMIT-shaped cube with an explicit active window (C5).  Full sizes are checked through size-independent
properties plus the oracle on a few columns of the same cube (the oracle needs ~1 s per 20k-line column).
"""
import numpy as np
import pytest

from oracle import cmf_oracle as orc
from srcfinder_b200 import ColumnwiseMF, cmf_cube, synth

pytestmark = pytest.mark.gpu

TIGHT_SIGMA = 1e-7


def _abscf(active):
    return synth.load_ch4_library()[active[0] - 1:active[1], 2]


def test_c3_bad_pixel_flightline_properties():
    """configs[2]: 598 x 425 x 20000 with 0.5 % nodata pixels, NaN / negative / inf single-band defects and
    saturated spectra.  The valid mask must equal the reference rule (:282) bit for bit, dropped pixels keep
    nodata, the filter stays normalised, and two columns agree with the oracle."""
    import torch
    L, S, active = 20000, 598, [351, 422]
    ab = _abscf(active)
    slab = synth.make_slab_torch(L, S, active[0], active[1], "cuda", seed=3)
    bad = synth.inject_bad_pixels_torch(slab, 3)
    want_mask = (torch.isfinite(slab) & ~(slab < 0)).all(dim=1)
    assert torch.equal(~want_mask, bad)
    torch.cuda.synchronize()
    with ColumnwiseMF(L, 425, S, active, ab) as eng:
        eng.bind_device(slab.data_ptr())
        eng.run()
        r = eng.results()
        nvalid = eng.nvalid()
    want = want_mask.cpu().numpy()
    assert np.array_equal(r["mask"], want)
    assert np.array_equal(nvalid, want.sum(axis=0))
    assert np.all(r["mf"][~want] == -9999.0) and np.all(r["mf"][want] != -9999.0)
    assert np.array_equal(r["colnum"], want.sum(axis=0).astype(np.float64))
    assert np.all(r["status"] == 0)
    t = ab[None, :] * r["mu"]
    assert np.allclose(np.sum(r["weights"] * t, axis=1), 1.0e5, rtol=1e-9)
    assert np.all(np.abs(r["colavg"]) < 1e-7 * r["colstd"])
    # column statistics are those of the written scores (:388-391)
    for c in (0, 17, 597):
        v = r["mf"][want[:, c], c]
        assert r["colstd"][c] == pytest.approx(np.std(v), rel=1e-10)
    c0 = 300
    host = np.zeros((L, 425, 2), dtype=np.float32)
    host[:, active[0] - 1:active[1], :] = slab[:, :, c0:c0 + 2].cpu().numpy()
    ref = orc.cmf_cube(host, ab, active)
    assert np.array_equal(ref["mask"], want[:, c0:c0 + 2])
    assert np.array_equal(ref["alpha_index"], r["alpha_index"][c0:c0 + 2])
    for j in range(2):
        ok = ref["mask"][:, j]
        err = np.max(np.abs(ref["mf"][ok, j] - r["mf"][ok, c0 + j])) / ref["colstd"][j]
        assert err < TIGHT_SIGMA


def test_c4_flightline_batch_on_one_gpu():
    """configs[3] at one GPU: a batch of flightlines streamed from pinned host cubes through two contexts
    (cmf_run_host, CMF_RUN_ASYNC) gives, flightline by flightline, the bits of the one-at-a-time path; the
    rank's share is dist.flightline_shard()."""
    import torch
    from srcfinder_b200 import dist
    nflight, L, S, active = 6, 700, 40, [351, 422]
    ab = _abscf(active)
    mine = dist.flightline_shard(nflight, 1, 0)
    assert mine == list(range(nflight))
    cubes = [torch.from_numpy(synth.make_cube(L, S, seed=200 + f, bad_pixels=(f % 2 == 1))).pin_memory()
             for f in mine]
    outs = [torch.empty((L, S), dtype=torch.float64).pin_memory() for _ in mine]
    aidx = [torch.empty(S, dtype=torch.int32).pin_memory() for _ in mine]
    engs = [ColumnwiseMF(L, 425, S, active, ab) for _ in range(2)]
    for i in range(len(mine)):
        e = engs[i % 2]
        e.sync()
        e.run_host(cubes[i].data_ptr(), outs[i].data_ptr(), None, aidx[i].data_ptr(), wait=False)
    for e in engs:
        e.sync()
        e.close()
    for i in (0, 3, 5):
        one = cmf_cube(cubes[i].numpy(), ab, active)
        assert np.array_equal(outs[i].numpy(), one["mf"], equal_nan=True)
        assert np.array_equal(aidx[i].numpy(), one["alpha_index"])
    ref = orc.cmf_cube(np.ascontiguousarray(cubes[1].numpy()[:, :, :4]), ab, active)
    got = outs[1].numpy()[:, :4]
    assert np.array_equal(got != -9999.0, ref["mask"])
    for c in range(4):
        ok = ref["mask"][:, c]
        assert np.max(np.abs(got[ok, c] - ref["mf"][ok, c])) / ref["colstd"][c] < TIGHT_SIGMA


def emit_case(L, S, seed=5):
    """EMIT-shaped inputs (configs[4]): 285 bands on a 381-2493 nm grid, the CH4 library resampled onto it,
    active window = the bands whose centres fall in 2129-2485 nm."""
    wl = np.linspace(381.0, 2493.0, 285)
    lib = synth.resample_library(wl)
    inside = np.where((wl >= 2129.0) & (wl <= 2485.0))[0]
    active = [int(inside[0]) + 1, int(inside[-1]) + 1]
    cube = synth.make_cube(L, S, bands=285, seed=seed, lib=lib)
    return cube, lib, active


def test_c5_emit_shape_explicit_window():
    """configs[4]: 1242 cols x 285 ch x 1280 lines.  The window is an argument here (the reference hard-codes
    AVIRIS-NG indices, :186-191); 16 spread columns are checked against the oracle, all of them through the
    filter's normalisation and zero-mean properties."""
    L, S = 1280, 1242
    cube, lib, active = emit_case(L, S)
    D = active[1] - active[0] + 1
    assert 40 <= D <= 56
    ab = lib[active[0] - 1:active[1], 2]
    with ColumnwiseMF(L, 285, S, active, ab) as eng:
        eng.upload(cube)
        eng.run()
        r = eng.results()
    assert r["mask"].all() and np.all(r["status"] == 0)
    t = ab[None, :] * r["mu"]
    assert np.allclose(np.sum(r["weights"] * t, axis=1), 1.0e5, rtol=1e-9)
    assert np.all(np.abs(r["colavg"]) < 1e-7 * r["colstd"])
    cols = np.linspace(0, S - 1, 16).astype(int)
    ref = orc.cmf_cube(np.ascontiguousarray(cube[:, :, cols]), ab, active)
    assert np.array_equal(ref["alpha_index"], r["alpha_index"][cols])
    for j, c in enumerate(cols):
        err = np.max(np.abs(ref["mf"][:, j] - r["mf"][:, c])) / ref["colstd"][j]
        assert err < TIGHT_SIGMA, "column %d: %.3g sigma" % (c, err)


def test_c5_emit_cli_active_argument(tmp_path):
    """The drop-in CLI takes the window through --active (extension) and writes the reference's products."""
    from srcfinder_b200 import envi, robust_mf
    L, S = 320, 24
    cube, lib, active = emit_case(L, S, seed=6)
    inp, out = str(tmp_path / "emit_rdn"), str(tmp_path / "emit_mf")
    libpath = synth.write_library_txt(str(tmp_path / "emit_ch4_unit.txt"), lib)
    mm = envi.create_image(inp, {"samples": S, "lines": L, "bands": 285, "data type": 4, "interleave": "bil",
                                 "byte order": 0, "data ignore value": -9999})
    mm[:] = cube
    mm.flush()
    rc = robust_mf.main(["--active", "%d,%d" % tuple(active), "--rgb_bands", "35,20,10", inp, libpath, out])
    assert rc == 0
    prod = np.asarray(envi.open_memmap(out))
    assert prod.shape == (L, S, 4)
    ref = orc.cmf_cube(cube, lib[active[0] - 1:active[1], 2], active)
    for c in range(S):
        assert np.max(np.abs(prod[:, c, 3] - ref["mf"][:, c])) / ref["colstd"][c] < TIGHT_SIGMA
    assert np.array_equal(prod[:, :, 0], cube[:, 35, :].astype(np.float64))
    hdr = envi.read_header(out + ".hdr")
    # the header reader splits brace values on commas (as spectral does), so the window comes back in two items
    parms = ",".join(hdr["model parameters"]).replace(" ", "")
    assert "active_bands=[%d,%d]" % tuple(active) in parms
