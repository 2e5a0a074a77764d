"""Background modes found on the device (-k > 1): PCA projection, deterministic k-means, -r and -f.

The reference's MiniBatchKMeans is unseeded (cmf/robust_mf.py:312) -- its own partition differs from run to
run -- so the partition rule is this implementation's (csrc/k_cluster.cu) and is checked against its numpy
restatement (oracle/cluster_oracle.py) bit for bit; everything downstream of the labels is checked against
the reference restatement (oracle/cmf_oracle.py) and the goldens of seeded reference runs.
"""
import numpy as np
import pytest

from oracle import cluster_oracle as clo
from oracle import cmf_oracle as orc
from srcfinder_b200 import ColumnwiseMF, cmf_cube, synth

pytestmark = pytest.mark.gpu

ACTIVE = [351, 422]


def _abscf(active=ACTIVE):
    return synth.load_ch4_library()[active[0] - 1:active[1], 2]


def _two_population_cube(L, S, seed, bright=(120, 170)):
    """Synthetic cube with a darker and a brighter surface (line blocks) and a small very bright block."""
    cube = synth.make_cube(L, S, seed=seed, bad_pixels=True)
    ok = cube > 0
    half = slice(L // 2, L)
    cube[half] = np.where(ok[half], cube[half] * 1.6, cube[half])
    l0, l1 = bright
    cube[l0:l1] = np.where(ok[l0:l1], cube[l0:l1] * 3.0, cube[l0:l1])
    return cube


@pytest.mark.parametrize("k,pcadim", [(3, 6), (2, 4), (5, 8)])
def test_partition_matches_its_numpy_restatement(k, pcadim):
    L, S = 1200, 6
    cube = _two_population_cube(L, S, seed=91)
    got = cmf_cube(cube, _abscf(), ACTIVE, kmodes=k, pcadim=pcadim)
    for c in range(S):
        use = got["mask"][:, c]
        x = np.float64(cube[use, ACTIVE[0] - 1:ACTIVE[1], c])
        y_ref, _ = clo.pca_projections(x, pcadim)
        y_dev = got["pca"][c][use]
        # the projections agree with numpy's eigh to rounding (relative to the largest component)
        assert np.max(np.abs(y_dev - y_ref)) <= 1e-8 * np.max(np.abs(y_ref))
        assert np.all(got["pca"][c][~use] == 0.0)
        # the k-means is exact integer arithmetic on the quantised projections: labels identical
        lab, iters = clo.kmeans_labels(y_dev, k)
        assert np.array_equal(got["labels"][use, c], lab), "column %d" % c
        assert got["kmeans_iters"][c] == iters
        assert np.all(got["labels"][~use, c] == 0)


@pytest.mark.parametrize("reject,regfull", [(True, False), (True, True), (False, True)])
def test_clustered_run_equals_labelled_run_and_oracle(reject, regfull):
    """The scores of a clustered run are those of the reference's mode loop (:339-386) for the labels the
    device found: identical to a run that is handed the same labels, and equal to the oracle's."""
    L, S, k = 1500, 5, 3
    cube = _two_population_cube(L, S, seed=92, bright=(400, 440))
    ab = _abscf()
    rmin = orc.min_cluster_samples(ACTIVE) if reject else 0
    got = cmf_cube(cube, ab, ACTIVE, kmodes=k, reject_min=rmin, regfull=regfull)
    again = cmf_cube(cube, ab, ACTIVE, labels=got["labels"], reject_min=rmin, regfull=regfull)
    for key in ("mf", "alpha_index", "colstd", "cluster_id", "alpha_image"):
        assert np.array_equal(got[key], again[key], equal_nan=True), key
    ref = orc.cmf_cube(cube, ab, ACTIVE, labels=got["labels"], reject_min=rmin or None, regfull=regfull)
    assert np.array_equal(got["mf"] == -9999.0, ref["mf"] == -9999.0)
    if reject:
        assert (got["cluster_id"] < 0).any(), "the 40-line block must be rejected (bgminsamp = 85)"
    for c in range(S):
        ok = ref["mf"][:, c] != -9999.0
        err = np.max(np.abs(got["mf"][ok, c] - ref["mf"][ok, c])) / np.std(ref["mf"][ok, c])
        assert err <= 1e-6, "column %d: %.3g sigma" % (c, err)
        assert got["colstd"][c] == pytest.approx(ref["colstd"][c], rel=1e-8)
    assert np.array_equal(got["alpha_index"], ref["alpha_index"])


def test_regfull_with_given_labels_against_oracle():
    """-f with labels handed in: T = covariance of the whole column (looshrinkage's I_reg, :100, :131, :358)."""
    L, S, k = 1100, 4, 3
    cube = synth.make_cube(L, S, seed=93, bad_pixels=True)
    ab = _abscf()
    labels = (np.arange(L)[:, None] * k // L).astype(np.int32).repeat(S, 1)
    ref = orc.cmf_cube(cube, ab, ACTIVE, labels=labels, regfull=True)
    got = cmf_cube(cube, ab, ACTIVE, labels=labels, regfull=True)
    plain = cmf_cube(cube, ab, ACTIVE, labels=labels)
    assert np.array_equal(got["alpha_index"], ref["alpha_index"])
    for c in range(S):
        ok = ref["mask"][:, c]
        err = np.max(np.abs(got["mf"][ok, c] - ref["mf"][ok, c])) / ref["colstd"][c]
        assert err <= 1e-6
    assert not np.array_equal(got["mf"], plain["mf"])     # the regulariser changes the fit


def test_cli_multimodal_products(tmp_path):
    """robust_mf -k 3 -r -f -m: header string, _bgmeta bands and scores of the drop-in CLI."""
    from srcfinder_b200 import envi, robust_mf
    L, S = 900, 4
    cube = _two_population_cube(L, S, seed=94, bright=(100, 140))
    inp, out = str(tmp_path / "scene_rdn"), str(tmp_path / "scene_mf")
    lib = synth.write_library_txt(str(tmp_path / "ang_ch4_unit.txt"))
    mm = envi.create_image(inp, {"samples": S, "lines": L, "bands": 425, "data type": 4, "interleave": "bil",
                                 "byte order": 0, "data ignore value": -9999})
    mm[:] = cube
    mm.flush()
    assert robust_mf.main(["-k", "3", "-r", "-f", "-m", inp, lib, out]) == 0
    hdr = envi.read_header(out + ".hdr")
    # brace values come back as comma-split items (as spectral's reader returns them)
    assert ", ".join(hdr["model parameters"]) == ("modelname=looshrinkage, bgmodel=multimodal, bgmodes=3, pcadim=6, "
                                                  "reject=True, regfull=True, aminexp=-10.0, amaxexp=0.0, astep=0.05, "
                                                  "reflectance=False, active_bands=[351, 422]")
    prod = np.asarray(envi.open_memmap(out))
    bg = np.asarray(envi.open_memmap(out + "_bgmeta"))
    want = cmf_cube(cube, _abscf(), ACTIVE, kmodes=3, reject_min=85, regfull=True)
    assert np.array_equal(prod[..., 3], want["mf"], equal_nan=True)
    assert np.array_equal(bg[..., 0], np.where(want["mask"], want["cluster_id"], 0))
    assert np.array_equal(bg[..., 1], np.where(want["mask"], want["alpha_image"], 0))
    assert (bg[..., 0] < 0).any()


def test_c3_robust_flightline_with_rejection():
    """configs[2] at full size: 598 x 425 x 20000 with bad pixels, k = 3 modes with outlier rejection, all on
    the device.  Properties: every pixel the mask keeps and the partition does not reject is scored, rejected
    and invalid pixels keep nodata, rerun is bitwise identical; two columns agree with the oracle given the
    device's labels."""
    import torch
    L, S = 20000, 598
    ab = _abscf()
    slab = synth.make_slab_torch(L, S, ACTIVE[0], ACTIVE[1], "cuda", seed=4)
    slab[9000:9060] *= 8.0                         # a 60-line anomalous block: below bgminsamp = 85
    synth.inject_bad_pixels_torch(slab, 4)
    torch.cuda.synchronize()
    with ColumnwiseMF(L, 425, S, ACTIVE, ab) as eng:
        eng.bind_device(slab.data_ptr())
        eng.set_clustering(3, pcadim=6, reject_min=85)
        eng.run()
        r1 = eng.results()
        cid, labels, iters = eng.cluster_id(), eng.labels(), eng.kmeans_iters()
        eng.run()
        r2 = eng.results()
    assert np.array_equal(r1["mf"], r2["mf"], equal_nan=True)
    assert iters.max() <= 100 and iters.min() >= 1
    rejected = cid < 0
    assert rejected[9000:9060].mean() > 0.8 and rejected.mean() < 0.02
    assert np.array_equal(r1["mf"] == -9999.0, ~r1["mask"] | rejected)
    c0 = 250
    host = np.zeros((L, 425, 2), dtype=np.float32)
    host[:, ACTIVE[0] - 1:ACTIVE[1], :] = slab[:, :, c0:c0 + 2].cpu().numpy()
    ref = orc.cmf_cube(host, ab, ACTIVE, labels=labels[:, c0:c0 + 2], reject_min=85)
    for j in range(2):
        ok = ref["mf"][:, j] != -9999.0
        assert np.array_equal(ok, r1["mf"][:, c0 + j] != -9999.0)
        err = np.max(np.abs(ref["mf"][ok, j] - r1["mf"][ok, c0 + j])) / np.std(ref["mf"][ok, j])
        assert err < 1e-6


def test_all_clusters_rejected_keeps_flagged_ids_in_bgmeta():
    """When every cluster of a column is below bgminsamp (possible only without a label 0: -0 == 0 never flips,
    cmf/robust_mf.py:323-324) the reference warns, proceeds WITHOUT rejection (:330-332) -- but _bgmeta band 0 was
    already written with the negated ids inside the counting loop (:326-327).  Scores as if nothing was rejected,
    cluster image with negative ids."""
    L, S = 400, 3
    cube = synth.make_cube(L, S, seed=95)
    active = [351, 422]
    ab = synth.load_ch4_library()[active[0] - 1:active[1], 2]
    labels = np.ones((L, S), dtype=np.int32)
    labels[L // 2:, :] = 2
    labels[:, 2] = np.where(np.arange(L) < 150, 0, 1)           # a column WITH label 0: ordinary rejection rules
    big = 10 * L                                                 # every cluster is "too small"
    got = cmf_cube(cube, ab, active, labels=labels, reject_min=big)
    ref = orc.cmf_cube(cube, ab, active, labels=labels, reject_min=big)
    free = cmf_cube(cube, ab, active, labels=labels, reject_min=0)
    assert np.array_equal(got["mf"] == -9999.0, ref["mf"] == -9999.0)
    for c in range(S):
        ok = ref["mf"][:, c] != -9999.0
        err = np.max(np.abs(got["mf"][ok, c] - ref["mf"][ok, c])) / np.std(ref["mf"][ok, c])
        assert err <= 1e-6
    # columns 0, 1: all rejected -> scored like the run without rejection, ids negative in the image
    assert np.array_equal(got["mf"][:, :2], free["mf"][:, :2])
    assert np.array_equal(got["cluster_id"][:, :2], -labels[:, :2])
    # column 2: label 0 survives, label 1 is rejected (negative id, scores stay nodata)
    assert np.array_equal(got["cluster_id"][:, 2], np.where(labels[:, 2] == 1, -1, 0))
    assert np.all(got["mf"][labels[:, 2] == 1, 2] == -9999.0)
