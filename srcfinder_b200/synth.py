"""Synthetic AVIRIS-NG-shaped radiance cubes.

float32 **BIL** cubes ``(lines, bands, samples)``: a smooth per-band mean radiance curve,
a rank-5 correlated background (first factor = illumination scaling), white sensor noise,
per-column gain jitter, an injected CH4 plume built from the unit-absorption library
(``cmf/ang_ch4_unit_3col_425chan.txt`` in the reference, shipped here as
``srcfinder_b200/data/ch4_unit_425.npy``), and optional bad pixels.

Two generators share the same recipe:
  * :func:`make_cube`        -- numpy, host memory, every band (tests, goldens, CPU baseline)
  * :func:`make_slab_torch`  -- torch, any device, only a band window (bench at flightline size)
"""
from __future__ import annotations

import os

import numpy as np

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")

NODATA = -9999.0


def load_ch4_library():
    """(425, 3) float64: [1-based channel, wavelength nm, unit absorption]."""
    return np.load(os.path.join(_DATA, "ch4_unit_425.npy"))


def write_library_txt(path, lib=None):
    """Write the 3-column text library the CLI takes as LIBRARY (name must contain 'ch4'/'co2')."""
    lib = load_ch4_library() if lib is None else lib
    with open(path, "w") as fh:
        for ch, wl, a in lib:
            fh.write("%03d  %8.2f %.12f\n" % (int(ch), wl, a))
    return path


def resample_library(wavelengths_nm, lib=None):
    """Unit absorption linearly interpolated onto another sensor's band centres (EMIT, C5)."""
    lib = load_ch4_library() if lib is None else lib
    wl = np.asarray(wavelengths_nm, dtype=np.float64)
    a = np.interp(wl, lib[:, 1], lib[:, 2], left=0.0, right=0.0)
    return np.stack([np.arange(1, len(wl) + 1, dtype=np.float64), wl, a], axis=1)


def mean_radiance(wavelengths_nm):
    """Smooth solar-like radiance curve (uW/nm/sr/cm2): ~8 in the VNIR, 0.3-0.7 at 2.1-2.5 um."""
    wl = np.asarray(wavelengths_nm, dtype=np.float64)
    rise = 1.0 - np.exp(-np.maximum(wl - 340.0, 0.0) / 60.0)
    curve = 0.15 + 8.0 * np.exp(-np.maximum(wl - 480.0, 0.0) / 600.0) * rise
    for centre, width, depth in ((1380.0, 40.0, 0.85), (1880.0, 50.0, 0.9), (940.0, 25.0, 0.4)):
        curve = curve * (1.0 - depth * np.exp(-((wl - centre) / width) ** 2))
    return curve


def _factors(wavelengths_nm, mu, rng):
    """(5, B) loading matrix: illumination (25 % of mu) + four smooth 2 % shapes."""
    wl = np.asarray(wavelengths_nm, dtype=np.float64)
    x = (wl - wl.min()) / (wl.max() - wl.min())
    F = [0.25 * mu]
    for k in range(1, 5):
        phase = rng.uniform(0, 2 * np.pi)
        F.append(0.02 * mu * np.cos(np.pi * k * x * 3.0 + phase))
    return np.stack(F, axis=0)


def plume_map(lines, samples, rng, nplumes=2, peak=(500.0, 5000.0)):
    """(L, S) ppm*m enhancement: a few elliptical Gaussian footprints, <= ~1 % of pixels."""
    ll, ss = np.meshgrid(np.arange(lines, dtype=np.float64),
                         np.arange(samples, dtype=np.float64), indexing="ij")
    out = np.zeros((lines, samples))
    for _ in range(nplumes):
        cl, cs = rng.uniform(0.1, 0.9) * lines, rng.uniform(0.1, 0.9) * samples
        sl = max(2.0, 0.03 * lines * rng.uniform(0.5, 1.5))
        sc = max(1.0, 0.02 * samples * rng.uniform(0.5, 1.5))
        amp = rng.uniform(*peak)
        out += amp * np.exp(-0.5 * (((ll - cl) / sl) ** 2 + ((ss - cs) / sc) ** 2))
    out[out < 1.0] = 0.0
    return out


def make_cube(lines, samples, bands=425, seed=1, lib=None, plume=True, bad_pixels=False,
              noise=0.004, return_truth=False):
    """float32 BIL cube ``(lines, bands, samples)`` per SURVEY.md 8(d).

    ``bad_pixels`` adds (C3): 0.5 % whole pixels = NODATA, 0.1 % single-band NaN inside the
    CH4 window, 0.1 % negative values, 0.05 % saturated (> 6.0) pixels.
    """
    rng = np.random.default_rng(seed)
    lib = load_ch4_library() if lib is None else lib
    assert lib.shape[0] == bands, "library rows must match band count"
    wl, absorb = lib[:, 1], lib[:, 2]
    mu = mean_radiance(wl)
    F = _factors(wl, mu, rng)                                        # (5, B)
    gain = rng.uniform(0.7, 1.3, size=samples)                       # per-column jitter
    z = rng.standard_normal((lines, samples, F.shape[0]))
    z[..., 0] = np.clip(z[..., 0], -3.0, 3.0)                        # keep radiance > 0
    cube = mu[None, :, None] + np.einsum("lsk,kb->lbs", z, F)
    cube *= gain[None, None, :]
    ppmm = plume_map(lines, samples, rng) if plume else np.zeros((lines, samples))
    if plume:
        cube *= np.exp(absorb[None, :, None] * (ppmm[:, None, :] / 1.0e5))
    cube += noise * rng.standard_normal(cube.shape)
    np.maximum(cube, 1.0e-4, out=cube)
    cube = np.ascontiguousarray(cube, dtype=np.float32)      # einsum hands back a strided view
    if bad_pixels:
        npx = lines * samples
        act0, act1 = 350, 422                                        # 0-based CH4 window
        idx = rng.choice(npx, size=max(1, int(0.005 * npx)), replace=False)
        cube[idx // samples, :, idx % samples] = NODATA
        idx = rng.choice(npx, size=max(1, int(0.001 * npx)), replace=False)
        cube[idx // samples, rng.integers(act0, act1, size=idx.size), idx % samples] = np.nan
        idx = rng.choice(npx, size=max(1, int(0.001 * npx)), replace=False)
        cube[idx // samples, rng.integers(act0, act1, size=idx.size), idx % samples] = -0.01
        idx = rng.choice(npx, size=max(1, int(0.0005 * npx)), replace=False)
        cube[idx // samples, 320:425, idx % samples] = 6.5
        idx = rng.choice(npx, size=max(1, int(0.0002 * npx)), replace=False)
        cube[idx // samples, rng.integers(act0, act1, size=idx.size), idx % samples] = np.inf
    if return_truth:
        return cube, ppmm
    return cube


def make_slab_torch(lines, samples, band_lo, band_hi, device, seed=2, lib=None, plume=True,
                    noise=0.004, dtype=None, chunk_lines=1024, out=None, bad_pixels=False):
    """Same recipe, only 1-based bands ``band_lo..band_hi``, generated on ``device`` with torch.

    Returns a float32 tensor ``(lines, D, samples)`` (the active slab of a BIL cube).
    Different RNG stream than :func:`make_cube` -- full-size runs are checked through
    size-independent properties, not against stored goldens.
    """
    import torch

    lib = load_ch4_library() if lib is None else lib
    wl, absorb = lib[:, 1], lib[:, 2]
    mu_full = mean_radiance(wl)
    rng = np.random.default_rng(seed)
    F_full = _factors(wl, mu_full, rng)
    sl = slice(band_lo - 1, band_hi)
    D = band_hi - band_lo + 1
    mu = torch.tensor(mu_full[sl], dtype=torch.float32, device=device)
    F = torch.tensor(F_full[:, sl], dtype=torch.float32, device=device)
    ab = torch.tensor(absorb[sl], dtype=torch.float32, device=device)
    gain = torch.tensor(rng.uniform(0.7, 1.3, size=samples), dtype=torch.float32, device=device)
    ppmm = torch.tensor(plume_map(lines, samples, rng) if plume else np.zeros((lines, samples)),
                        dtype=torch.float32, device=device)
    gen = torch.Generator(device=device)
    gen.manual_seed(int(seed))
    if out is None:
        out = torch.empty((lines, D, samples), dtype=torch.float32, device=device)
    for l0 in range(0, lines, chunk_lines):
        l1 = min(lines, l0 + chunk_lines)
        z = torch.randn((l1 - l0, samples, F.shape[0]), generator=gen, device=device)
        z[..., 0].clamp_(-3.0, 3.0)
        blk = mu[None, :, None] + torch.einsum("lsk,kb->lbs", z, F)
        blk *= gain[None, None, :]
        if plume:
            blk *= torch.exp(ab[None, :, None] * (ppmm[l0:l1, None, :] / 1.0e5))
        blk += noise * torch.randn(blk.shape, generator=gen, device=device)
        blk.clamp_(min=1.0e-4)
        out[l0:l1] = blk
    if bad_pixels:
        inject_bad_pixels_torch(out, seed)
    return out


def inject_bad_pixels_torch(slab, seed):
    """C3 bad pixels on an active slab ``(L, D, S)`` in place, same rates as :func:`make_cube`: 0.5 % whole
    pixels = NODATA, 0.1 % single-band NaN, 0.1 % single-band negative, 0.05 % saturated (6.5 in every band),
    0.02 % single-band +inf.  Returns the (L, S) bool image of pixels the filter must drop."""
    import torch

    L, D, S = slab.shape
    rng = np.random.default_rng(1000 + int(seed))
    npx = L * S
    dev = slab.device

    def pick(frac):
        idx = rng.choice(npx, size=max(1, int(frac * npx)), replace=False)
        return (torch.as_tensor(idx // S, device=dev), torch.as_tensor(idx % S, device=dev),
                torch.as_tensor(rng.integers(0, D, size=idx.size), device=dev))

    bad = torch.zeros((L, S), dtype=torch.bool, device=dev)
    l, s, _ = pick(0.005)
    slab[l, :, s] = NODATA
    bad[l, s] = True
    for frac, val in ((0.001, float("nan")), (0.001, -0.01), (0.0002, float("inf"))):
        l, s, b = pick(frac)
        slab[l, b, s] = val
        bad[l, s] = True
    l, s, _ = pick(0.0005)
    cur = slab[l, :, s]
    # saturate only healthy values: a defect planted above must survive (the pixel stays dropped)
    slab[l, :, s] = torch.where(torch.isfinite(cur) & ~(cur < 0), torch.full_like(cur, 6.5), cur)
    return bad
