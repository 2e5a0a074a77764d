"""TEST INFRASTRUCTURE ONLY (never imported by srcfinder_b200/): CPU restatement of the steps either side of
the matched filter that SURVEY.md 8(f) ranks next.

Pinning: the reference modules cannot be imported (argv parsing / file I/O at import time, GDAL, rasterio), so
``oracle/make_product_golden.py`` cuts the functions named below out of the reference sources with ``ast`` and
executes their UNMODIFIED text on seeded inputs; ``tests/golden/{flags,profile,filtdet,cnnnorm}_*.npz`` hold those
outputs and ``tests/test_products_oracle.py`` checks every restatement here against them.

* ``pixel_flags``  -- spectrometer_masks/masks_sds.py: get_saturation_mask (:133-150), get_spec_mask (:152-163),
  get_dark_mask (:165-180), get_cloud_mask (:182-230), same numpy expressions on a (lines, samples, bands) view.
* ``column_profile`` -- triage/cmf_profile.py:110-133, the same numpy calls (float32 image, nanmean / nanstd /
  nanmin / nanmax or nanmedian / nearest-rank nanpercentile); ``extrema`` is srcfinder_util.py:647-658.
* ``detection_prefilter`` -- srcfinder_util.py: kde (:1383-1387) and the head of filtdet (:1428-1436).
* ``cnn_input`` -- cnn/cnn_pred_pipeline.py: ClampCH4 (:19-30) + transforms.Normalize (:126-157), float32.
"""
import warnings

import numpy as np

SATURATED, SPECULAR, DARK, CLOUD = 1, 2, 4, 8


def pixel_flags(cube_lbs, wave, threshold=6.0, waverange=(1945, 2485), dark_threshold=0.104, cldthreshold=(15.0,),
                bandrange=(15, 60, 175), visible_mask_growing_threshold=9.0):
    data = np.transpose(cube_lbs, (0, 2, 1))                 # the reference reads BIP blocks (l, s, b)
    wave = np.asarray(wave)
    # get_saturation_mask (:149)
    is_saturated = (data[..., np.logical_and(wave >= waverange[0], wave <= waverange[1])] > threshold).any(axis=-1)
    # get_spec_mask (:159-162)
    test2 = data[:, :, 25] > visible_mask_growing_threshold
    is_spec = np.logical_and(is_saturated == 1, test2 == 1)
    # get_dark_mask (:175-179)
    test = data[:, :, 352]
    is_dark = np.logical_and((test < dark_threshold) == 1, (test <= -9999) == 0)
    # get_cloud_mask (:196-228)
    rdn1, rdn2, rdn3 = data[:, :, bandrange[0]], data[:, :, bandrange[1]], data[:, :, bandrange[2]]
    is_bright = rdn1 > cldthreshold[0]
    wide, tall = rdn1.shape
    x_rdn_a = np.zeros((wide, tall, 2), dtype=np.float32)
    x_rdn_b = np.zeros((wide, tall, 2), dtype=np.float32)
    x_rdn_a[:, :, 0], x_rdn_a[:, :, 1] = rdn1, rdn2
    x_rdn_b[:, :, 0], x_rdn_b[:, :, 1] = rdn2, rdn3
    x_diff_a, x_diff_b = np.diff(x_rdn_a), np.diff(x_rdn_b)
    y_diff_a = wave[bandrange[0]] - wave[bandrange[1]]
    y_diff_b = wave[bandrange[1]] - wave[bandrange[2]]
    y_arr_a = np.ones((wide, tall, 1), dtype=np.float32) * y_diff_a * -1
    y_arr_b = np.ones((wide, tall, 1), dtype=np.float32) * y_diff_b * -1
    slope_a_bool = (x_diff_a / y_arr_a < 0)[:, :, 0]
    slope_b_bool = (x_diff_b / y_arr_b < 0)[:, :, 0]
    # the third positional argument of np.logical_and is `out`: slope_b does not enter the result (:228)
    is_cloud = np.logical_and(is_bright == 1, slope_a_bool == 1, slope_b_bool == 1)
    return (is_saturated * SATURATED + is_spec * SPECULAR + is_dark * DARK + is_cloud * CLOUD).astype(np.uint8)


def extrema(a, p=1.0, axis=None):
    """srcfinder_util.py:647-658 (``interpolation='nearest'`` is spelled ``method=`` in numpy >= 1.22)."""
    if p == 1.0:
        return np.nanmin(a, axis=axis), np.nanmax(a, axis=axis)
    assert 0.0 < p < 1.0
    return (np.nanpercentile(a, axis=axis, q=(1 - p) * 100, method="nearest"),
            np.nanpercentile(a, axis=axis, q=p * 100, method="nearest"))


def column_profile(mf_ls, nodata=-9999, use_robust_stats=False):
    """triage/cmf_profile.py:110-130 on the score band ``mf_ls`` (lines, samples) of a product."""
    nodatav = np.float32(nodata)
    cmf = np.float32(np.array(mf_ls, copy=True))
    cmfnodata = (cmf == nodatav) | np.isnan(cmf)
    cmfmask = ~cmfnodata & (cmf > 0)
    cmf[~cmfmask] = np.nan
    colnum = np.count_nonzero(cmfmask, axis=0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if use_robust_stats:
            colavg = np.nanmedian(cmf, axis=0)
            colstd = np.nanmedian(np.abs(cmf - colavg), axis=0)
            colmin, colmax = extrema(cmf, p=0.95, axis=0)
            names = ("npix", "med", "mad", "p05", "p95")
        else:
            colavg = np.nanmean(cmf, axis=0)
            colstd = np.nanstd(cmf, axis=0)
            colmin = np.nanmin(cmf, axis=0)
            colmax = np.nanmax(cmf, axis=0)
            names = ("npix", "avg", "std", "min", "max")
    return dict(zip(names, (colnum, colavg, colstd, colmin, colmax)))


def detection_prefilter(ch4mf, k=50, mfmin=500, mfmax=1500):
    """srcfinder_util.py:1383-1387 (kde) and :1428-1436 (filtdet head): (detkde, ch4min, detmask)."""
    from scipy.ndimage import gaussian_filter
    detkde = ch4mf.copy()
    ch4min = ch4mf >= mfmin
    imgkde = gaussian_filter(detkde, sigma=k, truncate=1)
    imgkde = (imgkde - imgkde.min()) / (imgkde.max() - imgkde.min())
    detkde = detkde * imgkde
    detkde = np.clip((detkde - mfmin) / (mfmax - mfmin), 0.0, 1.0)
    return detkde, ch4min, detkde > 0


def cnn_input(ch4mf, vmin=0, vmax=4000, mean=110.6390, std=183.9152):
    """cnn/cnn_pred_pipeline.py:19-30, 126-157 in float32 (torch.clamp, then sub / div as torchvision does)."""
    x = np.float32(ch4mf)
    c = np.where(np.isnan(x), x, np.minimum(np.maximum(x, np.float32(vmin)), np.float32(vmax)))
    return (c - np.float32(mean)) / np.float32(std)
