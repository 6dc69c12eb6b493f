"""CPU restatement of the remaining Assemble_Dice / test_dice.py options.
Test infrastructure — see oracle/__init__.py.

* ``match_histograms`` — skimage.exposure.match_histograms(image, reference) for single-channel arrays
  (util/assemble_dice.py:150-151).  scikit-image 0.18.3 is the reference's pin and is neither installed nor vendored:
  this restates its published algorithm (skimage/exposure/histogram_matching.py::_match_cumulative_cdf) with the numpy
  calls it is made of run as written.  PARITY UNPINNED for this function (no skimage to run against) — like
  assemble.rescale_intensity.
* ``save_projections`` — test_dice.py:159-177 (np.amax with the script's hard-coded windows).
* ``normalize`` / ``standardize`` / ``get_psnr`` / ``psnr_report`` — util/util.py:56-71,107-115 and test_dice.py:
  229-253, restated; PINNED against the reference's own util.util functions in oracle/make_golden.py::golden_report.
"""
from __future__ import annotations

import math

import numpy as np


def match_histograms(image: np.ndarray, reference: np.ndarray) -> np.ndarray:
    src_values, src_unique_indices, src_counts = np.unique(image.ravel(), return_inverse=True, return_counts=True)
    tmpl_values, tmpl_counts = np.unique(reference.ravel(), return_counts=True)
    src_quantiles = np.cumsum(src_counts) / image.size
    tmpl_quantiles = np.cumsum(tmpl_counts) / reference.size
    interp_a_values = np.interp(src_quantiles, tmpl_quantiles, tmpl_values)
    return interp_a_values[src_unique_indices].reshape(image.shape)


def save_projections(fake_volume, real_volume=None):
    out = {"fake_xy": np.amax(fake_volume, axis=0), "fake_xz": np.amax(fake_volume[:, 800:1100, :], axis=1),
           "fake_yz": np.amax(fake_volume[:, :, 200:500], axis=2)}
    if real_volume is not None:
        out.update({"real_xy": np.amax(real_volume, axis=0), "real_xz": np.amax(real_volume, axis=1),
                    "real_yz": np.amax(real_volume, axis=2)})
    return out


def normalize(img_np, data_type=float):
    img_min, img_max = np.min(img_np), np.max(img_np)
    new_min = 0
    new_max = {np.uint8: 2 ** 8 - 1, np.uint16: 2 ** 16 - 1}.get(data_type, 1)
    return ((img_np - img_min) * ((new_max - new_min) / (img_max - img_min)) + new_min).astype(data_type)


def standardize(img_np):
    return (img_np - np.mean(img_np)) / np.std(img_np)


def get_psnr(source, target, data_range):
    target, source = target.astype(float), source.astype(float)
    mse = np.mean((target - source) ** 2)
    return 20 * math.log(data_range, 10) - 10 * math.log(mse, 10)


def psnr_report(real_volume, fake_volume, gt_volume):
    """test_dice.py:239-253: returns (psnr_input_gt, psnr_output_gt, real8, fake8, gt8)"""
    vols = [real_volume, fake_volume, gt_volume]
    for _ in range(2):
        vols = [normalize(standardize(v), data_type=np.uint8) for v in vols]
    return get_psnr(vols[0], vols[2], 2 ** 8 - 1), get_psnr(vols[1], vols[2], 2 ** 8 - 1), vols[0], vols[1], vols[2]
