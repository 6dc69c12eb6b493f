"""The reporting tail of the reference's test_dice.py on the device: ``--save_projections`` (test_dice.py:159-177) and
the whole-volume PSNR report against a ground-truth volume (``--dataroot_gt``, test_dice.py:229-270 with
util/util.py:56-71 ``normalize``, :107-108 ``standardize``, :110-115 ``get_psnr``).

The assembled volumes are 1.5-9 GB: np.amax over an axis, np.mean / np.std and the per-voxel standardise -> rescale ->
uint8 passes are memory-bound sweeps that the reference runs single-threaded in numpy (float64 temporaries of 8x the
volume).  Here each is one pass over the uint16 / uint8 volume in HBM (nc_amax_axis, nc_volume_moments,
nc_standardize_normalize_u8, nc_sqdiff_u8).  No CPU fallback.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch

from ._lib import NeuroclearError, call, i64, ptr, stream_ptr


def _device_volume(vol, device):
    if isinstance(vol, np.ndarray):
        vol = torch.from_numpy(np.ascontiguousarray(vol))
    if vol.dtype not in (torch.uint16, torch.uint8) or vol.dim() != 3:
        raise NeuroclearError("expected a (Z,Y,X) uint16 / uint8 volume")
    device = torch.device(device)
    if device.type != "cuda":
        raise NeuroclearError("the report path runs on a CUDA device (no CPU fallback)")
    return vol.to(device).contiguous()


def max_projection(vol, axis: int, lo=None, hi=None, device="cuda") -> np.ndarray:
    """np.amax(vol[lo:hi along axis], axis) — returns a host array of the volume's dtype."""
    v = _device_volume(vol, device)
    z, y, x = v.shape
    ext = v.shape[axis]
    a0, a1 = slice(lo, hi).indices(ext)[:2]          # numpy slice semantics (negative / open bounds)
    shape = tuple(s for i, s in enumerate(v.shape) if i != axis)
    with torch.cuda.device(v.device):
        out = torch.empty(shape, dtype=v.dtype, device=v.device)
        call("nc_amax_axis", ptr(v), v.element_size(), z, y, x, axis, a0, a1, ptr(out), stream_ptr())
        return out.cpu().numpy()


def save_projections(fake_volume, real_volume=None, device="cuda"):
    """test_dice.py:159-177: the three maximum-intensity projections written by --save_projections, with the
    reference's hard-coded windows for the output volume ([:, 800:1100, :] for xz, [:, :, 200:500] for yz) and full
    range for the input volume.  Returns {'fake_xy', 'fake_xz', 'fake_yz'[, 'real_xy', 'real_xz', 'real_yz']}."""
    f = _device_volume(fake_volume, device)
    out = {"fake_xy": max_projection(f, 0), "fake_xz": max_projection(f, 1, 800, 1100),
           "fake_yz": max_projection(f, 2, 200, 500)}
    if real_volume is not None:
        r = _device_volume(real_volume, device)
        out.update({"real_xy": max_projection(r, 0), "real_xz": max_projection(r, 1), "real_yz": max_projection(r, 2)})
    return out


def _moments(v):
    """exact (n, sum, sum of squares, min, max) of a device uint8 / uint16 volume as Python ints"""
    out = torch.empty(4, dtype=torch.int64, device=v.device)
    call("nc_volume_moments", ptr(v), v.element_size(), i64(v.numel()), ptr(out), stream_ptr())
    s, q, mn, mx = [int(t) & 0xFFFFFFFFFFFFFFFF for t in out.cpu().tolist()]
    return v.numel(), s, q, mn, mx


def standardize_normalize_u8(v: torch.Tensor) -> torch.Tensor:
    """util.normalize(util.standardize(v), data_type=np.uint8) on the device, bit for bit: the mean is exact (integer
    sum / n, as numpy's: sums of integers are exact in float64), the standard deviation reproduces np.std's float64
    pairwise summation order (nc_pairwise_sqdev_sum) — its last bit decides the second normalisation pass of
    test_dice.py:243-249, where every voxel sits on an integer boundary of the truncating cast."""
    from . import _lib
    with torch.cuda.device(v.device):
        n, s, q, mn, mx = _moments(v)
        mean = s / n                                            # correctly rounded; == np.mean(v)
        if n * q - s * s <= 0:
            raise NeuroclearError("standardize: constant volume (the reference divides by zero)")
        scratch = torch.empty(_lib.load().nc_pairwise_sqdev_scratch_doubles(n), dtype=torch.float64, device=v.device)
        acc = torch.empty(1, dtype=torch.float64, device=v.device)
        call("nc_pairwise_sqdev_sum", ptr(v), v.element_size(), i64(n), mean, ptr(scratch), ptr(acc), stream_ptr())
        std = math.sqrt(float(acc.item()) / n)                  # == np.std(v)
        smin, smax = (mn - mean) / std, (mx - mean) / std       # standardisation is monotone
        scale = (255 - 0) / (smax - smin)
        params = torch.tensor([mean, std, smin, scale], dtype=torch.float64).to(v.device)
        out = torch.empty(v.shape, dtype=torch.uint8, device=v.device)
        call("nc_standardize_normalize_u8", ptr(v), v.element_size(), i64(n), ptr(params), ptr(out), stream_ptr())
    return out


def get_psnr(source: torch.Tensor, target: torch.Tensor, data_range) -> float:
    """util.get_psnr (util/util.py:110-115) for two uint8 device volumes; the mean squared error is exact."""
    if source.shape != target.shape or source.dtype != torch.uint8 or target.dtype != torch.uint8:
        raise NeuroclearError("get_psnr: two uint8 volumes of the same shape are expected")
    with torch.cuda.device(source.device):
        acc = torch.empty(1, dtype=torch.int64, device=source.device)
        call("nc_sqdiff_u8", ptr(source.contiguous()), ptr(target.contiguous()), i64(source.numel()), ptr(acc),
             stream_ptr())
        mse = (int(acc.item()) & 0xFFFFFFFFFFFFFFFF) / source.numel()
    return 20 * math.log(data_range, 10) - 10 * math.log(mse, 10)


def psnr_report(real_volume, fake_volume, gt_volume, device="cuda", name="", web_dir=None):
    """test_dice.py:229-270: each volume is standardised and normalised to uint8 TWICE (as written there), then
    PSNR(input, gt) and PSNR(output, gt) with data range 255.  Returns (psnr_input_gt, psnr_output_gt, message) and
    appends the message to <web_dir>/metrics.txt when web_dir is given."""
    vols = []
    for v in (real_volume, fake_volume, gt_volume):
        d = _device_volume(v, device)
        vols.append(standardize_normalize_u8(standardize_normalize_u8(d)))
    real8, fake8, gt8 = vols
    psnr_input_gt = get_psnr(real8, gt8, 2 ** 8 - 1)
    psnr_output_gt = get_psnr(fake8, gt8, 2 ** 8 - 1)
    bar = "---------------------------------------------------------"
    message = "Experiment Name: " + name + "\n" + bar + "\n" + "\nWhole_volume\n" + bar + "\n" + \
        "Network Input vs. Groundtruth\n" + "(psnr: %.4f) \n" % psnr_input_gt + bar + "\n" + \
        "Network Output vs. Groundtruth\n" + "(psnr: %.4f) \n" % psnr_output_gt + bar
    if web_dir is not None:
        os.makedirs(web_dir, exist_ok=True)
        with open(os.path.join(web_dir, "metrics.txt"), "a") as f:
            f.write("%s\n" % message)
    return psnr_input_gt, psnr_output_gt, message
