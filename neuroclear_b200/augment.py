"""Drop-in for the reference's training dataset (data/singlevolume_dataset.py + the transform pipeline of
data/base_dataset.py:87-131) for the README's --preprocess
random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel.

The reference keeps the uint16 volume on the host and, for EVERY item, rotates every z-slice of the whole volume with
cv2.warpAffine, crops the largest axis-aligned rectangle, takes a random crop, normalises and flips — 1.67 s per
iteration in its own screenshot, 60x the 27 ms the B200 training iteration takes.  Here the volume lives in HBM and
nc_augment_crop_u16 evaluates the same (bit-exact) rotation only at the voxels of the crop.  The host side below
reproduces the reference's geometry (rotate_image's canvas, largest_rotated_rect, crop_around_center) and its
random draws in the same order (`random.randint` for the angle and the crop position, `random.shuffle` +
`np.random.uniform` for the flips), so a seeded run yields the same crops as the reference's dataset.
"""
from __future__ import annotations

import math
import random

import numpy as np
import torch

from ._lib import NeuroclearError, call, ptr, stream_ptr

AB_SCALE, ROUND_DELTA = 1024, 16           # cv2: AB_BITS 10, INTER_BITS 5 -> round_delta = 1024 / 32 / 2


def _rotation_matrix_2d(center, angle_deg):
    """cv2.getRotationMatrix2D(center, angle, 1.0)"""
    a = math.radians(angle_deg)
    alpha, beta = math.cos(a), math.sin(a)
    return np.array([[alpha, beta, (1 - alpha) * center[0] - beta * center[1]],
                     [-beta, alpha, beta * center[0] + (1 - alpha) * center[1]]], dtype=np.float64)


def _rotate_plan(width, height, angle_deg):
    """rotate_image (base_dataset.py:306-372): forward 2x3 matrix and the enlarged canvas size"""
    rot = np.vstack([_rotation_matrix_2d((width / 2, height / 2), angle_deg), [0, 0, 1]])
    r2 = rot[0:2, 0:2]
    w2, h2 = width * 0.5, height * 0.5
    corners = [np.array(c) @ r2 for c in ([-w2, h2], [w2, h2], [-w2, -h2], [w2, -h2])]
    xs, ys = [c[0] for c in corners], [c[1] for c in corners]
    new_w = int(abs(max(x for x in xs if x > 0) - min(x for x in xs if x < 0)))
    new_h = int(abs(max(y for y in ys if y > 0) - min(y for y in ys if y < 0)))
    trans = np.array([[1, 0, int(new_w * 0.5 - w2)], [0, 1, int(new_h * 0.5 - h2)], [0, 0, 1]], dtype=np.float64)
    return (trans @ rot)[0:2, :], new_w, new_h


def _largest_rotated_rect(w, h, angle):
    """base_dataset.py:375-408, as written there"""
    quadrant = int(math.floor(angle / (math.pi / 2))) & 3
    sign_alpha = angle if ((quadrant & 1) == 0) else math.pi - angle
    alpha = (sign_alpha % math.pi + math.pi) % math.pi
    bb_w = w * math.cos(alpha) + h * math.sin(alpha)
    bb_h = w * math.sin(alpha) + h * math.cos(alpha)
    gamma = math.atan2(bb_w, bb_w)
    delta = math.pi - alpha - gamma
    d = (h if (w < h) else w) * math.cos(alpha)
    a = d * math.sin(alpha) / math.sin(delta)
    y = a * math.cos(gamma)
    x = y * math.tan(gamma)
    return bb_w - 2 * x, bb_h - 2 * y


def rotate_clean_window(height, width, angle_deg):
    """__rotate_clean (base_dataset.py:433-443): (forward matrix, x1, x2, y1, y2) of the kept window of the canvas"""
    m, new_w, new_h = _rotate_plan(width, height, angle_deg)
    rw, rh = _largest_rotated_rect(width, height, math.radians(angle_deg))
    cx, cy = int(new_w * 0.5), int(new_h * 0.5)
    rw, rh = min(rw, new_w), min(rh, new_h)
    x1, x2, y1, y2 = int(cx - rw * 0.5), int(cx + rw * 0.5), int(cy - rh * 0.5), int(cy + rh * 0.5)
    return m, max(x1, 0), min(x2, new_w), max(y1, 0), min(y2, new_h)


def _inverse_map_tables(fwd, xs, ys):
    """cv2.warpAffine's fixed-point inverse map for destination columns xs / rows ys (int32 arrays)"""
    m = np.array(fwd, dtype=np.float64)
    d = m[0, 0] * m[1, 1] - m[0, 1] * m[1, 0]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[1, 1] * d, m[0, 0] * d
    m[0, 0], m[0, 1], m[1, 0], m[1, 1] = a11, m[0, 1] * -d, m[1, 0] * -d, a22
    b1 = -m[0, 0] * m[0, 2] - m[0, 1] * m[1, 2]
    b2 = -m[1, 0] * m[0, 2] - m[1, 1] * m[1, 2]
    m[0, 2], m[1, 2] = b1, b2
    rint = lambda v: np.rint(v).astype(np.int64)
    adelta, bdelta = rint(m[0, 0] * xs * AB_SCALE), rint(m[1, 0] * xs * AB_SCALE)
    x0 = rint((m[0, 1] * ys + m[0, 2]) * AB_SCALE) + ROUND_DELTA
    y0 = rint((m[1, 1] * ys + m[1, 2]) * AB_SCALE) + ROUND_DELTA
    return [t.astype(np.int32) for t in (x0, y0, adelta, bdelta)]


class RotatedCropSampler:
    """The uint16 volume in HBM + nc_augment_crop_u16."""

    def __init__(self, volume, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NeuroclearError("RotatedCropSampler needs a CUDA device (no CPU fallback)")
        if isinstance(volume, np.ndarray):
            volume = torch.from_numpy(volume)
        if volume.dtype != torch.uint16 or volume.dim() != 3:
            raise NeuroclearError("RotatedCropSampler expects a (Z,Y,X) uint16 volume")
        self.vol = volume.to(self.device).contiguous()

    def rotated_shape(self, angle_deg):
        _, x1, x2, y1, y2 = rotate_clean_window(self.vol.shape[1], self.vol.shape[2], angle_deg)
        return self.vol.shape[0], max(y2 - y1, 0), max(x2 - x1, 0)

    def crop(self, angle_deg, crop_pos, crop_size, flip_axes=()):
        """float32 CUDA (1, 1, cz, cy, cx): get_transform(opt, params)(volume) of the reference for this angle, crop
        position (in the rotated + rectangle-cropped volume) and flipped axes (0 = z, 1 = y, 2 = x)."""
        z, y, x = (int(v) for v in crop_pos)
        cz, cy, cx = (int(v) for v in crop_size)
        Z, H, W = self.vol.shape
        m, x1, x2, y1, y2 = rotate_clean_window(H, W, angle_deg)
        if z < 0 or z + cz > Z or y < 0 or y1 + y + cy > y2 or x < 0 or x1 + x + cx > x2:
            raise NeuroclearError("crop %s at %s does not fit the rotated volume %s" %
                                  ((cz, cy, cx), (z, y, x), (Z, y2 - y1, x2 - x1)))
        xs = np.arange(x1 + x, x1 + x + cx, dtype=np.int64)
        ys = np.arange(y1 + y, y1 + y + cy, dtype=np.int64)
        tabs = _inverse_map_tables(m, xs, ys)
        mask = sum(1 << int(a) for a in set(flip_axes))
        with torch.cuda.device(self.device):
            dev_tabs = [torch.from_numpy(t).to(self.device) for t in tabs]
            out = torch.empty((1, 1, cz, cy, cx), dtype=torch.float32, device=self.device)
            call("nc_augment_crop_u16", ptr(self.vol), Z, H, W, z, cz, cy, cx, ptr(dev_tabs[0]), ptr(dev_tabs[1]),
                 ptr(dev_tabs[2]), ptr(dev_tabs[3]), mask, ptr(out), stream_ptr())
        return out


class SingleVolumeDataset:
    """reference data/singlevolume_dataset.py: ten items per epoch, every item a fresh random rotation / crop / flip
    of the one training volume.  `volume` replaces the TIFF read (skimage.io.imread, out of scope)."""

    def __init__(self, opt, volume, device=None):
        pre = getattr(opt, "preprocess", "")
        for needed in ("random3Drotate", "randomcrop", "randomflip"):
            if needed not in pre:
                raise NotImplementedError("the B200 data path implements --preprocess random3Drotate_randomcrop_"
                                          "randomflip_addColorChannel_addBatchChannel (README training command)")
        gpu_ids = list(getattr(opt, "gpu_ids", [0])) or [0]
        self.sampler = RotatedCropSampler(volume, device if device is not None else "cuda:%d" % gpu_ids[0])
        self.crop_size = tuple(int(c) for c in opt.crop_size)
        self.A_path = getattr(opt, "dataroot", "volume")

    def __len__(self):
        return 10

    def __getitem__(self, index):
        # the reference's draws, in its order: __randomrotate_clean_3D_xy, __randomcrop, __randomflip
        angle = random.randint(0, 359)
        Z, Hr, Wr = self.sampler.rotated_shape(angle)
        cz, cy, cx = self.crop_size
        if Z < cz or Hr < cy or Wr < cx:
            raise NeuroclearError("crop %s larger than the rotated volume %s" % (self.crop_size, (Z, Hr, Wr)))
        z, y, x = random.randint(0, Z - cz), random.randint(0, Hr - cy), random.randint(0, Wr - cx)
        axis_list = [0, 1, 2]
        random.shuffle(axis_list)
        flips = []
        for _ in range(3):
            if np.random.uniform(0, 1) < 0.5:
                flips.append(axis_list.pop())
        return {"A": self.sampler.crop(angle, (z, y, x), self.crop_size, flips), "A_paths": self.A_path}
