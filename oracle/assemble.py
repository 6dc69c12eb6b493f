"""Assemble_Dice restated on the CPU.  Test infrastructure — see oracle/__init__.py.

Follows util/assemble_dice.py:130-213 of the reference: border crop, sequential overlap-add of cube/8 with a
count mask, (sum/mask)*8, optional percentile stretch, *65535, truncating cast, un-pad crop.
"""
from __future__ import annotations

import numpy as np

from .geometry import DiceGeometry


def crop_border(cube: np.ndarray, border: int) -> np.ndarray:
    """assemble_dice.py:137-145 (squeeze + [bc:-bc]^3; bc >= 1 is mandatory in the reference)."""
    c = np.asarray(cube, dtype=np.float32).squeeze()
    return c[border:-border, border:-border, border:-border]


def blend_sequential(cubes, geo: DiceGeometry):
    """assemble_dice.py:161-184 literally: returns (visual_ret float32 padded, mask_ret float32)."""
    r = geo.roi
    vis = np.zeros(geo.padded, dtype=np.float32)
    mask = np.zeros(geo.padded, dtype=np.float32)
    ones = np.ones((r, r, r), dtype=np.float32)
    for index, cube in enumerate(cubes):
        z, y, x = geo.origin(index)
        vis[z:z + r, y:y + r, x:x + r] += cube / 8
        mask[z:z + r, y:y + r, x:x + r] += ones
    vis = (vis / mask) * 8
    return vis, mask


def analytic_count(geo: DiceGeometry) -> np.ndarray:
    """Separable overlap count n(p) in {1,2,4,8}; equals mask_ret."""
    per_axis = []
    for p, k in zip(geo.padded, geo.steps):
        q = np.arange(p)
        m = np.ones(p, dtype=np.float32)
        for j in range(1, k):
            m[(q >= j * geo.step) & (q < j * geo.step + geo.overlap)] = 2
        per_axis.append(m)
    return per_axis[0][:, None, None] * per_axis[1][None, :, None] * per_axis[2][None, None, :]


def rescale_intensity(image: np.ndarray, in_range) -> np.ndarray:
    """skimage.exposure.rescale_intensity(image, in_range=(imin, imax)) for a float32 image.

    scikit-image 0.18.3 is the reference's pin (conda_environment/neuroclear_env.yml:203); it is neither installed
    nor vendored, so this restates its published algorithm (exposure.py, rescale_intensity):
        imin, imax = map(float, in_range); out range for a float image with imin >= 0 is (0, 1)
        image = np.clip(image, imin, imax)
        image = (image - imin) / (imax - imin)      # python-float scalars: float32 array arithmetic
        return np.asarray(image * (omax - omin) + omin, dtype=float32)
    PARITY UNPINNED for this function only (no skimage to run against); the numpy arithmetic it is made of is
    exercised as written.
    """
    imin, imax = float(in_range[0]), float(in_range[1])
    omin, omax = (0.0, 1.0) if imin >= 0 else (-1.0, 1.0)
    image = np.clip(image, np.float32(imin), np.float32(imax))
    if imin != imax:
        image = (image - np.float32(imin)) / np.float32(imax - imin)
        return np.asarray(image * np.float32(omax - omin) + np.float32(omin), dtype=np.float32)
    return np.clip(image, omin, omax).astype(np.float32)


def finish(vis: np.ndarray, geo: DiceGeometry, normalize_intensity: bool, sat_level=(0.25, 99.75),
           imtype: str = "uint16"):
    """assemble_dice.py:190-213.  Returns (volume of original size, (p1, p99) or None)."""
    pcts = None
    if normalize_intensity:
        p1, p99 = np.percentile(vis, sat_level)
        pcts = (float(p1), float(p99))
        vis = rescale_intensity(vis, in_range=(p1, p99))
    else:
        vis = vis.copy()
    if imtype == "uint8":
        vis *= 255
        vis = vis.astype(np.uint8)
    else:
        vis *= 2 ** 16 - 1
        vis = vis.astype(np.uint16)
    pads = [p - n for p, n in zip(geo.padded, geo.size)]
    return vis[:-pads[0], :-pads[1], :-pads[2]], pcts


def assemble(net_outputs, geo: DiceGeometry, normalize_intensity: bool, sat_level=(0.25, 99.75)):
    """Full Assemble_Dice: list of (1,1,E,E,E) or (E,E,E) network outputs in cube order -> uint16 (Z,Y,X)."""
    cubes = [crop_border(c, geo.border) for c in net_outputs]
    vis, _ = blend_sequential(cubes, geo)
    return finish(vis, geo, normalize_intensity, sat_level)
