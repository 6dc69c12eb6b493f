"""Dice extraction restated on the CPU.  Test infrastructure — see oracle/__init__.py.

``dice_cube_direct`` follows the reference literally (materialised zero pad + reflect pad + slice,
util/util.py:212, data/diceImage_dataset.py:95-96,108-120, data/base_dataset.py:134-143,291-301);
``dice_cube_gather`` is the closed form the CUDA kernel implements.  Both must agree bit for bit.
"""
from __future__ import annotations

import numpy as np

from .geometry import DiceGeometry


class DirectDicer:
    def __init__(self, volume: np.ndarray, geo: DiceGeometry):
        assert volume.shape == geo.size
        pad = [(0, p - n) for p, n in zip(geo.padded, geo.size)]
        padded = np.pad(volume, pad_width=pad)                                # pad_for_dicing: zeros, far end
        bc = geo.border
        self.image = np.pad(padded, ((bc, bc),) * 3, mode="reflect")          # DiceCube.__init__
        self.geo = geo

    def cube_raw(self, index: int) -> np.ndarray:
        g = self.geo
        z, y, x = g.origin(index)
        e = g.edge
        return self.image[z:z + e, y:y + e, x:x + e]

    def cube(self, index: int) -> np.ndarray:
        """-> float32 (1, E, E, E) exactly as DiceImageDataSet.__getitem__()['A']."""
        raw = self.cube_raw(index)
        denom = 2 ** 16 * 1.0 - 1 if raw.dtype == np.uint16 else 2 ** 8 * 1.0 - 1
        normd = (raw / denom).astype(float)                                   # __normalize (float64)
        return np.expand_dims(normd, 0).astype(np.float32)                    # __addColorChannel, __toTensor


def _reflect(j: np.ndarray, n: int) -> np.ndarray:
    j = np.where(j < 0, -j, j)
    return np.where(j >= n, 2 * (n - 1) - j, j)


def dice_cube_gather(volume: np.ndarray, geo: DiceGeometry, index: int) -> np.ndarray:
    """Closed form: reflect index into the zero-padded volume, fp32 divide by 65535."""
    oz, oy, ox = geo.origin(index)
    e, bc = geo.edge, geo.border
    idx = []
    for o, p in zip((oz, oy, ox), geo.padded):
        idx.append(_reflect(np.arange(o - bc, o - bc + e), p))
    iz, iy, ix = idx
    Z, Y, X = geo.size
    out = np.zeros((e, e, e), dtype=np.float32)
    vz, vy, vx = iz < Z, iy < Y, ix < X
    sub = volume[np.ix_(iz[vz], iy[vy], ix[vx])].astype(np.float32)
    out[np.ix_(vz, vy, vx)] = sub
    denom = np.float32(65535.0) if volume.dtype == np.uint16 else np.float32(255.0)
    return (out / denom)[None]
