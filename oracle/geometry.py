"""Dicing geometry (integer math).  Test infrastructure — see oracle/__init__.py.

Follows util/util.py:196-215 (pad_for_dicing), data/diceImage_dataset.py:82-106 (DiceCube) and
util/assemble_dice.py:21-25,60-77 of the reference.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class DiceGeometry:
    size: tuple          # original (Z, Y, X)
    padded: tuple        # after pad_for_dicing
    steps: tuple         # cubes per axis (z, y, x)
    roi: int
    overlap: int
    border: int

    @property
    def step(self) -> int:
        return self.roi - self.overlap

    @property
    def edge(self) -> int:
        return self.roi + 2 * self.border

    @property
    def n_cubes(self) -> int:
        return self.steps[0] * self.steps[1] * self.steps[2]

    def index_to_cube(self, index: int):
        """diceImage_dataset.py:99-106 / assemble_dice.py:60-66: x fastest, then y, then z."""
        nz, ny, nx = self.steps
        x = index % nx
        y = (index % (nx * ny)) // nx
        z = index // (nx * ny)
        return z, y, x

    def origin(self, index: int):
        """assemble_dice.py:68-77: cube origin in padded-volume coordinates."""
        z, y, x = self.index_to_cube(index)
        return z * self.step, y * self.step, x * self.step


def dice_geometry(size, roi: int, overlap: int, border: int = 0) -> DiceGeometry:
    step = roi - overlap
    padded, steps = [], []
    for n in size:
        counts = (n + overlap) // step                 # util/util.py:203-205
        pad = step * counts + roi - n                  # util/util.py:207-209 (always >= 1)
        p = n + pad
        padded.append(p)
        steps.append((p - overlap) // step)            # diceImage_dataset.py:90-92
    return DiceGeometry(tuple(int(s) for s in size), tuple(padded), tuple(steps), roi, overlap, border)
