"""Volume.get_projection / get_slice of the apollo model restated (numpy RNG order preserved).
Test infrastructure — see oracle/__init__.py.  Follows models/axial_to_lateral_gan_apollo_model.py:322-354.
"""
from __future__ import annotations

import numpy as np
import torch


def get_projection(vol: torch.Tensor, depth: int, axis: int, rng=np.random):
    """vol (N,1,D,H,W).  start = randint(0, shape[-1]-depth); max over `depth` planes along `axis`."""
    start = rng.randint(0, vol.shape[-1] - depth)
    if axis == 0:
        slab = vol[:, :, start:start + depth, :, :]
    elif axis == 1:
        slab = vol[:, :, :, start:start + depth, :]
    else:
        slab = vol[:, :, :, :, start:start + depth]
    return torch.max(slab, axis + 2)[0], start


def get_slice(vol: torch.Tensor, axis: int, rng=np.random):
    index = rng.randint(0, vol.shape[-1])
    if axis == 0:
        return vol[:, :, index, :, :], index
    if axis == 1:
        return vol[:, :, :, index, :], index
    return vol[:, :, :, :, index], index
