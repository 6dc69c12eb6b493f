"""Drop-in for ``Volume`` of the reference's apollo model (models/axial_to_lateral_gan_apollo_model.py:322-354):
random single-plane slices and randomised-depth max-intensity projections fed to the 2-D discriminators.
Same host ``np.random`` draws in the same order as the reference; the projection runs ``nc_mip_fwd`` and its
autograd backward ``nc_mip_bwd`` (gradient to the first arg-max plane, like ``torch.max(dim)[0]``)."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import NeuroclearError, call, ptr, stream_ptr


class _MipFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol, axis, start, depth):
        n, c, d, h, w = vol.shape
        v = vol.detach().contiguous()
        shape = (h, w) if axis == 0 else (d, w) if axis == 1 else (d, h)
        proj = torch.empty((n, c) + shape, dtype=torch.float32, device=vol.device)
        arg = torch.empty((n, c) + shape, dtype=torch.int32, device=vol.device)
        with torch.cuda.device(vol.device):
            for i in range(n * c):
                call("nc_mip_fwd", ptr(v.view(n * c, d, h, w)[i]), d, h, w, axis, start, depth,
                     ptr(proj.view(n * c, *shape)[i]), ptr(arg.view(n * c, *shape)[i]), stream_ptr())
        ctx.save_for_backward(arg)
        ctx.meta = (vol.shape, axis)
        return proj

    @staticmethod
    def backward(ctx, gout):
        (arg,) = ctx.saved_tensors
        (n, c, d, h, w), axis = ctx.meta
        g = gout.contiguous().float()
        gvol = torch.zeros((n, c, d, h, w), dtype=torch.float32, device=gout.device)
        shape = arg.shape[2:]
        with torch.cuda.device(gout.device):
            for i in range(n * c):
                call("nc_mip_bwd", ptr(g.view(n * c, *shape)[i]), ptr(arg.view(n * c, *shape)[i]), d, h, w, axis,
                     ptr(gvol.view(n * c, d, h, w)[i]), stream_ptr())
        return gvol, None, None, None


class Volume:
    def __init__(self, vol, device):
        self.volume = vol.to(device)
        if not self.volume.is_cuda:
            raise NeuroclearError("Volume (B200): the volume must live on a CUDA device (no CPU fallback)")
        self.num_slice = vol.shape[-1]

    def get_slice(self, slice_axis):
        i = np.random.randint(self.num_slice)                       # apollo_model.py:329
        if slice_axis == 0:
            return self.volume[:, :, i, :, :]
        if slice_axis == 1:
            return self.volume[:, :, :, i, :]
        return self.volume[:, :, :, :, i]

    def get_projection(self, depth, slice_axis):
        start = np.random.randint(0, self.num_slice - depth)        # apollo_model.py:340
        if self.volume.dtype != torch.float32:
            raise NeuroclearError("Volume.get_projection expects float32")
        # the reference slices [start : start + depth] along the axis: on a non-cubic crop (num_slice is the LAST
        # extent, :325) python slicing silently truncates at the end of the axis, and an empty slab makes torch.max fail
        size = self.volume.shape[slice_axis + 2]
        depth = min(int(depth), size - int(start))
        if depth <= 0:
            raise NeuroclearError("Volume.get_projection: empty slab (start %d on an axis of %d planes)" % (start, size))
        return _MipFn.apply(self.volume, slice_axis, int(start), depth)

    def get_volume(self):
        return self.volume
