"""The reference's AxialToLateralGANAthenaModel (models/axial_to_lateral_gan_athena_model.py): the
apollo cycle with SIX 2-D discriminators (xy / xz / yz for the isotropic output and for the reconstruction), each
applied to EVERY slice of the cube along its axis (:286-296, ``iter_f``), no projections and no random draws.

The reference runs S discriminator passes per ``iter_f`` and writes their outputs into a zero volume; the mean of the
LSGAN loss over that volume is the mean over the S prediction maps.  Here the S slices are one batch (S, 1, H, W) of a
single nc_patchgan_fwd / _bwd call (InstanceNorm statistics are per image, so batching changes no value): 18 calls per
iteration instead of 18 S.  Same kernels, same protocol as the apollo mirror."""
from __future__ import annotations

import itertools
import os
from collections import OrderedDict

import torch

from . import discriminator, networks
from ._lib import NeuroclearError
from .apollo_d_path import FusedAdam, allreduce_mean_gradients
from .apollo_model import AxialToLateralGANApolloModel

LOSS_NAMES = ["D_A_xy", "D_A_xz", "D_A_yz", "G_A", "G_A_xy", "G_A_xz", "G_A_yz", "cycle_A", "D_B_xy", "D_B_xz",
              "D_B_yz", "G_B", "G_B_xy", "G_B_xz", "G_B_yz"]                                     # athena_model.py:89-90
D_NAMES = ["D_A_yz", "D_A_xy", "D_A_xz", "D_B_yz", "D_B_xy", "D_B_xz"]                           # creation order :133-149
PLANE_TO_SLICE_AXIS = {"xy": 0, "xz": 1, "yz": 2}                                                # :100


def slices_as_batch(volume, slice_axis):
    """(1, 1, D, H, W) -> every slice along `slice_axis` as a batch (S, 1, a, b) (Volume.get_slice(i, axis), :305-319)"""
    if volume.dim() != 5 or volume.shape[0] != 1 or volume.shape[1] != 1:
        raise NeuroclearError("athena iter_f expects a (1, 1, D, H, W) cube")
    v = volume[0, 0]
    if slice_axis == 1:
        v = v.permute(1, 0, 2)
    elif slice_axis == 2:
        v = v.permute(2, 0, 1)
    return v.contiguous()[:, None]


class AxialToLateralGANAthenaModel(AxialToLateralGANApolloModel):
    def __init__(self, opt, device=None, group=None, distributed=None):
        import torch.distributed as dist
        self.opt = opt
        gpu_ids = list(getattr(opt, "gpu_ids", [0])) or [0]
        self.device = torch.device(device if device is not None else "cuda:%d" % gpu_ids[0])
        self.group = group
        self.distributed = dist.is_initialized() if distributed is None else distributed
        ids = [self.device.index if self.device.index is not None else torch.cuda.current_device()]
        self.loss_names = list(LOSS_NAMES)
        plane = list(getattr(opt, "conversion_plane", ["yz", "xy"]))
        remain = [a for a in PLANE_TO_SLICE_AXIS if a != plane[0] and a != plane[1]][0]
        self.source_sl_axis, self.target_sl_axis = PLANE_TO_SLICE_AXIS[plane[0]], PLANE_TO_SLICE_AXIS[plane[1]]
        self.remain_sl_axis = PLANE_TO_SLICE_AXIS[remain]
        s = float(sum(opt.lambda_plane))
        self.lambda_plane_target, self.lambda_plane_source, self.lambda_plane_ref = [f / s for f in opt.lambda_plane]
        self.netG_A = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                        opt.init_type, opt.init_gain, ids, dimension=3)
        self.netG_B = networks.define_G(opt.output_nc, opt.input_nc, opt.ngf, opt.netG_B, opt.norm,
                                        not opt.no_dropout, opt.init_type, opt.init_gain, ids, dimension=3)
        for n in D_NAMES:
            nc = opt.output_nc if n.startswith("D_A") else opt.input_nc
            setattr(self, "net" + n, discriminator.define_D(nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm,
                                                            opt.init_type, opt.init_gain, False, ids, dimension=2))
        self.criterionGAN = discriminator.GANLoss(opt.gan_mode).to(self.device)
        self.criterionCycle = discriminator.L1Loss()
        self.optimizer_G = FusedAdam(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()),
                                     lr=opt.lr, betas=(opt.beta1, 0.999))
        order = ["D_A_yz", "D_A_xy", "D_A_xz", "D_B_yz", "D_B_xy", "D_B_xz"]                      # :158-159
        self.optimizer_D = FusedAdam(itertools.chain(*[getattr(self, "net" + n).parameters() for n in order]),
                                     lr=opt.lr, betas=(opt.beta1, 0.999))
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.model_names = ["G_A", "G_B", "D_A_xy", "D_A_xz", "D_A_yz", "D_B_xy", "D_B_xz", "D_B_yz"]
        self.visual_names = ["real", "fake", "rec"]
        self.schedulers = []
        self.metric = 0
        self.save_dir = os.path.join(getattr(opt, "checkpoints_dir", "./checkpoints"), getattr(opt, "name", "athena"))

    def _ds(self):
        return [getattr(self, "net" + n) for n in D_NAMES]

    def set_input(self, input):                                # :166-180 (no random draw)
        a_to_b = getattr(self.opt, "direction", "AtoB") == "AtoB"
        self.real = input["A" if a_to_b else "B"].to(self.device)
        self.image_paths = input["A_paths" if a_to_b else "B_paths"]
        self.cube_shape = self.real.shape
        self.num_slice = self.cube_shape[-3]

    def iter_f(self, input, function, slice_axis):             # :286-296
        return function(slices_as_batch(input, slice_axis))

    def backward_D_basic(self, netD, real, fake, slice_axis_real, slice_axis_fake):     # :189-214
        pred_real = self.iter_f(real, netD, slice_axis_real)
        pred_fake = self.iter_f(fake.detach(), netD, slice_axis_fake)
        loss_D = (self.criterionGAN(pred_real, True) + self.criterionGAN(pred_fake, False)) * 0.5
        loss_D.backward()
        return loss_D

    def backward_G(self):                                      # :237-256
        g, t, s, r = self.criterionGAN, self.target_sl_axis, self.source_sl_axis, self.remain_sl_axis
        self.loss_G_A_xy = g(self.iter_f(self.fake, self.netD_A_xy, t), True) * self.lambda_plane_target
        self.loss_G_A_yz = g(self.iter_f(self.fake, self.netD_A_yz, s), True) * self.lambda_plane_source
        self.loss_G_A_xz = g(self.iter_f(self.fake, self.netD_A_xz, r), True) * self.lambda_plane_ref
        self.loss_G_A = self.loss_G_A_xy + self.loss_G_A_yz + self.loss_G_A_xz
        self.loss_G_B_xy = g(self.iter_f(self.rec, self.netD_B_xy, t), True) * (1 / 3)
        self.loss_G_B_yz = g(self.iter_f(self.rec, self.netD_B_yz, s), True) * (1 / 3)
        self.loss_G_B_xz = g(self.iter_f(self.rec, self.netD_B_xz, r), True) * (1 / 3)
        self.loss_G_B = self.loss_G_B_xy + self.loss_G_B_yz + self.loss_G_B_xz
        self.loss_cycle_A = self.criterionCycle(self.rec, self.real) * self.opt.lambda_A
        self.loss_G = self.loss_G_A + self.loss_G_B + self.loss_cycle_A
        self.loss_G.backward()

    def optimize_parameters(self):                             # :258-282
        t, s, r = self.target_sl_axis, self.source_sl_axis, self.remain_sl_axis
        self.forward()
        for net in self._ds():
            for p in net.parameters():
                p.requires_grad_(False)
        self.optimizer_G.zero_grad()
        self.backward_G()
        if self.distributed:
            allreduce_mean_gradients(self.optimizer_G.params, self.group)
        self.optimizer_G.step()
        for net in self._ds():
            for p in net.parameters():
                p.requires_grad_(True)
        self.optimizer_D.zero_grad()
        self.loss_D_A_xy = self.backward_D_basic(self.netD_A_xy, self.real, self.fake, t, t)
        self.loss_D_A_yz = self.backward_D_basic(self.netD_A_yz, self.real, self.fake, t, s)
        self.loss_D_A_xz = self.backward_D_basic(self.netD_A_xz, self.real, self.fake, t, r)
        self.loss_D_B_xy = self.backward_D_basic(self.netD_B_xy, self.real, self.rec, t, t)
        self.loss_D_B_yz = self.backward_D_basic(self.netD_B_yz, self.real, self.rec, s, s)
        self.loss_D_B_xz = self.backward_D_basic(self.netD_B_xz, self.real, self.rec, r, r)
        if self.distributed:
            allreduce_mean_gradients(self.optimizer_D.params, self.group)
        self.optimizer_D.step()

    def get_current_losses(self):
        return OrderedDict((n, float(getattr(self, "loss_" + n))) for n in self.loss_names)
