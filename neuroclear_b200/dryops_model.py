"""The reference's AxialToLateralGANDryopsModel (models/axial_to_lateral_gan_dryops_model.py): "the Apollo model for
an ablation case with no backward path" — G_A only, the two D_A discriminators, no G_B / D_B / cycle loss
(SURVEY.md §8 f4).  Same kernels, same protocol and host np.random draw order as the apollo mirror."""
from __future__ import annotations

import os
from collections import OrderedDict

import torch

from . import networks
from .apollo_d_path import ApolloDiscriminatorPath, FusedAdam, allreduce_mean_gradients
from .apollo_model import AxialToLateralGANApolloModel

LOSS_NAMES = ["D_A_lateral", "D_A_axial", "G_A", "G_A_lateral", "G_A_axial"]          # dryops_model.py:48


class AxialToLateralGANDryopsModel(AxialToLateralGANApolloModel):
    def __init__(self, opt, device=None, group=None, distributed=None):
        import torch.distributed as dist
        self.opt = opt
        gpu_ids = list(getattr(opt, "gpu_ids", [0])) or [0]
        self.device = torch.device(device if device is not None else "cuda:%d" % gpu_ids[0])
        self.group = group
        self.distributed = dist.is_initialized() if distributed is None else distributed
        ids = [self.device.index if self.device.index is not None else torch.cuda.current_device()]
        self.loss_names = list(LOSS_NAMES)
        self.netG_A = networks.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                        opt.init_type, opt.init_gain, ids, dimension=3)          # :79-81
        self.dpath = ApolloDiscriminatorPath(opt, self.device, group=group, distributed=self.distributed,
                                             with_B=False)                                        # :88-96
        self.netD_A_axial, self.netD_A_lateral = self.dpath.netD_A_axial, self.dpath.netD_A_lateral
        self.optimizer_G = FusedAdam(self.netG_A.parameters(), lr=opt.lr, betas=(opt.beta1, 0.999))   # :102
        self.optimizer_D = self.dpath.optimizer_D
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.model_names = ["G_A", "D_A_lateral", "D_A_axial"]                                    # :76
        self.visual_names = ["real", "fake"]
        self.schedulers = []
        self.metric = 0
        self.save_dir = os.path.join(getattr(opt, "checkpoints_dir", "./checkpoints"), getattr(opt, "name", "dryops"))

    def forward(self):                                       # :132-134
        self.fake = self.netG_A(self.real)

    def backward_G(self):                                    # :208-222
        self.loss_G = self.dpath.generator_losses(self.real, self.fake, None)
        self.loss_G.backward()

    def optimize_parameters(self):                           # :224-245
        self.forward()
        self.optimizer_G.zero_grad()
        self.backward_G()
        if self.distributed:
            allreduce_mean_gradients(self.optimizer_G.params, self.group)
        self.optimizer_G.step()
        self.dpath.optimize_D(self.real, self.fake.detach(), None)

    def test(self):
        with torch.no_grad():
            self.forward()

    def get_current_visuals(self):
        return OrderedDict((k, getattr(self, k)) for k in ("real", "fake"))
