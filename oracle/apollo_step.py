"""One training iteration of the reference's AxialToLateralGANApolloModel (models/axial_to_lateral_gan_apollo_model.py:
set_input :142-160, forward :162-167, backward_G :255-283, backward_D_* :169-253, optimize_parameters :285-307)
restated on the CPU from the oracle's pieces (unet, deeplinear, discriminator, mip) with torch autograd and
torch.optim.Adam.  Test infrastructure — see oracle/__init__.py.  Pinned against the fixture recorded from the
reference model (tests/golden/apollo_step_32.npz) in tests/test_oracle_golden.py; bench.py times it as the CPU baseline
of the training iteration.
"""
from __future__ import annotations

import numpy as np
import torch

from . import deeplinear, discriminator, mip, unet

D_NAMES = ["D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"]      # creation order, apollo_model.py:99-123


class ApolloStep:
    def __init__(self, sds: dict, lr=1e-4, beta1=0.1, lambda_A=5.0, lambda_plane=(1, 1, 1), projection_depth=10,
                 min_projection_depth=2, randomize_projection_depth=True):
        """sds: {'G_A', 'G_B', 'D_A_axial', 'D_A_lateral', 'D_B_axial', 'D_B_lateral'} -> state_dict"""
        self.p = {n: {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()} for n, sd in sds.items()}
        s = float(sum(lambda_plane))
        self.l_target, self.l_slice, self.l_proj = [f / s for f in lambda_plane]
        self.lambda_A = lambda_A
        self.max_depth, self.min_depth, self.randomize = projection_depth, min_projection_depth, randomize_projection_depth
        g = [t for n in ("G_A", "G_B") for t in self.p[n].values()]
        d = [t for n in D_NAMES for t in self.p[n].values()]
        self.opt_G = torch.optim.Adam(g, lr=lr, betas=(beta1, 0.999))
        self.opt_D = torch.optim.Adam(d, lr=lr, betas=(beta1, 0.999))
        self.loss = {}

    def set_input(self, real):
        self.real = real
        self.depth = (np.random.randint(max(2, self.min_depth), self.max_depth + 1) if self.randomize
                      else self.max_depth)

    def forward(self):
        self.fake = unet.unet_deconv_forward(self.real, self.p["G_A"], grad=True)
        self.rec = deeplinear.deep_linear_forward(self.fake, self.p["G_B"])

    def _D(self, name, img):
        return discriminator.discriminator_forward(img, self.p[name])

    def _proj(self, vol, name, axis):
        return self._D(name, mip.get_projection(vol, self.depth, axis)[0])

    def _slice(self, vol, name, axis):
        return self._D(name, mip.get_slice(vol, axis)[0])

    def backward_G(self):
        g, L = discriminator.lsgan_loss, self.loss
        L["G_A_lateral"] = g(self._proj(self.fake, "D_A_lateral", 0), True) * self.l_target
        L["G_A_axial"] = g(self._proj(self.fake, "D_A_axial", 1), True) * self.l_slice + \
            g(self._proj(self.fake, "D_A_axial", 2), True) * self.l_slice
        L["G_A"] = L["G_A_lateral"] + L["G_A_axial"] * 0.5
        L["G_B_lateral"] = g(self._slice(self.rec, "D_B_lateral", 0), True) * self.l_target
        L["G_B_axial"] = g(self._slice(self.rec, "D_B_axial", 1), True) * self.l_slice + \
            g(self._slice(self.rec, "D_B_axial", 2), True) * self.l_slice
        L["G_B"] = L["G_B_lateral"] + L["G_B_axial"] * 0.5
        L["cycle"] = (self.rec - self.real).abs().mean() * self.lambda_A
        (L["G_A"] + L["G_B"] + L["cycle"]).backward()

    def _backward_D(self, name, real_img, fake_img):
        g = discriminator.lsgan_loss
        loss = (g(self._D(name, real_img), True) + g(self._D(name, fake_img), False)) * 0.5
        loss.backward()
        return loss

    def _D_projection(self, name, real, fake, ax_real, ax_fake):
        r = mip.get_slice(real, ax_real)[0]
        f = mip.get_projection(fake.detach(), self.depth, ax_fake)[0]
        return self._backward_D(name, r, f)

    def _D_slice(self, name, real, fake, ax_real, ax_fake):
        r = mip.get_slice(real, ax_real)[0]
        f = mip.get_slice(fake.detach(), ax_fake)[0]
        return self._backward_D(name, r, f)

    def set_requires_grad_D(self, flag):
        for n in D_NAMES:
            for t in self.p[n].values():
                t.requires_grad_(flag)

    def backward_D(self):
        """the four backward_D_* calls of optimize_parameters (:302-306), in order"""
        L = self.loss
        L["D_A_lateral"] = self._D_projection("D_A_lateral", self.real, self.fake, 0, 0)
        a1 = self._D_projection("D_A_axial", self.real, self.fake, 0, 1)
        a2 = self._D_projection("D_A_axial", self.real, self.fake, 0, 2)
        L["D_A_axial"] = (a1 + a2) * 0.5
        L["D_B_lateral"] = self._D_slice("D_B_lateral", self.real, self.rec, 0, 0)
        b1 = self._D_slice("D_B_axial", self.real, self.rec, 1, 1)
        b2 = self._D_slice("D_B_axial", self.real, self.rec, 2, 2)
        L["D_B_axial"] = (b1 + b2) * 0.5

    def optimize_parameters(self):
        self.forward()
        self.set_requires_grad_D(False)
        self.opt_G.zero_grad()
        self.backward_G()
        self.opt_G.step()
        self.set_requires_grad_D(True)
        self.opt_D.zero_grad()
        self.backward_D()
        self.opt_D.step()
        return {k: float(v.detach()) for k, v in self.loss.items()}
