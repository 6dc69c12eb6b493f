"""The training step of the reference's AxialToLateralGANApolloModel (models/axial_to_lateral_gan_apollo_model.py)
on the B200 path: G_A = unet_deconv, G_B = deep_linear_gen (README training command), four 2-D PatchGAN
discriminators on slices / max-intensity projections, LSGAN + L1 cycle loss, two Adam optimisers.

Same protocol as the reference model: set_input(dict) -> optimize_parameters() -> get_current_losses(), attributes
real / fake / rec / loss_*, the same host np.random draw order (projection depth in set_input, then slice / start
indices in backward_G, backward_D_*).  Everything numeric runs on the hand-written kernels: the generators through
networks.define_G (tcgen05 convs forward and backward), the projection / discriminator path through
ApolloDiscriminatorPath, the updates through nc_adam_step.  One process drives one GPU; under torch.distributed the
gradients of both optimisers are averaged over the ranks with one flat-bucket all-reduce each (data parallel,
SURVEY.md §8e)."""
from __future__ import annotations

import itertools
import os
from collections import OrderedDict

import torch

from . import networks
from .apollo_d_path import ApolloDiscriminatorPath, FusedAdam, allreduce_mean_gradients

def _unwrap(net):
    return net.module if isinstance(net, torch.nn.DataParallel) else net


def save_networks(nets, save_dir, epoch):
    """base_model.py:146-162: '<epoch>_net_<name>.pth' holding the CPU state_dict of the unwrapped module (the
    reference moves the module to the CPU and back; copying the tensors gives the same file)."""
    os.makedirs(save_dir, exist_ok=True)
    for name, net in nets.items():
        sd = {k: v.detach().cpu() for k, v in _unwrap(net).state_dict().items()}
        torch.save(sd, os.path.join(save_dir, "%s_net_%s.pth" % (epoch, name)))


def load_networks(nets, save_dir, epoch, device="cpu"):
    """base_model.py:178-201 (the InstanceNorm key patching there concerns checkpoints older than PyTorch 0.4)."""
    for name, net in nets.items():
        path = os.path.join(save_dir, "%s_net_%s.pth" % (epoch, name))
        print("loading the model from %s" % path)
        sd = torch.load(path, map_location=str(device))
        if hasattr(sd, "_metadata"):
            del sd._metadata
        _unwrap(net).load_state_dict(sd)


LOSS_NAMES = ["D_A_lateral", "D_A_axial", "G_A", "G_A_lateral", "G_A_axial", "cycle",
              "D_B_lateral", "D_B_axial", "G_B", "G_B_lateral", "G_B_axial"]      # apollo_model.py:61-62


class AxialToLateralGANApolloModel:
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
                                        opt.init_type, opt.init_gain, ids, dimension=3)          # :91-93
        self.netG_B = networks.define_G(opt.output_nc, opt.input_nc, opt.ngf, opt.netG_B, opt.norm,
                                        not opt.no_dropout, opt.init_type, opt.init_gain, ids, dimension=3)   # :95-97
        self.dpath = ApolloDiscriminatorPath(opt, self.device, group=group, distributed=self.distributed)   # :99-136
        for n in ("D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"):
            setattr(self, "net" + n, getattr(self.dpath, "net" + n))
        self.optimizer_G = FusedAdam(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()),
                                     lr=opt.lr, betas=(opt.beta1, 0.999))                        # :131-132
        self.optimizer_D = self.dpath.optimizer_D
        self.optimizers = [self.optimizer_G, self.optimizer_D]
        self.model_names = ["G_A", "G_B", "D_A_lateral", "D_A_axial", "D_B_lateral", "D_B_axial"]     # :85
        self.visual_names = ["real", "fake", "rec"]
        self.schedulers = []
        self.metric = 0
        self.save_dir = os.path.join(getattr(opt, "checkpoints_dir", "./checkpoints"), getattr(opt, "name", "apollo"))

    # ---- BaseModel protocol used by train_onecube.py (models/base_model.py:81-136,146-201)
    def setup(self, opt):
        self.schedulers = [networks.get_scheduler(o, opt) for o in self.optimizers]
        if getattr(opt, "continue_train", False):
            suffix = "iter_%d" % opt.load_iter if getattr(opt, "load_iter", 0) > 0 else opt.epoch
            self.load_networks(suffix)
        self.print_networks(getattr(opt, "verbose", False))

    def print_networks(self, verbose=False):
        """base_model.py:203-219"""
        print("---------- Networks initialized -------------")
        for name in self.model_names:
            net = getattr(self, "net" + name)
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (name, sum(p.numel() for p in net.parameters()) / 1e6))
        print("-----------------------------------------------")

    def update_learning_rate(self):
        for sch in self.schedulers:
            if getattr(self.opt, "lr_policy", "constant") == "plateau":
                sch.step(self.metric)
            else:
                sch.step()

    def eval(self):
        for name in self.model_names:
            getattr(self, "net" + name).eval()

    def compute_visuals(self):
        pass

    def get_image_paths(self):
        return self.image_paths

    def save_networks(self, epoch):
        save_networks({n: getattr(self, "net" + n) for n in self.model_names}, self.save_dir, epoch)

    def load_networks(self, epoch):
        load_networks({n: getattr(self, "net" + n) for n in self.model_names}, self.save_dir, epoch, self.device)

    # ---- :142-160
    def set_input(self, input):
        a_to_b = getattr(self.opt, "direction", "AtoB") == "AtoB"
        src = input["A" if a_to_b else "B"]
        # a PINNED host crop is copied asynchronously (the reference's blocking .to() drains the stream once per
        # iteration and keeps the host from enqueueing ahead); pageable tensors keep the blocking copy
        self.real = src.to(self.device, non_blocking=bool(getattr(src, "is_pinned", lambda: False)()))
        self.image_paths = input["A_paths" if a_to_b else "B_paths"]
        self.projection_depth = self.dpath.draw_projection_depth()

    # ---- :162-167
    def forward(self):
        self.fake = self.netG_A(self.real)
        self.rec = self.netG_B(self.fake)

    def test(self):
        with torch.no_grad():
            self.forward()

    # ---- :255-283
    def backward_G(self):
        self.loss_G = self.dpath.generator_losses(self.real, self.fake, self.rec)
        self.loss_G.backward()

    #: True: the discriminators' forward / backward of the D update (small, latency-bound kernels) are enqueued on a
    #: side stream BEFORE the generators' backward and run in its shadow; the two Adam steps keep the reference's
    #: order.  Legal because the D gradients depend only on real / fake / rec and the CURRENT D weights (the G update
    #: does not touch them), and backward_G only READS the D weights.  np.random draws keep the reference's order
    #: (generator-side draws first, then the D-side draws).  False: strictly sequential, as written in the reference.
    overlap_d_step = True

    # ---- :285-307
    def optimize_parameters(self):
        self.forward()
        self.optimizer_G.zero_grad()
        if not self.overlap_d_step:
            self.backward_G()                   # the Ds are frozen inside generator_losses (set_requires_grad False)
            if self.distributed:
                allreduce_mean_gradients(self.optimizer_G.params, self.group)
            self.optimizer_G.step()
            self.dpath.optimize_D(self.real, self.fake.detach(), self.rec.detach())
            return
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_d_stream", None) is None:
            self._d_stream = torch.cuda.Stream(self.device)
        side = self._d_stream
        self.loss_G = self.dpath.generator_losses(self.real, self.fake, self.rec)     # draws + D forwards (frozen Ds)
        fake, rec = self.fake.detach(), self.rec.detach()
        side.wait_stream(main)                  # fake / rec / the generator-side D passes are enqueued before this
        with torch.cuda.stream(side):
            for t in (self.real, fake, rec):
                t.record_stream(side)
            self.dpath.d_gradients(self.real, fake, rec)      # zero_grad + the six D losses' backward, no update yet
        self.loss_G.backward()                  # on the main stream, concurrently with the side stream
        if self.distributed:
            allreduce_mean_gradients(self.optimizer_G.params, self.group)
        self.optimizer_G.step()
        main.wait_stream(side)
        self.dpath.d_update()

    def get_current_losses(self):
        out = OrderedDict()
        for name in self.loss_names:
            out[name] = float(getattr(self.dpath, "loss_" + name))
        return out

    def get_current_visuals(self):
        return OrderedDict((k, getattr(self, k)) for k in ("real", "fake", "rec"))
