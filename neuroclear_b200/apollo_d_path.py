"""The projection / discriminator path of the reference's ``AxialToLateralGANApolloModel``
(models/axial_to_lateral_gan_apollo_model.py) on the B200 kernels:

* the four 2-D PatchGAN discriminators D_A_axial, D_A_lateral, D_B_axial, D_B_lateral (:99-123), created in the
  reference's order so that seeded initialisation matches;
* ``iter_f`` / ``proj_f`` (:310-320) on ``Volume`` — same host ``np.random`` draws in the same order;
* the six discriminator losses ``backward_D_*`` (:169-253) and the discriminator update (:297-307) with a fused
  Adam kernel (lr, betas=(beta1, 0.999));
* the generator-side adversarial terms of ``backward_G`` (:255-276) as a differentiable function of ``fake`` and
  ``rec`` (so that a generator's autograd can continue from them), and the cycle L1 term (:279).

The generators' own forward/backward (G_A = unet_deconv training, G_B = deep_linear_gen) are not part of this
module.
"""
from __future__ import annotations

import itertools

import numpy as np
import torch

from . import discriminator
from ._lib import call, f32, stream_ptr
from .projection import Volume


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas) semantics (no weight decay / amsgrad) with ONE kernel launch per parameter
    group (nc_adam_step_multi).  A real torch Optimizer: param_groups / state / zero_grad / state_dict and the
    torch.optim.lr_scheduler classes (networks.get_scheduler) work as with the reference's torch.optim.Adam."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps))
        self.step_count = 0
        self._table = None
        self._table_event = None

    @property
    def params(self):
        return [p for g in self.param_groups for p in g["params"]]

    def _launch(self, table_dev, offset, count, group, step):
        """one launch over `count` entries of the device table starting at entry `offset`"""
        import ctypes as C
        call("nc_adam_step_multi", C.c_void_p(table_dev.data_ptr() + 40 * offset), count, f32(group["lr"]),
             f32(group["betas"][0]), f32(group["betas"][1]), f32(group["eps"]), step, stream_ptr())

    @torch.no_grad()
    def step(self, closure=None):
        import numpy as np
        self.step_count += 1
        if self._table_event is not None:
            self._table_event.synchronize()      # the pinned table of the previous step has been consumed
        rows, spans, keep = [], [], []
        for group in self.param_groups:
            by_step = {}      # torch.optim.Adam counts steps PER PARAMETER (one skipped for lack of a gradient lags)
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"], st["exp_avg"], st["exp_avg_sq"] = 0, torch.zeros_like(p), torch.zeros_like(p)
                st["step"] += 1
                g = p.grad.contiguous()
                keep.append(g)
                by_step.setdefault(st["step"], []).append(
                    [p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(), p.numel()])
            for step, entries in sorted(by_step.items()):
                spans.append((len(rows), len(entries), group, step))
                rows += entries
        if not rows:
            return
        dev = self.param_groups[0]["params"][0].device
        if self._table is None or self._table.shape[0] < len(rows):
            self._table = torch.empty((len(rows), 5), dtype=torch.int64)
            if dev.type == "cuda":
                self._table = self._table.pin_memory()
        tab = self._table[:len(rows)]
        tab.numpy()[:] = np.array(rows, dtype=np.int64)
        with torch.cuda.device(dev) if dev.type == "cuda" else _nullcontext():
            tab_dev = tab.to(dev, non_blocking=True)
            for first, count, group, step in spans:
                self._launch(tab_dev, first, count, group, step)
            if dev.type == "cuda":
                self._table_event = torch.cuda.Event()
                self._table_event.record()
        for group in self.param_groups:   # the kernel wrote through raw pointers: bump the version counters
            for p in group["params"]:
                if p.grad is not None:
                    torch.autograd.graph.increment_version(p)


class _params_require_grad:
    """requires_grad_(True) on the discriminators' parameters for the duration of the block, restoring the previous
    flags afterwards (the tape of a pending generator backward recorded them as frozen)."""

    def __init__(self, nets):
        self.params = [p for n in nets for p in n.parameters()]

    def __enter__(self):
        self.prev = [p.requires_grad for p in self.params]
        for p in self.params:
            p.requires_grad_(True)

    def __exit__(self, *a):
        for p, f in zip(self.params, self.prev):
            p.requires_grad_(f)
        return False


class _nullcontext:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


#: when set to a list, allreduce_mean_gradients appends a (start, end) CUDA-event pair around every bucket
#: all-reduce (bench.py's `train_step_dp.allreduce_ms`)
ALLREDUCE_EVENTS = None


def allreduce_mean_gradients(params, group=None):
    """Data-parallel training: average the gradients of `params` over the ranks of `group` with ONE
    all-reduce of a flat fp32 bucket (11 M discriminator parameters = 44 MB: a single NCCL call over NVLink).
    Every rank must hold a gradient for the same parameters; parameters without gradient are skipped everywhere."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    with_grad = [p for p in params if p.grad is not None]
    if world == 1 or not with_grad:
        return
    flat = torch.cat([p.grad.reshape(-1) for p in with_grad])
    timed = ALLREDUCE_EVENTS is not None and flat.is_cuda
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if timed:
        e1.record()
        ALLREDUCE_EVENTS.append((e0, e1))
    flat.div_(world)
    off = 0
    for p in with_grad:
        n = p.grad.numel()
        p.grad.copy_(flat[off:off + n].view_as(p.grad))
        off += n


class ApolloDiscriminatorPath:
    """with_B=False: the ablation of models/axial_to_lateral_gan_dryops_model.py — no G_B, no D_B_*, no cycle term."""

    def __init__(self, opt, device, group=None, distributed=None, with_B=True):
        import torch.distributed as dist
        self.opt = opt
        self.with_B = with_B
        self.device = torch.device(device)
        self.group = group
        self.distributed = dist.is_initialized() if distributed is None else distributed
        gpu_ids = [self.device.index if self.device.index is not None else 0]
        mk = lambda nc: discriminator.define_D(nc, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, opt.init_type,
                                               opt.init_gain, False, gpu_ids, dimension=2)
        self.netD_A_axial = mk(opt.output_nc)      # creation order of apollo_model.py:99-123
        self.netD_A_lateral = mk(opt.output_nc)
        if with_B:
            self.netD_B_axial = mk(opt.input_nc)
            self.netD_B_lateral = mk(opt.input_nc)
        self.criterionGAN = discriminator.GANLoss(opt.gan_mode).to(self.device)
        self.criterionCycle = discriminator.L1Loss()
        nets = [self.netD_A_axial, self.netD_A_lateral] + ([self.netD_B_axial, self.netD_B_lateral] if with_B else [])
        self.optimizer_D = FusedAdam(itertools.chain(*[n.parameters() for n in nets]),
                                     lr=opt.lr, betas=(opt.beta1, 0.999))
        s = float(sum(opt.lambda_plane))
        self.lambda_plane_target, self.lambda_slice, self.lambda_proj = [f / s for f in opt.lambda_plane]
        self.lateral_axis, self.axial_1_axis, self.axial_2_axis = 0, 1, 2
        self.randomize_projection_depth = opt.randomize_projection_depth
        self.projection_depth = opt.projection_depth
        #: True: the passes of one discriminator inside optimize_D / generator_losses run as ONE batched pass
        #: (InstanceNorm statistics are per image, so a slice's prediction does not depend on its batch mates) and the
        #: six D losses are back-propagated by one backward(): 8 discriminator passes per iteration instead of 18, and
        #: a third of the host work.  Same np.random draws in the same order, same loss values up to fp32 summation
        #: order.  False: one pass per slice, one backward per loss — literally the reference's call sequence.
        self.batched = True

    def discriminators(self):
        return [self.netD_A_lateral, self.netD_A_axial] + \
            ([self.netD_B_lateral, self.netD_B_axial] if self.with_B else [])

    # ---- set_input (:142-160): the projection depth is drawn once per step
    def draw_projection_depth(self):
        if self.randomize_projection_depth:
            self.projection_depth = np.random.randint(max(2, self.opt.min_projection_depth),
                                                      self.opt.projection_depth + 1)
        return self.projection_depth

    # ---- :310-320
    def iter_f(self, input, function, slice_axis):
        return function(Volume(input, self.device).get_slice(slice_axis))

    def proj_f(self, input, function, slice_axis):
        return function(Volume(input, self.device).get_projection(self.projection_depth, slice_axis))

    # ---- :169-229
    def backward_D_slice(self, netD, real, fake, slice_axis_real, slice_axis_fake):
        pred_real = self.iter_f(real, netD, slice_axis_real)
        pred_fake = self.iter_f(fake.detach(), netD, slice_axis_fake)
        loss_D = (self.criterionGAN(pred_real, True) + self.criterionGAN(pred_fake, False)) * 0.5
        loss_D.backward()
        return loss_D

    def backward_D_projection(self, netD, real, fake, slice_axis_real, slice_axis_fake):
        pred_real = self.iter_f(real, netD, slice_axis_real)
        pred_fake = self.proj_f(fake.detach(), netD, slice_axis_fake)
        loss_D = (self.criterionGAN(pred_real, True) + self.criterionGAN(pred_fake, False)) * 0.5
        loss_D.backward()
        return loss_D

    # ---- :231-253
    def backward_D_A_lateral(self, real, fake):
        self.loss_D_A_lateral = self.backward_D_projection(self.netD_A_lateral, real, fake, 0, 0)

    def backward_D_A_axial(self, real, fake):
        self.loss_D_A_axial_1 = self.backward_D_projection(self.netD_A_axial, real, fake, 0, 1)
        self.loss_D_A_axial_2 = self.backward_D_projection(self.netD_A_axial, real, fake, 0, 2)
        self.loss_D_A_axial = (self.loss_D_A_axial_1 + self.loss_D_A_axial_2) * 0.5

    def backward_D_B_lateral(self, real, rec):
        self.loss_D_B_lateral = self.backward_D_slice(self.netD_B_lateral, real, rec, 0, 0)

    def backward_D_B_axial(self, real, rec):
        self.loss_D_B_axial_1 = self.backward_D_slice(self.netD_B_axial, real, rec, 1, 1)
        self.loss_D_B_axial_2 = self.backward_D_slice(self.netD_B_axial, real, rec, 2, 2)
        self.loss_D_B_axial = (self.loss_D_B_axial_1 + self.loss_D_B_axial_2) * 0.5

    # ---- batched form of the four backward_D_* calls (same draws, same order; see self.batched)
    @staticmethod
    def _stack(images):
        """[(1,1,a,b)...] -> (N,1,a,b); falls back to None when the images differ in size (non-cubic crops)"""
        if any(im.shape != images[0].shape for im in images):
            return None
        return torch.cat(images, 0)

    def _backward_D_batched(self, real, fake, rec):
        r, f = Volume(real, self.device), Volume(fake.detach(), self.device)
        d = self.projection_depth
        # draws in the reference's order: (real slice, fake projection) per backward_D_projection call (:231-238) ...
        a_lat = [r.get_slice(0), f.get_projection(d, 0)]
        a_ax = [r.get_slice(0), f.get_projection(d, 1), r.get_slice(0), f.get_projection(d, 2)]
        groups = [(self.netD_A_lateral, a_lat), (self.netD_A_axial, a_ax)]
        if self.with_B:                      # ... then (real slice, rec slice) per backward_D_slice call (:240-253)
            c = Volume(rec.detach(), self.device)
            b_lat = [r.get_slice(0), c.get_slice(0)]
            b_ax = [r.get_slice(1), c.get_slice(1), r.get_slice(2), c.get_slice(2)]
            groups += [(self.netD_B_lateral, b_lat), (self.netD_B_axial, b_ax)]
        losses = []
        for net, images in groups:
            batch = self._stack(images)
            if batch is None:
                return False
            per_image = discriminator.batched_lsgan_losses(net(batch), [i % 2 == 0 for i in range(len(images))])
            losses.append((per_image[0::2] + per_image[1::2]) * 0.5)      # loss_D of each (real, fake) pair
        torch.cat(losses).sum().backward()                                # = the six loss_D.backward() calls
        self.loss_D_A_lateral = losses[0][0]
        self.loss_D_A_axial_1, self.loss_D_A_axial_2 = losses[1][0], losses[1][1]
        self.loss_D_A_axial = (self.loss_D_A_axial_1 + self.loss_D_A_axial_2) * 0.5
        if self.with_B:
            self.loss_D_B_lateral = losses[2][0]
            self.loss_D_B_axial_1, self.loss_D_B_axial_2 = losses[3][0], losses[3][1]
            self.loss_D_B_axial = (self.loss_D_B_axial_1 + self.loss_D_B_axial_2) * 0.5
        return True

    def backward_D_all(self, real, fake, rec):
        """the four backward_D_* calls of optimize_parameters (:302-306): gradients of the six D losses into .grad"""
        cubic = real.shape[-1] == real.shape[-2] == real.shape[-3]
        if not (self.batched and cubic and self._backward_D_batched(real, fake, rec)):
            self.backward_D_A_lateral(real, fake)
            self.backward_D_A_axial(real, fake)
            if self.with_B:
                self.backward_D_B_lateral(real, rec)
                self.backward_D_B_axial(real, rec)

    # ---- discriminator half of optimize_parameters (:297-307), in two parts so that the gradient computation can be
    # enqueued on a side stream while the generators' backward runs (apollo_model.overlap_d_step)
    def d_gradients(self, real, fake, rec):
        # NOTE: requires_grad of the D parameters is NOT touched here — the generators' backward (which needs them
        # frozen) may still be pending on the tape; _PatchGANFn computes parameter gradients on request
        self.optimizer_D.zero_grad()
        with _params_require_grad(self.discriminators()):
            self.backward_D_all(real, fake, rec)

    def d_update(self):
        if self.device.type == "cuda":      # the gradients were produced on a side stream and are consumed on this one
            cur = torch.cuda.current_stream(self.device)
            for p in self.optimizer_D.params:
                if p.grad is not None:
                    p.grad.record_stream(cur)
        if self.distributed:    # one crop per GPU; gradients averaged over the ranks before the update
            allreduce_mean_gradients(self.optimizer_D.params, self.group)
        self.optimizer_D.step()

    def optimize_D(self, real, fake, rec):
        for net in self.discriminators():
            for p in net.parameters():
                p.requires_grad_(True)
        self.optimizer_D.zero_grad()
        self.backward_D_all(real, fake, rec)
        if self.distributed:    # one crop per GPU; gradients averaged over the ranks before the update
            allreduce_mean_gradients(self.optimizer_D.params, self.group)
        self.optimizer_D.step()

    def _generator_losses_batched(self, real, fake, rec):
        """backward_G's adversarial terms with the two axial passes of each discriminator batched (draw order kept)"""
        bl = discriminator.batched_lsgan_losses
        f = Volume(fake, self.device)
        d = self.projection_depth
        p_lat = f.get_projection(d, 0)                                   # :259
        p_ax = torch.cat([f.get_projection(d, 1), f.get_projection(d, 2)], 0)            # :260-263
        self.loss_G_A_lateral = bl(self.netD_A_lateral(p_lat), [True])[0] * self.lambda_plane_target
        ax = bl(self.netD_A_axial(p_ax), [True, True])
        self.loss_G_A_axial = ax[0] * self.lambda_slice + ax[1] * self.lambda_slice
        self.loss_G_A = self.loss_G_A_lateral + self.loss_G_A_axial * 0.5
        if not self.with_B:
            self.loss_G = self.loss_G_A
            return self.loss_G
        c = Volume(rec, self.device)
        s_lat = c.get_slice(0)                                            # :269
        s_ax = torch.cat([c.get_slice(1), c.get_slice(2)], 0)             # :271-274
        self.loss_G_B_lateral = bl(self.netD_B_lateral(s_lat), [True])[0] * self.lambda_plane_target
        bx = bl(self.netD_B_axial(s_ax), [True, True])
        self.loss_G_B_axial = bx[0] * self.lambda_slice + bx[1] * self.lambda_slice
        self.loss_G_B = self.loss_G_B_lateral + self.loss_G_B_axial * 0.5
        self.loss_cycle = self.criterionCycle(rec, real) * self.opt.lambda_A
        self.loss_G = self.loss_G_A + self.loss_G_B + self.loss_cycle
        return self.loss_G

    # ---- generator-side terms of backward_G (:255-281); Ds are frozen (set_requires_grad(..., False), :291-292)
    def generator_losses(self, real, fake, rec):
        for net in self.discriminators():
            for p in net.parameters():
                p.requires_grad_(False)
        cubic = fake.shape[-1] == fake.shape[-2] == fake.shape[-3]
        if self.batched and cubic:
            return self._generator_losses_batched(real, fake, rec)
        g = self.criterionGAN
        self.loss_G_A_lateral = g(self.proj_f(fake, self.netD_A_lateral, 0), True) * self.lambda_plane_target
        self.loss_G_A_axial = g(self.proj_f(fake, self.netD_A_axial, 1), True) * self.lambda_slice + \
            g(self.proj_f(fake, self.netD_A_axial, 2), True) * self.lambda_slice
        self.loss_G_A = self.loss_G_A_lateral + self.loss_G_A_axial * 0.5
        if not self.with_B:                                   # dryops_model.py:209-222
            self.loss_G = self.loss_G_A
            return self.loss_G
        self.loss_G_B_lateral = g(self.iter_f(rec, self.netD_B_lateral, 0), True) * self.lambda_plane_target
        self.loss_G_B_axial = g(self.iter_f(rec, self.netD_B_axial, 1), True) * self.lambda_slice + \
            g(self.iter_f(rec, self.netD_B_axial, 2), True) * self.lambda_slice
        self.loss_G_B = self.loss_G_B_lateral + self.loss_G_B_axial * 0.5
        self.loss_cycle = self.criterionCycle(rec, real) * self.opt.lambda_A
        self.loss_G = self.loss_G_A + self.loss_G_B + self.loss_cycle
        return self.loss_G
