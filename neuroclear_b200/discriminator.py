"""Drop-ins for the reference's 2-D PatchGAN discriminator path (training):

* ``define_D(..., netD='basic', dimension=2)`` / ``NLayerDiscriminator`` — reference models/networks.py:199-247,
  1009-1067 — same ``model.{0,2,5,8,11}.{weight,bias}`` state_dict, same initialisation;
* ``GANLoss('lsgan')`` — networks.py:252-319 — and an L1 criterion (``torch.nn.L1Loss``, apollo_model.py:128).

Forward AND backward run the hand-written kernels of ``csrc/disc2d.cu`` through ``torch.autograd.Function``s
(PyTorch supplies the tape and the tensors, none of the arithmetic).  The ``nn.Conv2d`` children are parameter
containers only.  CUDA only: there is no CPU fallback.
"""
from __future__ import annotations


import torch
import torch.nn as nn

from ._lib import NeuroclearError, call, f32, i64, ptr, stream_ptr
from .networks import get_norm_layer, init_net

LRELU_SLOPE = 0.2
IN_EPS = 1e-5


def _check(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise NeuroclearError("%s must be a float32 CUDA tensor (no CPU fallback)" % name)
    return t.contiguous()


class _Conv2dK4(torch.autograd.Function):
    """Conv2d(k4, p1, stride) [+ fused LeakyReLU]: nc_conv2d_k4_fwd / _dgrad / _wgrad (+ nc_lrelu_bwd)."""

    @staticmethod
    def forward(ctx, x, w, b, stride, slope):
        x, w = _check(x, "input"), _check(w, "weight")
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        ho, wo = (h - 2) // stride + 1, (wd - 2) // stride + 1
        y = torch.empty((n, cout, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("nc_conv2d_k4_fwd", ptr(x), ptr(w), ptr(b), n, cin, h, wd, cout, stride, f32(slope), ptr(y), stream_ptr())
        ctx.save_for_backward(x, w, y if slope != 1.0 else None)
        ctx.meta = (stride, slope, b is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y = ctx.saved_tensors
        stride, slope, has_b = ctx.meta
        n, cin, h, wd = x.shape
        cout = w.shape[0]
        dy = dy.contiguous()
        with torch.cuda.device(x.device):
            s = stream_ptr()
            if slope != 1.0:
                g = torch.empty_like(dy)
                call("nc_lrelu_bwd", ptr(dy), ptr(y), i64(dy.numel()), f32(slope), ptr(g), s)
                dy = g
            dx = dw = db = None
            if ctx.needs_input_grad[0]:
                dx = torch.empty_like(x)
                call("nc_conv2d_k4_dgrad", ptr(dy), ptr(w), n, cin, h, wd, cout, stride, ptr(dx), s)
            if ctx.needs_input_grad[1] or (has_b and ctx.needs_input_grad[2]):
                dw = torch.empty_like(w)
                db = torch.empty(cout, dtype=torch.float32, device=x.device) if has_b else None
                call("nc_conv2d_k4_wgrad", ptr(x), ptr(dy), n, cin, h, wd, cout, stride, ptr(dw), ptr(db), s)
        return dx, dw, db, None, None


class _InstanceNormLReLU(torch.autograd.Function):
    """InstanceNorm2d(affine=False) + LeakyReLU(slope): nc_in2d_lrelu_fwd / _bwd."""

    @staticmethod
    def forward(ctx, x, slope):
        x = _check(x, "input")
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        mr = torch.empty((n * c, 2), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            call("nc_in2d_lrelu_fwd", ptr(x), n * c, h * w, f32(IN_EPS), f32(slope), ptr(y), ptr(mr), stream_ptr())
        ctx.save_for_backward(x, mr)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mr = ctx.saved_tensors
        n, c, h, w = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        with torch.cuda.device(x.device):
            call("nc_in2d_lrelu_bwd", ptr(dy), ptr(x), ptr(mr), n * c, h * w, f32(ctx.slope), ptr(dx), stream_ptr())
        return dx, None


class _Loss(torch.autograd.Function):
    """mode 0: mean((p - target)^2); mode 1: mean(|p - q|): nc_loss_fwd / nc_loss_bwd."""

    @staticmethod
    def forward(ctx, p, q, target, mode):
        p = _check(p, "prediction")
        q = _check(q, "target") if q is not None else None
        out = torch.empty((), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            call("nc_loss_fwd", ptr(p), ptr(q), f32(target), i64(p.numel()), mode, ptr(out), stream_ptr())
        ctx.save_for_backward(p, q)
        ctx.meta = (target, mode)
        return out

    @staticmethod
    def backward(ctx, g):
        p, q = ctx.saved_tensors
        target, mode = ctx.meta
        g = g.contiguous().float()
        dp = torch.empty_like(p)
        with torch.cuda.device(p.device):
            call("nc_loss_bwd", ptr(p), ptr(q), f32(target), i64(p.numel()), mode, ptr(g), ptr(dp), stream_ptr())
        return dp, (-dp if (q is not None and ctx.needs_input_grad[1]) else None), None, None


class _BatchedLsganLoss(torch.autograd.Function):
    """Per-image LSGAN losses of a batched prediction map: out[i] = mean((pred[i] - target[i])^2), one nc_loss_fwd /
    nc_loss_bwd per image on a slice of the batch (no torch slicing ops on the tape)."""

    @staticmethod
    def forward(ctx, pred, targets):
        import ctypes as C
        pred = _check(pred, "prediction")
        n = pred.shape[0]
        per = pred.numel() // n
        out = torch.empty(n, dtype=torch.float32, device=pred.device)
        with torch.cuda.device(pred.device):
            for i in range(n):
                call("nc_loss_fwd", C.c_void_p(pred.data_ptr() + 4 * per * i), None, f32(targets[i]), i64(per), 0,
                     C.c_void_p(out.data_ptr() + 4 * i), stream_ptr())
        ctx.save_for_backward(pred)
        ctx.targets = tuple(targets)
        return out

    @staticmethod
    def backward(ctx, g):
        import ctypes as C
        pred, = ctx.saved_tensors
        n = pred.shape[0]
        per = pred.numel() // n
        g = g.contiguous().float()
        dp = torch.empty_like(pred)
        with torch.cuda.device(pred.device):
            for i in range(n):
                call("nc_loss_bwd", C.c_void_p(pred.data_ptr() + 4 * per * i), None, f32(ctx.targets[i]), i64(per), 0,
                     C.c_void_p(g.data_ptr() + 4 * i), C.c_void_p(dp.data_ptr() + 4 * per * i), stream_ptr())
        return dp, None


def batched_lsgan_losses(prediction, targets_are_real):
    """GANLoss('lsgan') of every image of a batched prediction map -> float32 vector (N,)"""
    return _BatchedLsganLoss.apply(prediction, tuple(1.0 if t else 0.0 for t in targets_are_real))


class _PatchGANFn(torch.autograd.Function):
    """The whole discriminator in one call per direction (nc_patchgan_fwd / nc_patchgan_bwd): the 18 passes of a
    training iteration are launch-latency bound, so the layer loop lives inside the library, not in Python."""

    @staticmethod
    def forward(ctx, ndf, n_layers, x, *params):
        import ctypes as C
        from . import _lib
        x = _check(x, "input")
        n, cin, h, w = x.shape
        if cin != 1:
            raise NeuroclearError("the PatchGAN path takes 1-channel images")
        ws_floats = _lib.load().nc_patchgan_ws_floats(n, h, w, ndf, n_layers)
        if ws_floats < 0:
            raise NeuroclearError(_lib.load().nc_last_error().decode())
        L = n_layers + 2
        ho, wo = h, w
        for i in range(L):
            s = 2 if i < n_layers else 1
            ho, wo = (ho - 2) // s + 1, (wo - 2) // s + 1
        params = [_check(p, "parameter") for p in params]
        ws = torch.empty(ws_floats, dtype=torch.float32, device=x.device)
        pred = torch.empty((n, 1, ho, wo), dtype=torch.float32, device=x.device)
        wts = (C.c_void_p * L)(*[p.data_ptr() for p in params[0::2]])
        bss = (C.c_void_p * L)(*[p.data_ptr() for p in params[1::2]])
        with torch.cuda.device(x.device):
            call("nc_patchgan_fwd", ptr(x), n, h, w, ndf, n_layers, wts, bss, ptr(ws), ptr(pred), stream_ptr())
        ctx.save_for_backward(x, ws, *params)
        ctx.meta = (ndf, n_layers)
        return pred

    @staticmethod
    def backward(ctx, dpred):
        import ctypes as C
        x, ws, *params = ctx.saved_tensors
        ndf, n_layers = ctx.meta
        L = n_layers + 2
        n, _, h, w = x.shape
        dpred = dpred.contiguous()
        need_x = ctx.needs_input_grad[2]
        need_p = any(ctx.needs_input_grad[3:])
        dx = torch.empty_like(x) if need_x else None
        grads = [torch.empty_like(p) for p in params] if need_p else [None] * len(params)
        wts = (C.c_void_p * L)(*[p.data_ptr() for p in params[0::2]])
        dws = (C.c_void_p * L)(*[g.data_ptr() for g in grads[0::2]]) if need_p else None
        dbs = (C.c_void_p * L)(*[g.data_ptr() for g in grads[1::2]]) if need_p else None
        with torch.cuda.device(x.device):
            call("nc_patchgan_bwd", ptr(x), ptr(dpred), n, h, w, ndf, n_layers, wts, ptr(ws), ptr(dx), dws, dbs,
                 stream_ptr())
        return (None, None, dx) + tuple(grads)


class NLayerDiscriminator(nn.Module):
    """reference networks.py:1009-1067 with dimension=2 and InstanceNorm (use_bias=True); forward/backward on
    the kernels above.  Layer plan: (conv s2 + LReLU), (conv s2, IN, LReLU) x (n_layers-1), (conv s1, IN, LReLU),
    conv s1 -> 1 channel."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, dimension=2):
        super().__init__()
        if dimension != 2:
            raise NotImplementedError("the B200 discriminator path implements the 2-D PatchGAN of the apollo model")
        if use_sigmoid:
            raise NotImplementedError("use_sigmoid=True is not on the apollo path (apollo_model.py:109)")
        norm_layer = norm_layer or get_norm_layer("instance", 2)
        probe = norm_layer(1)
        if not isinstance(probe, nn.InstanceNorm2d) or probe.affine or probe.track_running_stats:
            raise NotImplementedError("the discriminator path is built for --norm instance (affine=False)")
        kw, padw = 4, 1
        seq = [nn.Conv2d(input_nc, ndf, kw, 2, padw), nn.LeakyReLU(LRELU_SLOPE, True)]
        self._plan = [(0, 2, "lrelu")]                     # (index in model, stride, what follows)
        nf = 1
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(2 ** n, 8)
            self._plan.append((len(seq), 2, "in"))
            seq += [nn.Conv2d(ndf * nf_prev, ndf * nf, kw, 2, padw, bias=True), norm_layer(ndf * nf),
                    nn.LeakyReLU(LRELU_SLOPE, True)]
        nf_prev, nf = nf, min(2 ** n_layers, 8)
        self._plan.append((len(seq), 1, "in"))
        seq += [nn.Conv2d(ndf * nf_prev, ndf * nf, kw, 1, padw, bias=True), norm_layer(ndf * nf),
                nn.LeakyReLU(LRELU_SLOPE, True)]
        self._plan.append((len(seq), 1, None))
        seq += [nn.Conv2d(ndf * nf, 1, kw, 1, padw)]
        self.model = nn.Sequential(*seq)
        self._ndf, self._n_layers, self._input_nc = ndf, n_layers, input_nc
        #: False runs the network layer by layer through the per-layer autograd Functions (same kernels)
        self.fused_pass = True

    def forward(self, input):
        if not input.is_cuda:
            raise NeuroclearError("NLayerDiscriminator (B200): input is on the CPU; there is no CPU fallback")
        x = input.float()
        if self.fused_pass and self._input_nc == 1:
            return _PatchGANFn.apply(self._ndf, self._n_layers, x, *self.parameters())
        for idx, stride, follow in self._plan:
            conv = self.model[idx]
            x = _Conv2dK4.apply(x, conv.weight, conv.bias, stride, LRELU_SLOPE if follow == "lrelu" else 1.0)
            if follow == "in":
                x = _InstanceNormLReLU.apply(x, LRELU_SLOPE)
        return x


class NLayerDiscriminatorSN(nn.Module):
    """reference networks.py:1069-1110 (netD 'basic_SN' / 'n_layers_SN'): the PatchGAN without
    normalisation layers, every convolution wrapped in torch.nn.utils.spectral_norm.  The children ARE the
    reference's (spectral_norm(nn.Conv2d) — same state_dict: model.N.{bias, weight_orig, weight_u, weight_v});
    torch's hook performs the power iteration and produces weight_orig / sigma (a handful of mat-vecs on a
    (Cout, 16 Cin) matrix: parameter plumbing), and every convolution + LeakyReLU, forward and backward, runs on
    nc_conv2d_k4_* with that weight."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm_layer=None, use_sigmoid=False, dimension=2):
        super().__init__()
        if dimension != 2:
            raise NotImplementedError("the B200 discriminator path implements 2-D PatchGANs")
        if use_sigmoid:
            raise NotImplementedError("use_sigmoid=True is not on the B200 path")
        kw, padw, sn = 4, 1, nn.utils.spectral_norm
        seq = [sn(nn.Conv2d(input_nc, ndf, kw, 2, padw)), nn.LeakyReLU(LRELU_SLOPE, True)]
        self._plan = [(0, 2, True)]                          # (index in model, stride, LeakyReLU follows)
        nf = 1
        for n in range(1, n_layers):
            nf_prev, nf = nf, min(2 ** n, 8)
            self._plan.append((len(seq), 2, True))
            seq += [sn(nn.Conv2d(ndf * nf_prev, ndf * nf, kw, 2, padw, bias=False)), nn.LeakyReLU(LRELU_SLOPE, True)]
        nf_prev, nf = nf, min(2 ** n_layers, 8)
        self._plan.append((len(seq), 1, True))
        seq += [sn(nn.Conv2d(ndf * nf_prev, ndf * nf, kw, 1, padw, bias=False)), nn.LeakyReLU(LRELU_SLOPE, True)]
        self._plan.append((len(seq), 1, False))
        seq += [sn(nn.Conv2d(ndf * nf, 1, kw, 1, padw))]
        self.model = nn.Sequential(*seq)

    def forward(self, input):
        if not input.is_cuda:
            raise NeuroclearError("NLayerDiscriminatorSN (B200): input is on the CPU; there is no CPU fallback")
        x = input.float()
        for idx, stride, lrelu in self._plan:
            conv = self.model[idx]
            for hook in conv._forward_pre_hooks.values():     # SpectralNorm: power iteration, weight = W / sigma
                hook(conv, (x,))
            x = _Conv2dK4.apply(x, conv.weight, conv.bias, stride, LRELU_SLOPE if lrelu else 1.0)
        return x


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02, use_sigmoid=False,
             gpu_ids=[], dimension=3):
    """reference networks.py:199-247; 'basic' / 'n_layers' (the PatchGAN) and their spectral-norm variants."""
    norm_layer = get_norm_layer(norm_type=norm, dimension=dimension)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, 3, norm_layer, use_sigmoid, dimension)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm_layer, use_sigmoid, dimension)
    elif netD == "basic_SN":
        net = NLayerDiscriminatorSN(input_nc, ndf, 3, norm_layer, use_sigmoid, dimension)
    elif netD == "n_layers_SN":
        net = NLayerDiscriminatorSN(input_nc, ndf, n_layers_D, norm_layer, use_sigmoid, dimension)
    else:
        raise NotImplementedError("Discriminator model name [%s] is not on the B200 path" % netD)
    return init_net(net, init_type, init_gain, gpu_ids)


class GANLoss(nn.Module):
    """reference networks.py:252-319, gan_mode 'lsgan': MSE between the prediction map and a constant label."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        if gan_mode != "lsgan":
            raise NotImplementedError("gan mode %s is not on the B200 path (README uses lsgan)" % gan_mode)
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        # host copies: reading the (device-resident) buffers back would synchronise the stream on every loss call
        self._labels = (float(target_fake_label), float(target_real_label))

    def forward(self, prediction, target_is_real):
        return _Loss.apply(prediction, None, self._labels[1 if target_is_real else 0], 0)


class L1Loss(nn.Module):
    """torch.nn.L1Loss() as used for the cycle term (apollo_model.py:128,279)."""

    def forward(self, input, target):
        return _Loss.apply(input, target, 0.0, 1)
