"""Drop-in for the reference's ``models/networks.py`` on the unet_deconv path.

``define_G(..., netG='unet_deconv')`` (reference networks.py:140-197) returns an ``nn.Module`` with exactly the
reference's state_dict (28 tensors, same keys / shapes / OIDHW + IODHW layouts), initialised by the same
``init_weights`` rule (networks.py:88-119), whose ``forward`` runs the hand-written sm_100a kernels through the C
ABI.  The ``nn.Conv3d`` / ``nn.ConvTranspose3d`` children are parameter containers only — they keep checkpoints,
optimisers and ``print_networks`` working — and are never called.
"""
from __future__ import annotations

import functools

import torch
import torch.nn as nn
from torch.nn import init

from ._lib import NeuroclearError
from .unet_engine import UnetDeconvEngine


class Identity(nn.Module):
    def forward(self, x):
        return x


def get_norm_layer(norm_type="instance", dimension=3):
    """reference networks.py:20-44"""
    inorm = {1: nn.InstanceNorm1d, 2: nn.InstanceNorm2d, 3: nn.InstanceNorm3d}[dimension]
    bnorm = {1: nn.BatchNorm1d, 2: nn.BatchNorm2d, 3: nn.BatchNorm3d}[dimension]
    if norm_type == "batch":
        return functools.partial(bnorm, affine=True, track_running_stats=True)
    if norm_type == "instance":
        return functools.partial(inorm, affine=False, track_running_stats=False)
    if norm_type in ("spectral", "none"):
        return lambda x: Identity()
    raise NotImplementedError("normalization layer [%s] is not found" % norm_type)


def get_scheduler(optimizer, opt):
    """reference networks.py:50-86: linear | constant | step | plateau | cosine on the torch schedulers (host logic;
    the optimisers of this package are torch Optimizers, see apollo_d_path.FusedAdam)."""
    from torch.optim import lr_scheduler
    if opt.lr_policy == "linear":
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    if opt.lr_policy == "constant":
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda epoch: 1.0)
    if opt.lr_policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == "plateau":
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5)
    if opt.lr_policy == "cosine":
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.n_epochs, eta_min=0)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


def init_weights(net, init_type="normal", init_gain=0.02):
    """reference networks.py:88-119 (same class-name matching, same initialisers)."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)

    print("initialize network with %s" % init_type)
    net.apply(init_func)
    for m in net.modules():            # the initialisers write through .data: no version bump — say so explicitly
        if hasattr(m, "invalidate"):
            m.invalidate()


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[]):
    """reference networks.py:122-137.  One process drives one GPU here, so a non-empty gpu_ids moves the net to
    gpu_ids[0] and wraps it in a single-device DataParallel: BaseModel.save_networks / load_networks
    (base_model.py:158-160,189-190) unwrap ``.module`` exactly as with the reference."""
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[0])
        net = torch.nn.DataParallel(net, [gpu_ids[0]])
    init_weights(net, init_type, init_gain=init_gain)
    return net


def _conv_block(n_convs, cin, cout, norm_layer):
    """Same child indices as double_conv / triple_conv / last_conv (networks.py:413-476): conv at 0, 3, 6."""
    layers = []
    for i in range(n_convs):
        layers += [nn.Conv3d(cin if i == 0 else cout, cout, 3, 1, 1), norm_layer(cout), nn.ReLU()]
    return layers


class _ConvStack(nn.Module):
    def __init__(self, n_convs, cin, cout, norm_layer):
        super().__init__()
        self.convolution = nn.Sequential(*_conv_block(n_convs, cin, cout, norm_layer))


def _weights_signature(params):
    """(data_ptr, version) per tensor PLUS an exact checksum of the values: the int64 sum of the fp32 bit patterns of
    all parameters (one cat + one reduction on the device, 8 bytes to the host).  The version counter alone is not
    enough: writes through ``p.data`` — the idiom of init_weights, EMA updates, ``p.data.copy_()`` checkpoint loaders —
    do not bump it, and the engines would silently keep stale packed weights."""
    ps = [p.detach() for p in params]
    with torch.no_grad():
        flat = torch.cat([p.reshape(-1).to(torch.float32) for p in ps])
        chk = int(flat.view(torch.int32).sum(dtype=torch.int64).item())
    return tuple((p.data_ptr(), p._version) for p in ps) + (chk,)


class _UnetDeconvFn(torch.autograd.Function):
    """Unet_deconv under autograd: forward / backward run on the training engine (neuroclear_b200.unet_train);
    PyTorch only carries the tape.  The input receives no gradient (G_A's input is the data crop)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.train_engine()
        with torch.cuda.device(x.device):
            y = eng.forward(x.detach().to(torch.float32)[:, 0].contiguous())
        ctx.module, ctx.eng, ctx.saved = module, eng, eng.saved
        eng.saved = None
        return y[:, None]

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        eng.saved, ctx.saved = ctx.saved, None
        with torch.cuda.device(dy.device):
            grads = eng.backward(dy[:, 0].contiguous())
        return (None, None) + tuple(grads[k] for k, _ in ctx.module.named_parameters())


class Unet_deconv(nn.Module):
    """reference networks.py:478-538 on sm_100a kernels: inference engine under no_grad / eval, training engine
    (forward keeping activations + full backward) under autograd."""

    def __init__(self, input_nc, output_nc, norm_layer=None, dimension=3):
        super().__init__()
        if dimension != 3 or input_nc != 1 or output_nc != 1:
            raise NotImplementedError("the B200 path implements the 3-D, 1-channel unet_deconv of the reference")
        norm_layer = norm_layer or get_norm_layer("instance", 3)
        probe = norm_layer(1)
        if not isinstance(probe, nn.InstanceNorm3d) or probe.affine or probe.track_running_stats:
            raise NotImplementedError("unet_deconv on B200 is built for --norm instance (affine=False), as in the "
                                      "reference's README commands")
        nc = input_nc * 64
        self.double_conv1 = _ConvStack(2, input_nc, nc, norm_layer)
        self.double_conv2 = _ConvStack(2, nc, nc * 2, norm_layer)
        self.bottom_layer = _ConvStack(3, nc * 2, nc * 4, norm_layer)
        self.t_conv2 = nn.ConvTranspose3d(nc * 4, nc * 2, 2, 2)
        self.ex_double_conv2 = _ConvStack(2, nc * 4, nc * 2, norm_layer)
        self.t_conv1 = nn.ConvTranspose3d(nc * 2, nc, 2, 2)
        self.ex_conv1_1 = _ConvStack(1, nc * 2, nc, norm_layer)
        self.one_by_one = nn.Conv3d(nc, output_nc, 1, 1, 0)
        self.one_by_one_2 = nn.Conv3d(output_nc, output_nc, 1, 1, 0)
        self._engine = None
        self._engine_sig = None
        self._train_engine = None

    # the packed fp16 weight cache follows the parameters.  Inference: repack when (data_ptr, _version, checksum of
    # the values) changed — in-place updates, load_state_dict, device moves and writes through .data are all seen.
    # Training: the weights change every iteration, so the training engine re-packs on EVERY forward (12 small
    # launches) and needs no signature at all.
    def _signature(self):
        return _weights_signature(self.parameters())

    def invalidate(self):
        """Drop the packed-weight caches (call after writing parameters through a side door)."""
        self._engine_sig = None

    def engine(self) -> UnetDeconvEngine:
        p = next(self.parameters())
        if not p.is_cuda:
            raise NeuroclearError("Unet_deconv (B200): parameters are on the CPU; there is no CPU fallback — "
                                  "move the network to a CUDA device")
        sig = self._signature()
        if self._engine is None or self._engine.device != p.device:
            self._engine, self._engine_sig = UnetDeconvEngine(p.device), None
        if self._engine_sig != sig:
            self._engine.load_state_dict(self.state_dict())
            self._engine_sig = sig
        return self._engine

    def train_engine(self):
        from .unet_train import UnetDeconvTrainEngine
        p = next(self.parameters())
        if not p.is_cuda:
            raise NeuroclearError("Unet_deconv (B200): parameters are on the CPU; there is no CPU fallback — "
                                  "move the network to a CUDA device")
        if self._train_engine is None or self._train_engine.device != p.device:
            self._train_engine = UnetDeconvTrainEngine(p.device)
        self._train_engine.load_state_dict(self.state_dict())
        return self._train_engine

    def forward(self, inputs):
        if inputs.dim() != 5 or inputs.shape[1] != 1:
            raise NeuroclearError("Unet_deconv expects (N, 1, D, H, W)")
        if not inputs.is_cuda:
            raise NeuroclearError("Unet_deconv (B200): input is on the CPU; there is no CPU fallback")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if inputs.requires_grad:
                raise NotImplementedError("the B200 unet_deconv path does not propagate a gradient to its input "
                                          "(it is the first generator of the cycle, fed with data)")
            return _UnetDeconvFn.apply(self, inputs, *self.parameters())
        eng = self.engine()
        with torch.cuda.device(inputs.device):
            x = inputs.detach().to(torch.float32)[:, 0].contiguous()
            return eng.forward(x)[:, None]


class Unet_vanilla(nn.Module):
    """reference networks.py:540-608 (define_G 'unet_vanilla'): the 4-level U-Net on the same sm_100a kernels
    (neuroclear_b200.unet_vanilla_engine).  Same children / state_dict as the reference.  Inference path (no_grad /
    parameters frozen); training this generator is not on the B200 path."""

    def __init__(self, input_nc, output_nc, norm_layer=None, dimension=3):
        super().__init__()
        if dimension != 3 or input_nc != 1 or output_nc != 1:
            raise NotImplementedError("the B200 path implements the 3-D, 1-channel unet_vanilla of the reference")
        norm_layer = norm_layer or get_norm_layer("instance", 3)
        probe = norm_layer(1)
        if not isinstance(probe, nn.InstanceNorm3d) or probe.affine or probe.track_running_stats:
            raise NotImplementedError("unet_vanilla on B200 is built for --norm instance (affine=False)")
        nc = input_nc * 64
        self.double_conv1 = _ConvStack(2, input_nc, nc, norm_layer)
        self.double_conv2 = _ConvStack(2, nc, nc * 2, norm_layer)
        self.double_conv3 = _ConvStack(2, nc * 2, nc * 4, norm_layer)
        self.bottom_layer = _ConvStack(2, nc * 4, nc * 8, norm_layer)
        self.t_conv3 = nn.ConvTranspose3d(nc * 8, nc * 4, 2, 2)
        self.ex_double_conv3 = _ConvStack(2, nc * 8, nc * 4, norm_layer)
        self.t_conv2 = nn.ConvTranspose3d(nc * 4, nc * 2, 2, 2)
        self.ex_double_conv2 = _ConvStack(2, nc * 4, nc * 2, norm_layer)
        self.t_conv1 = nn.ConvTranspose3d(nc * 2, nc, 2, 2)
        self.ex_conv1_1 = _ConvStack(2, nc * 2, nc, norm_layer)
        self.one_by_one = nn.Conv3d(nc, output_nc, 1, 1, 0)
        self._engine = None
        self._engine_sig = None

    def invalidate(self):
        self._engine_sig = None

    def engine(self):
        from .unet_vanilla_engine import UnetVanillaEngine
        p = next(self.parameters())
        if not p.is_cuda:
            raise NeuroclearError("Unet_vanilla (B200): parameters are on the CPU; there is no CPU fallback")
        if self._engine is None or self._engine.device != p.device:
            self._engine, self._engine_sig = UnetVanillaEngine(p.device), None
        sig = _weights_signature(self.parameters())
        if self._engine_sig != sig:
            self._engine.load_state_dict(self.state_dict())
            self._engine_sig = sig
        return self._engine

    def forward(self, inputs):
        if inputs.dim() != 5 or inputs.shape[1] != 1:
            raise NeuroclearError("Unet_vanilla expects (N, 1, D, H, W)")
        if not inputs.is_cuda:
            raise NeuroclearError("Unet_vanilla (B200): input is on the CPU; there is no CPU fallback")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("training unet_vanilla is not on the B200 path (use it under torch.no_grad(); "
                                      "the README trains unet_deconv)")
        with torch.cuda.device(inputs.device):
            return self.engine().forward(inputs.detach().to(torch.float32)[:, 0].contiguous())[:, None]


class _DeepLinearFn(torch.autograd.Function):
    """DeepLinearGenerator under autograd on neuroclear_b200.deeplinear_engine (input gradient included)."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng = module.engine(training=True)
        with torch.cuda.device(x.device):
            y = eng.forward(x.detach().to(torch.float32)[:, 0].contiguous())
        ctx.module, ctx.eng, ctx.saved = module, eng, eng.saved
        eng.saved = None
        return y[:, None]

    @staticmethod
    def backward(ctx, dy):
        eng = ctx.eng
        eng.saved, ctx.saved = ctx.saved, None
        with torch.cuda.device(dy.device):
            dx, grads = eng.backward(dy[:, 0].contiguous())
        return (None, dx[:, None]) + tuple(grads[k] for k, _ in ctx.module.named_parameters())


class DeepLinearGenerator(nn.Module):
    """reference networks.py:893-917 (same children, same state_dict) on sm_100a kernels."""

    def __init__(self, input_nc, output_nc):
        super().__init__()
        if input_nc != 1 or output_nc != 1:
            raise NotImplementedError("the B200 path implements the 1-channel deep_linear_gen of the reference")
        self.first_layer = nn.Conv3d(1, 64, 7, padding=3, bias=False)
        self.feature_block = nn.Sequential(nn.Conv3d(64, 64, 5, padding=2, bias=False),
                                           nn.Conv3d(64, 64, 3, padding=1, bias=False),
                                           nn.Conv3d(64, 32, 1, bias=False), nn.Conv3d(32, 16, 1, bias=False))
        self.final_layer = nn.Conv3d(16, 1, 1, bias=False)
        self._engine = None
        self._engine_sig = None

    def engine(self, training=False):
        from .deeplinear_engine import DeepLinearEngine
        p = next(self.parameters())
        if not p.is_cuda:
            raise NeuroclearError("DeepLinearGenerator (B200): parameters are on the CPU; there is no CPU fallback")
        if self._engine is None or self._engine.device != p.device:
            self._engine, self._engine_sig = DeepLinearEngine(p.device), None
        if training:                         # weights change every iteration — always re-pack, no host check
            self._engine.load_state_dict(self.state_dict())
            self._engine_sig = None
            return self._engine
        sig = _weights_signature(self.parameters())
        if self._engine_sig != sig:
            self._engine.load_state_dict(self.state_dict())
            self._engine_sig = sig
        return self._engine

    def invalidate(self):
        self._engine_sig = None

    def forward(self, input):
        if input.dim() != 5 or input.shape[1] != 1:
            raise NeuroclearError("DeepLinearGenerator expects (N, 1, D, H, W)")
        if not input.is_cuda:
            raise NeuroclearError("DeepLinearGenerator (B200): input is on the CPU; there is no CPU fallback")
        if torch.is_grad_enabled() and (input.requires_grad or any(p.requires_grad for p in self.parameters())):
            return _DeepLinearFn.apply(self, input, *self.parameters())
        with torch.cuda.device(input.device):
            return self.engine().forward(input.detach().to(torch.float32)[:, 0].contiguous(), keep=False)[:, None]


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02,
             gpu_ids=[], kernel_size=9, given_psf=None, noise_setting=None, dimension=3):
    """reference networks.py:140-197; the two generators on the hot path are provided."""
    norm_layer = get_norm_layer(norm_type=norm, dimension=dimension)
    if netG == "unet_deconv":
        net = Unet_deconv(1, output_nc, norm_layer=norm_layer, dimension=dimension)  # input_nc hard-coded 1 (:174)
    elif netG == "unet_vanilla":
        net = Unet_vanilla(1, output_nc, norm_layer=norm_layer, dimension=dimension)  # networks.py:175-176
    elif netG == "deep_linear_gen":
        net = DeepLinearGenerator(input_nc, output_nc)                               # networks.py:193-194
    else:
        raise NotImplementedError("Generator model name [%s] is not on the B200 path (unet_deconv, unet_vanilla and "
                                  "deep_linear_gen are)" % netG)
    return init_net(net, init_type, init_gain, gpu_ids)
