"""Unet_deconv forward restated with torch.nn.functional on the CPU in fp32.
Test infrastructure — see oracle/__init__.py.

The arithmetic of this path lives in a third-party dependency of the reference (PyTorch 1.10.2 / cuDNN 8.2,
conda_environment/neuroclear_env.yml:192); the restatement follows the reference's own call sites in
models/networks.py:413-538 layer by layer and takes the reference's state_dict unchanged (28 tensors,
SURVEY.md §8b).  It is pinned against the real reference module in oracle/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

STATE_DICT_SHAPES = {
    "double_conv1.convolution.0.weight": (64, 1, 3, 3, 3), "double_conv1.convolution.0.bias": (64,),
    "double_conv1.convolution.3.weight": (64, 64, 3, 3, 3), "double_conv1.convolution.3.bias": (64,),
    "double_conv2.convolution.0.weight": (128, 64, 3, 3, 3), "double_conv2.convolution.0.bias": (128,),
    "double_conv2.convolution.3.weight": (128, 128, 3, 3, 3), "double_conv2.convolution.3.bias": (128,),
    "bottom_layer.convolution.0.weight": (256, 128, 3, 3, 3), "bottom_layer.convolution.0.bias": (256,),
    "bottom_layer.convolution.3.weight": (256, 256, 3, 3, 3), "bottom_layer.convolution.3.bias": (256,),
    "bottom_layer.convolution.6.weight": (256, 256, 3, 3, 3), "bottom_layer.convolution.6.bias": (256,),
    "t_conv2.weight": (256, 128, 2, 2, 2), "t_conv2.bias": (128,),
    "ex_double_conv2.convolution.0.weight": (128, 256, 3, 3, 3), "ex_double_conv2.convolution.0.bias": (128,),
    "ex_double_conv2.convolution.3.weight": (128, 128, 3, 3, 3), "ex_double_conv2.convolution.3.bias": (128,),
    "t_conv1.weight": (128, 64, 2, 2, 2), "t_conv1.bias": (64,),
    "ex_conv1_1.convolution.0.weight": (64, 128, 3, 3, 3), "ex_conv1_1.convolution.0.bias": (64,),
    "one_by_one.weight": (1, 64, 1, 1, 1), "one_by_one.bias": (1,),
    "one_by_one_2.weight": (1, 1, 1, 1, 1), "one_by_one_2.bias": (1,),
}

N_PARAMS = 7_077_251  # README screenshot "7.077 M"; SURVEY.md §4


def random_state_dict(seed: int = 0, bias_std: float = 0.0) -> dict:
    """Kaiming fan_in normal weights like networks.init_weights('kaiming') (networks.py:88-119); biases are zero
    as in the reference's init unless bias_std > 0 (used by the fixtures so that bias handling is exercised)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in STATE_DICT_SHAPES.items():
        if k.endswith("bias"):
            sd[k] = torch.randn(shape, generator=g) * bias_std if bias_std > 0 else torch.zeros(shape)
        else:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]   # torch fan_in: size(1) * receptive field
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def _as_stored_fp16(t, on):
    """Value rounded to fp16 with a straight-through gradient: what a tensor the B200 path keeps in fp16 looks like.
    Only used to make gradient comparisons meaningful (see unet_deconv_gradients); never by the reference path."""
    return t + (t.half().float() - t).detach() if on else t


def _forced(t, stored, key):
    """Replace the VALUE of t by stored[key] (straight-through gradient): differentiate at a given linearisation
    point — the activations another implementation actually produced (see unet_deconv_gradients)."""
    if stored is None or key not in stored:
        return t
    return t + (stored[key].to(t.dtype) - t).detach()


def _conv_in_relu(x, sd, prefix, fp16_storage=False, stored=None, round_out=True):
    x = F.conv3d(x, _as_stored_fp16(sd[prefix + ".weight"], fp16_storage), sd[prefix + ".bias"], stride=1, padding=1)
    x = _forced(_as_stored_fp16(x, fp16_storage), stored, prefix)
    x = F.instance_norm(x, eps=1e-5)          # InstanceNorm3d(affine=False, track_running_stats=False)
    return _as_stored_fp16(F.relu(x), fp16_storage and round_out)


def _pool(x, fp16_storage):
    """MaxPool3d(2) (networks.py:491,494).  In storage mode x is the UNROUNDED activation: the B200 path pools the
    fp32 values and rounds the maximum, so the arg-max (= where the gradient goes) is taken before rounding —
    rounding first would create ties; the pooled VALUE is the same either way (rounding is monotonic)."""
    return _as_stored_fp16(F.max_pool3d(x, 2), fp16_storage)


def state_dict_checksum(sd: dict):
    return [float(sum(v.double().sum() for v in sd.values())), float(sum(v.double().abs().sum() for v in sd.values()))]


def unet_deconv_forward(x: torch.Tensor, sd: dict, taps: dict | None = None, grad: bool = False,
                        fp16_storage: bool = False, stored: dict | None = None) -> torch.Tensor:
    """x: float32 (N,1,D,H,W) with D,H,W % 4 == 0 -> float32 (N,1,D,H,W) in (0,1).  networks.py:512-538.
    grad=True records the autograd tape (unet_deconv_gradients); fp16_storage rounds every stored activation and
    every tensor-core weight to fp16 like the B200 path does (gradient comparisons only)."""
    def _conv_in_relu(x, sd, prefix, round_out=True):       # the module-level layer with the storage mode bound
        return globals()["_conv_in_relu"](x, sd, prefix, fp16_storage, stored, round_out)
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t
    with torch.set_grad_enabled(grad):
        c1 = _conv_in_relu(x, sd, "double_conv1.convolution.0")
        c1 = _conv_in_relu(c1, sd, "double_conv1.convolution.3", round_out=False)
        p1 = _pool(c1, fp16_storage)
        c1 = tap("conv1", _as_stored_fp16(c1, fp16_storage))
        c2 = _conv_in_relu(p1, sd, "double_conv2.convolution.0")
        c2 = _conv_in_relu(c2, sd, "double_conv2.convolution.3", round_out=False)
        p2 = _pool(c2, fp16_storage)
        c2 = tap("conv2", _as_stored_fp16(c2, fp16_storage))
        b = _conv_in_relu(p2, sd, "bottom_layer.convolution.0")
        b = _conv_in_relu(b, sd, "bottom_layer.convolution.3")
        b = tap("bottom", _conv_in_relu(b, sd, "bottom_layer.convolution.6"))
        t2 = tap("t_conv2", F.conv_transpose3d(b, _as_stored_fp16(sd["t_conv2.weight"], fp16_storage),
                                               sd["t_conv2.bias"], stride=2))
        t2 = _forced(_as_stored_fp16(t2, fp16_storage), stored, "t_conv2")
        e2 = _conv_in_relu(torch.cat([c2, t2], 1), sd, "ex_double_conv2.convolution.0")
        e2 = tap("ex_conv2", _conv_in_relu(e2, sd, "ex_double_conv2.convolution.3"))
        t1 = _as_stored_fp16(F.conv_transpose3d(e2, _as_stored_fp16(sd["t_conv1.weight"], fp16_storage),
                                                sd["t_conv1.bias"], stride=2), fp16_storage)
        t1 = _forced(t1, stored, "t_conv1")
        e1 = tap("ex_conv1", _conv_in_relu(torch.cat([c1, t1], 1), sd, "ex_conv1_1.convolution.0"))
        o = F.conv3d(e1, sd["one_by_one.weight"], sd["one_by_one.bias"])
        o = F.conv3d(o, sd["one_by_one_2.weight"], sd["one_by_one_2.bias"])
        return torch.sigmoid(o)


def unet_deconv_gradients(x: torch.Tensor, sd: dict, dout: torch.Tensor, fp16_storage: bool = False,
                          stored: dict | None = None):
    """Output and the gradient of sum(output * dout) w.r.t. every state_dict tensor — what autograd gives the
    reference module when backward_G (axial_to_lateral_gan_apollo_model.py:255-283) hands it dL/d fake = dout.

    fp16_storage: the gradient of a ReLU network is discontinuous in its activations — rounding a conv output to
    fp16 (or computing it in TF32, as the reference does on a GPU) flips the ReLU mask of the ~1e-4 of the elements
    that sit within rounding distance of zero, and every flip moves one gradient element by its full value: a
    relative L2 change of sqrt(flipped fraction) ~ 1-2 % PER LAYER, 7-8 % at the first layer of this network
    (measured with this very function).  Emulating the rounding is not enough either: two fp16 pipelines that
    differ in the last bit of an fp32 accumulation drift apart by about one fp16 ulp within a few layers, which
    flips a different set of masks.  To check the backward KERNELS tightly, the tests therefore differentiate at
    the GPU's own linearisation point: `stored` maps a conv layer's state_dict prefix (and "t_conv1"/"t_conv2") to
    the raw output the GPU forward actually stored; its value replaces the oracle's (straight-through gradient)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    y = unet_deconv_forward(x, leaves, grad=True, fp16_storage=fp16_storage, stored=stored)
    y.backward(dout)
    return y.detach(), {k: v.grad for k, v in leaves.items()}
