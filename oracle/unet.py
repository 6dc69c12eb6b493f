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


def _conv_in_relu(x, sd, prefix):
    x = F.conv3d(x, sd[prefix + ".weight"], sd[prefix + ".bias"], stride=1, padding=1)
    x = F.instance_norm(x, eps=1e-5)          # InstanceNorm3d(affine=False, track_running_stats=False)
    return F.relu(x)


def state_dict_checksum(sd: dict):
    return [float(sum(v.double().sum() for v in sd.values())), float(sum(v.double().abs().sum() for v in sd.values()))]


def unet_deconv_forward(x: torch.Tensor, sd: dict, taps: dict | None = None) -> torch.Tensor:
    """x: float32 (N,1,D,H,W) with D,H,W % 4 == 0 -> float32 (N,1,D,H,W) in (0,1).  networks.py:512-538."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t
        return t
    with torch.no_grad():
        c1 = _conv_in_relu(x, sd, "double_conv1.convolution.0")
        c1 = tap("conv1", _conv_in_relu(c1, sd, "double_conv1.convolution.3"))
        p1 = F.max_pool3d(c1, 2)
        c2 = _conv_in_relu(p1, sd, "double_conv2.convolution.0")
        c2 = tap("conv2", _conv_in_relu(c2, sd, "double_conv2.convolution.3"))
        p2 = F.max_pool3d(c2, 2)
        b = _conv_in_relu(p2, sd, "bottom_layer.convolution.0")
        b = _conv_in_relu(b, sd, "bottom_layer.convolution.3")
        b = tap("bottom", _conv_in_relu(b, sd, "bottom_layer.convolution.6"))
        t2 = tap("t_conv2", F.conv_transpose3d(b, sd["t_conv2.weight"], sd["t_conv2.bias"], stride=2))
        e2 = _conv_in_relu(torch.cat([c2, t2], 1), sd, "ex_double_conv2.convolution.0")
        e2 = tap("ex_conv2", _conv_in_relu(e2, sd, "ex_double_conv2.convolution.3"))
        t1 = F.conv_transpose3d(e2, sd["t_conv1.weight"], sd["t_conv1.bias"], stride=2)
        e1 = tap("ex_conv1", _conv_in_relu(torch.cat([c1, t1], 1), sd, "ex_conv1_1.convolution.0"))
        o = F.conv3d(e1, sd["one_by_one.weight"], sd["one_by_one.bias"])
        o = F.conv3d(o, sd["one_by_one_2.weight"], sd["one_by_one_2.bias"])
        return torch.sigmoid(o)
