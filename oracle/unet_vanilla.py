"""Unet_vanilla forward (reference models/networks.py:540-608) restated with torch.nn.functional on the CPU in fp32.
Test infrastructure — see oracle/__init__.py.  Takes the reference's state_dict unchanged; pinned against the real
reference module in oracle/make_golden.py::golden_unet_vanilla (tests/golden/unet_vanilla_small.npz)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CH = (64, 128, 256, 512)


def state_dict_shapes():
    s = {}
    def dc(name, cin, cout):
        s[name + ".convolution.0.weight"], s[name + ".convolution.0.bias"] = (cout, cin, 3, 3, 3), (cout,)
        s[name + ".convolution.3.weight"], s[name + ".convolution.3.bias"] = (cout, cout, 3, 3, 3), (cout,)
    dc("double_conv1", 1, 64), dc("double_conv2", 64, 128), dc("double_conv3", 128, 256), dc("bottom_layer", 256, 512)
    s["t_conv3.weight"], s["t_conv3.bias"] = (512, 256, 2, 2, 2), (256,)
    dc("ex_double_conv3", 512, 256)
    s["t_conv2.weight"], s["t_conv2.bias"] = (256, 128, 2, 2, 2), (128,)
    dc("ex_double_conv2", 256, 128)
    s["t_conv1.weight"], s["t_conv1.bias"] = (128, 64, 2, 2, 2), (64,)
    dc("ex_conv1_1", 128, 64)
    s["one_by_one.weight"], s["one_by_one.bias"] = (1, 64, 1, 1, 1), (1,)
    return s


def random_state_dict(seed: int = 0, bias_std: float = 0.0) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in state_dict_shapes().items():
        if k.endswith("bias"):
            sd[k] = torch.randn(shape, generator=g) * bias_std if bias_std > 0 else torch.zeros(shape)
        else:
            fan_in = shape[1] * shape[2] * shape[3] * shape[4]
            sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def _double_conv(x, sd, name):
    for i in (0, 3):
        x = F.conv3d(x, sd["%s.convolution.%d.weight" % (name, i)], sd["%s.convolution.%d.bias" % (name, i)], 1, 1)
        x = F.relu(F.instance_norm(x, eps=1e-5))
    return x


def unet_vanilla_forward(x: torch.Tensor, sd: dict) -> torch.Tensor:
    """x: (N,1,D,H,W) float32, D,H,W % 8 == 0 -> (N,1,D,H,W) in (0,1)"""
    with torch.no_grad():
        c1 = _double_conv(x, sd, "double_conv1")
        c2 = _double_conv(F.max_pool3d(c1, 2), sd, "double_conv2")
        c3 = _double_conv(F.max_pool3d(c2, 2), sd, "double_conv3")
        b = _double_conv(F.max_pool3d(c3, 2), sd, "bottom_layer")
        e3 = _double_conv(torch.cat([c3, F.conv_transpose3d(b, sd["t_conv3.weight"], sd["t_conv3.bias"], 2)], 1), sd,
                          "ex_double_conv3")
        e2 = _double_conv(torch.cat([c2, F.conv_transpose3d(e3, sd["t_conv2.weight"], sd["t_conv2.bias"], 2)], 1), sd,
                          "ex_double_conv2")
        e1 = _double_conv(torch.cat([c1, F.conv_transpose3d(e2, sd["t_conv1.weight"], sd["t_conv1.bias"], 2)], 1), sd,
                          "ex_conv1_1")
        return torch.sigmoid(F.conv3d(e1, sd["one_by_one.weight"], sd["one_by_one.bias"]))
