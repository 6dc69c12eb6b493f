"""NLayerDiscriminator (2-D PatchGAN) + lsgan GANLoss restated with torch.nn.functional on the CPU in fp32.
Test infrastructure — see oracle/__init__.py.  Follows models/networks.py:1009-1067 (dimension=2, InstanceNorm,
use_bias=True, n_layers=3) and networks.py:252-319; pinned against the reference module in oracle/make_golden.py.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

STATE_DICT_SHAPES = {
    "model.0.weight": (64, 1, 4, 4), "model.0.bias": (64,),
    "model.2.weight": (128, 64, 4, 4), "model.2.bias": (128,),
    "model.5.weight": (256, 128, 4, 4), "model.5.bias": (256,),
    "model.8.weight": (512, 256, 4, 4), "model.8.bias": (512,),
    "model.11.weight": (1, 512, 4, 4), "model.11.bias": (1,),
}
N_PARAMS = 2_762_689  # README screenshot "2.763 M"


def random_state_dict(seed: int = 0, bias_std: float = 0.05) -> dict:
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in STATE_DICT_SHAPES.items():
        if k.endswith("bias"):
            sd[k] = torch.randn(shape, generator=g) * bias_std
        else:
            sd[k] = torch.randn(shape, generator=g) * (2.0 / (shape[1] * 16)) ** 0.5
    return sd


def discriminator_forward(x: torch.Tensor, sd: dict) -> torch.Tensor:
    """x (N,1,S,S) -> (N,1,S/8-2,S/8-2) prediction map; differentiable (used for gradient parity too)."""
    y = F.leaky_relu(F.conv2d(x, sd["model.0.weight"], sd["model.0.bias"], stride=2, padding=1), 0.2)
    for idx, stride in ((2, 2), (5, 2), (8, 1)):
        y = F.conv2d(y, sd["model.%d.weight" % idx], sd["model.%d.bias" % idx], stride=stride, padding=1)
        y = F.leaky_relu(F.instance_norm(y, eps=1e-5), 0.2)
    return F.conv2d(y, sd["model.11.weight"], sd["model.11.bias"], stride=1, padding=1)


def lsgan_loss(pred: torch.Tensor, target_is_real: bool) -> torch.Tensor:
    label = torch.tensor(1.0 if target_is_real else 0.0)
    return F.mse_loss(pred, label.expand_as(pred))
