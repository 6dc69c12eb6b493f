"""DeepLinearGenerator (reference models/networks.py:893-917) restated with torch.nn.functional on the CPU in fp32.
Test infrastructure — see oracle/__init__.py.  Pinned against the reference module in oracle/make_golden.py."""
from __future__ import annotations

import torch
import torch.nn.functional as F

STATE_DICT_SHAPES = {
    "first_layer.weight": (64, 1, 7, 7, 7), "feature_block.0.weight": (64, 64, 5, 5, 5),
    "feature_block.1.weight": (64, 64, 3, 3, 3), "feature_block.2.weight": (32, 64, 1, 1, 1),
    "feature_block.3.weight": (16, 32, 1, 1, 1), "final_layer.weight": (1, 16, 1, 1, 1),
}
N_PARAMS = 647_120  # README screenshot "0.647 M"; SURVEY.md §2c


def random_state_dict(seed: int = 0) -> dict:
    """Kaiming fan_in normal weights like networks.init_weights('kaiming') (networks.py:88-119)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, shape in STATE_DICT_SHAPES.items():
        fan_in = shape[1] * shape[2] * shape[3] * shape[4]
        sd[k] = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
    return sd


def deep_linear_forward(x: torch.Tensor, sd: dict) -> torch.Tensor:
    """x: float32 (N,1,D,H,W) -> float32 (N,1,D,H,W).  networks.py:913-917."""
    h = F.conv3d(x, sd["first_layer.weight"], padding=3)
    h = F.conv3d(h, sd["feature_block.0.weight"], padding=2)
    h = F.conv3d(h, sd["feature_block.1.weight"], padding=1)
    h = F.conv3d(h, sd["feature_block.2.weight"])
    h = F.conv3d(h, sd["feature_block.3.weight"])
    return F.conv3d(h, sd["final_layer.weight"])


def deep_linear_gradients(x: torch.Tensor, sd: dict, dout: torch.Tensor):
    """Output, d/dx and the gradient of sum(output * dout) w.r.t. every weight."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    xl = x.detach().clone().requires_grad_(True)
    y = deep_linear_forward(xl, leaves)
    y.backward(dout)
    return y.detach(), xl.grad, {k: v.grad for k, v in leaves.items()}
