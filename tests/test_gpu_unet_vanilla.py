"""GPU: the sibling generator unet_vanilla (reference models/networks.py:540-608) on the same kernels,
against the fixture recorded from the reference module and against the oracle.  Same bar as unet_deconv:
max-abs <= 2e-2, PSNR >= 50 dB."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import unet_vanilla as uv

pytestmark = pytest.mark.gpu


def _check(got, ref):
    err = (got - ref).abs().max().item()
    mse = float(((got.double() - ref.double()) ** 2).mean())
    p = 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)
    assert err <= 2e-2 and p >= 50.0, "max-abs %.4g, PSNR %.2f dB" % (err, p)
    return err, p


@pytest.fixture(scope="module")
def net(cuda):
    from neuroclear_b200 import networks
    with redirect_stdout(io.StringIO()):
        n = networks.define_G(1, 1, 64, "unet_vanilla", "instance", False, "kaiming", 0.02, [0], dimension=3)
    assert {k: tuple(v.shape) for k, v in n.module.state_dict().items()} == uv.state_dict_shapes()
    n.module.load_state_dict(uv.random_state_dict(seed=3, bias_std=0.1))
    return n.eval()


def test_reference_fixture(net, cuda):
    f = np.load(os.path.join(GOLDEN, "unet_vanilla_small.npz"))
    for name in "ab":
        with torch.no_grad():
            y = net(torch.from_numpy(f["x_" + name]).to(cuda))
        assert y.shape == f["y_" + name].shape and y.dtype == torch.float32
        print(name, _check(y.cpu(), torch.from_numpy(f["y_" + name])))


def test_larger_cube_vs_oracle_and_error_behaviour(net, cuda):
    from neuroclear_b200._lib import NeuroclearError
    x = torch.rand((1, 1, 64, 72, 56), generator=torch.Generator().manual_seed(5)) ** 2
    ref = uv.unet_vanilla_forward(x, uv.random_state_dict(seed=3, bias_std=0.1))
    with torch.no_grad():
        y = net(x.to(cuda))
        assert torch.equal(y, net(x.to(cuda)))                        # deterministic
        with pytest.raises(NeuroclearError):
            net(torch.zeros((1, 1, 20, 16, 16), device=cuda))         # not divisible by 8
    print(_check(y.cpu(), ref))
    with pytest.raises(NotImplementedError):
        net(x.to(cuda))                                               # training is not on the B200 path
