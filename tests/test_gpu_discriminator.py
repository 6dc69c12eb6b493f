"""GPU: the 2-D PatchGAN discriminator path (define_D 'basic', GANLoss lsgan, L1) — forward and backward on the
hand-written kernels — against fixtures recorded from the REFERENCE module and against the CPU oracle's autograd."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import discriminator as odisc

pytestmark = pytest.mark.gpu

# a bias in front of InstanceNorm(affine=False) has an exactly-zero gradient (the mean subtraction cancels it)
NOISE_ONLY = {"model.2.bias", "model.5.bias", "model.8.bias"}


def _net(cuda, sd):
    from neuroclear_b200 import discriminator
    with redirect_stdout(io.StringIO()):
        net = discriminator.define_D(1, 64, "basic", norm="instance", use_sigmoid=False, init_type="kaiming",
                                     init_gain=0.02, gpu_ids=[0], dimension=2)
    assert list(net.module.state_dict().keys()) == list(odisc.STATE_DICT_SHAPES.keys())
    assert sum(p.numel() for p in net.parameters()) == odisc.N_PARAMS == 2_762_689
    net.module.load_state_dict(sd)
    return net


def test_golden_forward_backward(cuda):
    """Prediction map, lsgan loss, d/d(input) and d/d(parameters) recorded from the reference (44 x 36 image)."""
    from neuroclear_b200 import discriminator
    f = np.load(os.path.join(GOLDEN, "discriminator_44x36.npz"))
    net = _net(cuda, odisc.random_state_dict(seed=0))
    crit = discriminator.GANLoss("lsgan").to(cuda)
    x = torch.from_numpy(f["x"]).to(cuda).requires_grad_(True)
    pred = net(x)
    loss = crit(pred, True) * 0.5 + crit(pred, False) * 0.25
    loss.backward()
    assert pred.shape == f["pred"].shape
    assert np.abs(pred.detach().cpu().numpy() - f["pred"]).max() <= 1e-4 * max(1.0, np.abs(f["pred"]).max())
    assert abs(loss.item() - float(f["loss"])) <= 1e-5 * max(1.0, abs(float(f["loss"])))
    assert np.abs(x.grad.cpu().numpy() - f["dx"]).max() <= 1e-4 * max(1e-3, np.abs(f["dx"]).max())
    for k, prm in net.module.named_parameters():
        g = prm.grad.cpu().numpy()
        ref_s, ref_n = f["dsample_" + k], float(f["dnorm_" + k])
        if k in NOISE_ONLY:       # exactly zero in real arithmetic; the reference's value is rounding noise
            assert np.abs(g).max() <= 1e-6 and ref_n <= 1e-5, k
            continue
        assert abs(np.linalg.norm(g.astype(np.float64)) - ref_n) <= 1e-4 * ref_n, k
        assert np.abs(g.reshape(-1)[::61] - ref_s).max() <= 1e-4 * np.abs(ref_s).max(), k


@pytest.mark.parametrize("size", [(1, 108, 108), (2, 37, 52)])
def test_vs_oracle_autograd(cuda, size):
    """Full D step on apollo-sized inputs: 0.5*(mse(D(real),1) + mse(D(fake),0)).backward() (apollo_model.py:169-196)."""
    from neuroclear_b200 import discriminator
    n, h, w = size
    sd = odisc.random_state_dict(seed=3)
    net = _net(cuda, sd)
    crit = discriminator.GANLoss("lsgan").to(cuda)
    g = torch.Generator().manual_seed(h)
    real, fake = torch.rand((n, 1, h, w), generator=g), torch.rand((n, 1, h, w), generator=g)
    loss = 0.5 * (crit(net(real.to(cuda)), True) + crit(net(fake.to(cuda)), False))
    loss.backward()
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    lo = 0.5 * (odisc.lsgan_loss(odisc.discriminator_forward(real, sdo), True) +
                odisc.lsgan_loss(odisc.discriminator_forward(fake, sdo), False))
    lo.backward()
    assert abs(loss.item() - lo.item()) <= 1e-5 * max(1.0, abs(lo.item()))
    for k, prm in net.module.named_parameters():
        ref = sdo[k].grad
        if k in NOISE_ONLY:
            assert prm.grad.abs().max().item() <= 1e-6 and ref.abs().max().item() <= 1e-6, k
            continue
        assert (prm.grad.cpu() - ref).abs().max().item() <= 2e-4 * ref.abs().max().item(), k


def test_projection_feeds_discriminator_with_gradient(cuda):
    """proj_f of the apollo model (apollo_model.py:316-320): MIP of a cube -> D -> lsgan, gradient back to the cube."""
    from neuroclear_b200 import discriminator
    from neuroclear_b200.projection import Volume
    sd = odisc.random_state_dict(seed=5)
    net = _net(cuda, sd)
    crit = discriminator.GANLoss("lsgan").to(cuda)
    g = torch.Generator().manual_seed(11)
    vol = torch.rand((1, 1, 40, 40, 40), generator=g)
    v = vol.to(cuda).requires_grad_(True)
    np.random.seed(4)
    loss = crit(net(Volume(v, cuda).get_projection(6, 1)), True)
    loss.backward()
    vc = vol.clone().requires_grad_(True)
    np.random.seed(4)
    start = np.random.randint(0, 40 - 6)
    mip = torch.max(vc[:, :, :, start:start + 6, :], 3)[0]
    lo = odisc.lsgan_loss(odisc.discriminator_forward(mip, sd), True)
    lo.backward()
    assert abs(loss.item() - lo.item()) <= 1e-5 * max(1.0, lo.item())
    assert (v.grad.cpu() - vc.grad).abs().max().item() <= 2e-4 * vc.grad.abs().max().item()


def test_l1_loss(cuda):
    from neuroclear_b200 import discriminator
    g = torch.Generator().manual_seed(1)
    a, b = torch.randn((1, 1, 20, 24, 28), generator=g), torch.randn((1, 1, 20, 24, 28), generator=g)
    ad = a.to(cuda).requires_grad_(True)
    loss = discriminator.L1Loss()(ad, b.to(cuda)) * 5.0
    loss.backward()
    ac = a.clone().requires_grad_(True)
    lo = torch.nn.functional.l1_loss(ac, b) * 5.0
    lo.backward()
    assert abs(loss.item() - lo.item()) <= 1e-5 * lo.item() and torch.allclose(ad.grad.cpu(), ac.grad, atol=1e-9)
