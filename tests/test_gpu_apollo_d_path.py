"""GPU: the projection / discriminator path of AxialToLateralGANApolloModel against a fixture recorded from the
REFERENCE model (oracle/make_golden.py::golden_apollo_discriminator_path): same volumes, same D weights, same
np.random seed -> same six discriminator losses, same gradients, same Adam update, same generator-side losses and
gradients w.r.t. fake / rec."""
import io
import os
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import discriminator as odisc

pytestmark = pytest.mark.gpu

D_NAMES = ["D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"]
NOISE_ONLY = ("model.2.bias", "model.5.bias", "model.8.bias")   # exactly-zero gradients (bias in front of IN)


def _opt():
    return Namespace(gan_mode="lsgan", randomize_projection_depth=True, projection_depth=10, min_projection_depth=2,
                     lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ndf=64, netD="basic", n_layers_D=3,
                     norm="instance", init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, lambda_A=5.0)


@pytest.fixture(params=[True, False], ids=["batched", "one-pass-per-slice"])
def path(cuda, request):
    """both forms of the discriminator passes (ApolloDiscriminatorPath.batched) against the same reference fixture"""
    from neuroclear_b200.apollo_d_path import ApolloDiscriminatorPath
    with redirect_stdout(io.StringIO()):
        p = ApolloDiscriminatorPath(_opt(), cuda)
    p.batched = request.param
    for i, name in enumerate(D_NAMES):
        getattr(p, "net" + name).module.load_state_dict(odisc.random_state_dict(seed=10 + i))
    return p


def _vols(cuda):
    f = np.load(os.path.join(GOLDEN, "apollo_d_path_32.npz"))
    return f, [torch.from_numpy(f[k]).to(cuda) for k in ("real", "fake", "rec")]


def test_projection_depth_draw(path):
    np.random.seed(0)
    f = np.load(os.path.join(GOLDEN, "apollo_d_path_32.npz"))
    assert path.draw_projection_depth() == int(f["depth"])          # set_input's draw (apollo_model.py:160)


def test_discriminator_step_matches_reference(path, cuda):
    f, (real, fake, rec) = _vols(cuda)
    path.projection_depth = int(f["depth"])
    np.random.seed(7)
    path.optimize_D(real, fake, rec)
    for k in ("D_A_lateral", "D_A_axial", "D_B_lateral", "D_B_axial"):
        ref = float(f["loss_" + k])
        assert abs(float(getattr(path, "loss_" + k)) - ref) <= 1e-5 * max(1.0, abs(ref)), k
    for name in D_NAMES:
        for k, prm in getattr(path, "net" + name).module.named_parameters():
            ref_g = f["grad_%s.%s" % (name, k)]
            got_g = prm.grad.cpu().numpy().reshape(-1)[::61]
            if k in NOISE_ONLY:
                assert np.abs(got_g).max() <= 1e-6
            else:
                assert np.abs(got_g - ref_g).max() <= 2e-4 * np.abs(ref_g).max(), (name, k)
                # Adam: first step moves every weight by ~lr * sign(g); compare the updated parameters
                ref_p = f["after_%s.%s" % (name, k)]
                got_p = prm.detach().cpu().numpy().reshape(-1)[::61]
                moved = np.abs(ref_g) > 1e-3 * np.abs(ref_g).max()           # away from sign flips of tiny gradients
                assert np.abs(got_p - ref_p)[moved].max() <= 2e-6, (name, k)


def test_generator_side_losses_and_gradients(path, cuda):
    f, (real, fake, rec) = _vols(cuda)
    path.projection_depth = int(f["depth"])
    fake = fake.clone().requires_grad_(True)
    rec = rec.clone().requires_grad_(True)
    np.random.seed(9)
    path.generator_losses(real, fake, rec).backward()
    for k in ("G_A", "G_A_lateral", "G_A_axial", "G_B", "G_B_lateral", "G_B_axial", "cycle"):
        ref = float(f["loss_" + k])
        assert abs(float(getattr(path, "loss_" + k)) - ref) <= 1e-5 * max(1.0, abs(ref)), k
    for got, key in ((fake.grad, "dfake"), (rec.grad, "drec")):
        ref = f[key]
        assert np.abs(got.cpu().numpy() - ref).max() <= 2e-4 * np.abs(ref).max(), key
    assert all(p.grad is None for d in path.discriminators() for p in d.parameters())   # Ds were frozen


def test_fused_adam_matches_torch(cuda):
    from neuroclear_b200.apollo_d_path import FusedAdam
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(10007, generator=g)
    a = torch.nn.Parameter(p0.clone().to(cuda))
    b = torch.nn.Parameter(p0.clone())
    mine, ref = FusedAdam([a], lr=1e-4, betas=(0.1, 0.999)), torch.optim.Adam([b], lr=1e-4, betas=(0.1, 0.999))
    for _ in range(5):
        grad = torch.randn(10007, generator=g)
        a.grad, b.grad = grad.to(cuda), grad.clone()
        mine.step()
        ref.step()
    assert (a.detach().cpu() - b.detach()).abs().max().item() <= 1e-7
