"""GPU: the sibling models of SURVEY.md §8 f4 on the same kernels, against fixtures recorded from the REFERENCE classes
(oracle/make_golden.py::golden_sibling_models): AxialToLateralGANAthenaModel (six per-slice discriminators),
AxialToLateralGANDryopsModel (no G_B / D_B) and the spectral-norm PatchGAN (define_D 'basic_SN').
Tolerances as for the apollo step (tests/test_gpu_apollo_step.py): losses 2 %, updated parameters 2.1 lr, generator
gradient direction cos >= 0.97."""
import io
import os
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
LR, STRIDE = 1e-4, 61
NOISE_ONLY = ("double_conv1.convolution", "double_conv2.convolution", "bottom_layer.convolution",
              "ex_double_conv2.convolution", "ex_conv1_1.convolution")


def sample(t):
    flat = t.detach().float().cpu().reshape(-1)
    return (flat if flat.numel() <= 4096 else flat[::STRIDE]).numpy()


def _opt(**kw):
    o = dict(isTrain=True, gpu_ids=[0], gan_mode="lsgan", randomize_projection_depth=True, projection_depth=10,
             min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
             netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance",
             no_dropout=True, init_type="kaiming", init_gain=0.02, lr=LR, beta1=0.1, direction="AtoB", lambda_A=5.0)
    o.update(kw)
    return Namespace(**o)


def _compare(m, z, gens, ds):
    torch.cuda.synchronize()
    for k, v in m.get_current_losses().items():
        ref = float(z["loss_" + k])
        print("  loss_%-10s %.6f   reference %.6f" % (k, v, ref))
        assert abs(v - ref) <= 2e-2 * max(1.0, abs(ref)), k
    for name in gens:
        for k, p in getattr(m, "net" + name).module.named_parameters():
            assert np.abs(sample(p) - z["after_%s.%s" % (name, k)]).max() <= 2.1 * LR, (name, k)
            if k.endswith(".bias") and k.startswith(NOISE_ONLY):
                continue
            a, b = sample(p.grad).astype(np.float64), z["grad_%s.%s" % (name, k)].astype(np.float64)
            cos = float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))
            assert cos >= 0.97, (name, k, cos)
    for name in ds:
        for k, p in getattr(m, "net" + name).module.named_parameters():
            assert np.abs(sample(p) - z["after_%s.%s" % (name, k)]).max() <= 2.1 * LR, (name, k)


def test_athena_iteration_matches_reference_fixture(cuda):
    from neuroclear_b200.athena_model import AxialToLateralGANAthenaModel, D_NAMES
    from oracle import deeplinear, discriminator, unet
    z = np.load(os.path.join(GOLDEN, "athena_step_32.npz"))
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANAthenaModel(_opt(conversion_plane=["yz", "xy"]), cuda, distributed=False)
    m.netG_A.module.load_state_dict(unet.random_state_dict(seed=21, bias_std=0.05))
    m.netG_B.module.load_state_dict(deeplinear.random_state_dict(seed=22))
    for i, n in enumerate(D_NAMES):
        getattr(m, "net" + n).module.load_state_dict(discriminator.random_state_dict(seed=40 + i))
    m.set_input({"A": torch.from_numpy(z["real"]), "A_paths": "golden"})
    m.optimize_parameters()
    assert list(m.get_current_losses()) == [k[5:] for k in z.files if k.startswith("loss_")]
    _compare(m, z, ("G_A", "G_B"), D_NAMES)


def test_dryops_iteration_matches_reference_fixture(cuda):
    from neuroclear_b200.dryops_model import AxialToLateralGANDryopsModel
    from oracle import discriminator, unet
    z = np.load(os.path.join(GOLDEN, "dryops_step_32.npz"))
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANDryopsModel(_opt(), cuda, distributed=False)
    m.netG_A.module.load_state_dict(unet.random_state_dict(seed=21, bias_std=0.05))
    for i, n in enumerate(["D_A_axial", "D_A_lateral"]):
        getattr(m, "net" + n).module.load_state_dict(discriminator.random_state_dict(seed=30 + i))
    np.random.seed(3)
    m.set_input({"A": torch.from_numpy(z["real"]), "A_paths": "golden"})
    assert m.projection_depth == int(z["depth"])
    m.optimize_parameters()
    assert not hasattr(m, "netG_B") and set(m.get_current_losses()) == {k[5:] for k in z.files if k.startswith("loss_")}
    _compare(m, z, ("G_A",), ("D_A_axial", "D_A_lateral"))


def test_spectral_norm_patchgan_matches_reference_fixture(cuda):
    from neuroclear_b200 import discriminator
    z = np.load(os.path.join(GOLDEN, "discriminator_sn_44x36.npz"))
    torch.manual_seed(77)
    with redirect_stdout(io.StringIO()):
        net = discriminator.define_D(1, 64, "basic_SN", norm="instance", use_sigmoid=False, init_type="kaiming",
                                     init_gain=0.02, gpu_ids=[], dimension=2)
    for k, v in net.state_dict().items():                      # same initial state as the reference's (seed 77)
        assert float(v.double().sum()) == float(z["sdsum_" + k]), k
    net = net.to(cuda).train()
    x = torch.from_numpy(z["x"]).to(cuda).requires_grad_(True)
    pred = net(x)
    crit = discriminator.GANLoss("lsgan")
    loss = crit(pred, True) * 0.5 + crit(pred, False) * 0.25
    loss.backward()
    assert np.abs(pred.detach().cpu().numpy() - z["pred"]).max() <= 2e-5 * max(1.0, np.abs(z["pred"]).max())
    assert abs(float(loss) - float(z["loss"])) <= 1e-5 * max(1.0, abs(float(z["loss"])))
    scale = np.abs(z["dx"]).max()
    assert np.abs(x.grad.cpu().numpy() - z["dx"]).max() <= 1e-4 * scale
    for k, p in net.named_parameters():
        ref = z["grad_" + k]
        got = p.grad.detach().cpu().numpy().reshape(-1)[::7]
        assert np.abs(got - ref).max() <= 1e-3 * max(np.abs(ref).max(), 1e-12), k
    for k, v in net.state_dict().items():
        if k.endswith(("_u", "_v")):
            assert np.abs(v.cpu().numpy() - z["after_" + k]).max() <= 1e-5, k
