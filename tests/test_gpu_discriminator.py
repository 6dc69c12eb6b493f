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


def _conv_case(cuda, n, cin, cout, h, w, stride, seed):
    """nc_conv2d_k4_fwd / _dgrad / _wgrad alone against torch's fp32 CPU conv2d and its autograd; relative max errors."""
    from neuroclear_b200._lib import call, ptr, stream_ptr, f32
    g = torch.Generator().manual_seed(seed)
    x = torch.randn((n, cin, h, w), generator=g)
    wt = torch.randn((cout, cin, 4, 4), generator=g) * 0.1
    b = torch.randn((cout,), generator=g)
    xc, wc, bc = x.clone().requires_grad_(True), wt.clone().requires_grad_(True), b.clone().requires_grad_(True)
    yc = torch.nn.functional.conv2d(xc, wc, bc, stride=stride, padding=1)
    dy = torch.randn(yc.shape, generator=g)
    yc.backward(dy)
    xd, wd, bd, dyd = x.to(cuda), wt.to(cuda), b.to(cuda), dy.to(cuda)
    y = torch.empty(yc.shape, device=cuda)
    dx, dw, db = torch.empty_like(xd), torch.empty_like(wd), torch.empty_like(bd)
    s = stream_ptr()
    call("nc_conv2d_k4_fwd", ptr(xd), ptr(wd), ptr(bd), n, cin, h, w, cout, stride, f32(1.0), ptr(y), s)
    call("nc_conv2d_k4_dgrad", ptr(dyd), ptr(wd), n, cin, h, w, cout, stride, ptr(dx), s)
    call("nc_conv2d_k4_wgrad", ptr(xd), ptr(dyd), n, cin, h, w, cout, stride, ptr(dw), ptr(db), s)
    rel = lambda a, ref: (a.cpu() - ref).abs().max().item() / ref.abs().max().item()
    return (rel(y, yc.detach()), rel(dx, xc.grad), rel(dw, wc.grad), rel(db, bc.grad)), (y, dx, dw)


# (n, cin, cout, h, w, stride): the five layers of the 'basic' PatchGAN incl. odd image sizes (parity classes of
# different size in the stride-2 data gradient), the Cout = 1 / Cin = 1 kernels and channel counts off the tile size
CONV_CASES = [(2, 1, 64, 37, 52, 2), (1, 64, 128, 27, 23, 2), (2, 128, 256, 13, 14, 2), (1, 256, 512, 13, 13, 1),
              (3, 512, 1, 12, 9, 1), (1, 24, 40, 19, 21, 2), (1, 70, 6, 9, 10, 1), (1, 1, 8, 9, 11, 1),
              (2, 16, 1, 10, 9, 2), (1, 1, 1, 6, 7, 2)]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_k4_kernels_vs_torch(cuda, case):
    errs, _ = _conv_case(cuda, *case, seed=sum(case))
    assert max(errs) <= 2e-5, (case, errs)      # 3xTF32 with chunk-wise fp32 accumulation: measured <= 3e-6


def test_conv2d_k4_cluster_size_only_changes_summation_order(cuda, lib):
    """The split of the reduction over a thread-block cluster (debug hook) must not change the result beyond fp32
    summation order — long single-CTA reductions were 1e-3..3e-2 off before the chunk-wise accumulation."""
    case = (2, 128, 256, 27, 27, 2)
    try:
        outs = []
        for c in (1, 2, 4, 8, 0):
            lib.nc_debug_set_disc_cluster(c)
            errs, res = _conv_case(cuda, *case, seed=5)
            assert max(errs) <= 2e-5, (c, errs)
            outs.append(res)
    finally:
        lib.nc_debug_set_disc_cluster(0)
    for res in outs[1:]:
        for a, b in zip(res, outs[0]):
            assert (a - b).abs().max().item() <= 1e-5 * b.abs().max().item()
