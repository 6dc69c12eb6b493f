"""GPU: whole-cube Unet_deconv parity with the reference fp32 path (the oracle, pinned to the reference module).

Tolerance is BASELINE.json's: max-abs error <= 2e-2 and PSNR >= 50 dB on the [0,1] output."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import unet as ounet

pytestmark = pytest.mark.gpu

MAX_ABS, MIN_PSNR = 2e-2, 50.0


def psnr(a, b):
    mse = float(((a.double() - b.double()) ** 2).mean())
    return 99.0 if mse == 0 else 10 * np.log10(1.0 / mse)


def _check(got, ref):
    err = (got - ref).abs().max().item()
    p = psnr(got, ref)
    assert err <= MAX_ABS and p >= MIN_PSNR, "max-abs %.4g, PSNR %.2f dB" % (err, p)
    return err, p


@pytest.fixture(scope="module")
def net(cuda):
    from neuroclear_b200 import networks
    with redirect_stdout(io.StringIO()):
        n = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [0], dimension=3)
    assert isinstance(n, torch.nn.DataParallel) and hasattr(n, "module")    # base_model.py:158-160 contract
    n.module.load_state_dict(ounet.random_state_dict(seed=0, bias_std=0.1))
    n.eval()
    return n


def test_golden_small_cubes(net, cuda):
    """Inputs and outputs recorded from the reference module (tests/golden/unet_small.npz)."""
    f = np.load(os.path.join(GOLDEN, "unet_small.npz"))
    for name in "ab":
        with torch.no_grad():
            y = net(torch.from_numpy(f["x_" + name]).to(cuda))
        assert y.shape == f["y_" + name].shape and y.dtype == torch.float32
        _check(y.cpu(), torch.from_numpy(f["y_" + name]))


@pytest.mark.parametrize("shape", [(2, 1, 32, 40, 48), (1, 1, 52, 36, 44)])
def test_ragged_shapes_vs_oracle(net, cuda, shape):
    g = torch.Generator().manual_seed(shape[2])
    x = torch.rand(shape, generator=g) ** 4          # fluorescence-like: mostly dark, sparse bright
    ref = ounet.unet_deconv_forward(x, ounet.random_state_dict(seed=0, bias_std=0.1))
    with torch.no_grad():
        y = net(x.to(cuda))
    _check(y.cpu(), ref)


def test_full_size_cube_140(net, cuda):
    """One inference cube of the headline configuration (dice 120 + 2 x border 10 = 140^3)."""
    g = torch.Generator().manual_seed(140)
    x = torch.rand((1, 1, 140, 140, 140), generator=g)
    ref = ounet.unet_deconv_forward(x, ounet.random_state_dict(seed=0, bias_std=0.1))
    with torch.no_grad():
        y = net(x.to(cuda))
    err, p = _check(y.cpu(), ref)
    print("140^3 cube: max-abs %.4g PSNR %.2f dB" % (err, p))


def test_deterministic_and_batch_invariant(net, cuda):
    """InstanceNorm statistics use fixed-order reductions: bitwise repeatable, and a cube's result does not
    depend on which other cubes share its batch."""
    g = torch.Generator().manual_seed(8)
    x = torch.rand((3, 1, 24, 24, 24), generator=g).to(cuda)
    with torch.no_grad():
        a, b = net(x), net(x)
        single = net(x[1:2])
    assert torch.equal(a, b) and torch.equal(a[1:2], single)


def test_weight_updates_invalidate_packed_cache(net, cuda):
    x = torch.rand((1, 1, 16, 16, 16), generator=torch.Generator().manual_seed(1)).to(cuda)
    m = net.module
    with torch.no_grad():
        y0 = net(x)
        saved = m.one_by_one_2.bias.clone()
        m.one_by_one_2.bias.add_(1.0)                 # what an optimiser step does (in-place, bumps _version)
        y1 = net(x)
        m.one_by_one_2.bias.copy_(saved)
        y2 = net(x)
    assert not torch.equal(y0, y1) and torch.equal(y0, y2)
    # writes through .data do NOT bump _version (init_weights, EMA, p.data.copy_ loaders): the value checksum sees them
    with torch.no_grad():
        v0 = m.one_by_one_2.bias._version
        m.one_by_one_2.bias.data.add_(1.0)
        assert m.one_by_one_2.bias._version == v0
        y3 = net(x)
        m.one_by_one_2.bias.data.copy_(saved)
        y4 = net(x)
        w = m.double_conv2.convolution[3].weight
        w.data.neg_()                                 # a pure sign flip keeps every norm: the bit-pattern sum moves
        y5 = net(x)
        w.data.neg_()
    assert not torch.equal(y0, y3) and torch.equal(y0, y4) and not torch.equal(y0, y5)
    sd = {k: v.cpu() for k, v in m.state_dict().items()}                      # save_networks round trip
    m.cpu()
    m.cuda(0)
    m.load_state_dict(sd)
    with torch.no_grad():
        assert torch.equal(net(x), y0)


def test_rejects_bad_inputs(net, cuda):
    from neuroclear_b200._lib import NeuroclearError
    with torch.no_grad():
        with pytest.raises(NeuroclearError):
            net(torch.zeros((1, 1, 18, 16, 16), device=cuda))     # not divisible by 4 (the reference's cat fails too)
        with pytest.raises(NeuroclearError):
            net.module(torch.zeros((1, 1, 16, 16, 16)))           # CPU tensor: no fallback


def test_whole_cube_entry_is_bit_identical_and_graph_capturable(cuda):
    """nc_unet_deconv_infer_cube: one library call for the whole network == the per-layer path bit for
    bit; and, being allocation- and sync-free, it replays from a CUDA graph."""
    from neuroclear_b200.unet_engine import UnetDeconvEngine
    eng = UnetDeconvEngine(cuda)
    eng.load_state_dict(ounet.random_state_dict(seed=0, bias_std=0.1))
    x = torch.rand((2, 24, 40, 32), generator=torch.Generator().manual_seed(3)).to(cuda)
    want = eng.forward(x, crop=4).clone()
    got = eng.forward_cube(x, crop=4)
    assert got.shape == (2, 16, 32, 24) and torch.equal(got, want)
    # CUDA graph: capture once, replay on new input written into the same buffer
    xs = x.clone()
    ys = torch.empty_like(got)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        eng.forward_cube(xs, crop=4, out=ys)                 # warm-up on the capture stream (workspace, attributes)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        eng.forward_cube(xs, crop=4, out=ys)
    x2 = torch.rand((2, 24, 40, 32), generator=torch.Generator().manual_seed(4)).to(cuda)
    xs.copy_(x2)
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(ys, eng.forward(x2, crop=4))
