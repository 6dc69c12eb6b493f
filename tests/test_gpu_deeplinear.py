"""DeepLinearGenerator (G_B of the apollo model) on the GPU, forward and backward, against the fixture recorded
from the REFERENCE module (tests/golden/deeplinear_grad.npz) and the oracle's autograd on other shapes.  The network
is linear (no ReLU masks), so gradients are compared tightly: bf16 gradient tensors -> 2e-2 relative L2."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
STRIDE = 61


def rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-30))


def _engine(sd):
    from neuroclear_b200.deeplinear_engine import DeepLinearEngine
    eng = DeepLinearEngine("cuda")
    eng.load_state_dict(sd)
    return eng


def test_matches_reference_fixture():
    from oracle import deeplinear
    z = np.load(os.path.join(GOLD, "deeplinear_grad.npz"))
    sd = deeplinear.random_state_dict(seed=2)
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(z["w_checksum"][0])) < 1e-9
    eng = _engine(sd)
    y = eng.forward(torch.from_numpy(z["x"]).cuda()[:, 0].contiguous())
    y_ref = torch.from_numpy(z["y"])[:, 0]
    assert float((y.cpu() - y_ref).abs().max()) <= 5e-3 * float(y_ref.abs().max())
    dx, grads = eng.backward(torch.from_numpy(z["dout"]).cuda()[:, 0].contiguous())
    errs = {"dx": rel_l2(dx, torch.from_numpy(z["dx"])[:, 0])}
    for k, shape in deeplinear.STATE_DICT_SHAPES.items():
        assert tuple(grads[k].shape) == shape, k
        errs[k] = rel_l2(grads[k].float().cpu().reshape(-1)[::STRIDE], torch.from_numpy(z["gsample_" + k]))
    print("\n" + "\n".join("  %-28s rel L2 %.4f" % kv for kv in errs.items()))
    assert max(errs.values()) <= 2e-2, errs


@pytest.mark.parametrize("shape", [(2, 9, 17, 23), (1, 16, 8, 40)])
def test_matches_oracle(shape):
    from oracle import deeplinear
    sd = deeplinear.random_state_dict(seed=7)
    g = torch.Generator().manual_seed(3)
    x = torch.rand((shape[0], 1) + shape[1:], generator=g)
    dout = torch.randn(x.shape, generator=g) * 1e-5
    y_ref, dx_ref, g_ref = deeplinear.deep_linear_gradients(x, sd, dout)
    eng = _engine(sd)
    y = eng.forward(x.cuda()[:, 0].contiguous())
    assert float((y.cpu() - y_ref[:, 0]).abs().max()) <= 5e-3 * float(y_ref.abs().max())
    dx, grads = eng.backward(dout.cuda()[:, 0].contiguous())
    errs = {"dx": rel_l2(dx, dx_ref[:, 0])}
    errs.update({k: rel_l2(grads[k], g_ref[k]) for k in g_ref})
    print("\n" + "\n".join("  %-28s rel L2 %.4f" % kv for kv in errs.items()))
    assert max(errs.values()) <= 2e-2, errs


def test_module_autograd():
    """define_G('deep_linear_gen'): same state_dict as the reference, gradients to the weights AND the input."""
    import io
    from contextlib import redirect_stdout
    from neuroclear_b200 import networks
    from oracle import deeplinear
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "deep_linear_gen", "instance", False, "kaiming", 0.02, [0], dimension=3)
    mod = net.module
    assert {k: tuple(v.shape) for k, v in mod.state_dict().items()} == deeplinear.STATE_DICT_SHAPES
    sd = deeplinear.random_state_dict(seed=9)
    mod.load_state_dict(sd)
    g = torch.Generator().manual_seed(4)
    x = torch.rand((1, 1, 12, 12, 12), generator=g)
    xg = x.cuda().requires_grad_(True)
    y = net(xg)
    y.square().mean().backward()
    y_ref, dx_ref, g_ref = None, None, None
    xl = x.clone().requires_grad_(True)
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    deeplinear.deep_linear_forward(xl, leaves).square().mean().backward()
    assert rel_l2(xg.grad, xl.grad) <= 2e-2
    for k, p in mod.named_parameters():
        assert rel_l2(p.grad, leaves[k].grad) <= 2e-2, k
    with torch.no_grad():
        y0 = net(x.cuda())
        torch.optim.SGD(mod.parameters(), lr=1e-2).step()
        assert float((net(x.cuda()) - y0).abs().max()) > 1e-5        # packed-weight cache follows the update


def test_im2col_col2im_are_adjoint_and_stencils_match_conv():
    from neuroclear_b200 import _lib as L
    g = torch.Generator().manual_seed(5)
    nb, d, h, w = 2, 3, 9, 11
    x = torch.rand((nb, d, h, w), generator=g).cuda()
    cols = torch.empty((nb, d, h, w, 64), dtype=torch.bfloat16, device="cuda")
    L.call("nc_im2col49", L.ptr(x), nb, d, h, w, 1, L.ptr(cols), L.stream_ptr())
    ref = F.unfold(x.cpu().reshape(nb * d, 1, h, w), 7, padding=3).reshape(nb, d, 49, h, w).permute(0, 1, 3, 4, 2)
    assert torch.equal(cols[..., :49].float().cpu(), ref.bfloat16().float()) and float(cols[..., 49:].abs().max()) == 0
    gcol = (torch.randn((nb, d, h, w, 64), generator=g)).bfloat16().cuda()
    dx = torch.empty((nb, d, h, w), dtype=torch.float32, device="cuda")
    L.call("nc_col2im49", L.ptr(gcol), nb, d, h, w, L.ptr(dx), L.stream_ptr())
    lhs = float((cols[..., :49].double() * gcol[..., :49].double()).sum())          # <im2col(x), g>
    rhs = float((x.double() * dx.double()).sum())                                   # <x, col2im(g)>
    assert abs(lhs - rhs) <= 1e-3 * abs(lhs)
    # 64 -> 1 k3 stencil and its adjoint vs conv3d / conv_transpose3d
    hh = torch.randn((nb, 64, d, h, w), generator=g).half()
    K = torch.randn((64, 27), generator=g)
    out = torch.empty((nb, d, h, w), dtype=torch.float32, device="cuda")
    h_dev, K_dev = hh.permute(0, 2, 3, 4, 1).contiguous().cuda(), K.cuda()
    L.call("nc_stencil64to1_fwd", L.ptr(h_dev), L.ptr(K_dev), nb, d, h, w, L.ptr(out), L.stream_ptr())
    ref = F.conv3d(hh.float(), K.reshape(1, 64, 3, 3, 3), padding=1)[:, 0]
    assert float((out.cpu() - ref).abs().max()) <= 1e-3
    dout = torch.randn((nb, d, h, w), generator=g).cuda()
    dh = torch.empty((nb, d, h, w, 64), dtype=torch.bfloat16, device="cuda")
    L.call("nc_stencil64to1_bwd_data", L.ptr(dout), L.ptr(K_dev), nb, d, h, w, L.ptr(dh), L.stream_ptr())
    ref = F.conv_transpose3d(dout.cpu()[:, None], K.reshape(1, 64, 3, 3, 3), padding=1).permute(0, 2, 3, 4, 1)
    torch.cuda.synchronize()
    assert rel_l2(dh, ref) <= 5e-3
