"""Debug aid (GPU): compares every intermediate gradient tensor of UnetDeconvTrainEngine.backward with the oracle's
autograd (fp16-storage emulation) on the unet_grad fixture.  Not a test; prints one line per tensor."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import unet  # noqa: E402
from neuroclear_b200.unet_train import UnetDeconvTrainEngine  # noqa: E402

z = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "unet_grad.npz"))
sd = unet.random_state_dict(seed=4, bias_std=0.1)
x, dout = torch.from_numpy(z["x"]), torch.from_numpy(z["dout"])
st = lambda t: t + (t.half().float() - t).detach()
keep = {}


def cir(inp, p, name_in=None):
    if name_in and inp.requires_grad:
        inp.retain_grad()
        keep["d_in." + name_in] = inp
    y = st(F.conv3d(inp, st(sd[p + ".weight"]), sd[p + ".bias"], padding=1))
    y.retain_grad()
    keep["d_raw." + p] = y
    return st(F.relu(F.instance_norm(y, eps=1e-5)))


sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
c1 = cir(x, "double_conv1.convolution.0")
c1 = cir(c1, "double_conv1.convolution.3", "double_conv1.convolution.3")
p1 = F.max_pool3d(c1, 2)
c2 = cir(p1, "double_conv2.convolution.0", "double_conv2.convolution.0")
c2 = cir(c2, "double_conv2.convolution.3", "double_conv2.convolution.3")
p2 = F.max_pool3d(c2, 2)
b = cir(p2, "bottom_layer.convolution.0", "bottom_layer.convolution.0")
b = cir(b, "bottom_layer.convolution.3", "bottom_layer.convolution.3")
b = cir(b, "bottom_layer.convolution.6", "bottom_layer.convolution.6")
b.retain_grad(); keep["d_in.t_conv2"] = b
t2 = st(F.conv_transpose3d(b, st(sd["t_conv2.weight"]), sd["t_conv2.bias"], stride=2))
cat2 = torch.cat([c2, t2], 1)
e2 = cir(cat2, "ex_double_conv2.convolution.0", "ex_double_conv2.convolution.0")
e2 = cir(e2, "ex_double_conv2.convolution.3", "ex_double_conv2.convolution.3")
e2.retain_grad(); keep["d_in.t_conv1"] = e2
t1 = st(F.conv_transpose3d(e2, st(sd["t_conv1.weight"]), sd["t_conv1.bias"], stride=2))
cat1 = torch.cat([c1, t1], 1)
e1 = cir(cat1, "ex_conv1_1.convolution.0", "ex_conv1_1.convolution.0")
o = F.conv3d(e1, sd["one_by_one.weight"], sd["one_by_one.bias"])
o = F.conv3d(o, sd["one_by_one_2.weight"], sd["one_by_one_2.bias"])
torch.sigmoid(o).backward(dout)

eng = UnetDeconvTrainEngine("cuda")
eng.load_state_dict({k: v.detach() for k, v in sd.items()})
eng.forward(x.cuda()[:, 0].contiguous())
eng.debug = {}
eng.backward(dout.cuda()[:, 0].contiguous())
torch.cuda.synchronize()
order = ["d_raw.ex_conv1_1.convolution.0", "d_in.ex_conv1_1.convolution.0", "d_in.t_conv1",
         "d_raw.ex_double_conv2.convolution.3", "d_in.ex_double_conv2.convolution.3",
         "d_raw.ex_double_conv2.convolution.0", "d_in.ex_double_conv2.convolution.0", "d_in.t_conv2",
         "d_raw.bottom_layer.convolution.6", "d_in.bottom_layer.convolution.6", "d_raw.bottom_layer.convolution.3",
         "d_in.bottom_layer.convolution.3", "d_raw.bottom_layer.convolution.0", "d_in.bottom_layer.convolution.0",
         "d_raw.double_conv2.convolution.3", "d_in.double_conv2.convolution.3", "d_raw.double_conv2.convolution.0",
         "d_in.double_conv2.convolution.0", "d_raw.double_conv1.convolution.3", "d_in.double_conv1.convolution.3",
         "d_raw.double_conv1.convolution.0"]
for k in order:
    ref = keep[k].grad
    got = eng.debug[k].float().cpu().permute(0, 4, 1, 2, 3)
    if got.shape != ref.shape:
        print("%-46s shape %s vs %s" % (k, tuple(got.shape), tuple(ref.shape)))
        continue
    d = got - ref
    nflip = int(((got == 0) != (ref == 0)).sum())
    print("%-46s rel L2 %.4f  max-abs/max %.4f  zero-pattern mismatches %d of %d" % (
        k, float(d.norm() / ref.norm()), float(d.abs().max() / ref.abs().max()), nflip, ref.numel()))
