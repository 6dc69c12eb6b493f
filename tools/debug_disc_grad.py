"""Per-parameter gradient error of one D step (tests/test_gpu_discriminator.py::test_vs_oracle_autograd) for the
cluster sizes of nc_conv2d_k4_* (debug hook nc_debug_set_disc_cluster).  Usage: python tools/debug_disc_grad.py"""
import io
import os
import sys
from contextlib import redirect_stdout

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroclear_b200 import _lib, discriminator  # noqa: E402
from oracle import discriminator as odisc  # noqa: E402

cuda = torch.device("cuda", 0)
n, h, w = 1, 108, 108
sd = odisc.random_state_dict(seed=3)
g = torch.Generator().manual_seed(h)
real, fake = torch.rand((n, 1, h, w), generator=g), torch.rand((n, 1, h, w), generator=g)
sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
lo = 0.5 * (odisc.lsgan_loss(odisc.discriminator_forward(real, sdo), True) +
            odisc.lsgan_loss(odisc.discriminator_forward(fake, sdo), False))
lo.backward()
for c in (0, 1, 2, 4, 8):
    _lib.load().nc_debug_set_disc_cluster(c)
    with redirect_stdout(io.StringIO()):
        net = discriminator.define_D(1, 64, "basic", norm="instance", use_sigmoid=False, init_type="kaiming",
                                     init_gain=0.02, gpu_ids=[0], dimension=2)
    net.module.load_state_dict(sd)
    crit = discriminator.GANLoss("lsgan").to(cuda)
    loss = 0.5 * (crit(net(real.to(cuda)), True) + crit(net(fake.to(cuda)), False))
    loss.backward()
    out = ["cluster %d loss err %.2e" % (c, abs(loss.item() - lo.item()))]
    for k, prm in net.module.named_parameters():
        ref = sdo[k].grad
        out.append("%s %.1e" % (k.replace("model.", "m"), (prm.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)))
    print("  ".join(out))
