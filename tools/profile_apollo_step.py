"""Debug aid (GPU): wall-clock breakdown of one apollo training iteration by phase (host timers, synchronised)."""
import io
import os
import sys
import time
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroclear_b200 import _lib  # noqa: E402
from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 108
opt = Namespace(isTrain=True, gpu_ids=[0], gan_mode="lsgan", randomize_projection_depth=True, projection_depth=10,
                min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
                netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance",
                no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, direction="AtoB", lambda_A=5.0)
torch.manual_seed(0)
np.random.seed(0)
with redirect_stdout(io.StringIO()):
    m = AxialToLateralGANApolloModel(opt, "cuda", distributed=False)
crop = torch.rand((1, 1, S, S, S)).pin_memory()
acc = {}


def phase(name, fn):
    torch.cuda.synchronize()
    t0, n0 = time.perf_counter(), _lib.LAUNCHES
    out = fn()
    t1 = time.perf_counter()          # host time to ENQUEUE
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    a = acc.setdefault(name, [0.0, 0.0, 0])
    a[0] += t1 - t0
    a[1] += t2 - t0
    a[2] = _lib.LAUNCHES - n0
    return out


for it in range(6):
    if it == 2:
        acc.clear()
    phase("set_input", lambda: m.set_input({"A": crop, "A_paths": "x"}))
    phase("forward G_A", lambda: setattr(m, "fake", m.netG_A(m.real)))
    phase("forward G_B", lambda: setattr(m, "rec", m.netG_B(m.fake)))
    phase("zero_grad", lambda: m.optimizer_G.zero_grad())
    phase("generator_losses", lambda: setattr(m, "loss_G", m.dpath.generator_losses(m.real, m.fake, m.rec)))
    phase("loss_G.backward", lambda: m.loss_G.backward())
    phase("optimizer_G.step", lambda: m.optimizer_G.step())
    phase("optimize_D", lambda: m.dpath.optimize_D(m.real, m.fake.detach(), m.rec.detach()))
n = 4
print("%-20s %10s %10s %9s" % ("phase", "enqueue ms", "total ms", "launches"))
for k, (h, t, l) in acc.items():
    print("%-20s %10.2f %10.2f %9d" % (k, 1e3 * h / n, 1e3 * t / n, l))
print("%-20s %10.2f %10.2f" % ("sum", 1e3 * sum(v[0] for v in acc.values()) / n, 1e3 * sum(v[1] for v in acc.values()) / n))
