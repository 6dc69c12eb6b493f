"""Small end-to-end workload for compute-sanitizer: diced inference (cube edge 40: the remainder-pair kernel runs at
the 20^3 level), histogram matching + float64 blend, the report kernels, one apollo training iteration at 40^3
(tiled conv1_wgrad / stencil kernels, batched discriminators, side-stream D step)."""
import io
import os
import sys
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroclear_b200 import networks, report
from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel
from neuroclear_b200.dicing import Assemble_Dice, DiceImageDataSet
from neuroclear_b200.pipeline import DicedInference

dev = torch.device("cuda", 0)
torch.manual_seed(0)
with redirect_stdout(io.StringIO()):
    net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
vol = (np.random.default_rng(0).random((50, 45, 70)) ** 3 * 65535).astype(np.uint16)
out, _ = DicedInference(net.state_dict(), dev, 32, 8, 4, batch=3).run(vol)
print("inference", out.shape, int(out.max()))
opt = Namespace(dataroot="", dice_size=[32] * 3, overlap=8, border_cut=4, preprocess="addColorChannel",
                data_type="uint16", skip_real=True, histogram_match=True, normalize_intensity=True,
                sat_level=[0.25, 99.75], gpu_ids=[0])
ds = DiceImageDataSet(opt, volume=vol)
asm = Assemble_Dice(opt, ds)
net = net.to(dev).eval()
with torch.no_grad():
    for i in range(len(ds)):
        real = ds[i]["A"][None]
        asm.addToStack({"real": real, "fake": net(real)})
asm.assemble_all()
fake = asm.getDict()["fake"]
print("histogram-matched assembly", fake.shape)
print("projections", [report.max_projection(fake, a, device=dev).shape for a in range(3)])   # (the reference's
# hard-coded --save_projections windows are empty on a volume this small: np.amax raises there, and so do we)
print("psnr", report.psnr_report(vol, fake, np.roll(vol, 1, 0), dev)[:2])
mopt = Namespace(isTrain=True, gpu_ids=[0], gan_mode="lsgan", randomize_projection_depth=True, projection_depth=10,
                 min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
                 netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance",
                 no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, direction="AtoB", lambda_A=5.0)
with redirect_stdout(io.StringIO()):
    m = AxialToLateralGANApolloModel(mopt, dev, distributed=False)
np.random.seed(0)
for _ in range(2):
    m.set_input({"A": torch.rand((1, 1, 40, 40, 40)), "A_paths": "x"})
    m.optimize_parameters()
torch.cuda.synchronize()
print("training", {k: round(v, 4) for k, v in m.get_current_losses().items()})
