"""GPU timing of one Unet_deconv training pass (forward keeping activations + full backward) on a random crop,
CUDA events on the launching stream.  FLOPs: forward 1 327 618 per voxel, backward 2x (dgrad + wgrad;
the first layer has no dgrad).  Usage: python tools/bench_train_step.py [S=108] [iters=5]"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroclear_b200 import _lib, networks  # noqa: E402
from neuroclear_b200.unet_engine import FLOP_PER_VOXEL  # noqa: E402
from neuroclear_b200.unet_train import UnetDeconvTrainEngine  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 108
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
import io
from contextlib import redirect_stdout
with redirect_stdout(io.StringIO()):
    net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
eng = UnetDeconvTrainEngine("cuda")
eng.load_state_dict(net.state_dict())
x = torch.rand((1, S, S, S), device="cuda")
dout = torch.randn((1, S, S, S), device="cuda") * 1e-6
ev = lambda: torch.cuda.Event(enable_timing=True)
tf, tb = [], []
for i in range(iters + 2):
    e0, e1, e2 = ev(), ev(), ev()
    n0 = _lib.LAUNCHES
    e0.record()
    eng.forward(x)
    e1.record()
    eng.backward(dout)
    e2.record()
    torch.cuda.synchronize()
    if i >= 2:
        tf.append(e0.elapsed_time(e1))
        tb.append(e1.elapsed_time(e2))
    launches = _lib.LAUNCHES - n0
f, b = sum(tf) / len(tf), sum(tb) / len(tb)
flop_f = FLOP_PER_VOXEL * S ** 3
print(json.dumps({"crop": S, "fwd_ms": round(f, 3), "bwd_ms": round(b, 3), "launches": launches,
                  "fwd_tflops": round(flop_f / f / 1e9, 1), "bwd_tflops": round(2 * flop_f / b / 1e9, 1),
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)}))
