"""GPU timing of the full apollo training iteration (BASELINE.json configs[2]/[3]: train_onecube.py
axial_to_lateral_gan_apollo, unet_deconv + deep_linear_gen + basic D, batch 1, randomized projection depth 10) on a
random crop: set_input (H2D of the crop) + optimize_parameters(), CUDA events on the launching stream.
Usage: python tools/bench_apollo_step.py [S=108] [iters=5]
       python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
              tools/bench_apollo_step.py S iters        (data parallel: one crop per GPU, gradients all-reduced)"""
import io
import json
import os
import sys
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from neuroclear_b200 import _lib  # noqa: E402
from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel  # noqa: E402
from neuroclear_b200 import deeplinear_engine, unet_engine  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 108
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
opt = Namespace(isTrain=True, gpu_ids=[0], gan_mode="lsgan", randomize_projection_depth=True, projection_depth=10,
                min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
                netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance",
                no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, direction="AtoB", lambda_A=5.0)
import torch.distributed as dist  # noqa: E402

world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
opt.gpu_ids = [local]
torch.manual_seed(0)             # same initial weights on every rank
np.random.seed(rank)
with redirect_stdout(io.StringIO()):
    m = AxialToLateralGANApolloModel(opt, "cuda:%d" % local, distributed=world > 1)
torch.manual_seed(100 + rank)    # different crops
crops = [torch.rand((1, 1, S, S, S)).pin_memory() for _ in range(3)]
ev = lambda: torch.cuda.Event(enable_timing=True)
# warm-up, then (a) latency: a device synchronisation after every iteration; (b) throughput: `iters` iterations
# enqueued back to back, one synchronisation at the end (how the reference's loop runs, train_onecube.py:83-110)
for i in range(4):
    m.set_input({"A": crops[i % 3], "A_paths": "synthetic"})
    m.optimize_parameters()
torch.cuda.synchronize()
lat = []
for i in range(3):
    e0, e1 = ev(), ev()
    if world > 1:
        dist.barrier()
    e0.record()
    m.set_input({"A": crops[i % 3], "A_paths": "synthetic"})
    m.optimize_parameters()
    e1.record()
    torch.cuda.synchronize()
    lat.append(e0.elapsed_time(e1))
if world > 1:
    dist.barrier()
e0, e1 = ev(), ev()
n0 = _lib.LAUNCHES
e0.record()
for i in range(iters):
    m.set_input({"A": crops[i % 3], "A_paths": "synthetic"})
    m.optimize_parameters()
e1.record()
torch.cuda.synchronize()
launches = (_lib.LAUNCHES - n0) // iters
ms = e0.elapsed_time(e1) / iters
ms_sync = sorted(lat)[len(lat) // 2]
if world > 1:
    t = torch.tensor([ms], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
flop = 3 * (unet_engine.FLOP_PER_VOXEL + deeplinear_engine.FLOP_PER_VOXEL) * S ** 3      # reference layer FLOPs, fwd + 2x bwd
losses = m.get_current_losses()
if rank == 0:
  print(json.dumps({"crop": S, "n_gpus": world, "ms_per_iter": round(ms, 3), "ms_per_iter_synchronised": round(ms_sync, 3),
                  "crops_per_s": round(world * 1e3 / ms, 2), "launches": launches,
                  "reference_flop_tflops": round(world * flop / ms / 1e9, 1),
                  "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2 ** 30, 2),
                  "finite": all(np.isfinite(v) for v in losses.values())}))
if world > 1:
    dist.destroy_process_group()
