#!/usr/bin/env python
"""The cuDNN "library bar" on the same GPU: how fast does the reference's own
formulation of the hot path run when PyTorch / cuDNN executes it on a B200?  NOT product code — never imported by
neuroclear_b200; bench.py runs it as a subprocess at N = 1 and attaches the result as `library_bar`.

    python tools/library_bar.py [--json] [--edge 140] [--iters 5] [--train-crop 108] [--no-train]

1. Inference: torch.nn Unet_deconv (the architecture of the reference's models/networks.py:478-538, written with
   nn.Sequential below; random init) on ONE 140^3 cube, batch 1 as in test_dice.py:
     a. fp32 NCDHW, cudnn.benchmark = True (models/base_model.py:40-41), TF32 convolutions allowed — torch's default
        and therefore the reference "as shipped";
     b. the same with TF32 disabled (true fp32 arithmetic);
     c. bf16 weights + activations, channels_last_3d — the fastest library configuration.
   Reported per cube in ms and as reference-FLOP TFLOP/s (3 642 983 792 000 FLOP per 140^3 cube), next to the
   hand-written engine on the same cube (batch 1 and batch 9).
2. Training (when oracle/_ref, the byte-compiled reference, is present): one optimize_parameters() of the
   REFERENCE's AxialToLateralGANApolloModel on the GPU (fp32 + TF32 through cuDNN, its stock code path) at the crop
   size of BASELINE.json configs[2], against the package's AxialToLateralGANApolloModel.
"""
from __future__ import annotations

import argparse
import contextlib
import io
import json
import os
import sys
import time

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FLOP_PER_VOXEL = 1_327_618      # 2 x MACs of the 14 convolutions per network-input voxel


def _stack(n, cin, cout):
    layers = []
    for i in range(n):
        layers += [nn.Conv3d(cin if i == 0 else cout, cout, 3, 1, 1), nn.InstanceNorm3d(cout), nn.ReLU()]
    return nn.Sequential(*layers)


class TorchUnetDeconv(nn.Module):
    """double_conv / maxpool / double_conv / maxpool / triple_conv / convT k2s2 + cat / double_conv / convT + cat /
    conv / 1x1 / 1x1 / sigmoid — the layer table of SURVEY.md §2c (U1..U14)."""

    def __init__(self):
        super().__init__()
        self.d1, self.d2, self.bottom = _stack(2, 1, 64), _stack(2, 64, 128), _stack(3, 128, 256)
        self.t2, self.u2 = nn.ConvTranspose3d(256, 128, 2, 2), _stack(2, 256, 128)
        self.t1, self.u1 = nn.ConvTranspose3d(128, 64, 2, 2), _stack(1, 128, 64)
        self.head = nn.Sequential(nn.Conv3d(64, 1, 1), nn.Conv3d(1, 1, 1), nn.Sigmoid())
        self.pool = nn.MaxPool3d(2)

    def forward(self, x):
        c1 = self.d1(x)
        c2 = self.d2(self.pool(c1))
        b = self.bottom(self.pool(c2))
        e2 = self.u2(torch.cat([c2, self.t2(b)], 1))
        return self.head(self.u1(torch.cat([c1, self.t1(e2)], 1)))


def _time(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def inference_bar(edge, iters):
    dev = torch.device("cuda", 0)
    flop = FLOP_PER_VOXEL * edge ** 3
    out = {"cube_edge": edge, "flop_per_cube": flop}
    torch.backends.cudnn.benchmark = True
    torch.manual_seed(0)
    net = TorchUnetDeconv().to(dev).eval()
    x = torch.rand((1, 1, edge, edge, edge), device=dev)

    def entry(ms):
        return {"ms_per_cube": ms, "tflops": flop / ms / 1e9, "voxels_per_s_900cube_equiv": 900 ** 3 / (729 * ms * 1e-3)}

    with torch.no_grad():
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        y_tf32 = net(x)
        out["cudnn_fp32_tf32_ncdhw"] = entry(_time(lambda: net(x), iters))
        torch.backends.cudnn.allow_tf32 = False
        y_fp32 = net(x)
        out["cudnn_fp32_strict_ncdhw"] = entry(_time(lambda: net(x), max(2, iters // 2)))
        torch.backends.cudnn.allow_tf32 = True
        out["tf32_vs_fp32_max_abs"] = float((y_tf32 - y_fp32).abs().max())
        nb = net.to(torch.bfloat16).to(memory_format=torch.channels_last_3d)
        xb = x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last_3d)
        y_bf16 = nb(xb)
        out["cudnn_bf16_channels_last_3d"] = entry(_time(lambda: nb(xb), iters))
        out["bf16_vs_fp32_max_abs"] = float((y_bf16.float() - y_fp32).abs().max())
    del net, nb, y_tf32, y_fp32, y_bf16
    torch.cuda.empty_cache()

    # the hand-written engine on the same cube, same random-init architecture
    from neuroclear_b200 import networks
    from neuroclear_b200.unet_engine import UnetDeconvEngine
    with contextlib.redirect_stdout(io.StringIO()):
        ours = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    eng = UnetDeconvEngine(dev)
    eng.load_state_dict(ours.state_dict())
    for nb_ in (1, 9):
        xx = torch.rand((nb_, edge, edge, edge), device=dev)
        ms = _time(lambda: eng.forward(xx), iters) / nb_
        out["neuroclear_b200_batch%d" % nb_] = entry(ms)
    best = min(out[k]["ms_per_cube"] for k in ("cudnn_fp32_tf32_ncdhw", "cudnn_bf16_channels_last_3d"))
    out["speedup_vs_cudnn_as_shipped"] = out["cudnn_fp32_tf32_ncdhw"]["ms_per_cube"] / out["neuroclear_b200_batch9"]["ms_per_cube"]
    out["speedup_vs_best_cudnn"] = best / out["neuroclear_b200_batch9"]["ms_per_cube"]
    return out


def training_bar(crop, iters):
    """The reference's own apollo model (byte-compiled in oracle/_ref) on the GPU vs the package's."""
    from argparse import Namespace
    from oracle import reference_harness as rh       # measurement of the reference itself, not a product path
    if not rh.compiled_available():
        return {"unavailable": "oracle/_ref absent (python -m oracle.build_ref)"}
    dev = torch.device("cuda", 0)
    rh.install(compiled=True)
    from models.axial_to_lateral_gan_apollo_model import AxialToLateralGANApolloModel as RefModel
    base = dict(isTrain=True, checkpoints_dir="/tmp/nc_libbar_ckpt", name="libbar", preprocess="none",
                gan_mode="lsgan", image_dimension=3, randomize_projection_depth=True, projection_depth=10,
                min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
                netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance",
                no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, direction="AtoB",
                lambda_A=5.0, lr_policy="constant")
    out = {"crop": crop}
    torch.manual_seed(0)
    np.random.seed(0)
    real = torch.rand((1, 1, crop, crop, crop)).pin_memory()

    def run(model):
        def step():
            model.set_input({"A": real, "A_paths": "synthetic"})
            model.optimize_parameters()
        # 6 warm-up iterations: the caching allocator (side-stream record_stream) needs a few iterations before the
        # timed region stops calling cudaMalloc (bench.py's train_step uses the same count)
        return _time(step, iters, warmup=6)

    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefModel(Namespace(gpu_ids=[0], **base))
    out["reference_cudnn_fp32_tf32_ms_per_iter"] = run(ref)
    out["reference_peak_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    del ref
    torch.cuda.empty_cache()
    from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel
    with contextlib.redirect_stdout(io.StringIO()):
        ours = AxialToLateralGANApolloModel(Namespace(gpu_ids=[0], **base), dev, distributed=False)
    out["neuroclear_b200_ms_per_iter"] = run(ours)
    out["speedup_vs_cudnn_as_shipped"] = out["reference_cudnn_fp32_tf32_ms_per_iter"] / out["neuroclear_b200_ms_per_iter"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", action="store_true")
    ap.add_argument("--edge", type=int, default=140)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--train-crop", type=int, default=108)
    ap.add_argument("--no-train", action="store_true")
    args = ap.parse_args()
    t0 = time.time()
    res = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    try:
        res["inference"] = inference_bar(args.edge, args.iters)
    except Exception as e:  # noqa: BLE001
        res["inference"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if not args.no_train:
        try:
            res["training"] = training_bar(args.train_crop, args.iters)
        except Exception as e:  # noqa: BLE001
            res["training"] = {"error": "%s: %s" % (type(e).__name__, e)}
    res["seconds"] = time.time() - t0
    print(json.dumps(res) if args.json else json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
