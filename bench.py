#!/usr/bin/env python
"""bench.py — voxels/s of diced unet_deconv inference on a synthetic 900^3 16-bit volume (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over the whole volume (729 cubes of 140^3 -> blend -> percentile stretch ->
uint16).  `value` is timed with the uint16 volume already resident in HBM; `e2e` goes through the public API
(DicedInference.run) with the pinned-host -> device copy of the volume and the device -> host copy of the result
inside the timed region.  Timing: CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  One JSON line on stdout (rank 0).

Extra keys next to the contract's: `roofline` (live CUDA-event TFLOP/s of the tcgen05 conv launches), `cpu_baseline`
(the oracle on a bounded sample), `train_step` (N = 1 only; a secondary measurement of BASELINE.json configs[2]: the
full apollo training iteration on a 108^3 crop, after the headline timing; --no-train-sample skips it).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxels/sec diced unet_deconv inference, 900^3 vol"
UNIT = "voxels/s"
PUBLISHED_VOXELS_PER_S = 1.84e6   # 729 cubes at 1.84 it/s on the author's GPU (screenshot)
ROI, OVERLAP, BORDER = 120, 15, 10


def synthetic_volume(shape, seed=0, z0=0, z1=None):
    """uniform random uint16 volume (the reference's generator notebook is a missing blob).
    Seeded per z-plane, so a rank can generate exactly the planes [z0, z1) it owns without building the whole volume."""
    z1 = shape[0] if z1 is None else z1
    out = np.empty((z1 - z0, shape[1], shape[2]), dtype=np.uint16)
    for z in range(z0, z1):
        out[z - z0] = np.random.default_rng([seed, z]).integers(0, 65536, shape[1:], dtype=np.uint16)
    return out


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p.get("bf16_tflops_sustained", p["bf16_tflops"]), "measured (MEASURED_PEAKS.json, sustained bf16)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md, sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_baseline_sample(shape, threads=None):
    """The oracle (port of the reference's CPU path) on a bounded sample of the workload: one of the n cubes through
    dice + Unet_deconv fp32 forward, and the blend/percentile/rescale stage on a 2x2x2-cube (225^3 padded) volume
    scaled by cube count.  Returns (voxels_per_s, cores, sample description, seconds spent)."""
    from oracle import assemble, dice, geometry as ogeo, unet as ounet
    if threads:
        torch.set_num_threads(threads)
    cores = torch.get_num_threads()
    g = ogeo.dice_geometry(shape, ROI, OVERLAP, BORDER)
    sd = ounet.random_state_dict(seed=0)
    small = synthetic_volume((128, 128, 128), seed=1)
    gs = ogeo.dice_geometry(small.shape, ROI, OVERLAP, BORDER)
    t0 = time.perf_counter()
    x = torch.from_numpy(dice.dice_cube_gather(small, gs, 0))[None]
    y = ounet.unet_deconv_forward(x, sd)
    t_cube = time.perf_counter() - t0
    cube = assemble.crop_border(y.numpy(), BORDER)
    t0 = time.perf_counter()
    vis, _ = assemble.blend_sequential([cube] * gs.n_cubes, gs)
    assemble.finish(vis, gs, True)
    t_asm = (time.perf_counter() - t0) * g.n_cubes / gs.n_cubes
    total = g.n_cubes * t_cube + t_asm
    voxels = float(np.prod(shape))
    sample = ("1 of %d cubes (dice + Unet_deconv fp32 forward, %.2f s) + blend/percentile/rescale of an 8-cube volume "
              "scaled x%d/8 (%.2f s), extrapolated to the whole volume" % (g.n_cubes, t_cube, g.n_cubes, t_asm))
    return voxels / total, cores, sample, t_cube + t_asm * gs.n_cubes / g.n_cubes


def run_reference(args, shape, guard):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, t_wall = [], []
    for i in range(args.warmup + args.steps):
        v, cores, sample, secs = cpu_baseline_sample(shape)
        if i >= args.warmup:
            vals.append(v)
            t_wall.append(secs)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.prod(shape)) / value * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": value / PUBLISHED_VOXELS_PER_S, "dtype": "f32", "data": "synthetic",
        "config": workload_config(shape, None, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "reference = the oracle port of the reference's CPU path (torch CPU fp32 + numpy) on all host threads; "
                "the Python reference itself cannot travel to the GPU box; each step is a bounded sample, ms_per_step "
                "is the extrapolated whole-volume time",
    }
    guard.emit(json.dumps(line))


def train_step_cpu_baseline(sample_crop, crop):
    """The oracle's restatement of the reference training iteration (oracle/apollo_step.py, torch CPU fp32 autograd,
    pinned to the reference fixture) on the host cores, on a bounded sample: one iteration at sample_crop^3 after
    one warm-up, scaled to `crop`^3 by the voxel ratio (the generators' cost is linear in the voxel count)."""
    try:
        from oracle import apollo_step, deeplinear, discriminator, unet
        sds = {"G_A": unet.random_state_dict(seed=1), "G_B": deeplinear.random_state_dict(seed=2)}
        for i, n in enumerate(apollo_step.D_NAMES):
            sds[n] = discriminator.random_state_dict(seed=3 + i)
        step = apollo_step.ApolloStep(sds)
        np.random.seed(0)
        g = torch.Generator().manual_seed(0)
        ms = 0.0
        for it in range(2):
            step.set_input(torch.rand((1, 1, sample_crop, sample_crop, sample_crop), generator=g))
            t0 = time.perf_counter()
            step.optimize_parameters()
            ms = (time.perf_counter() - t0) * 1e3
        scaled = ms * (crop / sample_crop) ** 3
        return {"value": 1e3 / scaled, "unit": "iterations/s", "cores": torch.get_num_threads(), "kind": "port",
                "ms_per_iter_sample": ms, "ms_per_iter_scaled": scaled,
                "sample": "one optimize_parameters() of the oracle at %d^3, scaled to %d^3 by the voxel ratio"
                          % (sample_crop, crop)}
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def train_step_sample(dev, crop=108, iters=5, warmup=3, cpu_crop=0):
    """Secondary measurement (BASELINE.json configs[2], not the headline metric): one full training iteration of the
    apollo model (unet_deconv + deep_linear_gen + 4 basic Ds, batch 1, randomized projection depth 10) on a random
    crop — set_input (H2D of the crop from pinned memory) + optimize_parameters(), CUDA events on the launching
    stream.  Never allowed to take the headline line down: any failure is reported in place of the numbers."""
    try:
        import contextlib
        import io
        from argparse import Namespace
        from neuroclear_b200 import _lib
        from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel
        opt = Namespace(isTrain=True, gpu_ids=[dev.index or 0], gan_mode="lsgan", randomize_projection_depth=True,
                        projection_depth=10, min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1,
                        ngf=64, ndf=64, netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3,
                        norm="instance", no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1,
                        direction="AtoB", lambda_A=5.0)
        torch.manual_seed(0)
        np.random.seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            model = AxialToLateralGANApolloModel(opt, dev, distributed=False)
        crops = [torch.rand((1, 1, crop, crop, crop)).pin_memory() for _ in range(2)]
        times, launches = [], 0
        for i in range(warmup + iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n0 = _lib.LAUNCHES
            e0.record()
            model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
            model.optimize_parameters()
            e1.record()
            torch.cuda.synchronize()
            launches = _lib.LAUNCHES - n0
            if i >= warmup:
                times.append(e0.elapsed_time(e1))
        ms = sum(times) / len(times)
        finite = all(np.isfinite(v) for v in model.get_current_losses().values())
        out = {"metric": "apollo training iteration (G_A unet_deconv + G_B deep_linear_gen + 4 PatchGAN Ds, batch 1)",
               "crop": crop, "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "iters": iters, "warmup": warmup,
               "library_calls_per_iter": launches, "losses_finite": bool(finite),
               "data": "synthetic random crop, random-init weights"}
        if cpu_crop:
            out["cpu_baseline"] = train_step_cpu_baseline(cpu_crop, crop)
        return out
    except Exception as e:  # noqa: BLE001 - secondary measurement
        return {"error": "%s: %s" % (type(e).__name__, e)}


def workload_config(shape, batch, gpus):
    return {"workload": "test_dice.py unet_deconv inference, synthetic %dx%dx%d uint16 volume, dice %d overlap %d "
                        "border_cut %d, normalize_intensity (0.25, 99.75)" % (*shape, ROI, OVERLAP, BORDER),
            "cubes": None, "batch_cubes": batch, "parallelism": "cube-range x%d + z-slab assembly" % gpus,
            "l2": "inputs larger than L2 (1.46 GB volume, >2 GB of activations per batch)"}


# ------------------------------------------------------------------------------------------------ GPU arm
class StdoutGuard:
    """Libraries (NCCL's version banner, torch warnings) may print to fd 1; the driver wants exactly ONE JSON line
    there.  Everything written to fd 1 while the guard is active goes to stderr; emit() writes to the real stdout."""

    def __init__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)

    def emit(self, line: str):
        sys.stdout.flush()
        os.write(self.real, (line + "\n").encode())


def main():
    guard = StdoutGuard()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, nargs=3, default=[900, 900, 900])
    ap.add_argument("--batch", type=int, default=9)   # 729 = 81 x 9: no ragged last batch
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train-sample", action="store_true")
    args = ap.parse_args()
    shape = tuple(args.size)

    if args.impl == "reference":
        return run_reference(args, shape, guard)

    import torch.distributed as dist
    from neuroclear_b200 import _lib
    from neuroclear_b200.pipeline import DicedInference
    from neuroclear_b200 import networks
    from neuroclear_b200.unet_engine import FLOP_PER_VOXEL

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N>1)" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # random-init weights of the named architecture through the package's own define_G (kaiming, seed 0); the oracle
    # is NOT on this arm: it is only executed by cpu_baseline_sample() / --impl reference
    torch.manual_seed(0)
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    sd = net.state_dict()
    pipe = DicedInference(sd, dev, ROI, OVERLAP, BORDER, normalize_intensity=True, batch=args.batch)
    plan = pipe.plan(shape)
    geo = plan["geo"]
    z0, z1 = plan["in_planes"]
    # every rank generates (a real run would read from disk) only the input planes it needs
    slab_host = torch.from_numpy(synthetic_volume(shape, z0=z0, z1=z1)).pin_memory()
    vol_dev = slab_host.to(dev)
    o0, o1 = plan["out_planes"]
    out_host = torch.empty((o1 - o0, shape[1], shape[2]), dtype=torch.uint16).pin_memory()

    def timed(fn, steps, profile=False):
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        launches0 = _lib.LAUNCHES
        pipe.engine.profile = [] if profile else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop()
        prof, pipe.engine.profile = pipe.engine.profile, None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, clocks, _lib.LAUNCHES - launches0, prof

    resident = lambda: pipe.run_device(vol_dev, z0, shape)
    e2e = lambda: pipe.run_slab(slab_host, z0, shape, out_host)

    for _ in range(args.warmup):
        resident()
    ms, clocks, launches, prof = timed(resident, args.steps, profile=True)
    voxels = float(np.prod(shape))
    value = voxels * args.steps / (ms * 1e-3)

    # live roofline of the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs / event time
    peak, peak_src = measured_peaks()
    layers = {}
    for name, flops, a, b in prof:
        t = layers.setdefault(name, [0.0, 0.0, 0])
        t[0] += flops
        t[1] += a.elapsed_time(b)
        t[2] += 1
    tot_f = sum(v[0] for v in layers.values())
    tot_ms = sum(v[1] for v in layers.values())
    achieved = tot_f / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "tensor", "kernel": "conv3d_tc_kernel (9 Conv3d k3 layers U2..U12, per-launch average)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": peak_src, "launches": sum(v[2] for v in layers.values()),
                "share_of_step": tot_ms / ms,
                "layers": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms_per_launch": v[1] / v[2]}
                           for k, v in layers.items()}}

    for _ in range(1):
        e2e()
    ms_e, clocks_e, _, _ = timed(e2e, args.steps)
    e2e_value = voxels * args.steps / (ms_e * 1e-3)
    in_bytes = (z1 - z0) * shape[1] * shape[2] * 2
    out_bytes = (o1 - o0) * shape[1] * shape[2] * 2

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, cores, sample, _ = cpu_baseline_sample(shape)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    train = None
    if rank == 0 and world == 1 and not args.no_train_sample:
        train = train_step_sample(dev, cpu_crop=0 if args.no_cpu_baseline else 48)

    if rank == 0:
        cfg = workload_config(shape, args.batch, world)
        cfg["cubes"] = geo.n_cubes
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / PUBLISHED_VOXELS_PER_S, "dtype": "fp16", "data": "synthetic", "config": cfg,
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "ms_per_step": ms_e / args.steps, "clocks": clocks_e},
            "gpu_launches": launches,
            "roofline": roofline,
            "cpu_baseline": cpu,
            "train_step": train,
            "tensor_pipe_frac_whole_step": FLOP_PER_VOXEL * geo.n_cubes * geo.edge ** 3 * args.steps
                                            / (ms * 1e-3) / 1e12 / peak / world,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
