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
(the reference / oracle on a bounded sample), `out_sha256` (digest of the assembled uint16 volume, independent of how
many GPUs produced it: the bit-exact multi-GPU claim), `train_step` (N = 1; BASELINE.json configs[2]: the full apollo
training iteration on a 108^3 crop), `train_step_dp` (every N; configs[3]: the same iteration at 148^3 per GPU, data
parallel under NCCL, with the all-reduce time and the weak-scaling efficiency against the same ranks running alone),
`config5` (N = 8; configs[4]: the 1024x2048x2048 volume) and `library_bar` (N = 1; tools/library_bar.py: the
reference's torch.nn Unet_deconv through cuDNN on the same GPU).  All secondary measurements run AFTER the headline
timing, and a failure in one of them is reported in its place, never instead of the line.
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
def host_threads():
    """All host cores this process may use.  torch.distributed.run exports OMP_NUM_THREADS=1; the CPU arm must not
    inherit that (round 1's N>1 reference lines ran single-threaded)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    torch.set_num_threads(n)
    return torch.get_num_threads()


class CpuSample:
    """The reference's CPU path on a BOUNDED sample of the workload (BASELINE.json configs[0] geometry: the 8-cube
    128^3 volume, dice 120 / overlap 15 / border 10): `cubes` cubes through dice + Unet_deconv fp32 forward, and the
    blend / percentile / rescale stage of the 8-cube volume.  kind = "reference": the reference's own modules
    (oracle/_ref, byte-compiled by oracle/build_ref.py: data.DiceImageDataSet, models.networks.define_G,
    util.assemble_dice.Assemble_Dice); kind = "port": the oracle restatement, when oracle/_ref is absent."""

    def __init__(self):
        self.cores = host_threads()
        self.small = synthetic_volume((128, 128, 128), seed=1)
        from oracle import reference_harness as rh
        self.kind = "reference" if rh.compiled_available() else "port"
        if self.kind == "reference":
            import contextlib
            import io
            import tempfile
            rh.install(compiled=True)
            rh.set_volume(self.small)
            import data as refdata
            from models import networks as refnet
            from util.assemble_dice import Assemble_Dice
            self.opt = rh.dice_opt(rh.make_dataroot(tempfile.mkdtemp()), ROI, OVERLAP, BORDER, True)
            torch.manual_seed(0)
            with contextlib.redirect_stdout(io.StringIO()):
                self.ds = refdata.find_dataset_using_name("diceImage")(self.opt)
                self.net = refnet.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [],
                                           dimension=3).eval()
            self.Assemble_Dice = Assemble_Dice
            self.n_small = len(self.ds)
        else:
            from oracle import geometry as ogeo, unet as ounet
            self.gs = ogeo.dice_geometry(self.small.shape, ROI, OVERLAP, BORDER)
            self.sd = ounet.random_state_dict(seed=0)
            self.n_small = self.gs.n_cubes

    def cube(self, i):
        """seconds for dice + normalise + forward of cube i of the sample volume; returns (seconds, output)"""
        t0 = time.perf_counter()
        if self.kind == "reference":
            with torch.no_grad():
                y = self.net(self.ds[i % self.n_small]["A"][None])
        else:
            from oracle import dice, unet as ounet
            x = torch.from_numpy(dice.dice_cube_gather(self.small, self.gs, i % self.n_small))[None]
            y = ounet.unet_deconv_forward(x, self.sd)[None]
        return time.perf_counter() - t0, y

    def assemble(self, y):
        """seconds for the whole assembly stage of the 8-cube sample volume (every cube = y)"""
        import contextlib
        import io
        from collections import OrderedDict
        t0 = time.perf_counter()
        if self.kind == "reference":
            with contextlib.redirect_stdout(io.StringIO()):
                asm = self.Assemble_Dice(self.opt)
                for _ in range(self.n_small):
                    asm.addToStack(OrderedDict(real=y, fake=y))
                asm.assemble_all()
                asm.getDict()
        else:
            from oracle import assemble
            cube = assemble.crop_border(y[0].numpy(), BORDER)
            vis, _ = assemble.blend_sequential([cube] * self.n_small, self.gs)
            assemble.finish(vis, self.gs, True)
        return time.perf_counter() - t0

    def describe(self, n_cubes, t_cubes, t_asm):
        return ("%d of %d cubes (dice + Unet_deconv fp32 forward, median %.2f s of %s) + blend/percentile/rescale of "
                "an 8-cube volume (%.2f s, scaled x%d/8); voxels/s = volume voxels / (cubes x median cube time + "
                "scaled assembly time)" % (len(t_cubes), n_cubes, float(np.median(t_cubes)),
                                           "/".join("%.2f" % t for t in t_cubes[:6]), t_asm, n_cubes))


def cpu_baseline_sample(shape, n_time_cubes=3, sample=None):
    """Returns (voxels_per_s, cores, kind, sample description, seconds spent)."""
    sample = sample or CpuSample()
    n_cubes = n_cubes_of(shape)
    t_cubes, y = [], None
    for i in range(n_time_cubes):
        t, y = sample.cube(i)
        t_cubes.append(t)
    t_asm = sample.assemble(y)
    total = n_cubes * float(np.median(t_cubes)) + t_asm * n_cubes / sample.n_small
    return (float(np.prod(shape)) / total, sample.cores, sample.kind, sample.describe(n_cubes, t_cubes, t_asm),
            sum(t_cubes) + t_asm)


def run_reference(args, shape, guard):
    """`--impl reference`: the reference's CPU implementation of the path on the box's host cores.  Rank 0 alone works
    (the other ranks of a torchrun launch exit 0).  Every step is a bounded sample (one cube + the 8-cube assembly;
    >= 3 cubes over the run); `value` = whole-volume voxels/s from the MEDIAN cube time over all timed steps,
    `ms_per_step` = the measured wall time of a sample step (so steps x ms_per_step is what this run really took),
    `ms_per_step_whole_volume` = the extrapolated time for all 729 cubes."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    sample = CpuSample()
    n_cubes = n_cubes_of(shape)
    per_step = max(1, -(-3 // max(args.steps, 1)))
    t_cubes, t_asms, walls = [], [], []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        tc = []
        for j in range(per_step):
            t, y = sample.cube(i * per_step + j)
            tc.append(t)
        ta = sample.assemble(y)
        if i >= args.warmup:
            t_cubes += tc
            t_asms.append(ta)
            walls.append(time.perf_counter() - t0)
    t_asm = float(np.median(t_asms))
    total = n_cubes * float(np.median(t_cubes)) + t_asm * n_cubes / sample.n_small
    value = float(np.prod(shape)) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3, "ms_per_step_whole_volume": total * 1e3,
        "higher_is_better": True,
        "scaling": "strong", "vs_baseline": value / PUBLISHED_VOXELS_PER_S, "dtype": "f32", "data": "synthetic",
        "config": workload_config(shape, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": sample.cores, "kind": sample.kind,
                         "sample": sample.describe(n_cubes, t_cubes, t_asm)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": ("reference = the reference's own modules (byte-compiled into oracle/_ref by oracle/build_ref.py) "
                 if sample.kind == "reference" else
                 "reference = the oracle port of the reference's CPU path (oracle/_ref absent) ") +
                "on %d host threads; torch CPU fp32 + numpy; each step is a bounded sample and the whole-volume "
                "throughput is extrapolated by cube count (a full run would take %.0f min)" % (sample.cores, total / 60),
    }
    if sample.cores <= 1:
        line["warning"] = "CPU arm ran on a single thread"
    guard.emit(json.dumps(line))


def train_step_cpu_baseline(sample_crop, crop):
    """The oracle's restatement of the reference training iteration (oracle/apollo_step.py, torch CPU fp32 autograd,
    bit-identical to the reference fixture) on all host cores, on a bounded sample: one iteration at sample_crop^3
    (>= 64) after one warm-up, scaled to `crop`^3 by the voxel ratio (the generators' cost is linear in the voxels)."""
    try:
        from oracle import apollo_step, deeplinear, discriminator, unet
        cores = host_threads()
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
        return {"value": 1e3 / scaled, "unit": "iterations/s", "cores": cores, "kind": "port",
                "ms_per_iter_sample": ms, "ms_per_iter_scaled": scaled,
                "sample": "one optimize_parameters() of the oracle at %d^3, scaled to %d^3 by the voxel ratio"
                          % (sample_crop, crop)}
    except Exception as e:  # noqa: BLE001
        return {"error": "%s: %s" % (type(e).__name__, e)}


def apollo_opt(local):
    from argparse import Namespace
    return Namespace(isTrain=True, gpu_ids=[local], gan_mode="lsgan", randomize_projection_depth=True,
                     projection_depth=10, min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1,
                     ngf=64, ndf=64, netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3,
                     norm="instance", no_dropout=True, init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1,
                     direction="AtoB", lambda_A=5.0)


def time_apollo_iterations(dev, crop, iters, warmup, distributed, rank=0, barrier=None):
    """set_input (H2D of the crop from pinned memory) + optimize_parameters(), CUDA events on the launching stream.
    Returns (ms per iteration on this rank, library calls per iteration, all-reduce ms per iteration, model)."""
    import contextlib
    import io
    from neuroclear_b200 import _lib, apollo_d_path
    from neuroclear_b200.apollo_model import AxialToLateralGANApolloModel
    torch.manual_seed(0)                 # the same initial weights on every rank
    np.random.seed(rank)                 # each rank its own projection-depth / slice draws
    with contextlib.redirect_stdout(io.StringIO()):
        model = AxialToLateralGANApolloModel(apollo_opt(dev.index or 0), dev, distributed=distributed)
    g = torch.Generator().manual_seed(100 + rank)      # a different crop per rank
    crops = [torch.rand((1, 1, crop, crop, crop), generator=g).pin_memory() for _ in range(2)]
    # Steady-state throughput, as the reference's training loop runs it (train_onecube.py:83-110 never synchronises
    # between iterations): W warm-up iterations, then K iterations enqueued back to back between two CUDA events, ONE
    # synchronisation at the end.  (Synchronising after every iteration adds the ~2 ms the host needs to enqueue the
    # next forward to each iteration: that latency figure is returned as well.)
    launches, ar = 0, []
    for i in range(warmup):
        if barrier is not None:
            barrier()
        model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
        model.optimize_parameters()
    torch.cuda.synchronize()
    lat = []
    for i in range(3):                                      # per-iteration latency (synchronised every iteration)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
        model.optimize_parameters()
        e1.record()
        torch.cuda.synchronize()
        lat.append(e0.elapsed_time(e1))
    time_apollo_iterations.last_latency_ms = sorted(lat)[len(lat) // 2]      # median of three
    if barrier is not None:
        barrier()
    # three repeats of the K-iteration region, median reported (a single region occasionally catches a host hiccup:
    # the same binary measured 17.1 and 18.7 ms in two driver-style runs while the synchronised latency stayed put)
    times, ar, launches = [], [], 0
    for rep in range(3):
        if barrier is not None:
            barrier()
        apollo_d_path.ALLREDUCE_EVENTS = [] if distributed else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = _lib.LAUNCHES
        e0.record()
        for i in range(iters):
            model.set_input({"A": crops[i % 2], "A_paths": "synthetic"})
            model.optimize_parameters()
        e1.record()
        torch.cuda.synchronize()
        launches = (_lib.LAUNCHES - n0) // iters
        times.append(e0.elapsed_time(e1) / iters)
        if distributed:
            ar.append(sum(a.elapsed_time(b) for a, b in apollo_d_path.ALLREDUCE_EVENTS) / iters)
        apollo_d_path.ALLREDUCE_EVENTS = None
    mid = sorted(range(3), key=lambda j: times[j])[1]
    time_apollo_iterations.last_repeats_ms = list(times)
    return times[mid], launches, (ar[mid] if ar else 0.0), model


def train_step_sample(dev, crop=108, iters=10, warmup=6, cpu_crop=0):
    """Secondary measurement (BASELINE.json configs[2], not the headline metric): one full training iteration of the
    apollo model (unet_deconv + deep_linear_gen + 4 basic Ds, batch 1, randomized projection depth 10) on a random
    crop.  Never allowed to take the headline line down: any failure is reported in place of the numbers."""
    try:
        ms, launches, _, model = time_apollo_iterations(dev, crop, iters, warmup, False)
        finite = all(np.isfinite(v) for v in model.get_current_losses().values())
        out = {"metric": "apollo training iteration (G_A unet_deconv + G_B deep_linear_gen + 4 PatchGAN Ds, batch 1)",
               "crop": crop, "ms_per_iter": ms, "iters_per_s": 1e3 / ms, "iters": iters, "warmup": warmup,
               "library_calls_per_iter": launches, "losses_finite": bool(finite),
               "ms_per_iter_synchronised": time_apollo_iterations.last_latency_ms,
               "ms_per_iter_repeats": time_apollo_iterations.last_repeats_ms,
               "timing": "K iterations enqueued back to back between two CUDA events (steady-state throughput, as the "
                         "reference's loop runs), three such regions, median; ms_per_iter_synchronised = with a device "
                         "synchronisation after every iteration",
               "data": "synthetic random crop, random-init weights"}
        if cpu_crop:
            out["cpu_baseline"] = train_step_cpu_baseline(cpu_crop, crop)
        return out
    except Exception as e:  # noqa: BLE001 - secondary measurement
        return {"error": "%s: %s" % (type(e).__name__, e)}


def train_step_dp(dev, rank, world, barrier, crop=148, iters=10, warmup=6):
    """BASELINE.json configs[3]: the apollo iteration data parallel, one crop^3 per GPU, gradients of both optimisers
    averaged by one NCCL all-reduce each (the reference's only collective is DataParallel, models/networks.py:132-135).
    Every rank first runs the iteration ALONE (no process group use), then all ranks run it together; times are the
    max over ranks.  weak_scaling_efficiency = alone / together.  Also checks that all ranks hold bit-identical
    generator and discriminator weights after the data-parallel iterations."""
    import torch.distributed as dist

    def rmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    torch.cuda.reset_peak_memory_stats(dev)
    try:    # phase 1 has no collective inside: a failure on one rank is agreed on before anyone enters phase 2
        ms_alone, launches, _, model = time_apollo_iterations(dev, crop, iters, warmup, False, rank)
        err = None
    except Exception as e:  # noqa: BLE001 - secondary measurement
        ms_alone, err = -1.0, "%s: %s" % (type(e).__name__, e)
    if rmax(1.0 if err else 0.0) > 0:
        return {"error": err or "another rank failed in the single-GPU phase"}
    try:
        ms_alone = rmax(ms_alone)
        out = {"metric": "apollo training iteration, data parallel, crop %d^3 per GPU, batch 1 per GPU" % crop,
               "crop": crop, "n_gpus": world, "iters": iters, "warmup": warmup,
               "ms_per_iter_single_gpu": ms_alone, "library_calls_per_iter": launches}
        if world == 1:
            out.update({"ms_per_iter": ms_alone, "crops_per_s": 1e3 / ms_alone, "allreduce_ms_per_iter": 0.0,
                        "weak_scaling_efficiency": 1.0})
        else:
            del model
            torch.cuda.empty_cache()
            ms_dp, _, ar, model = time_apollo_iterations(dev, crop, iters, warmup, True, rank, barrier)
            ms_dp, ar_wait = rmax(ms_dp), rmax(ar)
            ar = -rmax(-ar)       # min over ranks: the last rank to arrive measures the collective itself; the others
            #                       also measure how long they waited for it (GPU-to-GPU speed differences)
            same = True
            for opt_ in model.optimizers:
                flat = torch.cat([p.detach().reshape(-1) for p in opt_.params])
                lo, hi = flat.clone(), flat.clone()
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                same = same and bool(torch.equal(lo, hi))
            out.update({"ms_per_iter": ms_dp, "crops_per_s": world * 1e3 / ms_dp, "allreduce_ms_per_iter": ar,
                        "allreduce_ms_per_iter_incl_wait_max": ar_wait,
                        "allreduce_bytes_per_iter": 4 * sum(p.numel() for o in model.optimizers for p in o.params),
                        "weak_scaling_efficiency": ms_alone / ms_dp, "weights_identical_across_ranks": same})
        out["losses_finite"] = bool(all(np.isfinite(v) for v in model.get_current_losses().values()))
        out["peak_mem_gb"] = torch.cuda.max_memory_allocated(dev) / 2 ** 30
        return out
    except Exception as e:  # noqa: BLE001 - secondary measurement
        return {"error": "%s: %s" % (type(e).__name__, e)}


def n_cubes_of(shape):
    """util/util.py:196-215 + diceImage_dataset.py:90-92: cubes per axis = (n + overlap) // step + 1."""
    step = ROI - OVERLAP
    return int(np.prod([(n + OVERLAP) // step + 1 for n in shape]))


def workload_config(shape, gpus):
    """Names the workload; identical on both arms (the driver compares the two dicts)."""
    return {"workload": "test_dice.py unet_deconv inference, synthetic %dx%dx%d uint16 volume, dice %d overlap %d "
                        "border_cut %d, normalize_intensity (0.25, 99.75)" % (*shape, ROI, OVERLAP, BORDER),
            "cubes": n_cubes_of(shape), "parallelism": "cube-range x%d + z-slab assembly" % gpus,
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


def volume_digest(planes: np.ndarray, z0: int, rank: int, world: int, dev):
    """sha256 over the concatenated per-plane sha256 digests of the assembled uint16 volume, in z order.  Every rank
    hashes its own output planes; the 32-byte digests are gathered on rank 0.  The value does not depend on how the
    volume was split, so it must be IDENTICAL at N = 1, 2, 4, 8 (north-star: multi-GPU results bit-exact)."""
    import hashlib
    import torch.distributed as dist
    mine = [(z0 + i, hashlib.sha256(np.ascontiguousarray(planes[i]).tobytes()).digest()) for i in range(planes.shape[0])]
    if world > 1:
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(mine, gathered, dst=0)
        if rank != 0:
            return None, 0
        mine = [x for part in gathered for x in part]
    mine.sort()
    assert [z for z, _ in mine] == list(range(len(mine))), "output planes are not a partition of the volume"
    h = hashlib.sha256()
    for _, d in mine:
        h.update(d)
    return h.hexdigest(), len(mine)


def library_bar():
    """tools/library_bar.py in a subprocess (its cuDNN workspaces and autotuning must not disturb this process):
    the reference's torch.nn Unet_deconv through cuDNN on this GPU, one 140^3 cube."""
    try:
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "library_bar.py"), "--json"],
                           capture_output=True, text=True, timeout=600)
        for ln in reversed(r.stdout.strip().splitlines()):
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": "no JSON from tools/library_bar.py (rc %d): %s" % (r.returncode, r.stderr[-300:])}
    except Exception as e:  # noqa: BLE001 - secondary measurement
        return {"error": "%s: %s" % (type(e).__name__, e)}


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
    ap.add_argument("--no-library-bar", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--config5", action="store_true", help="also measure the 1024x2048x2048 volume at N < 8")
    ap.add_argument("--streams", type=int, default=1,
                    help="experimental: cube batches alternate between this many CUDA streams (1 = serial batches)")
    ap.add_argument("--no-remainder-pairs", action="store_true",
                    help="A/B measurement: disable the remainder-pair conv kernel (same results, more MMA rows)")
    args = ap.parse_args()
    shape = tuple(args.size)

    if args.impl == "reference":
        return run_reference(args, shape, guard)

    import torch.distributed as dist
    from neuroclear_b200 import _lib
    from neuroclear_b200.pipeline import DicedInference
    from neuroclear_b200 import networks
    from neuroclear_b200.unet_engine import FLOP_PER_VOXEL

    if args.no_remainder_pairs:
        _lib.load().nc_debug_set_remainder_pairs(0)
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
    peak, peak_src = measured_peaks()

    def measure(shape, steps, warmup, with_roofline):
        """One volume shape: device-resident timing, e2e timing, the digest of the e2e output."""
        pipe = DicedInference(sd, dev, ROI, OVERLAP, BORDER, normalize_intensity=True, batch=args.batch,
                              streams=args.streams)
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
            for eng in pipe.engines:
                eng.profile = [] if profile else None
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record()
            for _ in range(steps):
                fn()
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            clocks = sampler.stop()
            prof = [p for eng in pipe.engines for p in (eng.profile or [])] if profile else None
            for eng in pipe.engines:
                eng.profile = None
            if world > 1:
                t = torch.tensor([ms], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            return ms, clocks, _lib.LAUNCHES - launches0, prof

        resident = lambda: pipe.run_device(vol_dev, z0, shape)
        e2e = lambda: pipe.run_slab(slab_host, z0, shape, out_host)

        for _ in range(warmup):
            resident()
        ms, clocks, launches, prof = timed(resident, steps, profile=with_roofline)
        voxels = float(np.prod(shape))
        res = {"value": voxels * steps / (ms * 1e-3), "ms": ms, "clocks": clocks, "launches": launches, "geo": geo,
               "cubes": geo.n_cubes}

        if with_roofline:
            # live roofline of the dominant kernel family (tcgen05 implicit-GEMM conv): algorithmic FLOPs / event time
            layers = {}
            for name, flops, a, b in prof:
                t = layers.setdefault(name, [0.0, 0.0, 0])
                t[0] += flops
                t[1] += a.elapsed_time(b)
                t[2] += 1
            tot_f = sum(v[0] for v in layers.values())
            tot_ms = sum(v[1] for v in layers.values())
            achieved = tot_f / (tot_ms * 1e-3) / 1e12 if tot_ms > 0 else 0.0
            traffic, traffic_src = None, None
            try:
                with open(os.path.join(ROOT, "profiles", "conv_traffic.json")) as f:
                    tj = json.load(f)
                traffic = tj.get("dram_bytes_per_launch")
                traffic_src = "static: " + tj.get("source", "ncu --set full capture, profiles/conv_traffic.json")
            except Exception:
                pass
            res["roofline"] = {
                "bound": "tensor", "kernel": "conv3d_tc_kernel (9 Conv3d k3 layers U2..U12, per-launch average)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src, "peak_source": peak_src,
                "launches": sum(v[2] for v in layers.values()), "share_of_step": tot_ms / ms,
                "layers": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12, "ms_per_launch": v[1] / v[2]}
                           for k, v in layers.items()}}

        e2e()
        ms_e, clocks_e, _, _ = timed(e2e, steps)
        res["e2e"] = {"value": voxels * steps / (ms_e * 1e-3), "unit": UNIT,
                      "h2d_bytes_per_step": (z1 - z0) * shape[1] * shape[2] * 2,
                      "d2h_bytes_per_step": (o1 - o0) * shape[1] * shape[2] * 2,
                      "ms_per_step": ms_e / steps, "clocks": clocks_e}
        res["out_sha256"], res["out_planes"] = volume_digest(out_host.numpy(), o0, rank, world, dev)
        del pipe, vol_dev
        torch.cuda.empty_cache()
        return res

    main_res = measure(shape, args.steps, args.warmup, True)
    geo, ms, value = main_res["geo"], main_res["ms"], main_res["value"]

    c5 = None
    if (world == 8 and not args.no_config5) or args.config5:
        try:
            shape5 = (1024, 2048, 2048)       # (Z, Y, X); BASELINE.json configs[4] "2048x2048x1024"
            r5 = measure(shape5, 2, 1, False)
            c5 = {"workload": workload_config(shape5, world)["workload"], "cubes": r5["cubes"], "steps": 2, "warmup": 1,
                  "value": r5["value"], "unit": UNIT, "ms_per_step": r5["ms"] / 2, "e2e": r5["e2e"],
                  "out_sha256": r5["out_sha256"], "clocks": r5["clocks"]}
        except Exception as e:  # noqa: BLE001 - secondary measurement
            c5 = {"error": "%s: %s" % (type(e).__name__, e)}

    train_dp = None
    if not args.no_train_sample:
        train_dp = train_step_dp(dev, rank, world, barrier)

    cpu = train = lib_bar = None
    if rank == 0 and world == 1:
        if not args.no_train_sample:
            train = train_step_sample(dev, cpu_crop=0 if args.no_cpu_baseline else 64)
        if not args.no_library_bar:
            torch.cuda.empty_cache()
            lib_bar = library_bar()
        if not args.no_cpu_baseline:
            v, cores, kind, sample, _ = cpu_baseline_sample(shape)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": value / PUBLISHED_VOXELS_PER_S, "dtype": "fp16", "data": "synthetic",
            "config": workload_config(shape, world), "batch_cubes": args.batch, "streams": args.streams,
            "clocks": main_res["clocks"],
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["launches"],
            "roofline": main_res["roofline"],
            "cpu_baseline": cpu,
            "out_sha256": main_res["out_sha256"],
            "train_step": train,
            "train_step_dp": train_dp,
            "config5": c5,
            "library_bar": lib_bar,
            "tensor_pipe_frac_whole_step": FLOP_PER_VOXEL * geo.n_cubes * geo.edge ** 3 * args.steps
                                            / (ms * 1e-3) / 1e12 / peak / world,
        }
        guard.emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
