"""Whole-volume diced inference (the loop of the reference's test_dice.py:70-121) as one device-resident pipeline:

    host uint16 volume --H2D--> dice (fused pad/reflect/normalise) --> Unet_deconv on batches of cubes
      --> border-cut outputs in an HBM queue --> [piece exchange over NVLink when sharded]
      --> gather-blend --> exact percentile (radix select, histogram all-reduce) --> rescale/cast/un-pad --D2H--> host

One process drives one GPU.  With a process group, cubes are split into balanced contiguous index ranges and the
assembled volume into balanced z-slabs (neuroclear_b200.sharding); results are value-identical for any world
size because the blend always runs in ascending cube index on the slab owner.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from . import sharding
from ._lib import NeuroclearError, call, i64
from .dicing import (PercentileSelect, blend_gather, dice_extract, dice_geometry, rescale_u16_crop)
from .unet_engine import UnetDeconvEngine


class ChunkedUpload:
    """A pinned host slab (planes [z0, z0 + n) of the volume) -> device slab, chunk by chunk on a copy stream.
    axis 0: chunks are z-plane ranges (contiguous copies; what a file reader produces); axis 1: chunks are y-row ranges
    of ALL planes (one pitched cudaMemcpy2DAsync each) — cubes are ordered x -> y -> z, so when a rank's cubes span
    only one or two z-layers (8 GPUs) the first batches need every plane but only the first rows.
    ensure(e): every chunk below coordinate e (a plane index for axis 0, a row index for axis 1) has been produced by
    `source` (reader thread) and its H2D copy is enqueued; wait_for(e): the CURRENT stream also waits for them."""

    def __init__(self, slab_host, vol_dev, z0, copy_stream, chunks, source=None, axis=0):
        import queue
        import threading
        self.host, self.dev, self.stream, self.axis = slab_host, vol_dev, copy_stream, axis
        self.z0 = z0 if axis == 0 else 0
        n = slab_host.shape[axis]
        step = max(1, -(-n // chunks))
        self.bounds = [(a, min(n, a + step)) for a in range(0, n, step)]
        self.events = []             # one per enqueued chunk
        self.waited = {}             # stream handle -> number of chunk events that stream already waits for
        self.error = None
        self.filled = None
        if source is not None:
            self.filled = queue.Queue()

            def reader():
                try:
                    for a, b in self.bounds:
                        source(a, b)
                        self.filled.put((a, b))
                except BaseException as e:  # noqa: BLE001 - re-raised on the consumer side
                    self.error = e
                    self.filled.put(None)
            self.thread = threading.Thread(target=reader, daemon=True)
            self.thread.start()

    def ensure(self, z_end):
        while len(self.events) < len(self.bounds) and \
                (not self.events or self.z0 + self.bounds[len(self.events) - 1][1] < z_end):
            a, b = self.bounds[len(self.events)]
            if self.filled is not None:
                got = self.filled.get()
                if got is None:
                    raise self.error
            with torch.cuda.stream(self.stream):
                if self.axis == 0:
                    self.dev[a:b].copy_(self.host[a:b], non_blocking=True)
                else:       # rows [a, b) of every plane: one pitched copy (torch would stage a non-contiguous slice)
                    nz, ny, nx = self.host.shape
                    row = nx * self.host.element_size()
                    call("nc_memcpy2d_h2d_async", C.c_void_p(self.dev.data_ptr() + a * row),
                         C.c_void_p(self.host.data_ptr() + a * row), i64(ny * row), i64((b - a) * row), i64(nz),
                         C.c_void_p(self.stream.cuda_stream))
                ev = torch.cuda.Event()
                ev.record(self.stream)
            self.events.append(ev)

    def wait_for(self, z_end):
        """the CURRENT stream waits for every chunk enqueued so far (each stream keeps its own cursor)"""
        self.ensure(z_end)
        cur = torch.cuda.current_stream()
        done = self.waited.get(cur.cuda_stream, 0)
        while done < len(self.events):
            cur.wait_event(self.events[done])
            done += 1
        self.waited[cur.cuda_stream] = done


class DicedInference:
    def __init__(self, state_dict, device, roi=120, overlap=15, border=10, normalize_intensity=True,
                 sat_level=(0.25, 99.75), batch=4, group=None, distributed=None, streams=1):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise NeuroclearError("DicedInference needs a CUDA device (no CPU fallback)")
        if border < 1 or overlap < 1 or 2 * overlap > roi:
            raise NeuroclearError("need border_cut >= 1 and 0 < overlap <= roi - overlap (as the reference's assembly)")
        self.roi, self.overlap, self.border = roi, overlap, border
        self.normalize_intensity, self.sat_level = normalize_intensity, tuple(sat_level)
        self.batch = batch
        self.group = group
        self.distributed = dist.is_initialized() if distributed is None else distributed
        self.rank = dist.get_rank(group) if self.distributed else 0
        self.world = dist.get_world_size(group) if self.distributed else 1
        #: EXPERIMENTAL (default 1 = strictly serial batches).  With streams = 2 cube batches alternate between two
        #: CUDA streams, each with its own engine workspace: while one batch sits in a tensor-bound convolution the
        #: other batch's memory-bound kernels (InstanceNorm apply / pool, head, dice — a quarter of a batch's time)
        #: run in its shadow.  Values are unchanged (batches are independent; same out_sha256), but on a power-capped
        #: B200 the gain is only 1 % (3025 -> 2994 ms on one box: the overlap raises the average power and the clock
        #: drops 1590 -> 1540 MHz), per-launch CUDA-event times stop being kernel times, and a run with three streams
        #: died with an unspecified launch failure that was not tracked down — so it is off by default.
        self.n_streams = max(1, int(streams))
        with torch.cuda.device(self.device):
            self.engines = [UnetDeconvEngine(self.device) for _ in range(self.n_streams)]
            for e in self.engines:
                e.load_state_dict(state_dict)
            self.engine = self.engines[0]
            self._streams = None
            self.select = PercentileSelect(self.device)
        self._plan_key = None
        self._copy_stream = None
        self.last = {}
        self.out_dtype = torch.uint16     # follows the input volume's dtype (--data_type uint16 | uint8)

    # ---------------------------------------------------------------- static plan for a volume shape
    def plan(self, size):
        key = tuple(size)
        if self._plan_key == key:
            return self._plan
        geo = dice_geometry(size, self.roi, self.overlap, self.border)
        cube_ranges = sharding.balanced_ranges(geo.n_cubes, self.world)
        slab_ranges = sharding.balanced_ranges(geo.padded[0], self.world)
        c0, c1 = cube_ranges[self.rank]
        p = dict(geo=geo, cube_ranges=cube_ranges, slab_ranges=slab_ranges, cubes=(c0, c1),
                 in_planes=sharding.input_plane_range(geo, c0, c1), slab=slab_ranges[self.rank])
        if self.world > 1:
            p["pieces"] = sharding.plan_pieces(geo, cube_ranges, slab_ranges)
            off, z0, _, total = sharding.piece_tables(geo, p["pieces"], self.rank, self.device)
            p["piece_off"], p["piece_z0"], p["recv_total"] = off, z0, total
        else:
            r3 = self.roi ** 3
            p["piece_off"] = torch.arange(geo.n_cubes, dtype=torch.int64, device=self.device) * r3
            p["piece_z0"] = torch.zeros(geo.n_cubes, dtype=torch.int32, device=self.device)
        s0, s1 = p["slab"]
        p["out_planes"] = (min(s0, geo.size[0]), min(s1, geo.size[0]))   # un-padded part of this rank's slab
        self._plan_key, self._plan = key, p
        return p

    # ---------------------------------------------------------------- stages (all async on the current stream)
    def upload(self, volume_host: torch.Tensor, plan):
        """volume_host: uint16 (Z,Y,X) host tensor (pinned for async copies).  Only this rank's planes move."""
        z0, z1 = plan["in_planes"]
        return volume_host[z0:z1].to(self.device, non_blocking=True), z0

    def upload_chunked(self, slab_host: torch.Tensor, plan, chunks=16, source=None):
        """H2D of this rank's input planes on a side stream, in z-chunks, so that the first cube batches start while
        the rest of the slab is still crossing PCIe.  Returns (device slab, ChunkedUpload): infer_cubes asks it for the
        planes a batch reads.  `source` (optional) is a callable (a, b) -> None that FILLS slab_host[a:b] (e.g. reads
        the planes from a file); it runs on a reader thread one chunk ahead, so disk, PCIe and compute overlap."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(self.device)
        z0, z1 = plan["in_planes"]
        vol_dev = torch.empty(slab_host.shape, dtype=slab_host.dtype, device=self.device)
        self._copy_stream.wait_stream(torch.cuda.current_stream())      # the allocation above is stream-ordered
        vol_dev.record_stream(self._copy_stream)      # allocated on the compute stream, written on the copy stream
        # from host memory: y-row chunks (the first cube rows start at once); from a reader: z-plane chunks (file order)
        axis = 1 if (source is None and slab_host.is_pinned() and slab_host.is_contiguous()) else 0
        up = ChunkedUpload(slab_host, vol_dev, z0, self._copy_stream, chunks, source, axis=axis)
        if source is None:
            up.ensure(1 << 30)                         # everything is already in host memory: enqueue all copies now
        return vol_dev, up

    def infer_cubes(self, vol_dev, vol_z0, plan, queue=None, ready=None):
        self.out_dtype = vol_dev.dtype
        geo = plan["geo"]
        c0, c1 = plan["cubes"]
        r, e = self.roi, geo.edge
        if queue is None:
            queue = torch.empty((c1 - c0, r, r, r), dtype=torch.float32, device=self.device)
        nbmax = min(self.batch, max(c1 - c0, 1))
        n_batches = -(-(c1 - c0) // nbmax)
        ns = min(self.n_streams, max(n_batches, 1))
        xbufs = [torch.empty((nbmax, e, e, e), dtype=torch.float32, device=self.device) for _ in range(ns)]
        main = torch.cuda.current_stream()
        if ns > 1:
            if self._streams is None:
                self._streams = [torch.cuda.Stream(self.device) for _ in range(self.n_streams)]
            lanes = self._streams[:ns]
            for st in lanes:
                st.wait_stream(main)              # vol_dev / queue / xbufs were allocated (and uploaded) on `main`
        else:
            lanes = [main]
        for bi, b0 in enumerate(range(c0, c1, nbmax)):
            nb = min(nbmax, c1 - b0)
            k = bi % ns
            with torch.cuda.stream(lanes[k]):
                if ready is not None:   # the uploaded chunks this batch's cubes read (border and reflection included)
                    ready.wait_for(sharding.input_plane_range(geo, b0, b0 + nb)[1] if ready.axis == 0
                                   else sharding.input_row_end(geo, b0, b0 + nb))
                x = dice_extract(vol_dev, vol_z0, geo, b0, nb, out=xbufs[k][:nb])
                self.engines[k].forward(x, crop=self.border, out=queue[b0 - c0:b0 - c0 + nb], nb_cap=nbmax)
        if ns > 1:
            for st in lanes:
                main.wait_stream(st)
        return queue

    def assemble(self, queue, plan):
        geo = plan["geo"]
        if self.world > 1:
            pieces = sharding.exchange_pieces(queue, plan["cubes"][0], geo, plan["pieces"], self.rank, self.group)
        else:
            pieces = queue.view(-1)
        s0, s1 = plan["slab"]
        vis = blend_gather(pieces, plan["piece_off"], plan["piece_z0"], geo, s0, s1 - s0)
        norm3 = None
        if self.normalize_intensity:
            n_total = geo.padded[0] * geo.padded[1] * geo.padded[2]
            norm3, self.last["percentiles"] = self.select.run(vis, n_total, self.sat_level, self.group,
                                                              distributed=self.world > 1)
        o0, o1 = plan["out_planes"]
        return rescale_u16_crop(vis, s0, geo, norm3, o0, o1 - o0, dtype=self.out_dtype)

    # ---------------------------------------------------------------- public API
    def run_device(self, vol_dev, vol_z0, size, ready=None):
        """Inputs already in HBM (or arriving: `ready` from upload_chunked) -> this rank's uint16 output planes in HBM."""
        with torch.cuda.device(self.device):
            plan = self.plan(size)
            queue = self.infer_cubes(vol_dev, vol_z0, plan, ready=ready)
            return self.assemble(queue, plan)

    def run_slab(self, slab_host: torch.Tensor, slab_z0: int, size, out_host: torch.Tensor = None):
        """Sharded entry point: `slab_host` holds only this rank's input planes plan(size)['in_planes'] of a (Z,Y,X)
        volume (pinned uint16 host tensor).  H2D, the whole path and D2H of this rank's output slab."""
        with torch.cuda.device(self.device):
            plan = self.plan(tuple(size))
            z0, z1 = plan["in_planes"]
            if slab_z0 != z0 or slab_host.shape[0] != z1 - z0:
                raise NeuroclearError("run_slab: expected input planes [%d, %d)" % (z0, z1))
            vol_dev, ready = self.upload_chunked(slab_host, plan)
            out_dev = self.run_device(vol_dev, z0, tuple(size), ready=ready)
            if out_host is None:
                out_host = torch.empty(out_dev.shape, dtype=out_dev.dtype, pin_memory=True)
            out_host.copy_(out_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out_host.numpy(), plan["out_planes"]

    def run_file(self, in_path, out_path, chunks=16):
        """test_dice.py end to end on files: read the (multi-page, uncompressed) TIFF volume `in_path`, run the path,
        write the result TIFF `out_path` (skimage.io.imread / tifffile.imsave in the reference, diceImage_dataset.py:35,
        test_dice.py:151).  Streaming: a reader thread fills the pinned input slab chunk by chunk while earlier chunks
        cross PCIe and the first cube batches already run; the result leaves in plane chunks whose D2H copies overlap
        the file writes of the chunks before them.  Sharded: every rank reads only the input planes it needs and
        writes its own output slab into the shared file; rank 0 adds the header and the page directory."""
        from . import volume_io
        tv = volume_io.TiffVolume(in_path)
        size = tv.shape
        with torch.cuda.device(self.device):
            plan = self.plan(size)
            z0, z1 = plan["in_planes"]
            dt = torch.uint16 if tv.dtype.itemsize == 2 else torch.uint8
            slab = torch.empty((z1 - z0,) + size[1:], dtype=dt).pin_memory()
            slab_np = slab.numpy()
            vol_dev, ready = self.upload_chunked(
                slab, plan, chunks, source=lambda a, b: tv.read(z0 + a, z0 + b, out=slab_np[a:b]))
            out_dev = self.run_device(vol_dev, z0, size, ready=ready)
            o0, o1 = plan["out_planes"]
            layout = volume_io.TiffLayout(size, np.dtype("u2") if out_dev.dtype == torch.uint16 else np.dtype("u1"))
            if self.rank == 0 and os.path.exists(out_path):
                os.remove(out_path)
            if self.world > 1:
                dist.barrier(self.group)
            # D2H in chunks through two pinned buffers; chunk k is written while chunk k + 1 is copied
            n = o1 - o0
            step = max(1, -(-n // chunks))
            bufs = [torch.empty((step,) + size[1:], dtype=out_dev.dtype).pin_memory() for _ in range(2)]
            evs = [None, None]
            self._copy_stream.wait_stream(torch.cuda.current_stream())
            bounds = [(a, min(n, a + step)) for a in range(0, n, step)]

            def issue(k):
                a, b = bounds[k]
                with torch.cuda.stream(self._copy_stream):
                    bufs[k % 2][:b - a].copy_(out_dev[a:b], non_blocking=True)
                    evs[k % 2] = torch.cuda.Event()
                    evs[k % 2].record(self._copy_stream)
            if bounds:
                issue(0)
            for k, (a, b) in enumerate(bounds):
                evs[k % 2].synchronize()
                if k + 1 < len(bounds):
                    issue(k + 1)
                volume_io.write_planes(out_path, layout, bufs[k % 2][:b - a].numpy(), o0 + a,
                                       write_directory=(self.rank == 0 and k == 0))
            if not bounds and self.rank == 0:
                volume_io.write_planes(out_path, layout, np.empty((0,) + size[1:], dtype=layout.dtype), 0,
                                       write_directory=True)
            out_dev.record_stream(self._copy_stream)
        if self.world > 1:
            dist.barrier(self.group)
        return layout

    def run(self, volume, out_host: torch.Tensor = None):
        """volume: uint16 (Z,Y,X) numpy array or host tensor.  Returns (planes, (z_begin, z_end)): this rank's
        slab of the assembled uint16 volume as a host numpy array (the whole volume when not sharded)."""
        if isinstance(volume, np.ndarray):
            volume = torch.from_numpy(volume)
        if volume.dtype not in (torch.uint16, torch.uint8) or volume.dim() != 3:
            raise NeuroclearError("DicedInference.run expects a (Z,Y,X) uint16 or uint8 volume")
        with torch.cuda.device(self.device):
            plan = self.plan(tuple(volume.shape))
            z0, z1 = plan["in_planes"]
            if volume.is_pinned():
                vol_dev, ready = self.upload_chunked(volume[z0:z1], plan)
            else:
                (vol_dev, _), ready = self.upload(volume, plan), None
            out_dev = self.run_device(vol_dev, z0, tuple(volume.shape), ready=ready)
            if out_host is None:
                out_host = torch.empty(out_dev.shape, dtype=out_dev.dtype, pin_memory=True)
            out_host.copy_(out_dev, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        return out_host.numpy(), plan["out_planes"]
