"""Drop-ins for the reference's ``data/diceImage_dataset.py`` and ``util/assemble_dice.py`` on the C ABI.

Same class names, constructor (``opt`` namespace), methods and attributes as the reference
(diceImage_dataset.py:9-79, assemble_dice.py:11-244), but the volume lives in HBM: cubes are cut by
``nc_dice_extract_u16`` (zero pad + reflect border + /65535 fused, no padded copies) and the assembly runs
``nc_blend_gather_f32`` / radix-select percentile / ``nc_rescale_u16_crop`` on the device.
"""
from __future__ import annotations

import os
import struct
from collections import OrderedDict
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib
from ._lib import I3, U4, NeuroclearError, call, i64, ptr, stream_ptr


# ---------------------------------------------------------------------------------------------- geometry
@dataclass(frozen=True)
class DiceGeometry:
    size: tuple      # original (Z, Y, X)
    padded: tuple    # util.pad_for_dicing result
    steps: tuple     # cubes per axis (z, y, x)
    roi: int
    overlap: int
    border: int

    @property
    def step(self):
        return self.roi - self.overlap

    @property
    def edge(self):
        return self.roi + 2 * self.border

    @property
    def n_cubes(self):
        return self.steps[0] * self.steps[1] * self.steps[2]

    def c_arrays(self):
        return I3(*self.size), I3(*self.padded), I3(*self.steps)


def dice_geometry(size, roi: int, overlap: int, border: int = 0) -> DiceGeometry:
    """util/util.py:196-215 + diceImage_dataset.py:90-92, computed by the library (nc_dice_geometry)."""
    padded, steps = I3(), I3()
    n = _lib.load().nc_dice_geometry(I3(*[int(s) for s in size]), roi, overlap, padded, steps)
    if n < 0:
        raise NeuroclearError(_lib.load().nc_last_error().decode())
    return DiceGeometry(tuple(int(s) for s in size), tuple(padded), tuple(steps), roi, overlap, border)


# ---------------------------------------------------------------------------------------------- device ops
def dice_extract(vol_dev: torch.Tensor, vol_z0: int, geo: DiceGeometry, cube_begin: int, count: int,
                 out: torch.Tensor = None) -> torch.Tensor:
    """vol_dev: uint16 / uint8 CUDA planes [vol_z0, vol_z0+n) of the original volume -> float32 (count, E, E, E)."""
    if not vol_dev.is_cuda:
        raise NeuroclearError("dice_extract: the volume must be in device memory (no CPU fallback)")
    assert vol_dev.dtype in (torch.uint16, torch.uint8) and vol_dev.is_contiguous()
    e = geo.edge
    if out is None:
        out = torch.empty((count, e, e, e), dtype=torch.float32, device=vol_dev.device)
    size, padded, steps = geo.c_arrays()
    fn = "nc_dice_extract_u16" if vol_dev.dtype == torch.uint16 else "nc_dice_extract_u8"
    call(fn, ptr(vol_dev), vol_z0, vol_dev.shape[0], size, padded, steps, geo.roi, geo.overlap, geo.border,
         i64(cube_begin), count, ptr(out), stream_ptr())
    return out


def blend_gather(pieces, piece_off, piece_z0, geo: DiceGeometry, z0: int, nz: int, out=None):
    if out is None:
        out = torch.empty((nz, geo.padded[1], geo.padded[2]), dtype=torch.float32, device=pieces.device)
    _, padded, steps = geo.c_arrays()
    call("nc_blend_gather_f32", ptr(pieces), ptr(piece_off), ptr(piece_z0), padded, steps, geo.roi, geo.overlap,
         z0, nz, ptr(out), stream_ptr())
    return out


def percentile_ranks(n_total: int, sat_level):
    """numpy 'linear' method index arithmetic (np.percentile -> _quantile): virtual index (n-1)*q/100 in fp64."""
    q = np.true_divide(np.asarray(sat_level, dtype=np.float64), 100)
    virt = (n_total - 1) * q
    prev = np.floor(virt)
    gamma = virt - prev
    lo = prev.astype(np.int64)
    hi = np.minimum(lo + 1, n_total - 1)
    return [int(lo[0]), int(hi[0]), int(lo[1]), int(hi[1])], float(gamma[0]), float(gamma[1])


class PercentileSelect:
    """Exact np.percentile(vis, (p_lo, p_hi)) on the device; optionally all-reduces histograms over a group."""

    def __init__(self, device):
        self.state = torch.zeros(64, dtype=torch.uint8, device=device)          # SelectState
        self.hist = torch.zeros(4 * 4096, dtype=torch.int64, device=device)     # uint64 [4][4096]
        self.out64 = torch.zeros(2, dtype=torch.float64, device=device)
        self.norm3 = torch.zeros(3, dtype=torch.float32, device=device)

    def run(self, data: torch.Tensor, n_total: int, sat_level, group=None, distributed=False):
        ranks, t_lo, t_hi = percentile_ranks(n_total, sat_level)
        s = stream_ptr()
        self.hist.zero_()
        call("nc_select_init", U4(*ranks), ptr(self.state), s)
        for p in range(3):
            call("nc_select_histogram", ptr(data), i64(data.numel()), p, ptr(self.state), ptr(self.hist), s)
            if distributed:
                torch.distributed.all_reduce(self.hist, group=group)
            call("nc_select_update", p, ptr(self.state), ptr(self.hist), s)
        call("nc_percentile_lerp", ptr(self.state), t_lo, t_hi, ptr(self.out64), ptr(self.norm3), s)
        return self.norm3, self.out64


def rescale_u16_crop(vis: torch.Tensor, vis_z0: int, geo: DiceGeometry, norm3, z_begin: int, z_count: int, out=None,
                     dtype=torch.uint16):
    """rescale + cast + un-pad; dtype uint16 (default) or uint8 (--data_type)."""
    if out is None:
        out = torch.empty((z_count, geo.size[1], geo.size[2]), dtype=dtype, device=vis.device)
    size, padded, _ = geo.c_arrays()
    fn = "nc_rescale_u16_crop" if out.dtype == torch.uint16 else "nc_rescale_u8_crop"
    call(fn, ptr(vis), vis_z0, padded, size, ptr(norm3), z_begin, z_count, ptr(out), stream_ptr())
    return out


# ---------------------------------------------------------------------------------------------- dataset
def _load_volume(path):
    """skimage.io.imread of the reference (diceImage_dataset.py:35): .npy, or a multi-page TIFF / BigTIFF through
    neuroclear_b200.volume_io (uncompressed pages, read straight into one array); compressed or tiled TIFFs fall back
    to cv2's codec."""
    if os.path.isdir(path):
        names = sorted(f for f in os.listdir(path) if not f.startswith(".") and
                       f.lower().endswith((".npy", ".tif", ".tiff")))
        if not names:
            raise FileNotFoundError("no .npy/.tif volume under %s" % path)
        path = os.path.join(path, names[0])
    if path.lower().endswith(".npy"):
        return np.load(path)
    from . import volume_io
    try:
        return volume_io.read_volume(path)
    except (NeuroclearError, KeyError, struct.error):       # not a plain strip TIFF: let cv2's codec try
        pass
    import cv2
    ok, pages = cv2.imreadmulti(path, flags=cv2.IMREAD_UNCHANGED)
    if not ok:
        raise IOError("cannot read %s" % path)
    return np.stack(pages)


class DiceImageDataSet:
    """reference data/diceImage_dataset.py:9-79 — one volume, one cube per index (x fastest, then y, then z)."""

    @staticmethod
    def modify_commandline_options(parser, is_train=False):
        parser.add_argument("--overlap", type=int, default=0,
                            help="set the size of overlapping region when dicing the dataset.")
        parser.add_argument("--border_cut", default=0, type=int,
                            help="specify how much border you want to remove in a cube-by-cube inference.")
        return parser

    def __init__(self, opt, volume: np.ndarray = None):
        self.opt = opt
        self.roi_size = opt.dice_size[0]
        self.overlap = opt.overlap
        self.border_cut = opt.border_cut
        vol = volume if volume is not None else _load_volume(opt.dataroot)
        if vol.dtype not in (np.uint16, np.uint8):
            raise NeuroclearError("the dice path takes uint16 or uint8 volumes (--data_type)")
        if "addColorChannel" not in getattr(opt, "preprocess", "addColorChannel"):
            raise NeuroclearError("DiceImageDataSet expects --preprocess addColorChannel as in the reference README")
        gpu_ids = getattr(opt, "gpu_ids", [0])
        self.device = torch.device("cuda", gpu_ids[0] if gpu_ids else 0)
        self.geo = dice_geometry(vol.shape, self.roi_size, self.overlap, self.border_cut)
        self.image_size_original = tuple(vol.shape)
        self.image_size = self.geo.padded
        self.host_volume = vol
        self._dev = None
        self.A_path = getattr(opt, "dataroot", "")

    def device_volume(self) -> torch.Tensor:
        if self._dev is None:
            self._dev = torch.from_numpy(np.ascontiguousarray(self.host_volume)).to(self.device)
        return self._dev

    def cubes(self, begin: int, count: int, out=None) -> torch.Tensor:
        """Batched accessor used by the fast path: float32 (count, E, E, E) on the device."""
        with torch.cuda.device(self.device):
            return dice_extract(self.device_volume(), 0, self.geo, begin, count, out)

    def __getitem__(self, index):
        if index < 0 or index >= len(self):
            raise IndexError(index)
        return {"A": self.cubes(index, 1), "A_paths": str(index)}   # (1, E, E, E): colour channel added

    def __len__(self):
        return self.geo.n_cubes

    def __iter__(self):
        """Iterating yields batch-1 items like the reference's DataLoader: A is (1, 1, E, E, E)."""
        for i in range(len(self)):
            item = self[i]
            yield {"A": item["A"][None], "A_paths": [item["A_paths"]]}

    def shape(self):
        return self.geo.steps

    def size(self):
        return self.image_size

    def size_original(self):
        return self.image_size_original


# ---------------------------------------------------------------------------------------------- assembly
class Assemble_Dice:
    """reference util/assemble_dice.py:11-244 — queue border-cut cubes, blend, normalise, cast, un-pad."""

    def __init__(self, opt, dataset: DiceImageDataSet = None):
        ds = dataset if dataset is not None else DiceImageDataSet(opt)
        self.geo = ds.geo
        self.device = ds.device
        self.image_size_original = ds.size_original()
        self.image_size = ds.size()
        self.border_cut = opt.border_cut
        self.roi_size = opt.dice_size[0]
        self.overlap = opt.overlap
        self.step = self.roi_size - self.overlap
        self.z_steps, self.y_steps, self.x_steps = self.geo.steps
        self.visual_names = ["real", "fake"]
        self.imtype = opt.data_type
        if self.imtype not in ("uint16", "uint8"):
            raise NeuroclearError("Assemble_Dice (B200): --data_type must be uint16 or uint8")
        self.skip_real = opt.skip_real
        self.histogram_match = bool(getattr(opt, "histogram_match", False))       # assemble_dice.py:35,150-151
        if self.histogram_match:
            print("We will match the histograms of output sub-volumes with input sub-volumes.")
        self.normalize_intensity = opt.normalize_intensity
        if self.normalize_intensity:
            self.p1, self.p99 = opt.sat_level
        if self.border_cut < 1:
            raise NeuroclearError("border_cut must be >= 1 (the reference's [bc:-bc] crop is empty for 0)")
        if self.overlap <= 0:
            raise NeuroclearError("overlap must be > 0 (the reference assembles all zeros otherwise)")
        self.len_cube_queue = self.geo.n_cubes
        self.visual_ret = OrderedDict()
        self.snapDict = OrderedDict()
        self.cube_queue = OrderedDict()
        self.percentiles = {}
        self._count = {}
        r = self.roi_size
        for name in self.visual_names:
            if self.skip_real and name == "real":
                continue
            # match_histograms returns float64 and the reference queues it as such: the blend then adds in float64
            dt = torch.float64 if (self.histogram_match and name == "fake") else torch.float32
            self.cube_queue[name] = torch.empty((self.len_cube_queue, r, r, r), dtype=dt, device=self.device)
            self._count[name] = 0
        self._hm_scratch = None

    def indexTo3DIndex(self, index):
        x = index % self.x_steps
        y = (index % (self.x_steps * self.y_steps)) // self.x_steps
        z = index // (self.x_steps * self.y_steps)
        return z, y, x

    def indexToCoordinates(self, index):
        z, y, x = self.indexTo3DIndex(index)
        return z * self.step, y * self.step, x * self.step

    # ---- flip test-time augmentation helpers (assemble_dice.py:79-128): pure tensor re-indexing + a 4-way mean,
    # device-agnostic torch ops (not on the hot path; test_dice.py does not call them)
    def varycubeinput(self, input):
        """[input, input flipped along z, along y, along x] as dataset-style dicts"""
        names = list(input.keys())
        vol, path = input[names[0]], input[names[1]]
        out = [input]
        for axis in range(2, vol.dim()):
            d = OrderedDict()
            d[names[0]] = vol.flip(axis)
            d[names[1]] = path
            out.append(d)
        return out

    def combinecube(self, visual_list):
        """un-flip the outputs of varycubeinput's copies and average them with the unflipped one"""
        keys = list(visual_list[0].keys())
        ndim = visual_list[0][keys[0]].dim()
        out = OrderedDict()
        for name in keys:
            stack = [visual_list[0][name]] + [v[name].flip(2 + i) for i, v in enumerate(visual_list[1:ndim - 1])]
            out[name] = torch.mean(torch.stack(stack, dim=0), dim=0)
        return out

    def addToStack(self, cube):
        """cube: dict with 'real' and 'fake' (1,1,E,E,E) tensors, as BaseModel.get_current_visuals() returns."""
        bc = self.border_cut
        cut = {}
        for name in self.visual_names:
            t = cube[name]                       # both keys are required, like the reference (:132-133)
            if not t.is_cuda:
                raise NeuroclearError("Assemble_Dice (B200) takes device tensors; there is no CPU assembly path")
            c = t.reshape(t.shape[-3:])[bc:-bc, bc:-bc, bc:-bc]
            assert tuple(c.shape) == (self.roi_size,) * 3, "the cube dimensions are invalid."
            cut[name] = c
        for name in self.visual_names:
            if self.skip_real and name == "real":
                continue
            i = self._count[name]
            if i >= self.len_cube_queue:
                raise NeuroclearError("more cubes added than the volume has")
            if self.histogram_match and name == "fake":      # :150-151, matched against the input cube
                self.match_histograms(cut["fake"], cut["real"], out=self.cube_queue[name][i])
            else:
                self.cube_queue[name][i].copy_(cut[name])
            self._count[name] = i + 1

    def match_histograms(self, fake, real, out=None):
        """skimage.exposure.match_histograms(fake, real) of one cube on the device -> float64, same shape."""
        with torch.cuda.device(self.device):
            f = fake.to(torch.float32).contiguous()
            r = real.to(torch.float32).contiguous()
            n = f.numel()
            if r.numel() != n:
                raise NeuroclearError("match_histograms: image and reference must have the same size")
            need = _lib.load().nc_hist_match_scratch_bytes(n)
            if self._hm_scratch is None or self._hm_scratch.numel() < need:
                self._hm_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
            if out is None:
                out = torch.empty(f.shape, dtype=torch.float64, device=self.device)
            call("nc_hist_match_f32", ptr(f), ptr(r), i64(n), ptr(self._hm_scratch), i64(self._hm_scratch.numel()),
                 ptr(out), stream_ptr())
        return out

    def queue_view(self, name="fake"):
        """Device queue (n_cubes, roi, roi, roi): the fast path lets the network head write into it directly."""
        return self.cube_queue[name]

    def mark_filled(self, name="fake"):
        self._count[name] = self.len_cube_queue

    def assemble_all(self):
        g = self.geo
        r3 = self.roi_size ** 3
        with torch.cuda.device(self.device):
            off = torch.arange(g.n_cubes, dtype=torch.int64, device=self.device) * r3
            z0 = torch.zeros(g.n_cubes, dtype=torch.int32, device=self.device)
            sel = PercentileSelect(self.device) if self.normalize_intensity else None
            for name, queue in self.cube_queue.items():
                if self._count[name] != self.len_cube_queue:
                    raise NeuroclearError("assemble_all: %d of %d cubes queued for '%s'" %
                                          (self._count[name], self.len_cube_queue, name))
                if queue.dtype == torch.float64:
                    vis = torch.empty(g.padded, dtype=torch.float32, device=self.device)
                    _, padded, steps = g.c_arrays()
                    call("nc_blend_gather_f64", ptr(queue), ptr(off), ptr(z0), padded, steps, g.roi, g.overlap, 0,
                         g.padded[0], ptr(vis), stream_ptr())
                else:
                    vis = blend_gather(queue.view(-1), off, z0, g, 0, g.padded[0])
                norm3 = None
                if self.normalize_intensity:
                    norm3, p64 = sel.run(vis, vis.numel(), (self.p1, self.p99))
                    self.percentiles[name] = p64
                out = rescale_u16_crop(vis, 0, g, norm3, 0, g.size[0],
                                       dtype=torch.uint16 if self.imtype == "uint16" else torch.uint8)
                del vis
                self.visual_ret[name] = out.cpu().numpy()
                if name in self.percentiles:
                    self.percentiles[name] = tuple(self.percentiles[name].cpu().tolist())

    def getSnapshots(self, index, slice_axis=2):
        for name in self.visual_ret:
            v = self.visual_ret[name]
            self.snapDict[name] = v[index] if slice_axis == 0 else v[:, index] if slice_axis == 1 else v[:, :, index]
        return self.snapDict

    def getDict(self):
        return self.visual_ret

    def getCubeQueue(self):
        return self.cube_queue
