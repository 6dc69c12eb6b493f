"""Imports the UNMODIFIED reference with stub modules for its absent third-party imports.
Test infrastructure — see oracle/__init__.py.

Two homes for the reference: /root/reference (sources; the build container only — oracle/make_golden.py pins the
restatements against it and writes tests/golden/) and oracle/_ref/neuroclear.zip (the same modules byte-compiled by
oracle/build_ref.py; git-ignored, travels to the GPU box).  bench.py / the -m gpu tests pass compiled=True: they
never read /root/reference.

Stubs: skimage{,.io,.exposure,.transform}, np.float.  skimage.exposure.rescale_intensity is
routed to oracle.assemble.rescale_intensity (the one unpinned piece).
"""
from __future__ import annotations

import os
import sys
import types
from argparse import Namespace

import numpy as np

REFERENCE_ROOT = "/root/reference"
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "neuroclear.zip")
_VOLUME = {"array": None}


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "models"))


def compiled_available() -> bool:
    return os.path.exists(COMPILED_ROOT)


def install(compiled: bool = False):
    """Put the reference on sys.path behind stubs; idempotent.  compiled=True: only oracle/_ref (never the sources)."""
    root = COMPILED_ROOT if compiled or not available() else REFERENCE_ROOT
    if "skimage" not in sys.modules:
        from . import assemble as _asm
        if not hasattr(np, "float"):
            np.float = float  # base_dataset.py:293, util/util.py:35 use the removed alias
        sk = types.ModuleType("skimage")
        io = types.ModuleType("skimage.io")
        ex = types.ModuleType("skimage.exposure")
        tr = types.ModuleType("skimage.transform")
        io.imread = lambda path: _VOLUME["array"]
        ex.rescale_intensity = lambda image, in_range: _asm.rescale_intensity(image, in_range)
        ex.match_histograms = lambda a, b: a
        sk.io, sk.exposure, sk.transform = io, ex, tr
        sys.modules.update({"skimage": sk, "skimage.io": io, "skimage.exposure": ex, "skimage.transform": tr})
    if root not in sys.path:
        sys.path.insert(0, root)
    return root


def set_volume(vol: np.ndarray):
    _VOLUME["array"] = vol


def dice_opt(dataroot: str, roi=120, overlap=15, border=10, normalize_intensity=True, sat_level=(0.25, 99.75)):
    return Namespace(dataroot=dataroot, dice_size=[roi] * 3, overlap=overlap, border_cut=border,
                     preprocess="addColorChannel", image_dimension=3, dataset_mode="diceImage", data_type="uint16",
                     skip_real=True, histogram_match=False, normalize_intensity=normalize_intensity,
                     sat_level=list(sat_level), batch_size=1, serial_batches=True, num_threads=0,
                     max_dataset_size=float("inf"))


def make_dataroot(tmpdir: str) -> str:
    """The reference lists files by extension (data/image_folder.py) and then calls our imread stub."""
    os.makedirs(tmpdir, exist_ok=True)
    open(os.path.join(tmpdir, "volume.tif"), "wb").close()
    return tmpdir
