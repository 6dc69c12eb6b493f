"""Volume I/O either side of the hot path: the reference reads its input with
skimage.io.imread (data/diceImage_dataset.py:35) and writes the result with tifffile.imsave (test_dice.py:151) —
whole 1.5-9 GB arrays through pageable memory.  This module reads / writes the same files — multi-page, uncompressed,
single-channel 8- or 16-bit TIFF / BigTIFF, one z-plane per page, as tifffile.imsave(path, volume) produces — plane
range by plane range, straight into / out of caller-provided (pinned) host buffers:

* a rank of a sharded run reads only the input planes it needs (sharding.input_plane_range) and writes only its
  output slab: the file layout is fixed by the volume shape alone (header | all pixel data | all IFDs), so every rank
  writes its own byte range of ONE shared file and rank 0 adds the header and the directory;
* no intermediate copies: readinto() a numpy view of the pinned tensor the H2D copy starts from.

Pure host code (struct + file I/O); no third-party TIFF library is needed or used.
"""
from __future__ import annotations

import os
import struct
import sys

import numpy as np

from ._lib import NeuroclearError

_CLASSIC_LIMIT = (1 << 32) - (1 << 20)


class TiffVolume:
    """Page table of a multi-page TIFF whose pages are the z-planes of one volume."""

    def __init__(self, path):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(16)
            if head[:2] == b"II":
                self.bo = "<"
            elif head[:2] == b"MM":
                self.bo = ">"
            else:
                raise NeuroclearError("%s: not a TIFF file" % path)
            magic = struct.unpack(self.bo + "H", head[2:4])[0]
            if magic == 42:
                self.big = False
                ifd = struct.unpack(self.bo + "I", head[4:8])[0]
            elif magic == 43:
                self.big = True
                ifd = struct.unpack(self.bo + "Q", head[8:16])[0]
            else:
                raise NeuroclearError("%s: unknown TIFF magic %d" % (path, magic))
            self.pages = []          # per page: list of (offset, nbytes) strips, in order
            shape = dtype = None
            while ifd:
                tags, ifd = self._read_ifd(f, ifd)
                if 256 not in tags or 257 not in tags:
                    raise NeuroclearError("%s: page without ImageWidth / ImageLength tags" % path)
                w, h = tags[256][0], tags[257][0]
                bits = tags.get(258, [1])[0]
                if tags.get(259, [1])[0] != 1:
                    raise NeuroclearError("%s: compressed TIFF pages are not supported (compression %d)"
                                          % (path, tags[259][0]))
                if tags.get(277, [1])[0] != 1 or bits not in (8, 16) or tags.get(339, [1])[0] not in (1,):
                    raise NeuroclearError("%s: only single-channel unsigned 8 / 16-bit pages are supported" % path)
                if 273 not in tags:
                    raise NeuroclearError("%s: tiled TIFF pages are not supported" % path)
                offs, cnts = tags[273], tags.get(279)
                if cnts is None:
                    cnts = [w * h * bits // 8]
                if sum(cnts) != w * h * bits // 8:
                    raise NeuroclearError("%s: strip byte counts do not add up to a %d x %d page" % (path, w, h))
                page_shape, page_dtype = (h, w), np.dtype("u%d" % (bits // 8))
                if shape is None:
                    shape, dtype = page_shape, page_dtype
                elif (page_shape, page_dtype) != (shape, dtype):
                    raise NeuroclearError("%s: pages differ in size or type" % path)
                self.pages.append(list(zip(offs, cnts)))
        if not self.pages:
            raise NeuroclearError("%s: no image pages" % path)
        self.shape = (len(self.pages),) + shape
        self.dtype = dtype

    def _read_ifd(self, f, off):
        bo = self.bo
        f.seek(off)
        if self.big:
            n = struct.unpack(bo + "Q", f.read(8))[0]
            raw = f.read(n * 20 + 8)
            esz, cfmt, vsz = 20, "Q", 8
        else:
            n = struct.unpack(bo + "H", f.read(2))[0]
            raw = f.read(n * 12 + 4)
            esz, cfmt, vsz = 12, "I", 4
        tags = {}
        for i in range(n):
            e = raw[i * esz:(i + 1) * esz]
            tag, typ = struct.unpack(bo + "HH", e[:4])
            cnt = struct.unpack(bo + cfmt, e[4:4 + vsz])[0]
            if tag not in (256, 257, 258, 259, 273, 277, 279, 339) or typ not in (1, 3, 4, 16):
                continue
            fmt = {1: "B", 3: "H", 4: "I", 16: "Q"}[typ]
            size = struct.calcsize(fmt) * cnt
            field = e[4 + vsz:]
            if size <= vsz:
                data = field[:size]
            else:
                here = f.tell()
                f.seek(struct.unpack(bo + cfmt, field)[0])
                data = f.read(size)
                f.seek(here)
            tags[tag] = list(struct.unpack(bo + fmt * cnt, data))
        nxt = struct.unpack(bo + cfmt, raw[n * esz:n * esz + vsz])[0]
        return tags, nxt

    def read(self, z0=0, z1=None, out=None):
        """planes [z0, z1) -> numpy array (z1-z0, Y, X) of the file's dtype in native byte order; `out` (e.g. the
        numpy view of a pinned torch tensor) is filled in place with readinto()."""
        z1 = self.shape[0] if z1 is None else z1
        if not (0 <= z0 <= z1 <= self.shape[0]):
            raise NeuroclearError("plane range [%d, %d) outside the %d planes of %s" % (z0, z1, self.shape[0], self.path))
        want = (z1 - z0,) + self.shape[1:]
        if out is None:
            out = np.empty(want, dtype=self.dtype)
        if tuple(out.shape) != want or out.dtype != self.dtype or not out.flags["C_CONTIGUOUS"]:
            raise NeuroclearError("read(): out must be a C-contiguous %s array of shape %s" % (self.dtype, want))
        flat = out.reshape(-1).view(np.uint8)
        plane = self.shape[1] * self.shape[2] * self.dtype.itemsize
        with open(self.path, "rb", buffering=0) as f:
            z = z0
            while z < z1:
                # a run of pages whose strips follow each other in the file is ONE read
                start = self.pages[z][0][0]
                pos, zz = start, z
                while zz < z1:
                    p, contiguous = pos, True
                    for o, c in self.pages[zz]:
                        if o != p:
                            contiguous = False
                            break
                        p += c
                    if not contiguous:
                        break
                    pos, zz = p, zz + 1
                if zz == z:          # this page's strips are scattered: strip by strip
                    dst = (z - z0) * plane
                    for o, c in self.pages[z]:
                        f.seek(o)
                        self._readinto(f, flat[dst:dst + c])
                        dst += c
                    z += 1
                else:
                    f.seek(start)
                    self._readinto(f, flat[(z - z0) * plane:(zz - z0) * plane])
                    z = zz
        if self.dtype.itemsize > 1 and (self.bo == "<") != (sys.byteorder == "little"):
            out.byteswap(inplace=True)
        return out

    @staticmethod
    def _readinto(f, view):
        mv, got = memoryview(view), 0
        while got < len(mv):
            n = f.readinto(mv[got:])
            if not n:
                raise NeuroclearError("unexpected end of file")
            got += n


def read_volume(path, z0=0, z1=None, out=None):
    """skimage.io.imread(path) for the volumes of this path (or a plane range of it)."""
    return TiffVolume(path).read(z0, z1, out)


# ---------------------------------------------------------------------------------------------------- writer
class TiffLayout:
    """Byte layout of the output file, a function of (shape, dtype) only: header | pixel data | IFDs."""

    def __init__(self, shape, dtype, bigtiff=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        if len(self.shape) != 3 or self.dtype not in (np.dtype("u1"), np.dtype("u2")):
            raise NeuroclearError("TIFF output: a (Z,Y,X) uint8 / uint16 volume is expected")
        z, y, x = self.shape
        self.plane = y * x * self.dtype.itemsize
        ntags = 9
        classic_size = 8 + z * self.plane + z * (2 + ntags * 12 + 4)
        self.big = classic_size > _CLASSIC_LIMIT if bigtiff is None else bool(bigtiff)
        self.header = 16 if self.big else 8
        self.ifd_size = (8 + ntags * 20 + 8) if self.big else (2 + ntags * 12 + 4)
        self.data_offset = self.header
        # TIFF 6.0: an IFD must begin on a word boundary (BigTIFF: 8 bytes); an odd Y*X*Z of uint8 planes would put
        # it on an odd offset, so pad (the gap bytes are never referenced)
        align = 8 if self.big else 2
        self.ifd_offset = -(-(self.header + z * self.plane) // align) * align
        self.file_size = self.ifd_offset + z * self.ifd_size

    def plane_offset(self, z):
        return self.data_offset + z * self.plane

    def directory_bytes(self):
        """header and the chain of IFDs (little endian)"""
        z, y, x = self.shape
        bits = self.dtype.itemsize * 8
        if self.big:
            head = struct.pack("<2sHHHQ", b"II", 43, 8, 0, self.ifd_offset)
        else:
            head = struct.pack("<2sHI", b"II", 42, self.ifd_offset)
        ifds = bytearray()
        for k in range(z):
            nxt = self.ifd_offset + (k + 1) * self.ifd_size if k + 1 < z else 0
            entries = [(256, x), (257, y), (258, bits), (259, 1), (262, 1), (273, self.plane_offset(k)), (277, 1),
                       (278, y), (279, self.plane)]
            if self.big:
                ifds += struct.pack("<Q", len(entries))
                for tag, val in entries:
                    typ = 16 if tag in (273, 279) else (3 if tag in (258, 259, 262, 277) else 4)
                    ifds += struct.pack("<HHQ", tag, typ, 1) + struct.pack("<Q", val)
                ifds += struct.pack("<Q", nxt)
            else:
                ifds += struct.pack("<H", len(entries))
                for tag, val in entries:
                    typ = 3 if tag in (258, 259, 262, 277) else 4
                    ifds += struct.pack("<HHI", tag, typ, 1) + (struct.pack("<HH", val, 0) if typ == 3
                                                                else struct.pack("<I", val))
                ifds += struct.pack("<I", nxt)
        return head, bytes(ifds)


def write_planes(path, layout: TiffLayout, planes, z0, write_directory=False):
    """Write planes [z0, z0 + len(planes)) of the volume into `path` at their final position.  Any number of
    processes may write disjoint plane ranges of the same file; exactly one of them passes write_directory=True."""
    planes = np.ascontiguousarray(planes)
    if planes.dtype != layout.dtype or tuple(planes.shape[1:]) != layout.shape[1:] or z0 < 0 \
            or z0 + planes.shape[0] > layout.shape[0]:
        raise NeuroclearError("write_planes: planes do not match the layout")
    if planes.dtype.itemsize > 1 and planes.dtype.byteorder == ">":
        planes = planes.astype(planes.dtype.newbyteorder("<"))
    fd = os.open(path, os.O_RDWR | os.O_CREAT, 0o644)
    try:
        if write_directory:
            head, ifds = layout.directory_bytes()
            os.pwrite(fd, head, 0)
            os.pwrite(fd, ifds, layout.ifd_offset)
        mv, off, done = memoryview(planes.reshape(-1).view(np.uint8)), layout.plane_offset(z0), 0
        while done < len(mv):
            done += os.pwrite(fd, mv[done:done + (1 << 30)], off + done)
    finally:
        os.close(fd)


def write_volume(path, volume, bigtiff=None):
    """tifffile.imsave(path, volume) for a (Z,Y,X) uint8 / uint16 volume: one uncompressed page per z-plane."""
    volume = np.asarray(volume)
    layout = TiffLayout(volume.shape, volume.dtype, bigtiff)
    if os.path.exists(path):
        os.remove(path)
    write_planes(path, layout, volume, 0, write_directory=True)
    return layout
