"""neuroclear_b200.volume_io (TIFF / BigTIFF plane-range reader and shared-file writer) against Pillow's TIFF codec."""
import os

import numpy as np
import pytest

from neuroclear_b200 import volume_io as vio
from neuroclear_b200._lib import NeuroclearError

Image = pytest.importorskip("PIL.Image")


def _volume(shape, dtype, seed=0):
    rng = np.random.default_rng(seed)
    return rng.integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)


@pytest.mark.parametrize("dtype", [np.uint16, np.uint8])
def test_written_file_is_read_back_by_pillow(tmp_path, dtype):
    vol = _volume((5, 13, 17), dtype)
    path = str(tmp_path / "out.tif")
    vio.write_volume(path, vol)
    with Image.open(path) as im:
        assert im.n_frames == 5
        for z in range(5):
            im.seek(z)
            assert np.array_equal(np.array(im), vol[z]), z


@pytest.mark.parametrize("dtype,mode", [(np.uint16, "I;16"), (np.uint8, "L")])
def test_reads_pillow_multipage_files_and_plane_ranges(tmp_path, dtype, mode):
    vol = _volume((6, 11, 9), dtype, seed=1)
    path = str(tmp_path / "in.tif")
    frames = [Image.fromarray(vol[z]) for z in range(6)]
    frames[0].save(path, save_all=True, append_images=frames[1:])
    tv = vio.TiffVolume(path)
    assert tv.shape == (6, 11, 9) and tv.dtype == np.dtype(dtype)
    assert np.array_equal(tv.read(), vol)
    out = np.empty((3, 11, 9), dtype=dtype)
    assert tv.read(2, 5, out=out) is out and np.array_equal(out, vol[2:5])
    assert np.array_equal(vio.read_volume(path, 5, 6), vol[5:6])


def test_bigtiff_round_trip_and_big_endian(tmp_path):
    vol = _volume((4, 7, 5), np.uint16, seed=2)
    path = str(tmp_path / "big.tif")
    layout = vio.write_volume(path, vol, bigtiff=True)
    assert layout.big and os.path.getsize(path) == layout.file_size
    assert np.array_equal(vio.read_volume(path), vol)
    # a big-endian classic file (as some microscopes write): build it by hand from the little-endian layout
    le = vio.TiffLayout(vol.shape, vol.dtype, bigtiff=False)
    head, ifds = le.directory_bytes()
    import struct
    be = bytearray(b"MM" + struct.pack(">HI", 42, le.ifd_offset))
    be += vol.astype(">u2").tobytes()
    for k in range(4):
        nxt = le.ifd_offset + (k + 1) * le.ifd_size if k < 3 else 0
        ent = [(256, 4, 5), (257, 4, 7), (258, 3, 16), (259, 3, 1), (262, 3, 1), (273, 4, le.plane_offset(k)),
               (277, 3, 1), (278, 4, 7), (279, 4, le.plane)]
        be += struct.pack(">H", len(ent))
        for tag, typ, val in ent:
            be += struct.pack(">HHI", tag, typ, 1) + (struct.pack(">HH", val, 0) if typ == 3 else struct.pack(">I", val))
        be += struct.pack(">I", nxt)
    p2 = str(tmp_path / "be.tif")
    open(p2, "wb").write(bytes(be))
    assert np.array_equal(vio.read_volume(p2), vol)


def test_ranks_write_disjoint_slabs_of_one_file(tmp_path):
    """the sharded pipeline: every rank writes its output z-slab at its final offset, rank 0 adds the directory"""
    vol = _volume((9, 10, 12), np.uint16, seed=3)
    path = str(tmp_path / "shared.tif")
    layout = vio.TiffLayout(vol.shape, vol.dtype)
    assert not layout.big
    for z0, z1, first in ((6, 9, False), (0, 2, True), (2, 6, False)):      # any order
        vio.write_planes(path, layout, vol[z0:z1], z0, write_directory=first)
    assert os.path.getsize(path) == layout.file_size
    assert np.array_equal(vio.read_volume(path), vol)
    with Image.open(path) as im:
        im.seek(7)
        assert np.array_equal(np.array(im), vol[7])
    assert vio.TiffLayout((1024, 2048, 2048), np.uint16).big            # config 5: 8.6 GB -> BigTIFF


def test_error_behaviour(tmp_path):
    vol = _volume((2, 8, 8), np.uint16)
    path = str(tmp_path / "lzw.tif")
    frames = [Image.fromarray(vol[z]) for z in range(2)]
    frames[0].save(path, save_all=True, append_images=frames[1:], compression="tiff_lzw")
    with pytest.raises(NeuroclearError):
        vio.TiffVolume(path)
    open(str(tmp_path / "junk.tif"), "wb").write(b"not a tiff at all")
    with pytest.raises(NeuroclearError):
        vio.TiffVolume(str(tmp_path / "junk.tif"))
    good = str(tmp_path / "good.tif")
    vio.write_volume(good, vol)
    with pytest.raises(NeuroclearError):
        vio.read_volume(good, 1, 5)
    with pytest.raises(NeuroclearError):
        vio.write_volume(good, vol.astype(np.float32))


def test_dataset_loader_reads_tiff_directories(tmp_path):
    """dicing._load_volume (the imread of DiceImageDataSet): first volume file of --dataroot, own reader for plain
    TIFFs, cv2 fallback for compressed ones"""
    from neuroclear_b200.dicing import _load_volume
    vol = _volume((3, 9, 7), np.uint16, seed=5)
    d = tmp_path / "data"
    d.mkdir()
    vio.write_volume(str(d / "a_volume.tif"), vol)
    assert np.array_equal(_load_volume(str(d)), vol)
    pytest.importorskip("cv2")
    frames = [Image.fromarray(vol[z]) for z in range(3)]
    frames[0].save(str(tmp_path / "lzw.tif"), save_all=True, append_images=frames[1:], compression="tiff_lzw")
    assert np.array_equal(_load_volume(str(tmp_path / "lzw.tif")), vol)


def _slab_writer(path, shape, z0, z1, first):
    vol = _volume(shape, np.uint16, seed=9)
    vio.write_planes(path, vio.TiffLayout(shape, np.uint16), vol[z0:z1], z0, write_directory=first)


def test_two_processes_write_one_file_concurrently(tmp_path):
    """what DicedInference.run_file does under torchrun: every rank writes its slab of the shared output file, using
    the balanced z-slab ranges of the sharding module"""
    import multiprocessing as mp
    from neuroclear_b200 import sharding
    shape = (23, 16, 20)
    path = str(tmp_path / "ranks.tif")
    ranges = sharding.balanced_ranges(shape[0], 3)
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_slab_writer, args=(path, shape, z0, z1, r == 0)) for r, (z0, z1) in enumerate(ranges)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert np.array_equal(vio.read_volume(path), _volume(shape, np.uint16, seed=9))


def test_ifd_chain_is_word_aligned_for_odd_uint8_volumes(tmp_path):
    """TIFF 6.0 requires IFDs on a word boundary; an odd Y*X*Z of uint8 planes used to put the chain on an odd offset."""
    from PIL import Image
    vol = np.random.default_rng(3).integers(0, 256, (3, 5, 7), dtype=np.uint8)       # 105 data bytes
    for big in (False, True):
        path = str(tmp_path / ("odd_%d.tif" % big))
        layout = vio.write_volume(path, vol, bigtiff=big)
        assert layout.ifd_offset % (8 if big else 2) == 0 and layout.file_size == os.path.getsize(path)
        assert np.array_equal(vio.read_volume(path), vol)
    im = Image.open(str(tmp_path / "odd_0.tif"))
    for z in range(3):
        im.seek(z)
        assert np.array_equal(np.array(im), vol[z])
