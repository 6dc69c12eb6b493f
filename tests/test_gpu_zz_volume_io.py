"""DicedInference.run_file: TIFF in -> diced inference -> TIFF out, equal to the in-memory path.  Composition of
parts that are each verified (volume_io on the CPU against Pillow, run_slab on the GPU).  First hardware run: the driver's round-1
GPU suite (passed); strict since round 2."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_file_to_file_equals_in_memory(tmp_path):
    from neuroclear_b200 import volume_io
    from neuroclear_b200.pipeline import DicedInference
    from oracle import unet as ounet
    vol = (np.random.default_rng(0).random((40, 41, 58)) ** 3 * 65535).astype(np.uint16)
    src, dst = str(tmp_path / "in.tif"), str(tmp_path / "out.tif")
    volume_io.write_volume(src, vol)
    pipe = DicedInference(ounet.random_state_dict(0, 0.1), "cuda:0", 24, 6, 4, batch=4)
    want, _ = pipe.run(vol)
    for chunks in (16, 3, 1):          # streaming granularity: reader thread / H2D / compute / D2H / write overlap
        layout = pipe.run_file(src, dst, chunks=chunks)
        got = volume_io.read_volume(dst)
        assert layout.shape == vol.shape and got.dtype == np.uint16
        assert np.array_equal(got, want), chunks
    # a pinned host volume takes the chunked-upload path of run() as well
    import torch
    pinned = torch.from_numpy(vol).pin_memory()
    again, _ = pipe.run(pinned)
    assert np.array_equal(again, want)


def test_streaming_reader_errors_surface(tmp_path):
    """a read error on the reader thread is re-raised by the consumer, not swallowed"""
    import pytest
    import torch
    from neuroclear_b200.pipeline import ChunkedUpload
    host = torch.zeros((8, 4, 4), dtype=torch.uint16).pin_memory()
    dev = torch.empty_like(host, device="cuda")

    def bad_source(a, b):
        if a >= 4:
            raise IOError("disk went away")
    up = ChunkedUpload(host, dev, 0, torch.cuda.Stream(), 4, bad_source)
    up.wait_for(4)
    with pytest.raises(IOError):
        up.wait_for(8)
