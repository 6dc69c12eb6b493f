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
    layout = pipe.run_file(src, dst)
    got = volume_io.read_volume(dst)
    assert layout.shape == vol.shape and got.dtype == np.uint16
    assert np.array_equal(got, want)
