"""GPU: the remaining Assemble_Dice / test_dice.py options against the oracle —
--histogram_match (bit-exact float64 per cube and through the whole assembly), --save_projections (bit-exact) and the
PSNR report (uint8 volumes bit-exact against the reference-recorded fixture, PSNR to 1e-9)."""
import os
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _opt(roi, ov, bc, hist, normalize=True, skip_real=True):
    return Namespace(dataroot="", dice_size=[roi] * 3, overlap=ov, border_cut=bc, preprocess="addColorChannel",
                     data_type="uint16", skip_real=skip_real, histogram_match=hist, normalize_intensity=normalize,
                     sat_level=[0.25, 99.75], gpu_ids=[0])


@pytest.mark.parametrize("n,kind", [(12 ** 3, "u16"), (120 ** 3, "u16"), (30 ** 3, "ties"), (10 ** 3, "constant")])
def test_match_histograms_bit_exact(cuda, n, kind):
    from neuroclear_b200.dicing import Assemble_Dice
    from oracle import postprocess
    rng = np.random.default_rng(n)
    fake = (1 / (1 + np.exp(-rng.normal(0, 1.5, n)))).astype(np.float32)
    if kind == "u16":
        real = (rng.integers(0, 65536, n, dtype=np.uint16) / 65535.0).astype(np.float32)    # dice output values
    elif kind == "ties":
        real = (rng.integers(0, 7, n) / 65535.0).astype(np.float32)
        fake = np.round(fake, 2)                                                             # ties on both sides
    else:
        real = np.full(n, 0.25, dtype=np.float32)
    ref = postprocess.match_histograms(fake, real)
    helper = Assemble_Dice.__new__(Assemble_Dice)
    helper.device, helper._hm_scratch = cuda, None
    got = helper.match_histograms(torch.from_numpy(fake).to(cuda), torch.from_numpy(real).to(cuda)).cpu().numpy()
    assert got.dtype == np.float64 and np.array_equal(got, ref)


def test_assembly_with_histogram_match_bit_exact(cuda):
    """Assemble_Dice(opt.histogram_match=True): per-cube matching against the input cube, a float64 queue, the
    float64-accumulating blend, percentile stretch, cast — against the oracle built from the same numpy calls."""
    from neuroclear_b200.dicing import Assemble_Dice, DiceImageDataSet
    from oracle import assemble, dice, geometry as ogeo, postprocess
    rng = np.random.default_rng(2)
    size, roi, ov, bc = (31, 40, 27), 12, 3, 2
    vol = rng.integers(0, 65536, size, dtype=np.uint16)
    og = ogeo.dice_geometry(size, roi, ov, bc)
    e = og.edge
    fakes = rng.random((og.n_cubes, 1, 1, e, e, e), dtype=np.float32)
    for normalize in (True, False):
        opt = _opt(roi, ov, bc, True, normalize)
        ds = DiceImageDataSet(opt, volume=vol)
        asm = Assemble_Dice(opt, ds)
        cubes = []
        for i in range(len(ds)):
            real = ds[i]["A"][None]
            asm.addToStack(OrderedDict(real=real, fake=torch.from_numpy(fakes[i]).to(cuda)))
            r = assemble.crop_border(dice.dice_cube_gather(vol, og, i), bc)
            cubes.append(postprocess.match_histograms(assemble.crop_border(fakes[i], bc), r))
        asm.assemble_all()
        got = asm.getDict()["fake"]
        # the reference's loop (assemble_dice.py:167-184) with float64 cubes: numpy adds in float64, stores float32
        vis = np.zeros(og.padded, dtype=np.float32)
        mask = np.zeros(og.padded, dtype=np.float32)
        for i, c in enumerate(cubes):
            z, y, x = og.origin(i)
            vis[z:z + roi, y:y + roi, x:x + roi] += c / 8
            mask[z:z + roi, y:y + roi, x:x + roi] += np.ones((roi, roi, roi), dtype=np.float32)
        vis = (vis / mask) * 8
        ref, _ = assemble.finish(vis, og, normalize)
        assert got.dtype == np.uint16 and np.array_equal(got, ref), normalize


def test_save_projections_bit_exact(cuda):
    from neuroclear_b200 import report
    from oracle import postprocess
    rng = np.random.default_rng(6)
    fake = rng.integers(0, 65536, (40, 1120, 520), dtype=np.uint16)     # large enough for the hard-coded windows
    real = rng.integers(0, 65536, (40, 1120, 520), dtype=np.uint16)
    got, ref = report.save_projections(fake, real, cuda), postprocess.save_projections(fake, real)
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].dtype == ref[k].dtype and np.array_equal(got[k], ref[k]), k
    small = rng.integers(0, 256, (9, 11, 13), dtype=np.uint8)
    for axis in range(3):
        assert np.array_equal(report.max_projection(small, axis, device=cuda), np.amax(small, axis=axis))
    assert np.array_equal(report.max_projection(small, 1, 3, 100, cuda), np.amax(small[:, 3:100, :], axis=1))


def test_psnr_report_matches_reference_fixture(cuda, tmp_path):
    from neuroclear_b200 import report
    z = np.load(os.path.join(GOLDEN, "report_psnr.npz"))
    # The reference normalises twice; its second pass maps a uint8 volume that already spans 0..255 onto itself, so
    # every voxel sits EXACTLY on an integer boundary of the truncating cast and the last bit of np.std's float64
    # pairwise sum decides it.  nc_pairwise_sqdev_sum reproduces numpy's association order: everything is bit-exact.
    for name in ("real", "fake", "gt"):
        d = torch.from_numpy(z[name]).to(cuda)
        got8 = report.standardize_normalize_u8(report.standardize_normalize_u8(d)).cpu().numpy()
        assert np.array_equal(got8, z[name + "8"]), name
    p_in, p_out, msg = report.psnr_report(z["real"], z["fake"], z["gt"], cuda, name="exp", web_dir=str(tmp_path))
    assert p_in == float(z["psnr_input_gt"]) and p_out == float(z["psnr_output_gt"])
    assert "(psnr: %.4f)" % p_out in msg and (tmp_path / "metrics.txt").read_text().startswith("Experiment Name: exp")


@pytest.mark.parametrize("shape,dtype", [((7, 9, 11), np.uint16), ((64, 72, 80), np.uint16), ((33, 47, 51), np.uint8),
                                         ((150, 200, 210), np.uint16), ((1, 1, 5), np.uint8)])
def test_std_reproduces_numpy_pairwise_sum_bit_for_bit(cuda, shape, dtype):
    from neuroclear_b200 import _lib
    from neuroclear_b200._lib import call, i64, ptr, stream_ptr
    v = np.random.default_rng(shape[0]).integers(0, np.iinfo(dtype).max + 1, shape).astype(dtype)
    d = torch.from_numpy(v).to(cuda)
    n = v.size
    mean = float(np.mean(v))
    scratch = torch.empty(_lib.load().nc_pairwise_sqdev_scratch_doubles(n), dtype=torch.float64, device=cuda)
    acc = torch.empty(1, dtype=torch.float64, device=cuda)
    call("nc_pairwise_sqdev_sum", ptr(d), d.element_size(), i64(n), mean, ptr(scratch), ptr(acc), stream_ptr())
    assert np.sqrt(float(acc.item()) / n) == float(np.std(v))
