"""GPU: the whole diced-inference path (test_dice.py loop) against the oracle pipeline, plus size-independent
properties at larger sizes."""
import io
from argparse import Namespace
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from oracle import assemble, dice, geometry as ogeo, unet as ounet

pytestmark = pytest.mark.gpu


def _opt(roi, ov, bc, normalize=True):
    return Namespace(dataroot="", dice_size=[roi] * 3, overlap=ov, border_cut=bc, preprocess="addColorChannel",
                     image_dimension=3, dataset_mode="diceImage", data_type="uint16", skip_real=True,
                     histogram_match=False, normalize_intensity=normalize, sat_level=[0.25, 99.75], gpu_ids=[0])


def _oracle_pipeline(vol, sd, roi, ov, bc, normalize):
    g = ogeo.dice_geometry(vol.shape, roi, ov, bc)
    outs = []
    for i in range(g.n_cubes):
        x = torch.from_numpy(dice.dice_cube_gather(vol, g, i))[None]
        outs.append(ounet.unet_deconv_forward(x, sd).numpy())
    blend, _ = assemble.blend_sequential([assemble.crop_border(o, bc) for o in outs], g)
    final, pcts = assemble.finish(blend, g, normalize)
    return outs, blend, final, pcts


def _volume(shape, seed=0, roi=24, ov=6, bc=4):
    """Fluorescence-like volume.  Value parity needs every cube to contain some data: a cube lying entirely in
    pad_for_dicing's zero padding is degenerate in the reference itself (InstanceNorm of constant channels
    amplifies fp32 rounding noise by 1/sqrt(eps) per layer — see DESIGN.md), so shapes with such cubes are only
    used for the finite-output test below."""
    step = roi - ov
    assert all(((n + ov) // step) * step - bc < n for n in shape), "shape has pure-padding cubes"
    rng = np.random.default_rng(seed)
    return (rng.random(shape) ** 3 * 65535).astype(np.uint16)


def test_drop_in_loop_like_test_dice(cuda):
    """The reference's own loop (test_dice.py:70-121) with our DiceImageDataSet / define_G / Assemble_Dice."""
    from neuroclear_b200 import networks
    from neuroclear_b200.dicing import Assemble_Dice, DiceImageDataSet
    roi, ov, bc = 24, 6, 4                                   # cube edge 32
    vol = _volume((40, 58, 38))
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    opt = _opt(roi, ov, bc)
    dataset = DiceImageDataSet(opt, volume=vol)
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [0], dimension=3)
    net.module.load_state_dict(sd)
    net.eval()
    asm = Assemble_Dice(opt, dataset)
    og = ogeo.dice_geometry(vol.shape, roi, ov, bc)
    assert asm.image_size == og.padded and (asm.z_steps, asm.y_steps, asm.x_steps) == og.steps
    assert len(dataset) == og.n_cubes and dataset.size_original() == vol.shape
    fakes = []
    for data in dataset:
        assert data["A"].shape == (1, 1, 32, 32, 32)
        with torch.no_grad():
            fake = net(data["A"])
        fakes.append(fake)
        asm.addToStack({"real": data["A"], "fake": fake})
    asm.assemble_all()
    got = asm.getDict()["fake"]
    outs, blend, final, pcts = _oracle_pipeline(vol, sd, roi, ov, bc, True)
    assert got.dtype == np.uint16 and got.shape == vol.shape
    # network tolerance (max-abs 2e-2 on [0,1]) ...
    err = max(float((f.cpu() - torch.from_numpy(o)).abs().max()) for f, o in zip(fakes, outs))
    assert err <= 2e-2, err
    # ... carried through the percentile stretch: y = (x - p1) / span moves by at most 4 err / span
    span = pcts[1] - pcts[0]
    assert np.abs(got.astype(np.int64) - final.astype(np.int64)).max() <= 4 * err / span * 65535 + 2
    # assembly INDEXING / LAYOUT is bit-exact: feed the oracle's own cube outputs through our assembly
    asm2 = Assemble_Dice(_opt(roi, ov, bc), dataset)
    for o in outs:
        t = torch.from_numpy(o).to(cuda)
        asm2.addToStack({"real": t, "fake": t})
    asm2.assemble_all()
    assert np.array_equal(asm2.getDict()["fake"], final)
    assert asm2.percentiles["fake"] == pcts


@pytest.mark.parametrize("normalize", [True, False])
def test_fused_pipeline_matches_oracle(cuda, normalize):
    from neuroclear_b200.pipeline import DicedInference
    roi, ov, bc = 24, 6, 4
    vol = _volume((40, 47, 61), seed=2)
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    pipe = DicedInference(sd, cuda, roi, ov, bc, normalize_intensity=normalize, batch=5)
    got, (z0, z1) = pipe.run(vol)
    outs, blend, final, pcts = _oracle_pipeline(vol, sd, roi, ov, bc, normalize)
    assert (z0, z1) == (0, vol.shape[0]) and got.shape == vol.shape and got.dtype == np.uint16
    scale = 4 * 65535 / (pcts[1] - pcts[0]) if normalize else 65535
    assert np.abs(got.astype(np.int64) - final.astype(np.int64)).max() <= 2e-2 * scale + 2
    # batching must not change a single bit
    pipe1 = DicedInference(sd, cuda, roi, ov, bc, normalize_intensity=normalize, batch=1)
    got1, _ = pipe1.run(vol)
    assert np.array_equal(got, got1)


def test_config1_128_cube_geometry_and_one_cube(cuda):
    """BASELINE config 1: 128^3 volume, dice 120 / overlap 15 / border 10 -> 8 cubes of 140^3."""
    from neuroclear_b200.pipeline import DicedInference
    vol = _volume((128, 128, 128), seed=4, roi=120, ov=15, bc=10)
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    pipe = DicedInference(sd, cuda, 120, 15, 10, normalize_intensity=True, batch=4)
    plan = pipe.plan(vol.shape)
    assert plan["geo"].padded == (225, 225, 225) and plan["geo"].steps == (2, 2, 2) and plan["geo"].edge == 140
    with torch.cuda.device(cuda):
        dev, z0 = pipe.upload(torch.from_numpy(vol), plan)
        queue = pipe.infer_cubes(dev, z0, plan)
    og = ogeo.dice_geometry(vol.shape, 120, 15, 10)
    x = torch.from_numpy(dice.dice_cube_gather(vol, og, 5))[None]
    ref = assemble.crop_border(ounet.unet_deconv_forward(x, sd).numpy(), 10)
    err = np.abs(queue[5].cpu().numpy() - ref).max()
    assert err <= 2e-2, err
    got, _ = pipe.run(vol)
    assert got.shape == vol.shape


def test_identity_network_roundtrip_large(cuda):
    """Size-independent property at a larger size: dice -> (border-cut identity) -> blend -> uint16 reproduces the
    input within 1 LSB — exercises reflect/zero padding, indexing and the blend for 4 x 4 x 5 cubes."""
    from neuroclear_b200.dicing import blend_gather, dice_extract, dice_geometry, rescale_u16_crop
    vol = np.random.default_rng(9).integers(0, 65536, (400, 330, 470), dtype=np.uint16)
    g = dice_geometry(vol.shape, 120, 15, 10)
    dev = torch.from_numpy(vol).to(cuda)
    queue = torch.empty((g.n_cubes, 120, 120, 120), dtype=torch.float32, device=cuda)
    for i in range(0, g.n_cubes, 8):
        n = min(8, g.n_cubes - i)
        queue[i:i + n] = dice_extract(dev, 0, g, i, n)[:, 10:-10, 10:-10, 10:-10]
    off = torch.arange(g.n_cubes, dtype=torch.int64, device=cuda) * 120 ** 3
    zz = torch.zeros(g.n_cubes, dtype=torch.int32, device=cuda)
    vis = blend_gather(queue.view(-1), off, zz, g, 0, g.padded[0])
    out = rescale_u16_crop(vis, 0, g, None, 0, g.size[0]).cpu().numpy()
    assert np.abs(out.astype(np.int64) - vol.astype(np.int64)).max() <= 1
    # the padded region of the blend is exactly zero, the overlap count never exceeds 8
    assert float(vis[vol.shape[0]:].abs().max()) == 0.0


def test_pure_padding_cubes_stay_finite(cuda):
    """(n + overlap) % step < overlap puts whole cubes into the zero padding (here z: 30 -> cubes at 0, 18, 36).
    Their reference output is amplified rounding noise (no parity possible); ours must simply be finite and the
    data region must still match the oracle where only data-bearing cubes contribute."""
    from neuroclear_b200.pipeline import DicedInference
    roi, ov, bc = 24, 6, 4
    rng = np.random.default_rng(6)
    vol = (rng.random((30, 40, 40)) ** 3 * 65535).astype(np.uint16)
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    pipe = DicedInference(sd, cuda, roi, ov, bc, normalize_intensity=False, batch=3)
    got, _ = pipe.run(vol)
    assert got.shape == vol.shape
    outs, blend, final, _ = _oracle_pipeline(vol, sd, roi, ov, bc, False)
    assert np.abs(got.astype(np.int64) - final.astype(np.int64)).max() <= 2e-2 * 65535 + 2


def test_tiny_volume_single_cube(cuda):
    """A volume smaller than one dice in every axis: pad_for_dicing makes exactly one 24^3 cube."""
    from neuroclear_b200.pipeline import DicedInference
    roi, ov, bc = 24, 6, 4
    vol = (np.random.default_rng(8).random((5, 7, 9)) * 65535).astype(np.uint16)
    sd = ounet.random_state_dict(seed=0, bias_std=0.1)
    pipe = DicedInference(sd, cuda, roi, ov, bc, normalize_intensity=False, batch=2)
    assert pipe.plan(vol.shape)["geo"].n_cubes == 1
    got, _ = pipe.run(vol)
    outs, blend, final, _ = _oracle_pipeline(vol, sd, roi, ov, bc, False)
    assert got.shape == vol.shape and np.abs(got.astype(np.int64) - final.astype(np.int64)).max() <= 2e-2 * 65535 + 2


def test_assembly_error_behaviour_matches_reference(cuda):
    """Same guard rails as util/assemble_dice.py: wrong cube size asserts, border_cut 0 / overlap 0 are rejected
    (the reference crashes / returns zeros for them), CPU tensors are refused (no fallback)."""
    from neuroclear_b200._lib import NeuroclearError
    from neuroclear_b200.dicing import Assemble_Dice, DiceImageDataSet
    vol = np.zeros((40, 40, 40), dtype=np.uint16)
    ds = DiceImageDataSet(_opt(24, 6, 4), volume=vol)
    asm = Assemble_Dice(_opt(24, 6, 4), ds)
    good = torch.zeros((1, 1, 32, 32, 32), device=cuda)
    with pytest.raises(AssertionError):
        asm.addToStack({"real": good, "fake": torch.zeros((1, 1, 30, 32, 32), device=cuda)})
    with pytest.raises(NeuroclearError):
        asm.addToStack({"real": good, "fake": good.cpu()})
    with pytest.raises(KeyError):
        asm.addToStack({"fake": good})                       # 'real' is required even with skip_real (:132-133)
    with pytest.raises(NeuroclearError):
        asm.assemble_all()                                   # queue not full
    for _ in range(ds.geo.n_cubes):
        asm.addToStack({"real": good, "fake": good})
    with pytest.raises(NeuroclearError):
        asm.addToStack({"real": good, "fake": good})         # more cubes than the volume has
    asm.assemble_all()
    assert asm.getDict()["fake"].shape == vol.shape and asm.getSnapshots(3, 0)["fake"].shape == (40, 40)
    with pytest.raises(NeuroclearError):
        Assemble_Dice(_opt(24, 6, 0), ds)
    with pytest.raises(NeuroclearError):
        Assemble_Dice(_opt(24, 0, 4), ds)
    with pytest.raises(NeuroclearError):
        DiceImageDataSet(_opt(24, 6, 4), volume=vol.astype(np.float32))   # only uint16 / uint8 volumes


def test_uint8_data_type(cuda):
    """--data_type uint8 (base_dataset.py:135-136, assemble_dice.py:195-200): /255 in, *255 + astype(uint8) out."""
    from neuroclear_b200.dicing import blend_gather, dice_extract, dice_geometry, rescale_u16_crop
    rng = np.random.default_rng(12)
    vol = rng.integers(0, 256, (40, 47, 61), dtype=np.uint8)
    g, og = dice_geometry(vol.shape, 24, 6, 4), ogeo.dice_geometry(vol.shape, 24, 6, 4)
    dev = torch.from_numpy(vol).to(cuda)
    cubes = dice_extract(dev, 0, g, 0, g.n_cubes)
    for i in (0, g.n_cubes // 2, g.n_cubes - 1):
        assert np.array_equal(cubes[i].cpu().numpy()[None], dice.dice_cube_gather(vol, og, i))
    fake = rng.random((g.n_cubes, 1, 32, 32, 32), dtype=np.float32)
    ref, pcts = assemble.assemble(list(fake), og, True)                      # uint16 path for the percentiles ...
    blend, _ = assemble.blend_sequential([assemble.crop_border(c, 4) for c in fake], og)
    ref8, _ = assemble.finish(blend, og, True, imtype="uint8")               # ... and the reference's uint8 branch
    q = torch.from_numpy(np.stack([assemble.crop_border(c, 4) for c in fake])).to(cuda)
    off = torch.arange(g.n_cubes, dtype=torch.int64, device=cuda) * 24 ** 3
    zz = torch.zeros(g.n_cubes, dtype=torch.int32, device=cuda)
    vis = blend_gather(q.view(-1), off, zz, g, 0, g.padded[0])
    from neuroclear_b200.dicing import PercentileSelect
    norm3, _ = PercentileSelect(cuda).run(vis, vis.numel(), (0.25, 99.75))
    out8 = rescale_u16_crop(vis, 0, g, norm3, 0, g.size[0], dtype=torch.uint8).cpu().numpy()
    assert out8.dtype == np.uint8 and np.array_equal(out8, ref8)
