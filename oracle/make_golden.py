"""Pins the oracle against the unmodified reference and writes tests/golden/*.npz.
Test infrastructure — see oracle/__init__.py.  Run in the build container only:  python -m oracle.make_golden

Every fixture stores the INPUT (or the seed that regenerates it) and the output of the REFERENCE code, never of
the oracle; the script also asserts oracle == reference on the spot.
"""
from __future__ import annotations

import io
import os
import sys
import tempfile
from collections import OrderedDict
from contextlib import redirect_stdout

import numpy as np
import torch

from . import assemble, dice, discriminator, geometry, mip, reference_harness as rh, unet

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_dataset_and_assembler(vol, roi, ov, bc, normalize):
    rh.install()
    rh.set_volume(vol)
    import data as refdata
    from util.assemble_dice import Assemble_Dice
    opt = rh.dice_opt(rh.make_dataroot(tempfile.mkdtemp()), roi, ov, bc, normalize)
    with redirect_stdout(io.StringIO()):
        ds = refdata.find_dataset_using_name("diceImage")(opt)
        asm = Assemble_Dice(opt)
    return ds, asm, opt


GRAD_SAMPLE_STRIDE = 61   # gradient tensors are committed as every 61st element


def golden_geometry():
    """Geometry KATs from the reference classes themselves."""
    rows = []
    for size, roi, ov, bc in [((128, 128, 128), 120, 15, 10), ((31, 40, 27), 12, 3, 2), ((50, 61, 33), 16, 4, 1),
                              ((105, 210, 120), 120, 15, 10), ((9, 9, 9), 8, 2, 1)]:
        vol = np.zeros(size, dtype=np.uint16)
        ds, asm, _ = _ref_dataset_and_assembler(vol, roi, ov, bc, False)
        g = geometry.dice_geometry(size, roi, ov, bc)
        assert tuple(ds.size()) == g.padded and tuple(ds.shape()) == g.steps and len(ds) == g.n_cubes, (size, g)
        assert (asm.z_steps, asm.y_steps, asm.x_steps) == g.steps
        for i in range(len(ds)):
            assert tuple(asm.indexToCoordinates(i)) == g.origin(i)
        rows.append(list(size) + [roi, ov, bc] + list(g.padded) + list(g.steps))
    # large shapes: formula only (SURVEY KATs: 900^3 -> 960^3 / 9x9x9; (1024,2048,2048) -> 10x20x20)
    for size in [(900, 900, 900), (1024, 2048, 2048)]:
        g = geometry.dice_geometry(size, 120, 15, 10)
        rows.append(list(size) + [120, 15, 10] + list(g.padded) + list(g.steps))
    assert rows[-2][6:] == [960, 960, 960, 9, 9, 9] and rows[-1][6:] == [1065, 2115, 2115, 10, 20, 20]
    np.savez_compressed(os.path.join(GOLD, "geometry.npz"), rows=np.array(rows, dtype=np.int64))


def golden_dice_assemble():
    """Non-cubic 31x40x27 uint16 volume, roi 12 / overlap 3 / border 2: reference cubes, blend, final volume."""
    rng = np.random.default_rng(0)
    size, roi, ov, bc = (31, 40, 27), 12, 3, 2
    vol = rng.integers(0, 65536, size, dtype=np.uint16)
    g = geometry.dice_geometry(size, roi, ov, bc)
    out = {}
    e = g.edge
    fake = rng.random((g.n_cubes, 1, e, e, e), dtype=np.float32)                     # stand-in network outputs
    for normalize in (True, False):
        ds, asm, _ = _ref_dataset_and_assembler(vol, roi, ov, bc, normalize)
        cubes = np.stack([ds[i]["A"].numpy() for i in range(len(ds))])             # (n,1,E,E,E) float32
        dd = dice.DirectDicer(vol, g)
        for i in range(g.n_cubes):
            assert np.array_equal(cubes[i], dd.cube(i)) and np.array_equal(cubes[i], dice.dice_cube_gather(vol, g, i))
        with redirect_stdout(io.StringIO()):
            for i in range(g.n_cubes):
                t = torch.from_numpy(fake[i][None])
                asm.addToStack(OrderedDict(real=t, fake=t))
            asm.assemble_all()
        ref_final = asm.getDict()["fake"]
        mine, pcts = assemble.assemble(list(fake), g, normalize)
        assert ref_final.dtype == np.uint16 and np.array_equal(ref_final, mine)
        out["final_norm" if normalize else "final_plain"] = ref_final
        if normalize:
            out["pcts"] = np.array(pcts)
        else:
            vis, mask = assemble.blend_sequential([assemble.crop_border(c, bc) for c in fake], g)
            assert np.array_equal(mask, asm.getMaskRet() if "real" in asm.mask_ret else mask)
            assert np.array_equal(mask, assemble.analytic_count(g))
            out["blend"] = vis
    out.update(volume=vol, cubes=cubes, fake=fake, params=np.array([roi, ov, bc]))
    np.savez_compressed(os.path.join(GOLD, "dice_assemble_31x40x27.npz"), **out)


def golden_unet():
    """Reference Unet_deconv (networks.define_G) vs oracle on a 16^3 and a 24x16x20 input.  The weights are
    oracle.unet.random_state_dict(seed=0, bias_std=0.1) loaded into the reference module (7 M parameters are
    regenerated from the seed in the tests instead of being committed; a checksum guards against RNG drift)."""
    rh.install()
    from models import networks
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    net.eval()
    ref_sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == unet.STATE_DICT_SHAPES
    assert list(ref_sd.keys()) == list(unet.STATE_DICT_SHAPES.keys())
    assert sum(v.numel() for v in ref_sd.values()) == unet.N_PARAMS
    sd = unet.random_state_dict(seed=0, bias_std=0.1)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(1)
    out = {}
    for name, shape in (("a", (1, 1, 16, 16, 16)), ("b", (1, 1, 24, 16, 20))):
        x = torch.rand(shape, generator=g)
        with torch.no_grad():
            y_ref = net(x)
        y_or = unet.unet_deconv_forward(x, sd)
        err = (y_ref - y_or).abs().max().item()
        assert err <= 1e-6, err
        out["x_" + name], out["y_" + name] = x.numpy(), y_ref.numpy()
    out["w_checksum"] = np.array(unet.state_dict_checksum(sd))
    np.savez_compressed(os.path.join(GOLD, "unet_small.npz"), **out)


def golden_unet_grad():
    """Reference Unet_deconv under autograd (train mode, as train_onecube.py runs it): loss = sum(net(x) * dout) on
    a 16x24x32 crop; records the output and every parameter gradient (sampled) and checks the oracle's gradients."""
    rh.install()
    from models import networks
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_deconv", "instance", False, "kaiming", 0.02, [], dimension=3)
    net.train()
    sd = unet.random_state_dict(seed=4, bias_std=0.1)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(8)
    x = torch.rand((1, 1, 16, 24, 32), generator=g)
    dout = torch.randn((1, 1, 16, 24, 32), generator=g) * 1e-3
    y = net(x)
    y.backward(dout)
    y_or, g_or = unet.unet_deconv_gradients(x, sd, dout)
    assert (y.detach() - y_or).abs().max().item() <= 1e-6
    out = {"x": x.numpy(), "dout": dout.numpy(), "y": y.detach().numpy(),
           "w_checksum": np.array(unet.state_dict_checksum(sd))}
    for k, prm in net.named_parameters():
        ref = prm.grad
        scale = ref.abs().max().item()
        assert (ref - g_or[k]).abs().max().item() <= 1e-4 * max(scale, 1e-7), k
        out["gnorm_" + k] = np.array([float(ref.double().norm()), scale])
        out["gsample_" + k] = ref.numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()
    np.savez_compressed(os.path.join(GOLD, "unet_grad.npz"), **out)


def golden_deeplinear():
    """Reference DeepLinearGenerator (networks.define_G('deep_linear_gen')) forward + autograd on a 12x20x16 crop vs
    the oracle; records output, input gradient and sampled weight gradients."""
    rh.install()
    from models import networks
    from oracle import deeplinear
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "deep_linear_gen", "instance", False, "kaiming", 0.02, [], dimension=3)
    ref_sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == deeplinear.STATE_DICT_SHAPES
    assert sum(v.numel() for v in ref_sd.values()) == deeplinear.N_PARAMS
    sd = deeplinear.random_state_dict(seed=2)
    net.load_state_dict(sd)
    g = torch.Generator().manual_seed(12)
    x = torch.rand((1, 1, 12, 20, 16), generator=g).requires_grad_(True)
    dout = torch.randn((1, 1, 12, 20, 16), generator=g) * 1e-3
    y = net(x)
    y.backward(dout)
    y_or, dx_or, g_or = deeplinear.deep_linear_gradients(x.detach(), sd, dout)
    assert (y.detach() - y_or).abs().max().item() <= 1e-5
    assert (x.grad - dx_or).abs().max().item() <= 1e-5 * dx_or.abs().max().item() + 1e-9
    out = {"x": x.detach().numpy(), "dout": dout.numpy(), "y": y.detach().numpy(), "dx": x.grad.numpy(),
           "w_checksum": np.array([float(sum(v.double().sum() for v in sd.values()))])}
    for k, prm in net.named_parameters():
        assert (prm.grad - g_or[k]).abs().max().item() <= 1e-4 * prm.grad.abs().max().item(), k
        out["gsample_" + k] = prm.grad.numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()
    np.savez_compressed(os.path.join(GOLD, "deeplinear_grad.npz"), **out)


def golden_mip():
    rh.install()
    from models.axial_to_lateral_gan_apollo_model import Volume
    g = torch.Generator().manual_seed(3)
    vol = torch.rand((1, 1, 12, 12, 12), generator=g)
    out = {"vol": vol.numpy()}
    np.random.seed(5)
    state = np.random.get_state()
    for axis in range(3):
        np.random.set_state(state)
        ref = Volume(vol, torch.device("cpu")).get_projection(4, axis)
        np.random.set_state(state)
        mine, start = mip.get_projection(vol, 4, axis)
        assert torch.equal(ref, mine)
        out[f"proj{axis}"], out[f"start{axis}"] = ref.numpy(), np.array(start)
        np.random.set_state(state)
        ref_s = Volume(vol, torch.device("cpu")).get_slice(axis)
        np.random.set_state(state)
        mine_s, idx = mip.get_slice(vol, axis)
        assert torch.equal(ref_s, mine_s)
    np.savez_compressed(os.path.join(GOLD, "mip_12.npz"), **out)




def golden_discriminator():
    """Reference define_D('basic', dimension=2) + GANLoss('lsgan') vs the oracle: prediction map, loss, and the
    gradients w.r.t. the input image and every parameter, on a 44 x 36 image."""
    rh.install()
    from models import networks
    with redirect_stdout(io.StringIO()):
        net = networks.define_D(1, 64, "basic", norm="instance", use_sigmoid=False, init_type="kaiming",
                                init_gain=0.02, gpu_ids=[], dimension=2)
    ref_sd = net.state_dict()
    assert {k: tuple(v.shape) for k, v in ref_sd.items()} == discriminator.STATE_DICT_SHAPES
    assert sum(v.numel() for v in ref_sd.values()) == discriminator.N_PARAMS
    sd = discriminator.random_state_dict(seed=0)
    net.load_state_dict(sd)
    crit = networks.GANLoss("lsgan")
    g = torch.Generator().manual_seed(2)
    x = torch.rand((1, 1, 44, 36), generator=g).requires_grad_(True)
    pred = net(x)
    loss = crit(pred, True) * 0.5 + crit(pred, False) * 0.25
    loss.backward()
    out = {"x": x.detach().numpy(), "pred": pred.detach().numpy(), "loss": np.array(loss.item()),
           "dx": x.grad.numpy()}
    grads = {k: prm.grad.numpy() for k, prm in net.named_parameters()}
    for k, gk in grads.items():      # 2.76 M gradient values: keep the norm and every 61st value of each tensor
        out["dnorm_" + k] = np.array(np.linalg.norm(gk.astype(np.float64)))
        out["dsample_" + k] = gk.reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()
    # oracle == reference
    sdo = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    xo = x.detach().clone().requires_grad_(True)
    po = discriminator.discriminator_forward(xo, sdo)
    lo = discriminator.lsgan_loss(po, True) * 0.5 + discriminator.lsgan_loss(po, False) * 0.25
    lo.backward()
    assert (po - pred).abs().max().item() <= 1e-6 and abs(lo.item() - loss.item()) <= 1e-7
    assert (xo.grad - x.grad).abs().max().item() <= 1e-7
    for k in sd:
        assert np.abs(sdo[k].grad.numpy() - grads[k]).max() <= 1e-6 * (1 + np.abs(grads[k]).max()), k
    np.savez_compressed(os.path.join(GOLD, "discriminator_44x36.npz"), **out)


def apollo_opt(**kw):
    from argparse import Namespace
    o = dict(isTrain=True, gpu_ids=[], checkpoints_dir="/tmp/nc_golden_ckpt", name="golden", preprocess="none",
             gan_mode="lsgan", image_dimension=3, randomize_projection_depth=True, projection_depth=10,
             min_projection_depth=2, lambda_plane=[1, 1, 1], input_nc=1, output_nc=1, ngf=64, ndf=64,
             netG="unet_deconv", netG_B="deep_linear_gen", netD="basic", n_layers_D=3, norm="instance", no_dropout=True,
             init_type="kaiming", init_gain=0.02, lr=1e-4, beta1=0.1, direction="AtoB", lambda_A=5.0,
             lr_policy="constant")
    o.update(kw)
    return Namespace(**o)


D_NAMES = ["D_A_axial", "D_A_lateral", "D_B_axial", "D_B_lateral"]   # creation order, apollo_model.py:99-123


def golden_apollo_discriminator_path():
    """The REFERENCE AxialToLateralGANApolloModel on a 32^3 crop (CPU): its real / fake / rec volumes, then
    (1) the discriminator half of optimize_parameters() (six D losses, Adam step) and (2) backward_G()'s losses and
    gradients w.r.t. fake and rec — both with a fixed np.random seed, D weights from the oracle's seeded state dicts."""
    rh.install()
    from models.axial_to_lateral_gan_apollo_model import AxialToLateralGANApolloModel
    torch.manual_seed(20211114)       # the generators keep their init_weights() draw (see main())
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANApolloModel(apollo_opt())
    for i, name in enumerate(D_NAMES):
        getattr(m, "net" + name).load_state_dict(discriminator.random_state_dict(seed=10 + i))
    np.random.seed(0)
    real = torch.rand((1, 1, 32, 32, 32), generator=torch.Generator().manual_seed(0))
    m.set_input({"A": real, "A_paths": "golden"})
    with torch.no_grad():
        m.forward()
    rec = (m.rec - m.rec.mean()) / m.rec.std() * 0.2 + 0.5      # G_B is un-normalised at init: keep rec in a sane range
    out = {"real": real.numpy(), "fake": m.fake.numpy(), "rec": rec.numpy(), "depth": np.array(m.projection_depth)}
    # ---- (2) generator-side terms first (Ds untouched): leaf copies of fake / rec collect the gradients
    m.fake = m.fake.detach().clone().requires_grad_(True)
    m.rec = rec.detach().clone().requires_grad_(True)
    m.set_requires_grad([m.netD_A_lateral, m.netD_A_axial, m.netD_B_lateral, m.netD_B_axial], False)
    np.random.seed(9)
    m.backward_G()
    for k in ("G_A", "G_A_lateral", "G_A_axial", "G_B", "G_B_lateral", "G_B_axial", "cycle"):
        out["loss_" + k] = np.array(float(getattr(m, "loss_" + k)))
    out["dfake"], out["drec"] = m.fake.grad.numpy(), m.rec.grad.numpy()
    # ---- (1) discriminator step
    m.set_requires_grad([m.netD_A_lateral, m.netD_A_axial, m.netD_B_lateral, m.netD_B_axial], True)
    m.optimizer_D.zero_grad()
    np.random.seed(7)
    m.backward_D_A_lateral()
    m.backward_D_A_axial()
    m.backward_D_B_lateral()
    m.backward_D_B_axial()
    for k in ("D_A_lateral", "D_A_axial", "D_B_lateral", "D_B_axial"):
        out["loss_" + k] = np.array(float(getattr(m, "loss_" + k)))
    for name in D_NAMES:
        for k, prm in getattr(m, "net" + name).named_parameters():
            out["grad_%s.%s" % (name, k)] = prm.grad.numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()
    m.optimizer_D.step()
    for name in D_NAMES:
        for k, prm in getattr(m, "net" + name).named_parameters():
            out["after_%s.%s" % (name, k)] = prm.detach().numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()
    np.savez_compressed(os.path.join(GOLD, "apollo_d_path_32.npz"), **out)


def golden_apollo_step():
    """One full optimize_parameters() of the REFERENCE AxialToLateralGANApolloModel (unet_deconv + deep_linear_gen +
    4 basic Ds, lsgan, lambda_A 5, randomized projection depth 10 — the README training command) on a 32^3 crop, CPU
    fp32, seeded weights and np.random: the 11 losses, fake / rec statistics, sampled generator gradients and the
    sampled parameters of all six networks after the step."""
    rh.install()
    from models.axial_to_lateral_gan_apollo_model import AxialToLateralGANApolloModel
    from oracle import deeplinear
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANApolloModel(apollo_opt())
    m.netG_A.load_state_dict(unet.random_state_dict(seed=21, bias_std=0.05))
    m.netG_B.load_state_dict(deeplinear.random_state_dict(seed=22))
    for i, name in enumerate(D_NAMES):
        getattr(m, "net" + name).load_state_dict(discriminator.random_state_dict(seed=30 + i))
    np.random.seed(3)
    real = torch.rand((1, 1, 32, 32, 32), generator=torch.Generator().manual_seed(5))
    m.set_input({"A": real, "A_paths": "golden"})
    m.optimize_parameters()
    out = {"real": real.numpy(), "depth": np.array(m.projection_depth),
           "fake_sample": m.fake.detach().numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy(),
           "rec_sample": m.rec.detach().numpy().reshape(-1)[::GRAD_SAMPLE_STRIDE].copy()}
    for k in m.loss_names:
        out["loss_" + k] = np.array(float(getattr(m, "loss_" + k)))
    def sample(t):          # every 61st element; small tensors (<= 4096 elements) in full
        flat = t.detach().numpy().reshape(-1)
        return (flat if flat.size <= 4096 else flat[::GRAD_SAMPLE_STRIDE]).copy()
    for name in ["G_A", "G_B"] + D_NAMES:
        for k, prm in getattr(m, "net" + name).named_parameters():
            out["after_%s.%s" % (name, k)] = sample(prm)
            if name.startswith("G_"):
                out["grad_%s.%s" % (name, k)] = sample(prm.grad)
    np.savez_compressed(os.path.join(GOLD, "apollo_step_32.npz"), **out)


def golden_augment():
    """The reference's training transform (data/base_dataset.py:87-131, README --preprocess
    random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel) with FIXED params (angle_3D, crop_pos,
    flip_axis — get_transform's `params` path) on a 20x48x56 uint16 volume, for seven angles; asserts
    oracle.augment.augment_crop (which only evaluates the crop) == reference (which rotates the whole volume),
    and oracle.augment.warp_affine_u16 == cv2.warpAffine, bit for bit."""
    rh.install()
    import cv2
    from argparse import Namespace
    from data.base_dataset import get_transform
    from oracle import augment
    rng = np.random.default_rng(7)
    vol = (rng.random((20, 48, 56)) ** 3 * 65535).astype(np.uint16)
    out = {"vol": vol}
    cases = []
    for i, angle in enumerate([0, 7, 45, 90, 133, 212, 359]):
        shape = augment.rotated_volume_shape(vol.shape, angle)
        crop = (12, min(16, shape[1]), min(16, shape[2]))
        pos = (int(rng.integers(0, 20 - crop[0] + 1)), int(rng.integers(0, shape[1] - crop[1] + 1)),
               int(rng.integers(0, shape[2] - crop[2] + 1)))
        flip = int(i % 3)
        opt = Namespace(preprocess="random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel",
                        image_dimension=3, crop_size=list(crop))
        ref = get_transform(opt, params={"angle_3D": angle, "crop_pos": pos, "flip_axis": flip})(vol)
        got = augment.augment_crop(vol, angle, pos, crop, flip)
        assert tuple(ref.shape) == got.shape == (1, 1) + crop, (angle, ref.shape, got.shape)
        assert np.array_equal(ref.numpy(), got), angle
        m, nw, nh = augment.rotate_plan(56, 48, angle)
        full = cv2.warpAffine(vol[3], m, (nw, nh), flags=cv2.INTER_LINEAR)
        assert np.array_equal(full, augment.warp_affine_u16(vol[3], m, np.arange(nw), np.arange(nh))), angle
        cases.append([angle, *pos, *crop, flip])
        out["crop_%d" % i] = ref.numpy()
    out["cases"] = np.array(cases)
    # the random path (params=None): same draws in the same order, seeded
    import random
    opt = Namespace(preprocess="random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel",
                    image_dimension=3, crop_size=[10, 12, 14])
    tf = get_transform(opt)
    drawn = []
    for i, seed in enumerate([11, 12, 13, 14]):
        random.seed(seed)
        np.random.seed(seed)
        ref = tf(vol)
        random.seed(seed)
        np.random.seed(seed)
        got, (angle, pos, flips) = augment.random_item(vol, (10, 12, 14))
        assert np.array_equal(ref.numpy(), got), (seed, angle, pos, flips)
        out["random_%d" % i] = ref.numpy()
        drawn.append([seed, angle, *pos, sum(1 << a for a in flips)])
    out["random_cases"] = np.array(drawn)
    np.savez_compressed(os.path.join(GOLD, "augment_20x48x56.npz"), **out)


def golden_unet_vanilla():
    """Reference define_G('unet_vanilla') (networks.py:540-608) on two small inputs; the oracle restatement must agree
    exactly (same torch CPU kernels)."""
    rh.install()
    from models import networks
    from . import unet_vanilla as uv
    with redirect_stdout(io.StringIO()):
        net = networks.define_G(1, 1, 64, "unet_vanilla", "instance", False, "kaiming", 0.02, [], dimension=3)
    assert {k: tuple(v.shape) for k, v in net.state_dict().items()} == uv.state_dict_shapes()
    sd = uv.random_state_dict(seed=3, bias_std=0.1)
    net.load_state_dict(sd)
    net.eval()
    out = {}
    g = torch.Generator().manual_seed(11)
    for name, shape in (("a", (1, 1, 16, 24, 32)), ("b", (2, 1, 24, 16, 16))):
        x = torch.rand(shape, generator=g) ** 3
        with torch.no_grad():
            y = net(x)
        assert torch.equal(y, uv.unet_vanilla_forward(x, sd))
        out["x_" + name], out["y_" + name] = x.numpy(), y.numpy()
    np.savez_compressed(os.path.join(GOLD, "unet_vanilla_small.npz"), **out)


def _sample(t):          # every 61st element; small tensors (<= 4096 elements) in full
    flat = t.detach().numpy().reshape(-1)
    return (flat if flat.size <= 4096 else flat[::GRAD_SAMPLE_STRIDE]).copy()


def golden_sibling_models():
    """SURVEY.md §8 f4 — one full optimize_parameters() of the REFERENCE's sibling models on a 32^3 crop (CPU fp32,
    seeded weights): AxialToLateralGANAthenaModel (six per-slice discriminators) and AxialToLateralGANDryopsModel (the
    ablation without G_B / D_B); and the spectral-norm PatchGAN (define_D 'basic_SN') forward + backward."""
    rh.install()
    from models.axial_to_lateral_gan_athena_model import AxialToLateralGANAthenaModel
    from models.axial_to_lateral_gan_dryops_model import AxialToLateralGANDryopsModel
    from models import networks
    from oracle import deeplinear
    real = torch.rand((1, 1, 32, 32, 32), generator=torch.Generator().manual_seed(7))
    # ---- athena
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANAthenaModel(apollo_opt(conversion_plane=["yz", "xy"], pool_size=50))
    names = ["D_A_yz", "D_A_xy", "D_A_xz", "D_B_yz", "D_B_xy", "D_B_xz"]
    m.netG_A.load_state_dict(unet.random_state_dict(seed=21, bias_std=0.05))
    m.netG_B.load_state_dict(deeplinear.random_state_dict(seed=22))
    for i, n in enumerate(names):
        getattr(m, "net" + n).load_state_dict(discriminator.random_state_dict(seed=40 + i))
    m.set_input({"A": real, "A_paths": "golden"})
    m.optimize_parameters()
    out = {"real": real.numpy()}
    for k in m.loss_names:
        out["loss_" + k] = np.array(float(getattr(m, "loss_" + k)))
    for n in ["G_A", "G_B"] + names:
        for k, prm in getattr(m, "net" + n).named_parameters():
            out["after_%s.%s" % (n, k)] = _sample(prm)
            if n.startswith("G_"):
                out["grad_%s.%s" % (n, k)] = _sample(prm.grad)
    np.savez_compressed(os.path.join(GOLD, "athena_step_32.npz"), **out)
    # ---- dryops
    with redirect_stdout(io.StringIO()):
        m = AxialToLateralGANDryopsModel(apollo_opt())
    m.netG_A.load_state_dict(unet.random_state_dict(seed=21, bias_std=0.05))
    for i, n in enumerate(["D_A_axial", "D_A_lateral"]):
        getattr(m, "net" + n).load_state_dict(discriminator.random_state_dict(seed=30 + i))
    np.random.seed(3)
    m.set_input({"A": real, "A_paths": "golden"})
    m.optimize_parameters()
    out = {"real": real.numpy(), "depth": np.array(m.projection_depth)}
    for k in m.loss_names:
        out["loss_" + k] = np.array(float(getattr(m, "loss_" + k)))
    for n in ["G_A", "D_A_axial", "D_A_lateral"]:
        for k, prm in getattr(m, "net" + n).named_parameters():
            out["after_%s.%s" % (n, k)] = _sample(prm)
            if n == "G_A":
                out["grad_%s.%s" % (n, k)] = _sample(prm.grad)
    np.savez_compressed(os.path.join(GOLD, "dryops_step_32.npz"), **out)
    # ---- spectral-norm PatchGAN: initial state (incl. the power-iteration vectors), one training-mode forward+backward
    torch.manual_seed(77)
    with redirect_stdout(io.StringIO()):
        net = networks.define_D(1, 64, "basic_SN", norm="instance", use_sigmoid=False, init_type="kaiming",
                                init_gain=0.02, gpu_ids=[], dimension=2)
    # the initial state is a function of the seed alone (same torch modules built in the same order on both sides):
    # the fixture keeps per-tensor float64 sums as its checksum instead of 11 MB of weights
    out = {"sdsum_" + k: np.array(float(v.double().sum())) for k, v in net.state_dict().items()}
    x = torch.rand((2, 1, 44, 36), generator=torch.Generator().manual_seed(2)).requires_grad_(True)
    net.train()
    pred = net(x)
    crit = networks.GANLoss("lsgan")
    loss = crit(pred, True) * 0.5 + crit(pred, False) * 0.25
    loss.backward()
    out.update({"x": x.detach().numpy(), "pred": pred.detach().numpy(), "loss": np.array(float(loss)),
                "dx": x.grad.numpy()})
    for k, prm in net.named_parameters():
        out["grad_" + k] = prm.grad.numpy().reshape(-1)[::7].copy()
    for k, v in net.state_dict().items():
        if k.endswith(("_u", "_v")):
            out["after_" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(GOLD, "discriminator_sn_44x36.npz"), **out)


def golden_report():
    """The PSNR report of test_dice.py:239-253 computed with the REFERENCE's util.util functions (normalize,
    standardize, get_psnr) on three small uint16 volumes, and the oracle's restatement checked against it."""
    rh.install()
    from util import util as refutil
    from . import postprocess
    rng = np.random.default_rng(17)
    shape = (20, 24, 28)
    gt = (rng.random(shape) ** 3 * 65535).astype(np.uint16)
    real = np.clip(gt.astype(np.float64) * 0.6 + rng.normal(0, 2500, shape) + 3000, 0, 65535).astype(np.uint16)
    fake = np.clip(gt.astype(np.float64) * 0.9 + rng.normal(0, 900, shape) + 500, 0, 65535).astype(np.uint16)
    vols = [real, fake, gt]
    for _ in range(2):                                                         # test_dice.py:243-249, as written
        vols = [refutil.normalize(refutil.standardize(v), data_type=np.uint8) for v in vols]
    p_in = refutil.get_psnr(vols[0], vols[2], 2 ** 8 - 1)
    p_out = refutil.get_psnr(vols[1], vols[2], 2 ** 8 - 1)
    mine = postprocess.psnr_report(real, fake, gt)
    assert mine[0] == p_in and mine[1] == p_out and all(np.array_equal(a, b) for a, b in zip(mine[2:], vols))
    np.savez_compressed(os.path.join(GOLD, "report_psnr.npz"), real=real, fake=fake, gt=gt, real8=vols[0],
                        fake8=vols[1], gt8=vols[2], psnr_input_gt=np.array(p_in), psnr_output_gt=np.array(p_out))


def main():
    if not rh.available():
        sys.exit("reference not mounted at /root/reference: golden vectors can only be regenerated in the build container")
    os.makedirs(GOLD, exist_ok=True)
    # Every function seeds what it draws (torch Generators, torch.manual_seed, np.random.seed, seeded state dicts), so
    # the fixtures do not depend on the order of the calls and every one of them regenerates bit-identically
    # (apollo_d_path_32.npz was re-recorded with the seeded function in round 2).
    golden_geometry()
    golden_dice_assemble()
    golden_unet()
    golden_mip()
    golden_discriminator()
    golden_apollo_discriminator_path()
    golden_unet_grad()
    golden_deeplinear()
    golden_apollo_step()
    golden_augment()
    golden_report()
    golden_unet_vanilla()
    golden_sibling_models()
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    main()
