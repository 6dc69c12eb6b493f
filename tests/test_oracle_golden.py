"""CPU: the oracle against the fixtures the REFERENCE produced (tests/golden, written by oracle/make_golden.py)."""
import os

import numpy as np
import torch

from conftest import GOLDEN
from oracle import assemble, dice, geometry, mip, unet


def test_geometry_kats():
    rows = np.load(os.path.join(GOLDEN, "geometry.npz"))["rows"]
    for row in rows:
        size, (roi, ov, bc), padded, steps = tuple(row[:3]), row[3:6], tuple(row[6:9]), tuple(row[9:12])
        g = geometry.dice_geometry(size, int(roi), int(ov), int(bc))
        assert g.padded == padded and g.steps == steps
    # SURVEY.md §4 known answers
    g = geometry.dice_geometry((900, 900, 900), 120, 15, 10)
    assert g.padded == (960, 960, 960) and g.steps == (9, 9, 9) and g.n_cubes == 729
    g = geometry.dice_geometry((128, 128, 128), 120, 15, 10)
    assert g.padded == (225, 225, 225) and g.steps == (2, 2, 2) and g.edge == 140
    g = geometry.dice_geometry((1024, 2048, 2048), 120, 15, 10)
    assert g.padded == (1065, 2115, 2115) and g.steps == (10, 20, 20) and g.n_cubes == 4000


def test_cube_order_x_fastest():
    g = geometry.dice_geometry((31, 40, 27), 12, 3, 2)
    assert g.index_to_cube(0) == (0, 0, 0) and g.index_to_cube(1) == (0, 0, 1)
    assert g.index_to_cube(g.steps[2]) == (0, 1, 0) and g.index_to_cube(g.steps[2] * g.steps[1]) == (1, 0, 0)


def _fixture():
    return np.load(os.path.join(GOLDEN, "dice_assemble_31x40x27.npz"))


def test_dice_matches_reference_cubes():
    f = _fixture()
    roi, ov, bc = (int(v) for v in f["params"])
    vol = f["volume"]
    g = geometry.dice_geometry(vol.shape, roi, ov, bc)
    dd = dice.DirectDicer(vol, g)
    for i in range(g.n_cubes):
        assert np.array_equal(f["cubes"][i], dd.cube(i))
        assert np.array_equal(f["cubes"][i], dice.dice_cube_gather(vol, g, i))


def test_u8_normalise_is_fp32_divide():
    v = np.arange(256, dtype=np.uint8)
    ref = torch.from_numpy((v / (2 ** 8 * 1.0 - 1)).astype(float)).float().numpy()     # base_dataset.py:135-136
    assert np.array_equal(ref, v.astype(np.float32) / np.float32(255.0))


def test_u16_normalise_is_fp32_divide():
    v = np.arange(65536, dtype=np.uint16)
    ref = torch.from_numpy((v / (2 ** 16 * 1.0 - 1)).astype(float)).float().numpy()   # base_dataset.py:134-143,291-295
    assert np.array_equal(ref, v.astype(np.float32) / np.float32(65535.0))


def test_assemble_matches_reference():
    f = _fixture()
    roi, ov, bc = (int(v) for v in f["params"])
    g = geometry.dice_geometry(f["volume"].shape, roi, ov, bc)
    cubes = [assemble.crop_border(c, bc) for c in f["fake"]]
    vis, mask = assemble.blend_sequential(cubes, g)
    assert np.array_equal(vis, f["blend"])
    assert np.array_equal(mask, assemble.analytic_count(g)) and mask.max() == 8.0
    plain, _ = assemble.finish(vis, g, False)
    assert plain.dtype == np.uint16 and np.array_equal(plain, f["final_plain"])
    norm, pcts = assemble.finish(vis, g, True)
    assert np.array_equal(norm, f["final_norm"]) and np.allclose(pcts, f["pcts"], rtol=0, atol=0)


def test_identity_roundtrip_one_lsb():
    """uint16 -> dice -> (identity network) -> assemble -> uint16 differs by at most 1 LSB."""
    rng = np.random.default_rng(1)
    vol = rng.integers(0, 65536, (20, 33, 25), dtype=np.uint16)
    g = geometry.dice_geometry(vol.shape, 12, 3, 2)
    outs = [dice.dice_cube_gather(vol, g, i) for i in range(g.n_cubes)]
    final, _ = assemble.assemble(outs, g, False)
    assert final.shape == vol.shape
    assert np.abs(final.astype(np.int64) - vol.astype(np.int64)).max() <= 1


def test_unet_matches_reference_module():
    f = np.load(os.path.join(GOLDEN, "unet_small.npz"))
    sd = unet.random_state_dict(seed=0, bias_std=0.1)
    assert np.allclose(unet.state_dict_checksum(sd), f["w_checksum"], rtol=1e-12), "torch RNG drift: regenerate golden"
    assert sum(v.numel() for v in sd.values()) == unet.N_PARAMS == 7_077_251
    for name in "ab":
        y = unet.unet_deconv_forward(torch.from_numpy(f["x_" + name]), sd).numpy()
        assert np.abs(y - f["y_" + name]).max() <= 1e-5


def test_mip_matches_reference():
    f = np.load(os.path.join(GOLDEN, "mip_12.npz"))
    vol = torch.from_numpy(f["vol"])
    np.random.seed(5)
    state = np.random.get_state()
    for axis in range(3):
        np.random.set_state(state)
        proj, start = mip.get_projection(vol, 4, axis)
        assert start == int(f[f"start{axis}"]) and np.array_equal(proj.numpy(), f[f"proj{axis}"])


def test_discriminator_matches_reference_module():
    from oracle import discriminator as odisc
    f = np.load(os.path.join(GOLDEN, "discriminator_44x36.npz"))
    sd = {k: v.clone().requires_grad_(True) for k, v in odisc.random_state_dict(seed=0).items()}
    assert sum(v.numel() for v in sd.values()) == odisc.N_PARAMS
    x = torch.from_numpy(f["x"]).requires_grad_(True)
    pred = odisc.discriminator_forward(x, sd)
    loss = odisc.lsgan_loss(pred, True) * 0.5 + odisc.lsgan_loss(pred, False) * 0.25
    loss.backward()
    assert np.abs(pred.detach().numpy() - f["pred"]).max() <= 1e-6 and abs(loss.item() - float(f["loss"])) <= 1e-7
    assert np.abs(x.grad.numpy() - f["dx"]).max() <= 1e-7
    for k in sd:
        assert np.abs(sd[k].grad.numpy().reshape(-1)[::61] - f["dsample_" + k]).max() <= 1e-6


def test_deeplinear_oracle_matches_reference_fixture():
    """oracle/deeplinear.py vs the output / gradients recorded from the reference DeepLinearGenerator"""
    from oracle import deeplinear
    z = np.load(os.path.join(GOLDEN, "deeplinear_grad.npz"))
    sd = deeplinear.random_state_dict(seed=2)
    y, dx, grads = deeplinear.deep_linear_gradients(torch.from_numpy(z["x"]), sd, torch.from_numpy(z["dout"]))
    assert float((y - torch.from_numpy(z["y"])).abs().max()) <= 1e-5
    assert float((dx - torch.from_numpy(z["dx"])).abs().max()) <= 1e-7
    for k in deeplinear.STATE_DICT_SHAPES:
        ref = torch.from_numpy(z["gsample_" + k])
        got = grads[k].reshape(-1)[::61]
        assert float((got - ref).abs().max()) <= 1e-4 * float(ref.abs().max()), k


def test_unet_gradient_oracle_matches_reference_fixture():
    """oracle/unet.py::unet_deconv_gradients vs the gradients recorded from the reference Unet_deconv (8^3 would be
    cheaper, but the fixture's 16x24x32 crop is what the GPU test uses; ~10 s on the CPU)"""
    from oracle import unet
    z = np.load(os.path.join(GOLDEN, "unet_grad.npz"))
    sd = unet.random_state_dict(seed=4, bias_std=0.1)
    y, grads = unet.unet_deconv_gradients(torch.from_numpy(z["x"]), sd, torch.from_numpy(z["dout"]))
    assert float((y - torch.from_numpy(z["y"])).abs().max()) <= 1e-6
    for k in unet.STATE_DICT_SHAPES:
        ref = torch.from_numpy(z["gsample_" + k])
        got = grads[k].reshape(-1)[::61]
        assert float((got - ref).abs().max()) <= 1e-4 * max(float(z["gnorm_" + k][1]), 1e-7), k


def test_apollo_step_oracle_matches_reference_fixture():
    """oracle/apollo_step.py (the whole training iteration restated) vs one optimize_parameters() of the reference
    model: identical losses, identical parameters of all six networks after both Adam updates."""
    from oracle import apollo_step, deeplinear, discriminator
    z = np.load(os.path.join(GOLDEN, "apollo_step_32.npz"))
    sds = {"G_A": unet.random_state_dict(seed=21, bias_std=0.05), "G_B": deeplinear.random_state_dict(seed=22)}
    for i, n in enumerate(apollo_step.D_NAMES):
        sds[n] = discriminator.random_state_dict(seed=30 + i)
    m = apollo_step.ApolloStep(sds)
    np.random.seed(3)
    m.set_input(torch.from_numpy(z["real"]))
    losses = m.optimize_parameters()
    assert m.depth == int(z["depth"])
    for k, v in losses.items():
        assert abs(v - float(z["loss_" + k])) <= 1e-6 * max(1.0, abs(v)), k
    for n in ["G_A", "G_B"] + apollo_step.D_NAMES:
        for k, t in m.p[n].items():
            flat = t.detach().numpy().reshape(-1)
            got = flat if flat.size <= 4096 else flat[::61]
            assert np.abs(got - z["after_%s.%s" % (n, k)]).max() <= 1e-7, (n, k)


def test_apollo_discriminator_path_oracle_matches_reference_fixture():
    """oracle/apollo_step.py's generator-side losses (backward_G) and discriminator half against the fixture recorded
    from the reference model on given real / fake / rec volumes (apollo_d_path_32.npz): losses, d loss / d fake,
    d loss / d rec, discriminator gradients and Adam-updated weights."""
    from oracle import apollo_step, deeplinear, discriminator
    z = np.load(os.path.join(GOLDEN, "apollo_d_path_32.npz"))
    sds = {"G_A": unet.random_state_dict(seed=0), "G_B": deeplinear.random_state_dict(seed=0)}      # unused here
    for i, n in enumerate(apollo_step.D_NAMES):
        sds[n] = discriminator.random_state_dict(seed=10 + i)
    m = apollo_step.ApolloStep(sds)
    m.real, m.depth = torch.from_numpy(z["real"]), int(z["depth"])
    m.fake = torch.from_numpy(z["fake"]).requires_grad_(True)
    m.rec = torch.from_numpy(z["rec"]).requires_grad_(True)
    m.set_requires_grad_D(False)
    np.random.seed(9)
    m.backward_G()
    for k in ("G_A", "G_A_lateral", "G_A_axial", "G_B", "G_B_lateral", "G_B_axial", "cycle"):
        assert abs(float(m.loss[k]) - float(z["loss_" + k])) <= 1e-6, k
    assert float((m.fake.grad - torch.from_numpy(z["dfake"])).abs().max()) <= 1e-9
    assert float((m.rec.grad - torch.from_numpy(z["drec"])).abs().max()) <= 1e-9
    m.set_requires_grad_D(True)
    m.fake, m.rec = m.fake.detach(), m.rec.detach()
    m.opt_D.zero_grad()
    np.random.seed(7)
    m.backward_D()
    for k in ("D_A_lateral", "D_A_axial", "D_B_lateral", "D_B_axial"):
        assert abs(float(m.loss[k]) - float(z["loss_" + k])) <= 1e-6, k
    for n in apollo_step.D_NAMES:
        for k, t in m.p[n].items():
            ref = z["grad_%s.%s" % (n, k)]
            assert np.abs(t.grad.numpy().reshape(-1)[::61] - ref).max() <= 1e-6 * (1 + np.abs(ref).max()), (n, k)
    m.opt_D.step()
    for n in apollo_step.D_NAMES:
        for k, t in m.p[n].items():
            assert np.abs(t.detach().numpy().reshape(-1)[::61] - z["after_%s.%s" % (n, k)]).max() <= 1e-7, (n, k)


def test_psnr_report_matches_reference_fixture():
    """oracle.postprocess (normalize / standardize / get_psnr, test_dice.py:239-253) against the values recorded from
    the reference's own util.util functions."""
    from oracle import postprocess
    z = np.load(os.path.join(GOLDEN, "report_psnr.npz"))
    p_in, p_out, r8, f8, g8 = postprocess.psnr_report(z["real"], z["fake"], z["gt"])
    assert p_in == float(z["psnr_input_gt"]) and p_out == float(z["psnr_output_gt"])
    assert np.array_equal(r8, z["real8"]) and np.array_equal(f8, z["fake8"]) and np.array_equal(g8, z["gt8"])


def test_match_histograms_restatement_properties():
    """skimage is absent (parity unpinned, oracle/postprocess.py): check the defining properties of CDF matching —
    the output takes values inside the template's range, is a monotone function of the input, and matching an array
    to itself (or to a permutation of itself) is the identity."""
    from oracle import postprocess
    rng = np.random.default_rng(4)
    a = rng.random((6, 7, 8)).astype(np.float32)
    b = (rng.integers(0, 4000, (6, 7, 8)) / 65535.0).astype(np.float32)
    m = postprocess.match_histograms(a, b)
    assert m.dtype == np.float64 and m.shape == a.shape and m.min() >= b.min() and m.max() <= b.max()
    order = np.argsort(a.ravel(), kind="stable")
    assert np.all(np.diff(m.ravel()[order]) >= 0)
    assert np.array_equal(postprocess.match_histograms(b, b), b.astype(np.float64))
    assert np.array_equal(postprocess.match_histograms(b, rng.permutation(b.ravel()).reshape(b.shape)), b.astype(np.float64))


def test_unet_vanilla_oracle_matches_reference_fixture():
    from oracle import unet_vanilla as uv
    f = np.load(os.path.join(GOLDEN, "unet_vanilla_small.npz"))
    sd = uv.random_state_dict(seed=3, bias_std=0.1)
    assert sum(v.numel() for v in sd.values()) == sum(int(np.prod(s)) for s in uv.state_dict_shapes().values())
    for name in "ab":
        y = uv.unet_vanilla_forward(torch.from_numpy(f["x_" + name]), sd)
        assert np.abs(y.numpy() - f["y_" + name]).max() <= 1e-6
