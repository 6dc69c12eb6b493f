"""GPU: the reference's OWN scripts on the B200 path.  tests/run_reference_script.py executes
test_dice.py / train_onecube.py from oracle/_ref/neuroclear.zip (the reference byte-compiled by oracle/build_ref.py)
twice: as shipped on the CPU (`--stock`, --gpu_ids -1) and with neuroclear_b200.dropin.install() on the GPU — same
command line, same checkpoint, same input volume.  Nothing else differs."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RUNNER = os.path.join(ROOT, "tests", "run_reference_script.py")
ARCHIVE = os.path.join(ROOT, "oracle", "_ref", "neuroclear.zip")


def _run(args, cwd):
    r = subprocess.run([sys.executable, RUNNER] + args, cwd=cwd, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(not os.path.exists(ARCHIVE), reason="oracle/_ref/neuroclear.zip absent (python -m oracle.build_ref)")
def test_reference_test_dice_script_runs_on_the_b200_path(tmp_path, cuda):
    from neuroclear_b200 import volume_io
    from oracle import unet as ounet
    vol = (np.random.default_rng(0).random((40, 41, 58)) ** 3 * 65535).astype(np.uint16)
    data = tmp_path / "data"
    data.mkdir()
    volume_io.write_volume(str(data / "volume.tif"), vol)
    ckpt = tmp_path / "ckpt" / "exp"
    ckpt.mkdir(parents=True)
    torch.save(ounet.random_state_dict(seed=0, bias_std=0.1), str(ckpt / "latest_net_G_A.pth"))
    common = ["--dataroot", str(data), "--name", "exp", "--checkpoints_dir", str(tmp_path / "ckpt"), "--model", "test",
              "--model_suffix", "_A", "--netG", "unet_deconv", "--norm", "instance", "--no_dropout",
              "--dataset_mode", "diceImage", "--preprocess", "addColorChannel", "--dice_size", "24", "24", "24",
              "--overlap", "6", "--border_cut", "4", "--data_type", "uint16", "--normalize_intensity", "--skip_real",
              "--save_volume", "--image_dimension", "3"]
    outs = {}
    for mode, extra in (("stock", ["--gpu_ids", "-1"]), ("b200", ["--gpu_ids", "0"])):
        res = tmp_path / ("results_" + mode)
        log = _run(["test_dice"] + (["--stock"] if mode == "stock" else []) + ["--"] + common + extra +
                   ["--results_dir", str(res)], str(tmp_path))
        assert "----Test done----" in log
        path = res / "exp" / "test_latest" / "volumes" / "output_volume_xy-view_epoch-latest.tif"
        outs[mode] = volume_io.read_volume(str(path))
    a, b = outs["stock"].astype(np.int64), outs["b200"].astype(np.int64)
    assert a.shape == vol.shape == b.shape and outs["b200"].dtype == np.uint16
    diff = np.abs(a - b)
    # the stretch (0.25, 99.75 percentiles of a narrow sigmoid output) amplifies the network error: same bound as smoke()
    print("test_dice.py stock CPU vs B200 drop-in: max |diff| %d LSB, mean %.2f LSB" % (diff.max(), diff.mean()))
    assert diff.max() <= 700 and diff.mean() <= 60


@pytest.mark.skipif(not os.path.exists(ARCHIVE), reason="oracle/_ref/neuroclear.zip absent (python -m oracle.build_ref)")
def test_reference_train_onecube_script_runs_on_the_b200_path(tmp_path, cuda):
    """Two iterations of train_onecube.py --model axial_to_lateral_gan_apollo: the reference's model class, options,
    loop and torch.optim.Adam, on our generators / discriminators / Volume / GPU data pipeline.  Same seeds in both
    runs; the data pipeline is bit-exact, so both runs see the same crops; the 11 losses of iteration 1 must agree
    within 2 % (both runs load the same six checkpoints through the reference's --continue_train)."""
    from neuroclear_b200 import volume_io
    from oracle import apollo_step, deeplinear, discriminator, unet as ounet
    vol = (np.random.default_rng(1).random((40, 72, 72)) ** 2 * 65535).astype(np.uint16)
    data = tmp_path / "data"
    data.mkdir()
    volume_io.write_volume(str(data / "volume.tif"), vol)
    # both runs start from the same checkpoints (--continue_train): init_weights draws differ between CPU and CUDA RNGs
    ckpt = tmp_path / "ckpt" / "exp"
    ckpt.mkdir(parents=True)
    sds = {"G_A": ounet.random_state_dict(seed=21, bias_std=0.05), "G_B": deeplinear.random_state_dict(seed=22)}
    for i, n in enumerate(apollo_step.D_NAMES):
        sds[n] = discriminator.random_state_dict(seed=30 + i)
    for n, sd in sds.items():
        torch.save(sd, str(ckpt / ("latest_net_%s.pth" % n)))
    common = ["--dataroot", str(data), "--name", "exp", "--checkpoints_dir", str(tmp_path / "ckpt"),
              "--model", "axial_to_lateral_gan_apollo", "--netG", "unet_deconv", "--netG_B", "deep_linear_gen",
              "--netD", "basic", "--norm", "instance", "--no_dropout", "--init_type", "kaiming", "--gan_mode", "lsgan",
              "--lambda_A", "5", "--lambda_plane", "1", "1", "1", "--randomize_projection_depth",
              "--projection_depth", "10", "--dataset_mode", "singlevolume", "--crop_size", "32", "32", "32",
              "--preprocess", "random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel",
              "--lr_policy", "constant", "--display_id", "-1", "--no_html", "--print_freq", "1000000",
              "--save_latest_freq", "1000000", "--display_freq", "1000000", "--continue_train"]
    got = {}
    for mode, extra in (("stock", ["--gpu_ids", "-1"]), ("b200", ["--gpu_ids", "0"])):
        log = _run(["train_onecube"] + (["--stock"] if mode == "stock" else []) + ["--iters", "2", "--"] + common + extra,
                   str(tmp_path))
        line = [ln for ln in log.splitlines() if ln.startswith("LOSSES_JSON ")][-1]
        got[mode] = json.loads(line[len("LOSSES_JSON "):])
        assert len(got[mode]) == 2
    for k, ref in got["stock"][0].items():
        v = got["b200"][0][k]
        print("  iteration 1 loss_%-12s stock %.6f   b200 %.6f" % (k, ref, v))
        assert abs(v - ref) <= 2e-2 * max(1.0, abs(ref)), k
    assert all(np.isfinite(v) for v in got["b200"][1].values())
