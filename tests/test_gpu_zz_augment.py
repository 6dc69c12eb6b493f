"""nc_augment_crop_u16 / neuroclear_b200.augment on the GPU against the fixture recorded from the reference's
transform pipeline (bit-exact).  The kernel's per-voxel arithmetic is already checked on the CPU
(tests/test_augment_host.py builds csrc/augment_math.h with gcc).  First hardware run: the driver's round-1 GPU suite (passed); strict since round 2."""
import os
import random

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_crops_match_reference_fixture_bit_exact():
    from neuroclear_b200.augment import RotatedCropSampler
    z = np.load(os.path.join(GOLD, "augment_20x48x56.npz"))
    s = RotatedCropSampler(z["vol"], "cuda")
    for i, c in enumerate(z["cases"]):
        angle, pz, py, px, cz, cy, cx, flip = (int(v) for v in c)
        got = s.crop(angle, (pz, py, px), (cz, cy, cx), [flip])
        assert got.dtype == torch.float32 and tuple(got.shape) == (1, 1, cz, cy, cx)
        assert np.array_equal(got.cpu().numpy(), z["crop_%d" % i]), angle


def test_dataset_draws_like_the_reference():
    from argparse import Namespace
    from neuroclear_b200.augment import SingleVolumeDataset
    z = np.load(os.path.join(GOLD, "augment_20x48x56.npz"))
    opt = Namespace(preprocess="random3Drotate_randomcrop_randomflip_addColorChannel_addBatchChannel",
                    crop_size=[10, 12, 14], gpu_ids=[0], dataroot="synthetic")
    ds = SingleVolumeDataset(opt, z["vol"])
    assert len(ds) == 10
    for i, c in enumerate(z["random_cases"]):
        random.seed(int(c[0]))
        np.random.seed(int(c[0]))
        item = ds[i]
        assert item["A_paths"] == "synthetic"
        assert np.array_equal(item["A"].cpu().numpy(), z["random_%d" % i]), int(c[0])
