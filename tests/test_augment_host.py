"""Training-data augmentation on the CPU: the oracle against the fixture recorded from the
reference's transform pipeline (tests/golden/augment_20x48x56.npz, oracle/make_golden.py::golden_augment), the
host-side geometry / fixed-point tables of neuroclear_b200.augment, and the per-voxel arithmetic of the CUDA kernel
itself — csrc/augment_math.h is shared between augment.cu and a plain-C build (tests/cuda/augment_host.c) that this
file compiles with gcc and runs against the same fixture, bit for bit."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from oracle import augment as oaug

FIX = os.path.join(GOLDEN, "augment_20x48x56.npz")


def _cases(z):
    for i, c in enumerate(z["cases"]):
        angle, pz, py, px, cz, cy, cx, flip = (int(v) for v in c)
        yield angle, (pz, py, px), (cz, cy, cx), [flip], z["crop_%d" % i]
    for i, c in enumerate(z["random_cases"]):
        seed, angle, pz, py, px, mask = (int(v) for v in c)
        yield angle, (pz, py, px), (10, 12, 14), [a for a in range(3) if mask >> a & 1], z["random_%d" % i]


def test_oracle_matches_reference_fixture():
    z = np.load(FIX)
    for angle, pos, crop, flips, ref in _cases(z):
        assert np.array_equal(oaug.augment_crop(z["vol"], angle, pos, crop, flips), ref), angle


def test_oracle_random_draw_order_matches_reference_fixture():
    z = np.load(FIX)
    for i, c in enumerate(z["random_cases"]):
        random.seed(int(c[0]))
        np.random.seed(int(c[0]))
        got, (angle, pos, flips) = oaug.random_item(z["vol"], (10, 12, 14))
        assert [angle, *pos, sum(1 << a for a in flips)] == [int(v) for v in c[1:]]
        assert np.array_equal(got, z["random_%d" % i])


def test_warp_affine_restatement_equals_cv2():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    img = rng.integers(0, 65536, (37, 41), dtype=np.uint16)
    for angle in (0, 1, 30, 45, 90, 200, 359):
        m, nw, nh = oaug.rotate_plan(41, 37, angle)
        ref = cv2.warpAffine(img, m, (nw, nh), flags=cv2.INTER_LINEAR)
        assert np.array_equal(ref, oaug.warp_affine_u16(img, m, np.arange(nw), np.arange(nh))), angle


def test_kernel_arithmetic_host_build_matches_reference_fixture(tmp_path):
    """augment_math.h (the kernel's per-voxel code) + neuroclear_b200.augment's host tables == reference, on the CPU"""
    from neuroclear_b200 import augment as gaug
    so = str(tmp_path / "augment_host.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                    os.path.join(ROOT, "tests", "cuda", "augment_host.c"), "-lm"], check=True)
    lib = C.CDLL(so)
    z = np.load(FIX)
    vol = np.ascontiguousarray(z["vol"])
    Z, H, W = vol.shape
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    for angle, pos, crop, flips, ref in _cases(z):
        m, x1, x2, y1, y2 = gaug.rotate_clean_window(H, W, angle)
        mo = oaug.rotate_clean_window(H, W, angle)
        assert np.array_equal(m, mo[0]) and (x1, x2, y1, y2) == mo[1:]
        xs = np.arange(x1 + pos[2], x1 + pos[2] + crop[2])
        ys = np.arange(y1 + pos[1], y1 + pos[1] + crop[1])
        tabs = [np.ascontiguousarray(t) for t in gaug._inverse_map_tables(m, xs, ys)]
        assert all(t.dtype == np.int32 for t in tabs)
        out = np.empty(crop, dtype=np.float32)
        lib.augment_crop_host(p(vol), H, W, pos[0], *crop, p(tabs[0]), p(tabs[1]), p(tabs[2]), p(tabs[3]),
                              sum(1 << a for a in flips), p(out))
        assert np.array_equal(out, ref[0, 0]), angle
