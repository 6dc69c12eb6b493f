"""tools/umma_model.py — the executable address model of TMA SWIZZLE_128B loads and tcgen05 shared-memory
descriptors (K-major and MN-major) — reproduces, with exactly the descriptor arithmetic of csrc/conv3d_tc.cu and
csrc/wgrad3d_tc.cu, a direct convolution and its weight gradient on ragged volumes (tiles hanging over every edge).
The same kernels are bit-exact on hardware (tests/cuda/probe_conv.cu, probe_grad.cu); this test keeps the model
honest so that new tilings can be checked on the CPU first."""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "tools"))
import umma_model as um  # noqa: E402


def _ints(rng, shape, lo=-3, hi=4):
    return rng.integers(lo, hi, shape).astype(np.float32)


def test_conv_forward_tiles_match_direct_convolution():
    rng = np.random.default_rng(0)
    NB, D, H, W, C, cout = 1, 3, 19, 11, 128, 16
    x = _ints(rng, (NB, D, H, W, C))
    w = _ints(rng, (cout, C, 3, 3, 3))
    ref = F.conv3d(torch.from_numpy(x).permute(0, 4, 1, 2, 3), torch.from_numpy(w), padding=1)
    ref = ref.permute(0, 2, 3, 4, 1).numpy()                                   # NDHWC
    for (d0, h0, w0) in ((0, 0, 0), (2, 16, 8), (0, 16, 0)):                    # interior corner, far corner, edge
        acc = um.conv3d_k3_tile(x, w, 0, d0, h0, w0, td=2)
        for j in range(2):
            for m in range(um.TW * um.TH):
                d, h, ww = d0 + j, h0 + m // um.TW, w0 + m % um.TW
                if d < D and h < H and ww < W:
                    assert np.array_equal(acc[j, m], ref[0, d, h, ww]), (d, h, ww)


def test_weight_gradient_tap_pairs_match_direct_gradient():
    rng = np.random.default_rng(1)
    NB, D, H, W = 1, 3, 19, 11
    x = _ints(rng, (NB, D, H, W, 64))
    dy = _ints(rng, (NB, D, H, W, 64))
    wt = torch.zeros((64, 64, 3, 3, 3), requires_grad=True)
    F.conv3d(torch.from_numpy(x).permute(0, 4, 1, 2, 3), wt, padding=1).backward(
        torch.from_numpy(dy).permute(0, 4, 1, 2, 3))
    ref = wt.grad.numpy()                                                       # (co, ci, kd, kh, kw)
    for kd in range(3):
        total = {t: np.zeros((64, 64), np.float32) for t in range(9)}
        for d in range(D):
            for h0 in range(0, H, um.TH):
                for w0 in range(0, W, um.TW):
                    part = um.wgrad_tile(x, dy, 0, d, h0, w0, kd, list(range(9)))
                    for t, v in part.items():
                        total[t] += v
        for t in range(9):
            assert np.array_equal(total[t], ref[:, :, kd, t // 3, t % 3].T), (kd, t)     # model gives (ci, co)


def test_linearised_plane_scheme_for_small_levels():
    """128 consecutive padded positions per tile instead of an 8 x 16 patch — every valid
    output equals the direct convolution, for tiles in the middle of the plane and hanging over its end"""
    rng = np.random.default_rng(2)
    NB, D, H, W, C, cout = 1, 3, 9, 35, 64, 8                                  # a 35-wide plane (the bottom level)
    x = _ints(rng, (NB, D, H, W, C))
    w = _ints(rng, (cout, C, 3, 3, 3))
    ref = F.conv3d(torch.from_numpy(x).permute(0, 4, 1, 2, 3), torch.from_numpy(w), padding=1)
    ref = ref.permute(0, 2, 3, 4, 1).numpy()
    L = W + 2
    checked = 0
    for p0 in (0, 128, 256):                                                   # 9 * 37 = 333 positions: 3 tiles
        acc = um.conv3d_k3_tile_linearised(x, w, 0, 1, p0, td=2)
        for j in range(2):
            for m in range(128):
                p = p0 + m
                h, ww, d = p // L, p % L - 1, 1 + j
                if h < H and 0 <= ww < W and d < D:
                    assert np.array_equal(acc[j, m], ref[0, d, h, ww]), (p0, m)
                    checked += 1
    assert checked == 2 * H * W                                                # every voxel of both planes, once


def test_swizzle_is_an_involution_on_chunks_and_keeps_rows():
    a = np.arange(0, 8192, 16)
    s = um.Smem.swizzle(a)
    assert np.array_equal(um.Smem.swizzle(s), a) and np.array_equal(s >> 7, a >> 7)
