"""Executable address model of the two Blackwell mechanisms the tensor-core kernels rest on, for checking a tiling /
descriptor scheme on the CPU before spending GPU minutes on it:

* a 5-D tiled TMA load with SWIZZLE_128B into shared memory (box of 64-channel rows, zero fill outside the tensor),
* the tcgen05 shared-memory matrix descriptor with the 128-byte swizzle, K-major (conv forward / data gradient) and
  MN-major (weight gradient), including the two facts established on hardware by tests/cuda/probe_conv.cu and
  probe_grad.cu: the XOR swizzle acts on ABSOLUTE shared-memory address bits (so any 128-byte-aligned start is a valid
  window into a TMA-written row array), and the leading-dimension offset of an MN-major operand may be any row shift
  (two filter taps stacked along M).

`conv3d_k3_tile` and `wgrad_tile` re-express, on top of that model, exactly the descriptor arithmetic of
csrc/conv3d_tc.cu (halo planes, one shifted window per tap) and csrc/wgrad3d_tc.cu (tap pairs along M, voxels along
K); tests/test_umma_model.py checks them against direct convolutions.  Pure numpy; not used by the product.
"""
from __future__ import annotations

import numpy as np

TW, TH = 8, 16


class Smem:
    """byte-addressed shared memory holding 16-bit elements (stored as float32 values for exact small integers)"""

    def __init__(self, nbytes):
        self.data = np.zeros(nbytes // 2, dtype=np.float32)

    @staticmethod
    def swizzle(addr):
        """SWIZZLE_128B: bits [4,7) ^= bits [7,10) of the absolute byte address"""
        return addr ^ (((addr >> 7) & 7) << 4)

    def write_row(self, row_addr, values64):
        """what TMA does with one 128-byte row (64 elements): 16-byte chunk j lands at chunk j ^ ((row_addr >> 7) & 7)"""
        assert row_addr % 128 == 0
        for j in range(8):
            phys = self.swizzle(row_addr + 16 * j)
            self.data[phys // 2: phys // 2 + 8] = values64[8 * j: 8 * j + 8]

    def read(self, logical_addr):
        """element(s) at LOGICAL byte address(es) as the tensor core's descriptor walk sees them"""
        return self.data[self.swizzle(np.asarray(logical_addr)) // 2]


def tma_load_plane(smem, dst, x, c0, w0, h0, d, nb, box_w, box_h):
    """cp.async.bulk.tensor.5d of a (64, box_w, box_h, 1, 1) box of x[NB, D, H, W, C]; rows ordered w fastest"""
    NB, D, H, W, C = x.shape
    for hh in range(box_h):
        for ww in range(box_w):
            h, w = h0 + hh, w0 + ww
            ok = 0 <= h < H and 0 <= w < W and 0 <= d < D and 0 <= nb < NB
            smem.write_row(dst + (hh * box_w + ww) * 128, x[nb, d, h, w, c0:c0 + 64] if ok else np.zeros(64, np.float32))


def read_k_major(smem, start, sbo, rows, k16):
    """K-major SW128 operand: `rows` rows x 16 elements of K step k16 (start already includes nothing of K)"""
    m, kk = np.arange(rows)[:, None], np.arange(16)[None, :]
    return smem.read(start + (m % 8) * 128 + (m // 8) * sbo + k16 * 32 + 2 * kk)


def read_mn_major(smem, start, lbo, sbo, mn):
    """MN-major SW128 operand: element (i, k) of an (mn x 16) slice of one K step; atoms of 64 along MN `lbo` bytes
    apart, K rows 128 bytes apart inside an 8-row group, groups `sbo` bytes apart (the caller advances `start` by 16
    rows per K step)"""
    i, k = np.arange(mn)[:, None], np.arange(16)[None, :]
    return smem.read(start + (i // 64) * lbo + (k % 8) * 128 + (k // 8) * sbo + (i % 64) * 2)


def conv3d_k3_tile(x, w, nb, d0, h0, w0, td=2, ks=3):
    """One output tile (TW x TH x td voxels, all Cout) of the k3 'same' conv exactly as conv3d_tc.cu addresses it:
    per 64-channel chunk, td+ks-1 halo planes of (TW+ks-1) x (TH+ks-1) rows; per tap a window starting
    (kh * HALO_W + kw) rows into the plane with 8-row-group stride HALO_W * 128 bytes."""
    NB, D, H, W, C = x.shape
    cout = w.shape[0]
    pad, hw, hh = ks // 2, TW + ks - 1, TH + ks - 1
    plane_bytes = (hw * hh * 128 + 1023) // 1024 * 1024
    acc = np.zeros((td, TW * TH, cout), dtype=np.float32)
    for c in range(C // 64):
        smem = Smem((td + ks - 1) * plane_bytes + 1024)
        for i in range(td + ks - 1):
            tma_load_plane(smem, i * plane_bytes, x, c * 64, w0 - pad, h0 - pad, d0 - pad + i, nb, hw, hh)
        for kd in range(ks):
            for kh in range(ks):
                for kw in range(ks):
                    bt = w[:, c * 64:c * 64 + 64, kd, kh, kw]                      # (cout, 64): B operand, K-major
                    for j in range(td):
                        start = (j + kd) * plane_bytes + (kh * hw + kw) * 128
                        for k16 in range(4):
                            a = read_k_major(smem, start, hw * 128, TW * TH, k16)   # (128, 16)
                            acc[j] += a @ bt[:, k16 * 16:(k16 + 1) * 16].T
    return acc            # [plane][row = mh * 8 + mw][cout]


def wgrad_tile(x, dy, nb, d, h0, w0, kd, taps, ks=3):
    """Contribution of one (x halo plane, dy tile) pair to dW[tap][ci][co] for the given (kh, kw) taps of one kd, as
    wgrad3d_tc.cu computes it: A = x MN-major with TWO taps stacked along M (second atom LBO bytes after the first),
    B = dy MN-major, K = the 128 voxels of the tile in 8 steps of 16 rows."""
    pad, hw, hh = ks // 2, TW + ks - 1, TH + ks - 1
    plane_bytes = (hw * hh * 128 + 128 + 1023) // 1024 * 1024
    smem = Smem(plane_bytes + TW * TH * 128 + 1024)
    tma_load_plane(smem, 0, x, 0, w0 - pad, h0 - pad, d + kd - pad, nb, hw, hh)
    tma_load_plane(smem, plane_bytes, dy, 0, w0, h0, d, nb, TW, TH)
    out = {}
    pairs = [(taps[i], taps[i + 1] if i + 1 < len(taps) else None) for i in range(0, len(taps), 2)]
    for ta, tb in pairs:
        sa = (ta // ks) * hw + ta % ks
        sb = (tb // ks) * hw + tb % ks if tb is not None else sa + 1      # dummy second half
        acc = np.zeros((128, 64), dtype=np.float32)
        for step in range(8):
            a = read_mn_major(smem, (2 * step * hw + sa) * 128, (sb - sa) * 128, hw * 128, 128)
            b = read_mn_major(smem, plane_bytes + step * 2048, 0, 1024, 64)
            acc += a @ b.T
        out[ta] = acc[:64]
        if tb is not None:
            out[tb] = acc[64:]
    return out            # tap -> (ci, co)


def conv3d_k3_tile_linearised(x, w, nb, d0, p0, td=2):
    """PLANNED tiling for the small levels, checked here before any GPU time is spent on it:
    a plane is staged as whole PADDED lines (W + 2 voxels, zero columns left and right by TMA's out-of-bounds fill), so
    that it is one linear row array with pitch L = W + 2; a tile is the 128 CONSECUTIVE padded positions
    p0 .. p0 + 127 (position p = h * L + (w + 1)); a tap is the window starting kh * L + kw rows further, its 8-row
    groups simply contiguous (SBO = 1024).  Outputs at the two pad columns are garbage and dropped by the caller.
    Returns acc[plane][128][cout]."""
    NB, D, H, W, C = x.shape
    cout = w.shape[0]
    L = W + 2
    h_first = p0 // L                                   # first output line touched by the tile
    n_lines = (p0 + 127) // L - h_first + 1 + 2         # + one line above and below
    plane_bytes = ((n_lines * L + 2) * 128 + 1023) // 1024 * 1024
    acc = np.zeros((td, 128, cout), dtype=np.float32)
    for c in range(C // 64):
        smem = Smem((td + 2) * plane_bytes + 1024)
        for i in range(td + 2):
            # row 0 of the buffer is a spare (kw = 0 of the first position looks one row back); lines start at row 1
            tma_load_plane(smem, i * plane_bytes + 128, x, c * 64, -1, h_first - 1, d0 - 1 + i, nb, L, n_lines)
        off = p0 - h_first * L                          # position of p0 inside its line block (buffer row 1 + L + off - 1 for tap (1,0)...)
        for kd in range(3):
            for kh in range(3):
                for kw in range(3):
                    bt = w[:, c * 64:c * 64 + 64, kd, kh, kw]
                    for j in range(td):
                        start = (j + kd) * plane_bytes + (1 + off + kh * L + kw - 1) * 128
                        for k16 in range(4):
                            a = read_k_major(smem, start, 1024, 128, k16)
                            acc[j] += a @ bt[:, k16 * 16:(k16 + 1) * 16].T
    return acc
