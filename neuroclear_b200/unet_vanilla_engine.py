"""Execution plan of the reference's 4-level ``Unet_vanilla`` (models/networks.py:540-608; define_G 'unet_vanilla') on
the same kernels as ``unet_deconv``: tcgen05 implicit-GEMM k3 convolutions, the TMA-store transposed
convolution, fused InstanceNorm + ReLU (+ MaxPool, + concat slice) passes and the 1x1 + sigmoid head.

    down  i = 0..2 : double_conv(c_{i-1} -> c_i), c = 64, 128, 256;  IN+ReLU -> cat_i[:, :c_i]  and MaxPool -> p_{i+1}
    bottom         : double_conv(256 -> 512)
    up    i = 2..0 : t_conv(c_{i+1} -> c_i) -> cat_i[:, c_i:];  double_conv(2 c_i -> c_i)
    head           : one_by_one (64 -> 1) + sigmoid            (no one_by_one_2: the head kernel gets w2 = 1, b2 = 0)

Layout: NDHWC fp16 activations, fp32 statistics, as in unet_engine.py.  InstanceNorm + ReLU between the two convs of a
double_conv is applied inside the consumer (in_mean_rstd) when it has >= 128 output channels — same rule as
unet_engine.  Inference only: training this generator is not on the B200 path (the README trains unet_deconv).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, i64, ptr, stream_ptr

IN_EPS = 1e-5
CH = (64, 128, 256, 512)
# 2 * MACs per network-input voxel: level i has 1/8^i of the voxels
FLOP_PER_VOXEL = 2 * (27 * (1 * 64 + 64 * 64 + 128 * 64 + 64 * 64) + 64
                      + (27 * (64 * 128 + 128 * 128 + 256 * 128 + 128 * 128) + 128 * 64 * 8) / 8
                      + (27 * (128 * 256 + 256 * 256 + 512 * 256 + 256 * 256) + 256 * 128 * 8) / 64
                      + (27 * (256 * 512 + 512 * 512) + 512 * 256 * 8) / 512)


def k3_layers():
    """(state_dict prefix, Cin, Cout) of the 13 tensor-core k3 convolutions, in execution order"""
    out = [("double_conv1.convolution.3", 64, 64)]
    for i in (1, 2):
        out += [("double_conv%d.convolution.0" % (i + 1), CH[i - 1], CH[i]), ("double_conv%d.convolution.3" % (i + 1), CH[i], CH[i])]
    out += [("bottom_layer.convolution.0", 256, 512), ("bottom_layer.convolution.3", 512, 512)]
    out += [("ex_double_conv3.convolution.0", 512, 256), ("ex_double_conv3.convolution.3", 256, 256),
            ("ex_double_conv2.convolution.0", 256, 128), ("ex_double_conv2.convolution.3", 128, 128),
            ("ex_conv1_1.convolution.0", 128, 64), ("ex_conv1_1.convolution.3", 64, 64)]
    return out


CT_LAYERS = [("t_conv3", 512, 256), ("t_conv2", 256, 128), ("t_conv1", 128, 64)]


class UnetVanillaEngine:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeuroclearError("UnetVanillaEngine needs a CUDA device (no CPU fallback)")
        _lib.load()
        self.packed, self.bias = {}, {}
        self.w_first = self.head = None
        self._ws_key = self._ws = None

    def load_state_dict(self, sd):
        dev = self.device
        f = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        lib = _lib.load()
        with torch.cuda.device(dev):
            w1 = f("double_conv1.convolution.0.weight").reshape(64, 27).contiguous()
            self.w_first = torch.empty(8192, dtype=torch.uint8, device=dev)
            call("nc_pack_weights_conv3d_cin1_k3", ptr(w1), ptr(self.w_first), stream_ptr())
            for prefix, cin, cout in k3_layers():
                w = f(prefix + ".weight")
                assert tuple(w.shape) == (cout, cin, 3, 3, 3), (prefix, tuple(w.shape))
                out = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=dev)
                call("nc_pack_weights_conv3d_k3", ptr(w), cout, cin, ptr(out), stream_ptr())
                self.packed[prefix] = out
            for prefix, cin, cout in CT_LAYERS:
                w = f(prefix + ".weight")
                assert tuple(w.shape) == (cin, cout, 2, 2, 2), (prefix, tuple(w.shape))
                out = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 1), dtype=torch.uint8, device=dev)
                call("nc_pack_weights_convT3d_k2s2", ptr(w), cin, cout, ptr(out), stream_ptr())
                self.packed[prefix] = out
                self.bias[prefix] = f(prefix + ".bias")
            one = torch.ones(1, device=dev)
            self.head = torch.cat([f("one_by_one.weight").reshape(64), f("one_by_one.bias").reshape(1), one,
                                   torch.zeros(1, device=dev)]).contiguous()

    def _workspace(self, nb, d, h, w):
        key = (nb, d, h, w)
        if self._ws_key == key:
            return self._ws
        self._ws = None
        dev, lib = self.device, _lib.load()
        vox = [(d >> i) * (h >> i) * (w >> i) for i in range(4)]
        e = lambda n, dt=torch.float16: torch.empty(n, dtype=dt, device=dev)
        ws = {}
        for i in range(4):
            ws["rawA%d" % i], ws["rawB%d" % i] = e(nb * vox[i] * CH[i]), e(nb * vox[i] * CH[i])
            if i < 3:
                ws["cat%d" % i] = e(nb * vox[i] * 2 * CH[i])
                ws["p%d" % (i + 1)] = e(nb * vox[i + 1] * CH[i])
        ws["a0"] = e(nb * vox[0] * 64)
        rows = [lib.nc_conv3d_k3_stats_rows(1, nb, d, h, w, 64) * 64]
        for i in range(4):
            rows.append(lib.nc_conv3d_k3_stats_rows(64, nb, d >> i, h >> i, w >> i, CH[i]) * CH[i])
        ws["stats"] = e(max(rows) * 2, torch.float32)
        ws["mrA"], ws["mrB"] = e(nb * 2 * 512, torch.float32), e(nb * 2 * 512, torch.float32)
        ws["fin"] = torch.zeros(lib.nc_in_stats_scratch_bytes(nb, 512), dtype=torch.uint8, device=dev)
        self._ws_key, self._ws = key, ws
        return ws

    def forward(self, x, crop: int = 0, out=None):
        """x: float32 CUDA (NB, D, H, W) contiguous, D,H,W % 8 == 0 -> float32 (NB, D-2c, H-2c, W-2c)."""
        if self.head is None:
            raise _lib.NeuroclearError("UnetVanillaEngine: weights not loaded")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
            raise _lib.NeuroclearError("UnetVanillaEngine.forward: x must be a contiguous float32 CUDA (NB,D,H,W) tensor")
        nb, d, h, w = x.shape
        if d % 8 or h % 8 or w % 8:
            raise _lib.NeuroclearError("Unet_vanilla needs D, H, W divisible by 8 (three 2x poolings + concat)")
        ws = self._workspace(nb, d, h, w)
        if out is None:
            out = torch.empty((nb, d - 2 * crop, h - 2 * crop, w - 2 * crop), dtype=torch.float32, device=x.device)
        s, lib = stream_ptr(), _lib.load()
        st = ws["stats"]
        dims = [(d >> i, h >> i, w >> i) for i in range(4)]

        def stats(cin, lv, c, mr):
            dd, hh, ww = dims[lv]
            rows = lib.nc_conv3d_k3_stats_rows(cin, nb, dd, hh, ww, c)
            call("nc_in_stats_finalize", ptr(st), nb, i64(rows // nb), c, i64(dd * hh * ww), IN_EPS, ptr(ws["fin"]),
                 ptr(mr), s)

        def conv(prefix, src, src_mr, lv, cin, cout, raw, raw_mr):
            dd, hh, ww = dims[lv]
            call("nc_conv3d_k3_fwd", ptr(src), ptr(src_mr), nb, dd, hh, ww, cin, ptr(self.packed[prefix]), cout,
                 ptr(raw), ptr(st), s)
            stats(cin, lv, cout, raw_mr)

        def apply(raw, mr, lv, c, dst, ld, coff, pooled=None):
            dd, hh, ww = dims[lv]
            call("nc_in_relu_apply", ptr(raw), ptr(mr), nb, dd, hh, ww, c, ptr(dst), ld, coff, ptr(pooled), s)

        def double_conv(name, src, lv, cin, cout):
            """two k3 convs; returns (raw output, its mean/rstd).  `src` is a normalised activation."""
            ra, rb = ws["rawA%d" % lv], ws["rawB%d" % lv]
            conv(name + ".convolution.0", src, None, lv, cin, cout, ra, ws["mrA"])
            if cout >= 128:          # IN + ReLU of the first conv inside the second one
                conv(name + ".convolution.3", ra, ws["mrA"], lv, cout, cout, rb, ws["mrB"])
            else:
                apply(ra, ws["mrA"], lv, cout, ws["a0"], cout, 0)
                conv(name + ".convolution.3", ws["a0"], None, lv, cout, cout, rb, ws["mrB"])
            return rb, ws["mrB"]

        # ---- contracting path
        call("nc_conv3d_cin1_k3_fwd", ptr(x), ptr(self.w_first), nb, d, h, w, 64, ptr(ws["rawA0"]), ptr(st), s)
        stats(1, 0, 64, ws["mrA"])
        apply(ws["rawA0"], ws["mrA"], 0, 64, ws["a0"], 64, 0)
        conv("double_conv1.convolution.3", ws["a0"], None, 0, 64, 64, ws["rawB0"], ws["mrB"])
        apply(ws["rawB0"], ws["mrB"], 0, 64, ws["cat0"], 128, 0, ws["p1"])
        for lv in (1, 2):
            raw, mr = double_conv("double_conv%d" % (lv + 1), ws["p%d" % lv], lv, CH[lv - 1], CH[lv])
            apply(raw, mr, lv, CH[lv], ws["cat%d" % lv], 2 * CH[lv], 0, ws["p%d" % (lv + 1)])
        raw, mr = double_conv("bottom_layer", ws["p3"], 3, 256, 512)
        # ---- expanding path
        for lv, tname, dname in ((2, "t_conv3", "ex_double_conv3"), (1, "t_conv2", "ex_double_conv2"),
                                 (0, "t_conv1", "ex_conv1_1")):
            act = ws["rawA%d" % (lv + 1)]                      # free: reuse as the normalised input of the t_conv
            apply(raw, mr, lv + 1, CH[lv + 1], act, CH[lv + 1], 0)
            dd, hh, ww = dims[lv + 1]
            call("nc_convT3d_k2s2_fwd", ptr(act), None, nb, dd, hh, ww, CH[lv + 1], ptr(self.packed[tname]),
                 ptr(self.bias[tname]), CH[lv], ptr(ws["cat%d" % lv]), 2 * CH[lv], CH[lv], s)
            raw, mr = double_conv(dname, ws["cat%d" % lv], lv, 2 * CH[lv], CH[lv])
        call("nc_head_1x1_sigmoid_fwd", ptr(raw), ptr(mr), ptr(self.head), nb, d, h, w, 64, crop, ptr(out), s)
        return out
