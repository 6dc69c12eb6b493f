"""Execution plan of Unet_deconv.forward (reference models/networks.py:512-538) on the C ABI.

Data layout in HBM (per batch of NB cubes of D x H x W voxels, L0 = full, L1 = /2, L2 = /4 resolution):

    x          fp32 (NB, L0)          dice output / network input
    raw0a/b    fp16 (NB, L0, 64)      raw conv outputs of U1, U2, U12 (ping-pong)
    a1         fp16 (NB, L0, 64)      IN+ReLU(U1)
    cat1       fp16 (NB, L0, 128)     [IN+ReLU(U2) | t_conv1]          = torch.cat([conv1, t_conv1], 1)
    p1         fp16 (NB, L1, 64)      maxpool1
    raw1a/b    fp16 (NB, L1, 128)     raw outputs of U3, U4, U9, U10
    cat2       fp16 (NB, L1, 256)     [IN+ReLU(U4) | t_conv2]          = torch.cat([conv2, t_conv2], 1)
    p2         fp16 (NB, L2, 128)     maxpool2
    raw2a/b    fp16 (NB, L2, 256)     raw outputs of U5, U6, U7
    mrA/mrB    fp32 (NB, 2, 256)      (mean, rstd) of the two most recent raw tensors
    y          fp32 (NB, L0 - 2*crop) sigmoid output, border already cut

All activations are NDHWC.  Where it pays, InstanceNorm + ReLU between two convolutions is applied INSIDE the
consumer conv (in_mean_rstd argument of nc_conv3d_k3_fwd): U3->U4, U5->U6->U7, U9->U10 (measured: +4 % on the
consumer, minus a whole read-raw/write-normalised pass).  A separate apply pass remains (a) for the two skip
connections, whose pass also produces the max-pooled tensor, (b) in front of U2 (the kd-stacked Cout-64 kernel
loses 26 % with the in-kernel transform — more than the pass costs) and (c) in front of the transposed convs
(they re-read their input once per n-tile, so the transform would be repeated 4-8 times).
The concat buffers make torch.cat a no-op (producers write channel slices).
Conv biases in front of InstanceNorm(affine=False) cancel exactly in the mean subtraction and are not applied.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, i64, ptr, stream_ptr

IN_EPS = 1e-5  # torch.nn.InstanceNorm3d default, used by networks.get_norm_layer (networks.py:33-34)

_K3_LAYERS = [  # (state_dict prefix, Cin, Cout)
    ("double_conv1.convolution.3", 64, 64),
    ("double_conv2.convolution.0", 64, 128),
    ("double_conv2.convolution.3", 128, 128),
    ("bottom_layer.convolution.0", 128, 256),
    ("bottom_layer.convolution.3", 256, 256),
    ("bottom_layer.convolution.6", 256, 256),
    ("ex_double_conv2.convolution.0", 256, 128),
    ("ex_double_conv2.convolution.3", 128, 128),
    ("ex_conv1_1.convolution.0", 128, 64),
]
_CT_LAYERS = [("t_conv2", 256, 128), ("t_conv1", 128, 64)]

# 2 * MACs of the 14 convolutions per network-input voxel: 3 642 983 792 000 per 140^3 cube
FLOP_PER_VOXEL = 1_327_618


class UnetDeconvEngine:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeuroclearError("UnetDeconvEngine needs a CUDA device (no CPU fallback)")
        _lib.load()
        self.packed = {}
        self.bias = {}
        self.w_first = None
        self.head = None
        self._ws_key = None
        self._ws = None
        #: when set to a list, forward() brackets every tensor-core conv launch with CUDA events on the launching
        #: stream and appends (layer, flops, start_event, end_event) — bench.py's live roofline measurement
        self.profile = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        dev = self.device
        f = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            w1 = f("double_conv1.convolution.0.weight").reshape(64, 27).contiguous()
            self.w_first = torch.empty(8192, dtype=torch.uint8, device=dev)
            call("nc_pack_weights_conv3d_cin1_k3", ptr(w1), ptr(self.w_first), stream_ptr())
            for prefix, cin, cout in _K3_LAYERS:
                w = f(prefix + ".weight")
                assert tuple(w.shape) == (cout, cin, 3, 3, 3), (prefix, tuple(w.shape))
                out = torch.empty(_lib.load().nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=dev)
                call("nc_pack_weights_conv3d_k3", ptr(w), cout, cin, ptr(out), stream_ptr())
                self.packed[prefix] = out
            for prefix, cin, cout in _CT_LAYERS:
                w = f(prefix + ".weight")
                assert tuple(w.shape) == (cin, cout, 2, 2, 2), (prefix, tuple(w.shape))
                out = torch.empty(_lib.load().nc_packed_weight_bytes(cout, cin, 1), dtype=torch.uint8, device=dev)
                call("nc_pack_weights_convT3d_k2s2", ptr(w), cin, cout, ptr(out), stream_ptr())
                self.packed[prefix] = out
                self.bias[prefix] = f(prefix + ".bias")
            self.head = torch.cat([f("one_by_one.weight").reshape(64), f("one_by_one.bias").reshape(1),
                                   f("one_by_one_2.weight").reshape(1), f("one_by_one_2.bias").reshape(1)]).contiguous()
            # no synchronisation: the fp32 staging copies are freed in stream order (same stream as the packers)

    # ------------------------------------------------------------------ workspace
    def _workspace(self, nb, d, h, w):
        key = (nb, d, h, w)
        if self._ws_key == key:
            return self._ws
        self._ws = None  # release before allocating the new one
        dev = self.device
        l0, l1, l2 = d * h * w, (d // 2) * (h // 2) * (w // 2), (d // 4) * (h // 4) * (w // 4)
        bf, f32 = torch.float16, torch.float32
        e = lambda n, dt: torch.empty(n, dtype=dt, device=dev)
        lib = _lib.load()
        rows = max(
            lib.nc_conv3d_k3_stats_rows(1, nb, d, h, w, 64) * 64,
            lib.nc_conv3d_k3_stats_rows(64, nb, d, h, w, 64) * 64,
            lib.nc_conv3d_k3_stats_rows(64, nb, d // 2, h // 2, w // 2, 128) * 128,
            lib.nc_conv3d_k3_stats_rows(128, nb, d // 4, h // 4, w // 4, 256) * 256,
        )
        ws = dict(
            raw0a=e(nb * l0 * 64, bf), raw0b=e(nb * l0 * 64, bf), a1=e(nb * l0 * 64, bf), cat1=e(nb * l0 * 128, bf), p1=e(nb * l1 * 64, bf),
            raw1a=e(nb * l1 * 128, bf), raw1b=e(nb * l1 * 128, bf), cat2=e(nb * l1 * 256, bf), p2=e(nb * l2 * 128, bf),
            raw2a=e(nb * l2 * 256, bf), raw2b=e(nb * l2 * 256, bf),
            stats=e(rows * 2, f32), mrA=e(nb * 2 * 256, f32), mrB=e(nb * 2 * 256, f32),
            fin=torch.zeros(lib.nc_in_stats_scratch_bytes(nb, 256), dtype=torch.uint8, device=dev),
        )
        self._ws_key, self._ws = key, ws
        return ws

    def workspace_bytes(self, nb, d, h, w):
        ws = self._workspace(nb, d, h, w)
        return sum(t.numel() * t.element_size() for t in ws.values())

    # ------------------------------------------------------------------ forward
    def forward(self, x, crop: int = 0, out=None, nb_cap=None):
        """x: float32 CUDA (NB, D, H, W) contiguous, D,H,W % 4 == 0 -> float32 (NB, D-2c, H-2c, W-2c)."""
        if self.head is None:
            raise _lib.NeuroclearError("UnetDeconvEngine: weights not loaded")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
            raise _lib.NeuroclearError("UnetDeconvEngine.forward: x must be a contiguous float32 CUDA (NB,D,H,W) tensor")
        nb, d, h, w = x.shape
        if d % 4 or h % 4 or w % 4:
            raise _lib.NeuroclearError("Unet_deconv needs D, H, W divisible by 4 (two 2x poolings + concat)")
        ws = self._workspace(max(nb, nb_cap or 0), d, h, w)
        if out is None:
            out = torch.empty((nb, d - 2 * crop, h - 2 * crop, w - 2 * crop), dtype=torch.float32, device=x.device)
        s = stream_ptr()
        lib = _lib.load()
        d1, h1, w1, d2, h2, w2 = d // 2, h // 2, w // 2, d // 4, h // 4, w // 4
        st, mrA, mrB = ws["stats"], ws["mrA"], ws["mrB"]

        def stats(cin, dd, hh, ww, c, mr):
            rows = lib.nc_conv3d_k3_stats_rows(cin, nb, dd, hh, ww, c)
            call("nc_in_stats_finalize", ptr(st), nb, i64(rows // nb), c, i64(dd * hh * ww), IN_EPS, ptr(ws["fin"]),
                 ptr(mr), s)

        def conv(prefix, src, src_mr, dd, hh, ww, cin, cout, raw, raw_mr):
            """src_mr = (mean, rstd) of `src` when src is a RAW conv output (IN+ReLU fused into this conv), else None"""
            if self.profile is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            call("nc_conv3d_k3_fwd", ptr(src), ptr(src_mr), nb, dd, hh, ww, cin, ptr(self.packed[prefix]), cout,
                 ptr(raw), ptr(st), s)
            if cout >= 128 and hh >= 16 and 1 <= hh % 16 <= 6:
                _lib.LAUNCHES += 1          # the remainder-pair kernel: a second launch inside the same call
            if self.profile is not None:
                e1.record()
                self.profile.append((prefix, 2.0 * nb * dd * hh * ww * cout * cin * 27, e0, e1))
            stats(cin, dd, hh, ww, cout, raw_mr)

        def apply(raw, mr, dd, hh, ww, c, dst, ld, coff, pooled=None):
            call("nc_in_relu_apply", ptr(raw), ptr(mr), nb, dd, hh, ww, c, ptr(dst), ld, coff, ptr(pooled), s)

        # ---- level 0 down: U1, U2; conv1 -> cat1[:, :64] and maxpool1
        call("nc_conv3d_cin1_k3_fwd", ptr(x), ptr(self.w_first), nb, d, h, w, 64, ptr(ws["raw0a"]), ptr(st), s)
        stats(1, d, h, w, 64, mrA)
        apply(ws["raw0a"], mrA, d, h, w, 64, ws["a1"], 64, 0)
        conv("double_conv1.convolution.3", ws["a1"], None, d, h, w, 64, 64, ws["raw0b"], mrB)
        apply(ws["raw0b"], mrB, d, h, w, 64, ws["cat1"], 128, 0, ws["p1"])
        # ---- level 1 down: U3, U4; conv2 -> cat2[:, :128] and maxpool2
        conv("double_conv2.convolution.0", ws["p1"], None, d1, h1, w1, 64, 128, ws["raw1a"], mrA)
        conv("double_conv2.convolution.3", ws["raw1a"], mrA, d1, h1, w1, 128, 128, ws["raw1b"], mrB)
        apply(ws["raw1b"], mrB, d1, h1, w1, 128, ws["cat2"], 256, 0, ws["p2"])
        # ---- bottom: U5, U6, U7
        conv("bottom_layer.convolution.0", ws["p2"], None, d2, h2, w2, 128, 256, ws["raw2a"], mrA)
        conv("bottom_layer.convolution.3", ws["raw2a"], mrA, d2, h2, w2, 256, 256, ws["raw2b"], mrB)
        conv("bottom_layer.convolution.6", ws["raw2b"], mrB, d2, h2, w2, 256, 256, ws["raw2a"], mrA)
        # ---- level 1 up: cat2 = [conv2 | t_conv2(IN+ReLU(U7))], U9, U10
        apply(ws["raw2a"], mrA, d2, h2, w2, 256, ws["raw2b"], 256, 0)          # raw2b is free: reuse as IN+ReLU(U7)
        call("nc_convT3d_k2s2_fwd", ptr(ws["raw2b"]), None, nb, d2, h2, w2, 256, ptr(self.packed["t_conv2"]),
             ptr(self.bias["t_conv2"]), 128, ptr(ws["cat2"]), 256, 128, s)
        conv("ex_double_conv2.convolution.0", ws["cat2"], None, d1, h1, w1, 256, 128, ws["raw1a"], mrA)
        conv("ex_double_conv2.convolution.3", ws["raw1a"], mrA, d1, h1, w1, 128, 128, ws["raw1b"], mrB)
        # ---- level 0 up: cat1 = [conv1 | t_conv1(IN+ReLU(U10))], U12
        apply(ws["raw1b"], mrB, d1, h1, w1, 128, ws["raw1a"], 128, 0)          # raw1a is free: reuse as IN+ReLU(U10)
        call("nc_convT3d_k2s2_fwd", ptr(ws["raw1a"]), None, nb, d1, h1, w1, 128, ptr(self.packed["t_conv1"]),
             ptr(self.bias["t_conv1"]), 64, ptr(ws["cat1"]), 128, 64, s)
        conv("ex_conv1_1.convolution.0", ws["cat1"], None, d, h, w, 128, 64, ws["raw0a"], mrA)
        # ---- head: IN + ReLU + 1x1x1 + 1x1x1 + sigmoid (+ border cut)
        call("nc_head_1x1_sigmoid_fwd", ptr(ws["raw0a"]), ptr(mrA), ptr(self.head), nb, d, h, w, 64, crop, ptr(out), s)
        return out

    #: kernels launched by one forward(): 1 + 9 convs, 2 convT, 10 finalize, 5 apply(+pool), 1 head
    LAUNCHES_PER_FORWARD = 1 + 9 + 2 + 10 + 5 + 1

    # ------------------------------------------------------------------ whole-network C entry
    def _weights_struct(self):
        """nc_unet_deconv_weights (include/neuroclear_b200.h): device pointers of the packed images"""
        import ctypes as C

        class W(C.Structure):
            _fields_ = [("first", C.c_void_p), ("k3", C.c_void_p * 9), ("ct", C.c_void_p * 2),
                        ("ct_bias", C.c_void_p * 2), ("head", C.c_void_p)]
        w = W()
        w.first = self.w_first.data_ptr()
        for i, (prefix, _, _) in enumerate(_K3_LAYERS):
            w.k3[i] = self.packed[prefix].data_ptr()
        for i, (prefix, _, _) in enumerate(_CT_LAYERS):
            w.ct[i] = self.packed[prefix].data_ptr()
            w.ct_bias[i] = self.bias[prefix].data_ptr()
        w.head = self.head.data_ptr()
        return w

    def forward_cube(self, x, crop: int = 0, out=None):
        """forward() through ONE library call, nc_unet_deconv_infer_cube (the layer loop runs inside the library; same
        kernels, same order, bit-identical results).  The call is allocation- and synchronisation-free, so it can be
        captured in a CUDA graph (torch.cuda.graph) with fixed x / out tensors and replayed."""
        import ctypes as C
        if self.head is None:
            raise _lib.NeuroclearError("UnetDeconvEngine: weights not loaded")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
            raise _lib.NeuroclearError("UnetDeconvEngine.forward_cube: x must be a contiguous float32 CUDA (NB,D,H,W) tensor")
        nb, d, h, w = x.shape
        lib = _lib.load()
        need = lib.nc_unet_deconv_workspace_bytes(nb, d, h, w)
        if need < 0:
            raise _lib.NeuroclearError(lib.nc_last_error().decode())
        if getattr(self, "_cube_ws_key", None) != (nb, d, h, w):
            self._cube_ws = None
            self._cube_ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            _lib.check(lib.nc_unet_deconv_workspace_init(ptr(self._cube_ws), i64(need), nb, d, h, w, stream_ptr()))
            self._cube_ws_key = (nb, d, h, w)
        if out is None:
            out = torch.empty((nb, d - 2 * crop, h - 2 * crop, w - 2 * crop), dtype=torch.float32, device=x.device)
        wts = self._weights_struct()
        call("nc_unet_deconv_infer_cube", ptr(x), nb, d, h, w, C.byref(wts), ptr(self._cube_ws), i64(need), crop,
             ptr(out), stream_ptr())
        _lib.LAUNCHES += self.LAUNCHES_PER_FORWARD - 1       # `call` counted one
        return out
