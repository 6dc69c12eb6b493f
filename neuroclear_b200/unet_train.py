"""Training-mode execution plan of Unet_deconv (reference models/networks.py:478-538 under autograd, as driven by
AxialToLateralGANApolloModel.backward_G, axial_to_lateral_gan_apollo_model.py:255-283): forward keeping what the
backward needs, and the full backward to every parameter gradient, on the C ABI.

Forward = the inference kernels (fp16 operands), but every layer's RAW conv output and its InstanceNorm statistics
stay resident (per layer: raw fp16 NDHWC + (mean, rstd)); normalised activations are NOT kept — the backward
re-creates them, already in bf16, right before the weight-gradient GEMM that reads them (one memory-bound pass,
~1 % of the layer's conv time).  Gradients flow in bf16 NDHWC (fp16 would underflow), accumulate in fp32.

Backward per k3 layer L (input a_{L-1}, raw output y_L):
    dA_L  --nc_in_relu_bwd-->  dY_L  --nc_conv3d_wgrad(a_{L-1}, dY_L)-->  dW_L
                                     --nc_conv3d_k3_dgrad(dY_L)-------->  dA_{L-1}
with the max-pool / skip-concat routing folded into nc_in_relu_bwd (mode 2), the head into mode 1, and the
transposed convs as space-to-depth + two GEMMs.  Biases in front of InstanceNorm(affine=False) have an exactly
zero gradient (the mean subtraction removes them); their .grad is returned as zeros.
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, i64, ptr, stream_ptr
from .unet_engine import IN_EPS, UnetDeconvEngine, _CT_LAYERS, _K3_LAYERS

# resolution level (0 = full, 1 = /2, 2 = /4) of the nine tensor-core k3 layers
_LEVEL = {"double_conv1.convolution.3": 0, "double_conv2.convolution.0": 1, "double_conv2.convolution.3": 1,
          "bottom_layer.convolution.0": 2, "bottom_layer.convolution.3": 2, "bottom_layer.convolution.6": 2,
          "ex_double_conv2.convolution.0": 1, "ex_double_conv2.convolution.3": 1, "ex_conv1_1.convolution.0": 0}
_FIRST = "double_conv1.convolution.0"


class UnetDeconvTrainEngine:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeuroclearError("UnetDeconvTrainEngine needs a CUDA device (no CPU fallback)")
        self.fwd = UnetDeconvEngine(self.device)   # packed forward weights (fp16 images)
        self.packed_dgrad = {}
        self.saved = None
        self.launches = 0
        #: set to a dict to keep every intermediate gradient tensor of the next backward() (tests / debugging)
        self.debug = None

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        self.fwd.load_state_dict(sd)
        dev = self.device
        lib = _lib.load()
        f = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        with torch.cuda.device(dev):
            for prefix, cin, cout in _K3_LAYERS:
                w = f(prefix + ".weight")
                out = torch.empty(lib.nc_packed_weight_bytes(cout, cin, 0), dtype=torch.uint8, device=dev)
                call("nc_pack_weights_conv3d_k3_dgrad", ptr(w), cout, cin, ptr(out), stream_ptr())
                self.packed_dgrad[prefix] = out
            for prefix, cin, cout in _CT_LAYERS:
                w = f(prefix + ".weight")
                out = torch.empty(8 * cout * cin * 2, dtype=torch.uint8, device=dev)
                call("nc_pack_weights_convT3d_k2s2_dgrad", ptr(w), cin, cout, ptr(out), stream_ptr())
                self.packed_dgrad[prefix] = out
            # no synchronisation: the fp32 staging copies are freed in stream order (same stream as the packers)

    # ------------------------------------------------------------------ forward (keeps raw outputs + statistics)
    def forward(self, x):
        """x: float32 CUDA (NB, D, H, W) contiguous, D,H,W % 4 == 0 -> float32 (NB, D, H, W)."""
        eng = self.fwd
        if eng.head is None:
            raise _lib.NeuroclearError("UnetDeconvTrainEngine: weights not loaded")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
            raise _lib.NeuroclearError("forward: x must be a contiguous float32 CUDA (NB,D,H,W) tensor")
        nb, d, h, w = x.shape
        if d % 4 or h % 4 or w % 4:
            raise _lib.NeuroclearError("Unet_deconv needs D, H, W divisible by 4 (two 2x poolings + concat)")
        dev, lib, s = self.device, _lib.load(), stream_ptr()
        dims = [(d, h, w), (d // 2, h // 2, w // 2), (d // 4, h // 4, w // 4)]
        vox = [a * b * c for a, b, c in dims]
        f16, f32 = torch.float16, torch.float32
        e = lambda n, dt: torch.empty(n, dtype=dt, device=dev)
        rows = max(lib.nc_conv3d_k3_stats_rows(1, nb, d, h, w, 64) * 64,
                   lib.nc_conv3d_k3_stats_rows(64, nb, d, h, w, 64) * 64,
                   lib.nc_conv3d_k3_stats_rows(64, nb, *dims[1], 128) * 128,
                   lib.nc_conv3d_k3_stats_rows(128, nb, *dims[2], 256) * 256)
        st = e(rows * 2, f32)
        fin = torch.zeros(lib.nc_in_stats_scratch_bytes(nb, 256), dtype=torch.uint8, device=dev)
        raw, mr = {}, {}
        tmp = [e(nb * vox[0] * 64, f16), e(nb * vox[1] * 128, f16), e(nb * vox[2] * 256, f16)]   # normalised inputs
        cat1, cat2 = e(nb * vox[0] * 128, f16), e(nb * vox[1] * 256, f16)
        p1, p2 = e(nb * vox[1] * 64, f16), e(nb * vox[2] * 128, f16)
        n0 = self._count()

        def stats(name, cin, lvl, c):
            dd, hh, ww = dims[lvl]
            r = lib.nc_conv3d_k3_stats_rows(cin, nb, dd, hh, ww, c)
            mr[name] = e(nb * 2 * c, f32)
            call("nc_in_stats_finalize", ptr(st), nb, i64(r // nb), c, i64(dd * hh * ww), IN_EPS, ptr(fin),
                 ptr(mr[name]), s)

        def conv(name, src, cin, cout):
            lvl = _LEVEL[name]
            dd, hh, ww = dims[lvl]
            raw[name] = e(nb * vox[lvl] * cout, f16)
            call("nc_conv3d_k3_fwd", ptr(src), None, nb, dd, hh, ww, cin, ptr(eng.packed[name]), cout,
                 ptr(raw[name]), ptr(st), s)
            stats(name, cin, lvl, cout)

        def apply(name, lvl, c, dst, ld, coff, pooled=None):
            dd, hh, ww = dims[lvl]
            call("nc_in_relu_apply", ptr(raw[name]), ptr(mr[name]), nb, dd, hh, ww, c, ptr(dst), ld, coff,
                 ptr(pooled), s)

        raw[_FIRST] = e(nb * vox[0] * 64, f16)
        call("nc_conv3d_cin1_k3_fwd", ptr(x), ptr(eng.w_first), nb, d, h, w, 64, ptr(raw[_FIRST]), ptr(st), s)
        stats(_FIRST, 1, 0, 64)
        apply(_FIRST, 0, 64, tmp[0], 64, 0)
        conv("double_conv1.convolution.3", tmp[0], 64, 64)
        apply("double_conv1.convolution.3", 0, 64, cat1, 128, 0, p1)
        conv("double_conv2.convolution.0", p1, 64, 128)
        apply("double_conv2.convolution.0", 1, 128, tmp[1], 128, 0)
        conv("double_conv2.convolution.3", tmp[1], 128, 128)
        apply("double_conv2.convolution.3", 1, 128, cat2, 256, 0, p2)
        conv("bottom_layer.convolution.0", p2, 128, 256)
        apply("bottom_layer.convolution.0", 2, 256, tmp[2], 256, 0)
        conv("bottom_layer.convolution.3", tmp[2], 256, 256)
        apply("bottom_layer.convolution.3", 2, 256, tmp[2], 256, 0)
        conv("bottom_layer.convolution.6", tmp[2], 256, 256)
        apply("bottom_layer.convolution.6", 2, 256, tmp[2], 256, 0)
        call("nc_convT3d_k2s2_fwd", ptr(tmp[2]), None, nb, *dims[2], 256, ptr(eng.packed["t_conv2"]),
             ptr(eng.bias["t_conv2"]), 128, ptr(cat2), 256, 128, s)
        conv("ex_double_conv2.convolution.0", cat2, 256, 128)
        apply("ex_double_conv2.convolution.0", 1, 128, tmp[1], 128, 0)
        conv("ex_double_conv2.convolution.3", tmp[1], 128, 128)
        apply("ex_double_conv2.convolution.3", 1, 128, tmp[1], 128, 0)
        call("nc_convT3d_k2s2_fwd", ptr(tmp[1]), None, nb, *dims[1], 128, ptr(eng.packed["t_conv1"]),
             ptr(eng.bias["t_conv1"]), 64, ptr(cat1), 128, 64, s)
        conv("ex_conv1_1.convolution.0", cat1, 128, 64)
        out = e(nb * vox[0], f32).view(nb, d, h, w)
        call("nc_head_1x1_sigmoid_fwd", ptr(raw["ex_conv1_1.convolution.0"]), ptr(mr["ex_conv1_1.convolution.0"]),
             ptr(eng.head), nb, d, h, w, 64, 0, ptr(out), s)
        # the transposed convs' outputs are only needed as (bf16) inputs of the weight gradients: keep the fp16 halves
        self.saved = dict(x=x, raw=raw, mr=mr, cat1=cat1, cat2=cat2, dims=dims, vox=vox, nb=nb)
        self.launches += self._count() - n0
        return out

    @staticmethod
    def _count():
        return _lib.LAUNCHES if hasattr(_lib, "LAUNCHES") else 0

    # ------------------------------------------------------------------ backward
    def backward(self, dout):
        """dout: float32 CUDA (NB, D, H, W) = dL/d output.  Returns {state_dict key: float32 gradient}."""
        sv = self.saved
        if sv is None:
            raise _lib.NeuroclearError("backward() without a preceding forward()")
        eng, dev, lib, s = self.fwd, self.device, _lib.load(), stream_ptr()
        raw, mr, dims, vox, nb = sv["raw"], sv["mr"], sv["dims"], sv["vox"], sv["nb"]
        dout = dout.to(torch.float32).contiguous()
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda n, dt: torch.empty(n, dtype=dt, device=dev)
        scratch = e(lib.nc_bwd_scratch_bytes(nb) // 4, f32)
        m12 = e(nb * 2 * 256, f32)
        grads = {}
        n0 = self._count()

        def act_bf16(name, lvl, c, dst, ld, coff, pooled=None):
            """relu(IN(raw[name])) as bf16 into a channel slice of dst (+ its max-pool)"""
            call("nc_in_relu_apply_bf16", ptr(raw[name]), ptr(mr[name]), nb, *dims[lvl], c, ptr(dst), ld, coff,
                 ptr(pooled), s)

        def in_bwd(name, lvl, c, mode, grad=None, ld=0, coff=0, du=None, w_head=None, dpool=None):
            d_raw = e(nb * vox[lvl] * c, bf)
            call("nc_in_relu_bwd", ptr(raw[name]), ptr(mr[name]), nb, *dims[lvl], c, mode, ptr(grad), ld, coff,
                 ptr(du), ptr(w_head), ptr(dpool), ptr(scratch), ptr(m12), ptr(d_raw), s)
            if self.debug is not None:
                self.debug["d_raw." + name] = d_raw.view(nb, *dims[lvl], c)
            return d_raw

        def wgrad(name, x_bf16, d_raw, lvl, cin, cout):
            sb = lib.nc_conv3d_wgrad_scratch_bytes(3, nb, *dims[lvl], cin, cout)
            ws = e(sb, torch.uint8)
            dw = e(cout * cin * 27, f32)
            call("nc_conv3d_wgrad", ptr(x_bf16), 1, ptr(d_raw), 1, nb, *dims[lvl], cin, cout, 3, ptr(ws), ptr(dw), s)
            grads[name + ".weight"] = dw.view(cout, cin, 3, 3, 3)
            grads[name + ".bias"] = torch.zeros(cout, dtype=f32, device=dev)   # exactly zero: bias before IN

        def dgrad(name, d_raw, lvl, cin, cout):
            dx = e(nb * vox[lvl] * cin, bf)
            call("nc_conv3d_k3_dgrad", ptr(d_raw), nb, *dims[lvl], cout, ptr(self.packed_dgrad[name]), cin, ptr(dx), s)
            if self.debug is not None:
                self.debug["d_in." + name] = dx.view(nb, *dims[lvl], cin)
            return dx

        def convT_bwd(name, x_bf16, dcat, ld, coff, lvl_in, cin, cout):
            """x: (.., cin) at level lvl_in; the layer's output gradient is dcat[..., coff:coff+cout] one level up"""
            g = e(nb * vox[lvl_in] * 8 * cout, bf)
            call("nc_space_to_depth_bf16", ptr(dcat), ld, coff, nb, *dims[lvl_in], cout, ptr(g), s)
            dx = e(nb * vox[lvl_in] * cin, bf)
            call("nc_conv3d_k1_bf16", ptr(g), nb, *dims[lvl_in], 8 * cout, ptr(self.packed_dgrad[name]), cin,
                 ptr(dx), s)
            sb = lib.nc_conv3d_wgrad_scratch_bytes(1, nb, *dims[lvl_in], cin, 8 * cout)
            ws = e(sb, torch.uint8)
            dw = e(8 * cout * cin, f32)
            call("nc_conv3d_wgrad", ptr(x_bf16), 1, ptr(g), 1, nb, *dims[lvl_in], cin, 8 * cout, 1, ptr(ws), ptr(dw), s)
            # (tap, co, ci) -> the parameter's (ci, co, 2, 2, 2): a 1 MB re-layout
            grads[name + ".weight"] = dw.view(8, cout, cin).permute(2, 1, 0).reshape(cin, cout, 2, 2, 2).contiguous()
            db = e(cout, f32)
            call("nc_colsum_bf16", ptr(dcat), ld, coff, nb, i64(vox[lvl_in] * 8), cout, ptr(scratch), ptr(db), s)
            grads[name + ".bias"] = db
            if self.debug is not None:
                self.debug["d_in." + name] = dx.view(nb, *dims[lvl_in], cin)
            return dx

        # ---- head: du, gradients of one_by_one / one_by_one_2
        L12 = "ex_conv1_1.convolution.0"
        du = e(nb * vox[0], f32)
        hg = e(68, f32)
        call("nc_head_1x1_sigmoid_bwd", ptr(raw[L12]), ptr(mr[L12]), ptr(eng.head), ptr(dout), nb, *dims[0], ptr(du),
             ptr(scratch), ptr(hg), s)
        grads["one_by_one.weight"] = hg[:64].reshape(1, 64, 1, 1, 1)
        grads["one_by_one.bias"] = hg[64:65]
        grads["one_by_one_2.weight"] = hg[65:66].reshape(1, 1, 1, 1, 1)
        grads["one_by_one_2.bias"] = hg[66:67]
        # ---- level 0 up: U12 <- cat1 = [a2 | t_conv1(a10)]
        d12 = in_bwd(L12, 0, 64, 1, du=du, w_head=eng.head)
        cat1b = e(nb * vox[0] * 128, bf)
        p1b = e(nb * vox[1] * 64, bf)
        act_bf16("double_conv1.convolution.3", 0, 64, cat1b, 128, 0, p1b)
        call("nc_cast_f16_bf16", ptr(sv["cat1"]), 128, 64, i64(nb * vox[0]), 64, ptr(cat1b), 128, 64, s)
        wgrad(L12, cat1b, d12, 0, 128, 64)
        dcat1 = dgrad(L12, d12, 0, 128, 64)
        del d12, cat1b, du
        # ---- t_conv1 <- a10
        L10, L9 = "ex_double_conv2.convolution.3", "ex_double_conv2.convolution.0"
        a = e(nb * vox[1] * 128, bf)
        act_bf16(L10, 1, 128, a, 128, 0)
        dA10 = convT_bwd("t_conv1", a, dcat1, 128, 64, 1, 128, 64)
        # ---- level 1 up: U10 <- a9, U9 <- cat2 = [a4 | t_conv2(a7)]
        d10 = in_bwd(L10, 1, 128, 0, grad=dA10, ld=128, coff=0)
        act_bf16(L9, 1, 128, a, 128, 0)
        wgrad(L10, a, d10, 1, 128, 128)
        dA9 = dgrad(L10, d10, 1, 128, 128)
        d9 = in_bwd(L9, 1, 128, 0, grad=dA9, ld=128, coff=0)
        cat2b = e(nb * vox[1] * 256, bf)
        p2b = e(nb * vox[2] * 128, bf)
        act_bf16("double_conv2.convolution.3", 1, 128, cat2b, 256, 0, p2b)
        call("nc_cast_f16_bf16", ptr(sv["cat2"]), 256, 128, i64(nb * vox[1]), 128, ptr(cat2b), 256, 128, s)
        wgrad(L9, cat2b, d9, 1, 256, 128)
        dcat2 = dgrad(L9, d9, 1, 256, 128)
        del d10, d9, dA10, dA9, cat2b
        # ---- t_conv2 <- a7, bottom: U7 <- a6, U6 <- a5, U5 <- p2
        L7, L6, L5 = "bottom_layer.convolution.6", "bottom_layer.convolution.3", "bottom_layer.convolution.0"
        b = e(nb * vox[2] * 256, bf)
        act_bf16(L7, 2, 256, b, 256, 0)
        dA7 = convT_bwd("t_conv2", b, dcat2, 256, 128, 2, 256, 128)
        d7 = in_bwd(L7, 2, 256, 0, grad=dA7, ld=256, coff=0)
        act_bf16(L6, 2, 256, b, 256, 0)
        wgrad(L7, b, d7, 2, 256, 256)
        dA6 = dgrad(L7, d7, 2, 256, 256)
        d6 = in_bwd(L6, 2, 256, 0, grad=dA6, ld=256, coff=0)
        act_bf16(L5, 2, 256, b, 256, 0)
        wgrad(L6, b, d6, 2, 256, 256)
        dA5 = dgrad(L6, d6, 2, 256, 256)
        d5 = in_bwd(L5, 2, 256, 0, grad=dA5, ld=256, coff=0)
        wgrad(L5, p2b, d5, 2, 128, 256)
        dp2 = dgrad(L5, d5, 2, 128, 256)
        del d7, d6, d5, dA7, dA6, dA5, b
        # ---- level 1 down: U4 (skip dcat2[:, :128] + pool route dp2) <- a3, U3 <- p1
        L4, L3 = "double_conv2.convolution.3", "double_conv2.convolution.0"
        d4 = in_bwd(L4, 1, 128, 2, grad=dcat2, ld=256, coff=0, dpool=dp2)
        act_bf16(L3, 1, 128, a, 128, 0)
        wgrad(L4, a, d4, 1, 128, 128)
        dA3 = dgrad(L4, d4, 1, 128, 128)
        d3 = in_bwd(L3, 1, 128, 0, grad=dA3, ld=128, coff=0)
        wgrad(L3, p1b, d3, 1, 64, 128)
        dp1 = dgrad(L3, d3, 1, 64, 128)
        del d4, d3, dA3, dcat2, dp2, a
        # ---- level 0 down: U2 (skip dcat1[:, :64] + pool route dp1) <- a1, U1 <- x
        L2 = "double_conv1.convolution.3"
        d2 = in_bwd(L2, 0, 64, 2, grad=dcat1, ld=128, coff=0, dpool=dp1)
        a1 = e(nb * vox[0] * 64, bf)
        act_bf16(_FIRST, 0, 64, a1, 64, 0)
        wgrad(L2, a1, d2, 0, 64, 64)
        dA1 = dgrad(L2, d2, 0, 64, 64)
        del d2, a1, dcat1, dp1
        d1 = in_bwd(_FIRST, 0, 64, 0, grad=dA1, ld=64, coff=0)
        dw1 = e(64 * 27, f32)
        call("nc_conv3d_cin1_k3_wgrad", ptr(sv["x"]), ptr(d1), 1, nb, *dims[0], ptr(scratch), ptr(dw1), s)
        grads[_FIRST + ".weight"] = dw1.view(64, 1, 3, 3, 3)
        grads[_FIRST + ".bias"] = torch.zeros(64, dtype=f32, device=dev)
        self.saved = None
        self.launches += self._count() - n0
        return grads
