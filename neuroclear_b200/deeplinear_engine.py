"""Execution plan of DeepLinearGenerator (reference models/networks.py:893-917; G_B of the apollo model,
axial_to_lateral_gan_apollo_model.py:95-97) on the C ABI, forward and backward.

    x fp32 --im2col49--> X49 fp16 --7-tap depth conv (tcgen05)--> h1 fp16 --k5 conv (tcgen05)--> h2 fp16
      --64->1 k3 stencil with K = fold(W3, W4, W5, W6)--> out fp32

The k3 layer and the three 1x1 layers are linear and followed only by pointwise ops, so they fold exactly into one
64->1 stencil (include/neuroclear_b200.h); the fold and its chain rule are ~1e5 flops done with torch on the device.
Backward: gradients in bf16 NDHWC, weight gradients by the tcgen05 split-K GEMM (nc_conv3d_wgrad), the data
gradients by the same conv kernels with flipped filters; returns d/dx too (G_B's input is G_A's output).
"""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call, i64, ptr, stream_ptr

KEYS = ("first_layer.weight", "feature_block.0.weight", "feature_block.1.weight", "feature_block.2.weight",
        "feature_block.3.weight", "final_layer.weight")
SHAPES = {"first_layer.weight": (64, 1, 7, 7, 7), "feature_block.0.weight": (64, 64, 5, 5, 5),
          "feature_block.1.weight": (64, 64, 3, 3, 3), "feature_block.2.weight": (32, 64, 1, 1, 1),
          "feature_block.3.weight": (16, 32, 1, 1, 1), "final_layer.weight": (1, 16, 1, 1, 1)}
#: 2 * MACs per voxel of the reference's six layers (1 630.4 GFLOP at 108^3)
FLOP_PER_VOXEL = 2 * (343 * 64 + 125 * 64 * 64 + 27 * 64 * 64 + 64 * 32 + 32 * 16 + 16)


def fold_tail(w3, w4, w5, w6):
    """K[ci][tap] = sum_co (W6 W5 W4)[co] * W3[co][ci][tap]  -> (64, 27)"""
    w_eff = (w6.reshape(1, 16) @ w5.reshape(16, 32) @ w4.reshape(32, 64)).reshape(64)
    return torch.einsum("o,oit->it", w_eff, w3.reshape(64, 64, 27)).contiguous()


class DeepLinearEngine:
    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise _lib.NeuroclearError("DeepLinearEngine needs a CUDA device (no CPU fallback)")
        _lib.load()
        self.w = None
        self.saved = None

    def load_state_dict(self, sd):
        dev = self.device
        f = lambda k: sd[k].detach().to(dev, torch.float32).contiguous()
        for k in KEYS:
            if tuple(sd[k].shape) != SHAPES[k]:
                raise _lib.NeuroclearError("DeepLinearGenerator: %s has shape %s" % (k, tuple(sd[k].shape)))
        with torch.cuda.device(dev):
            w1 = torch.zeros((64, 64, 7), dtype=torch.float32, device=dev)      # [co][kh*7+kw (49, zero padded)][kd]
            w1[:, :49] = f(KEYS[0]).reshape(64, 7, 49).permute(0, 2, 1)
            w2 = f(KEYS[1]).reshape(64, 64, 125)
            pk = {}
            for name, w, taps in (("l1", w1, 7), ("l2", w2, 125)):
                for dgrad in (0, 1):
                    out = torch.empty(64 * 64 * taps * 2, dtype=torch.uint8, device=dev)
                    call("nc_pack_weights_64", ptr(w), taps, dgrad, ptr(out), stream_ptr())
                    pk[name, dgrad] = out
            self.tail = [f(k) for k in KEYS[2:]]
            self.K = fold_tail(*self.tail)
            self.w = pk
            # no synchronisation: the fp32 staging copies are freed in stream order (same stream as the packers)

    # ------------------------------------------------------------------ forward
    def forward(self, x, keep=True):
        """x: float32 CUDA (NB, D, H, W) contiguous -> float32 (NB, D, H, W)"""
        if self.w is None:
            raise _lib.NeuroclearError("DeepLinearEngine: weights not loaded")
        if not (x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 4):
            raise _lib.NeuroclearError("forward: x must be a contiguous float32 CUDA (NB,D,H,W) tensor")
        nb, d, h, w = x.shape
        dev, s = self.device, stream_ptr()
        n = nb * d * h * w
        e = lambda dt: torch.empty(n * 64, dtype=dt, device=dev)
        x49 = e(torch.float16)
        call("nc_im2col49", ptr(x), nb, d, h, w, 0, ptr(x49), s)
        h1 = e(torch.float16)
        call("nc_conv3d_tc_64", ptr(x49), 0, nb, d, h, w, ptr(self.w["l1", 0]), 7, 1, ptr(h1), s)
        del x49
        h2 = e(torch.float16)
        call("nc_conv3d_tc_64", ptr(h1), 0, nb, d, h, w, ptr(self.w["l2", 0]), 5, 5, ptr(h2), s)
        out = torch.empty((nb, d, h, w), dtype=torch.float32, device=dev)
        call("nc_stencil64to1_fwd", ptr(h2), ptr(self.K), nb, d, h, w, ptr(out), s)
        self.saved = dict(x=x, h1=h1, h2=h2) if keep else None
        return out

    # ------------------------------------------------------------------ backward
    def backward(self, dout):
        """dout: float32 (NB, D, H, W).  Returns (dx float32 (NB, D, H, W), {state_dict key: gradient})."""
        sv = self.saved
        if sv is None:
            raise _lib.NeuroclearError("backward() without a preceding forward()")
        x, h1, h2 = sv["x"], sv["h1"], sv["h2"]
        nb, d, h, w = x.shape
        dev, s, lib = self.device, stream_ptr(), _lib.load()
        n = nb * d * h * w
        bf, f32 = torch.bfloat16, torch.float32
        e = lambda m, dt: torch.empty(m, dtype=dt, device=dev)
        dout = dout.to(f32).contiguous()
        scratch = e(lib.nc_bwd_scratch_bytes(nb) // 4, f32)
        grads = {}
        # ---- folded tail: dK (taps reversed: the kernel correlates x = dout against dy = h2), dh2
        dk_rev = e(64 * 27, f32)
        call("nc_conv3d_cin1_k3_wgrad", ptr(dout), ptr(h2), 0, nb, d, h, w, ptr(scratch), ptr(dk_rev), s)
        dK = dk_rev.view(64, 27).flip(1)
        leaves = [t.clone().requires_grad_(True) for t in self.tail]
        with torch.enable_grad():
            fold_tail(*leaves).backward(dK)
        for k, t in zip(KEYS[2:], leaves):
            grads[k] = t.grad
        dh2 = e(n * 64, bf)
        call("nc_stencil64to1_bwd_data", ptr(dout), ptr(self.K), nb, d, h, w, ptr(dh2), s)
        # ---- k5 layer
        h1b = e(n * 64, bf)
        call("nc_cast_f16_bf16", ptr(h1), 64, 0, i64(n), 64, ptr(h1b), 64, 0, s)
        ws = e(lib.nc_conv3d_wgrad_scratch_bytes(5, nb, d, h, w, 64, 64), torch.uint8)
        dw2 = e(64 * 64 * 125, f32)
        call("nc_conv3d_wgrad", ptr(h1b), 1, ptr(dh2), 1, nb, d, h, w, 64, 64, 5, ptr(ws), ptr(dw2), s)
        grads[KEYS[1]] = dw2.view(64, 64, 5, 5, 5)
        dh1 = h1b                                                                # h1b is dead after the wgrad: reuse
        call("nc_conv3d_tc_64", ptr(dh2), 1, nb, d, h, w, ptr(self.w["l2", 1]), 5, 5, ptr(dh1), s)
        # ---- k7 layer on the 49-channel im2col
        x49 = dh2                                                                # dh2 is dead: reuse
        call("nc_im2col49", ptr(x), nb, d, h, w, 1, ptr(x49), s)
        ws = e(lib.nc_conv3d_wgrad_scratch_bytes(71, nb, d, h, w, 64, 64), torch.uint8)
        dw1 = e(64 * 64 * 7, f32)
        call("nc_conv3d_wgrad", ptr(x49), 1, ptr(dh1), 1, nb, d, h, w, 64, 64, 71, ptr(ws), ptr(dw1), s)
        grads[KEYS[0]] = dw1.view(64, 64, 7)[:, :49].permute(0, 2, 1).reshape(64, 1, 7, 7, 7).contiguous()
        dx49 = x49
        call("nc_conv3d_tc_64", ptr(dh1), 1, nb, d, h, w, ptr(self.w["l1", 1]), 7, 1, ptr(dx49), s)
        dx = e(n, f32).view(nb, d, h, w)
        call("nc_col2im49", ptr(dx49), nb, d, h, w, ptr(dx), s)
        self.saved = None
        return dx, grads
