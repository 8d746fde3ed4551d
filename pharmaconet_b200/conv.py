"""Host side of the tcgen05 3x3x3 convolution (csrc/conv3d.cu): layout conversion, weight packing, the C-ABI call.

Replaces `BaseConv3d.forward` = act(norm(conv(x))) (src/pmnet/network/nn/layers.py:45-46) for the 96 -> 96, k = 3
shape that dominates the reference network. BatchNorm3d in eval mode is folded into a
per-channel scale and bias (`fold_bn`). torch provides device memory and streams only.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib

CH = 96


def to_c8(x: torch.Tensor) -> torch.Tensor:
    """[B, C, D, H, W] (any float dtype) -> bf16 [B, C/8, D, H, W, 8]."""
    B, Cc, D, H, W = x.shape
    assert Cc % 8 == 0
    return x.reshape(B, Cc // 8, 8, D, H, W).permute(0, 1, 3, 4, 5, 2).contiguous().to(torch.bfloat16)


def from_c8(y: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    """bf16 [B, C/8, D, H, W, 8] -> [B, C, D, H, W]."""
    B, C8, D, H, W, _ = y.shape
    return y.permute(0, 1, 5, 2, 3, 4).reshape(B, C8 * 8, D, H, W).to(dtype)


class Act:
    """A c8 activation as the convolution consumes it: `hi` = bf16 [B, C/8, D, H, W, 8]; `lo` = the bf16 residual of a
    two-term split (value = hi + lo to ~2^-16 relative) in the split-precision mode, else None."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor | None = None):
        self.hi, self.lo = hi, lo

    @property
    def shape(self):
        return self.hi.shape

    def __getitem__(self, idx) -> "Act":
        return Act(self.hi[idx], None if self.lo is None else self.lo[idx])

    def float_ncdhw(self) -> torch.Tensor:
        y = from_c8(self.hi)
        return y if self.lo is None else y + from_c8(self.lo)


def split_bf16(x: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
    """fp32 -> (hi, lo) bf16 with hi + lo = x up to 2^-17 relative."""
    x = x.float()
    hi = x.to(torch.bfloat16)
    return hi, (x - hi.float()).to(torch.bfloat16)


def to_act(x: torch.Tensor, split: bool) -> Act:
    """[B, C, D, H, W] fp32 -> Act (one or two bf16 c8 tensors)."""
    if not split:
        return Act(to_c8(x))
    hi, lo = split_bf16(x)
    return Act(to_c8(hi), to_c8(lo))


def pack_weights_k3(w: torch.Tensor) -> torch.Tensor:
    """Conv3d weight [C_out, C_in, 3, 3, 3] -> bf16 [27, C_in/8, C_out, 8]: for every tap (kd, kh, kw) the
    no-swizzle K-major image of the B operand (8 x 16 B core matrices, C_out rows)."""
    co, ci, kd, kh, kw = w.shape
    assert (kd, kh, kw) == (3, 3, 3) and ci % 8 == 0
    t = w.permute(2, 3, 4, 1, 0).reshape(27, ci // 8, 8, co)  # [tap][chunk][8][co]
    return t.permute(0, 1, 3, 2).contiguous().to(torch.bfloat16)


def fold_bn(conv_bias, bn_weight, bn_bias, running_mean, running_var, eps: float = 1e-5):
    """BatchNorm3d(eval) after a conv -> (scale, bias) per output channel, fp32."""
    scale = bn_weight.float() / torch.sqrt(running_var.float() + eps)
    bias = bn_bias.float() - running_mean.float() * scale
    if conv_bias is not None:
        bias = bias + conv_bias.float() * scale
    return scale.contiguous(), bias.contiguous()


def conv3d_k3_c96(
    x_c8: torch.Tensor,
    w_packed: torch.Tensor,
    scale: torch.Tensor,
    bias: torch.Tensor,
    relu: bool = True,
    head_w: torch.Tensor | None = None,
    head_b: float = 0.0,
    store_out: bool = True,
    planes_per_item: int = 0,
    max_ctas: int = 0,
):
    """y = act(scale * conv3d(x, w, padding=1) + bias) in c8 layout; optionally also the fused 1x1 -> 1 head
    `head = sum_c head_w[c] * y[c] + head_b` as fp32 [B, D, H, W]. Returns (y_c8 or None, head or None)."""
    if not x_c8.is_cuda:
        raise RuntimeError("conv3d_k3_c96 runs on CUDA devices only (no CPU fallback)")
    B, C8, D, H, W, e = x_c8.shape
    assert C8 * e == CH and e == 8 and x_c8.dtype == torch.bfloat16 and x_c8.is_contiguous()
    assert tuple(w_packed.shape) == (27, CH // 8, CH, 8) and w_packed.dtype == torch.bfloat16 and w_packed.is_contiguous()
    assert scale.dtype == torch.float32 and bias.dtype == torch.float32 and scale.numel() == CH and bias.numel() == CH
    L = _lib.lib()
    dev = x_c8.device
    with torch.cuda.device(dev):
        y = torch.empty_like(x_c8) if store_out else None
        head = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if head_w is not None else None
        if planes_per_item <= 0:
            # >= 4 work items per CTA when the problem allows it (static round-robin: keeps the last wave short);
            # every item re-loads 2 halo planes and refills the pipeline, so items are not made shorter than 8 planes
            tiles = B * -(-H // 16) * -(-W // 8)
            planes_per_item = D
            while planes_per_item > 8 and tiles * (D // planes_per_item) < 4 * 148:
                planes_per_item //= 2
            planes_per_item = max(2, planes_per_item)
        rc = L.pmnet_conv3d_k3_c96(
            x_c8.data_ptr(), w_packed.data_ptr(), scale.data_ptr(), bias.data_ptr(),
            y.data_ptr() if store_out else None,
            head_w.data_ptr() if head_w is not None else None, C.c_float(float(head_b)),
            head.data_ptr() if head is not None else None,
            B, D, H, W, int(relu), int(planes_per_item), int(max_ctas),
            C.c_void_p(torch.cuda.current_stream(dev).cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_conv3d_k3_c96")
    return y, head


def conv3d_k3_c96_x3(
    x: Act,
    w_hi: torch.Tensor,
    w_lo: torch.Tensor,
    scale: torch.Tensor,
    bias: torch.Tensor,
    relu: bool = True,
    head_w: torch.Tensor | None = None,
    head_b: float = 0.0,
    store_out: bool = True,
):
    """The same convolution with two-term bf16 splits of activations and weights: x w ~ x_hi w_hi + x_lo w_hi + x_hi w_lo
    (relative error 2^-16 instead of 2^-8), as three launches of the tcgen05 kernel that chain their fp32 accumulators
    through a global buffer; the last launch applies scale / bias / ReLU / head and emits the (hi, lo) pair of the next
    layer. Returns (Act or None, head or None)."""
    assert x.lo is not None
    B, C8, D, H, W, e = x.hi.shape
    L = _lib.lib()
    dev = x.hi.device
    with torch.cuda.device(dev):
        acc = torch.empty((B, D, H, W, CH), dtype=torch.float32, device=dev)
        y_hi = torch.empty_like(x.hi) if store_out else None
        y_lo = torch.empty_like(x.hi) if store_out else None
        head = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if head_w is not None else None
        tiles = B * -(-H // 16) * -(-W // 8)
        ppi = D
        while ppi > 8 and tiles * (D // ppi) < 4 * 148:
            ppi //= 2
        ppi = max(2, ppi)
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        passes = ((x.hi, w_hi, None, acc), (x.lo, w_hi, acc, acc), (x.hi, w_lo, acc, None))
        for xin, win, a_in, a_out in passes:
            last = a_out is None
            rc = L.pmnet_conv3d_k3_c96_pass(
                xin.data_ptr(), win.data_ptr(), scale.data_ptr(), bias.data_ptr(),
                y_hi.data_ptr() if (last and store_out) else None, y_lo.data_ptr() if (last and store_out) else None,
                a_in.data_ptr() if a_in is not None else None, a_out.data_ptr() if a_out is not None else None,
                head_w.data_ptr() if (last and head_w is not None) else None, C.c_float(float(head_b)),
                head.data_ptr() if (last and head is not None) else None,
                B, D, H, W, int(relu), int(ppi), 0, st,
            )  # fmt: skip
            _lib.check(rc, "pmnet_conv3d_k3_c96_pass")
    return (Act(y_hi, y_lo) if store_out else None), head
