"""Host side of the tcgen05 linear layer (csrc/gemm.cu): operand containers and the C-ABI call.

Replaces `nn.Linear` / `F.linear` of the reference network (swinv2.py:114-158, 346-363, 484-500; swin.py:19-44;
token_head.py:50-86; mask_head.py:128-168). torch provides device memory and streams only.
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib

ACT_NONE, ACT_GELU, ACT_RELU, ACT_SILU = 0, 1, 2, 3


class Operand:
    """A GEMM operand [rows, K], K contiguous: `hi` bf16 and, in the split-precision mode, the bf16 residual `lo`
    (value = hi + lo to 2^-17 relative)."""

    __slots__ = ("hi", "lo")

    def __init__(self, hi: torch.Tensor, lo: torch.Tensor | None = None):
        assert hi.dtype == torch.bfloat16 and hi.is_contiguous() and (lo is None or (lo.dtype == torch.bfloat16 and lo.is_contiguous()))
        self.hi, self.lo = hi, lo

    @classmethod
    def from_float(cls, x: torch.Tensor, split: bool) -> "Operand":
        x = x.float().contiguous()
        hi = x.to(torch.bfloat16)
        return cls(hi, (x - hi.float()).to(torch.bfloat16) if split else None)

    @property
    def shape(self):
        return self.hi.shape

    def view(self, *shape) -> "Operand":
        return Operand(self.hi.view(*shape), None if self.lo is None else self.lo.view(*shape))

    def float(self) -> torch.Tensor:
        return self.hi.float() if self.lo is None else self.hi.float() + self.lo.float()


def linear(
    a: Operand,
    w: Operand,
    bias: torch.Tensor | None = None,
    act: int = ACT_NONE,
    want_f32: bool = True,
    want_operand: bool = False,
):
    """y = act(a . w^T + bias). Returns (fp32 [M, N] or None, Operand or None); the Operand carries a low part iff the
    inputs do (split precision)."""
    if not a.hi.is_cuda:
        raise RuntimeError("pharmaconet_b200.gemm runs on CUDA devices only (no CPU fallback)")
    K = a.shape[-1]
    M = a.hi.numel() // K
    N = w.shape[0]
    assert w.shape[1] == K
    split = a.lo is not None
    assert split == (w.lo is not None), "both operands must be split, or neither"
    dev = a.hi.device
    lead = tuple(a.shape[:-1])
    with torch.cuda.device(dev):
        out32 = torch.empty(lead + (N,), dtype=torch.float32, device=dev) if want_f32 else None
        o_hi = torch.empty(lead + (N,), dtype=torch.bfloat16, device=dev) if want_operand else None
        o_lo = torch.empty(lead + (N,), dtype=torch.bfloat16, device=dev) if (want_operand and split) else None
        if bias is not None:
            assert bias.dtype == torch.float32 and bias.numel() == N
        rc = _lib.lib().pmnet_gemm_bf16(
            a.hi.data_ptr(), a.lo.data_ptr() if split else None, w.hi.data_ptr(), w.lo.data_ptr() if split else None,
            bias.data_ptr() if bias is not None else None, out32.data_ptr() if want_f32 else None,
            o_hi.data_ptr() if want_operand else None, o_lo.data_ptr() if o_lo is not None else None,
            M, N, K, int(act), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream),
        )  # fmt: skip
        _lib.check(rc, "pmnet_gemm_bf16")
    return out32, (Operand(o_hi, o_lo) if want_operand else None)
