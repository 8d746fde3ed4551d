"""The tcgen05 linear layer (csrc/gemm.cu through the C-ABI) against torch fp32 on the layer shapes of the network."""

import pytest
import torch

from pharmaconet_b200 import gemm

pytestmark = pytest.mark.gpu

SHAPES = [
    (4096, 288, 96),     # stage-0 qkv
    (4096, 96, 96),      # stage-0 proj
    (1000, 384, 96),     # fc1, M not a multiple of the tile
    (4096, 96, 384),     # fc2
    (512, 192, 768),     # PatchMerging reduction
    (333, 96, 264),      # PatchEmbed as a GEMM: K = 33 * 8 is not a multiple of the 64-wide K block
    (64, 3072, 768),     # stage-3 fc1
    (64, 768, 3072),     # stage-3 fc2
    (128, 96, 20736),    # 4^3 FPN convolution in im2col form
    (48, 96, 192),       # mask-head MLP, a handful of rows
    (200, 192, 192),     # token head
]


def _ref(a, w, bias, act):
    y = a.double() @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if act == gemm.ACT_GELU:
        y = torch.nn.functional.gelu(y)
    elif act == gemm.ACT_RELU:
        y = torch.relu(y)
    elif act == gemm.ACT_SILU:
        y = torch.nn.functional.silu(y)
    return y


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_bf16_operands_exact_products(M, N, K):
    """bf16-representable inputs: one pass is exact up to fp32 accumulation."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn((M, K), generator=g, device="cuda").bfloat16().float()
    w = (torch.randn((N, K), generator=g, device="cuda") / K**0.5).bfloat16().float()
    bias = torch.randn(N, generator=g, device="cuda")
    y, op = gemm.linear(gemm.Operand.from_float(a, False), gemm.Operand.from_float(w, False), bias, want_operand=True)
    ref = _ref(a, w, bias, 0)
    scale = float(ref.abs().max())
    assert float((y.double() - ref).abs().max()) <= 2e-5 * scale
    assert op.lo is None and float((op.hi.double() - ref).abs().max()) <= 2.0**-8 * scale


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("act", [gemm.ACT_NONE, gemm.ACT_GELU])
def test_gemm_split_precision(M, N, K, act):
    """fp32 inputs through the two-term split: 2^-16-level agreement with fp64; the emitted (hi, lo) pair carries it."""
    g = torch.Generator(device="cuda").manual_seed(M * 3 + N + K)
    a = torch.randn((M, K), generator=g, device="cuda")
    w = torch.randn((N, K), generator=g, device="cuda") / K**0.5
    bias = torch.randn(N, generator=g, device="cuda")
    y, op = gemm.linear(gemm.Operand.from_float(a, True), gemm.Operand.from_float(w, True), bias, act, want_operand=True)
    ref = _ref(a, w, bias, act)
    scale = float(ref.abs().max())
    # 2^-16 per product; the tensor core's fp32 accumulation adds ~2^-22 per K block (visible at K = 20736)
    tol = 6e-5 * max(1.0, (K / 4096) ** 0.5)
    assert float((y.double() - ref).abs().max()) <= tol * scale
    assert float((op.float().double() - ref).abs().max()) <= (tol + 2e-5) * scale
    # against one bf16 pass the split is ~100x closer
    y1, _ = gemm.linear(gemm.Operand.from_float(a, False), gemm.Operand.from_float(w, False), bias, act)
    assert float((y1.double() - ref).abs().max()) > 10 * float((y.double() - ref).abs().max())


def test_gemm_rejects_bad_arguments():
    a = gemm.Operand(torch.zeros((16, 20), dtype=torch.bfloat16, device="cuda"))
    w = gemm.Operand(torch.zeros((96, 20), dtype=torch.bfloat16, device="cuda"))
    with pytest.raises(RuntimeError):
        gemm.linear(a, w)  # K not a multiple of 8
