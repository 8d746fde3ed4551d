"""tcgen05 3x3x3 conv kernel against torch.nn.functional.conv3d in fp32 (the floating-point reference of this op).

Inputs and weights are rounded to bf16 first (the kernel's operand type), the reference accumulates in fp32 with
TF32 disabled; the kernel's output is bf16, so the bound is one bf16 rounding (2^-8 relative) plus fp32
accumulation-order noise."""

import pytest
import torch
import torch.nn.functional as F

from pharmaconet_b200 import conv

pytestmark = pytest.mark.gpu


def _ref(x, w, scale, bias, relu):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    y = F.conv3d(x, w, padding=1)
    y = y * scale.view(1, -1, 1, 1, 1) + bias.view(1, -1, 1, 1, 1)
    return torch.relu(y) if relu else y


def _case(B, D, H, W, relu=True, seed=0, **kw):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn((B, 96, D, H, W), generator=g, device="cuda").bfloat16().float()
    w = (torch.randn((96, 96, 3, 3, 3), generator=g, device="cuda") * 0.03).bfloat16().float()
    scale = torch.rand(96, generator=g, device="cuda") + 0.5
    bias = torch.randn(96, generator=g, device="cuda") * 0.2
    y_c8, _ = conv.conv3d_k3_c96(conv.to_c8(x), conv.pack_weights_k3(w), scale, bias, relu=relu, **kw)
    torch.cuda.synchronize()
    y = conv.from_c8(y_c8)
    ref = _ref(x, w, scale, bias, relu)
    err = (y - ref).abs()
    tol = 2.0**-8 * ref.abs() + 1e-3
    assert bool((err <= tol).all()), f"max err {err.max().item():.4g} at |ref| max {ref.abs().max().item():.3g}"
    return y, ref


def test_layout_roundtrip():
    x = torch.randn(2, 96, 4, 6, 8, device="cuda").bfloat16().float()
    assert torch.equal(conv.from_c8(conv.to_c8(x)), x)


@pytest.mark.parametrize(
    "shape",
    [(1, 2, 16, 8), (2, 8, 32, 16), (1, 4, 20, 12), (1, 8, 8, 8), (3, 4, 4, 4), (1, 16, 16, 16), (1, 6, 40, 24)],
)
def test_conv_matches_torch(shape):
    _case(*shape)


def test_conv_no_relu_and_item_splits():
    _case(1, 16, 32, 32, relu=False, planes_per_item=4)
    _case(1, 16, 32, 32, relu=True, planes_per_item=16, max_ctas=3)  # several items per CTA: ring wrap-around
    _case(2, 12, 16, 16, relu=True, planes_per_item=8)  # last item of a column is shorter (12 = 8 + 4)


def test_conv_64cube():
    _case(1, 64, 64, 64)


def test_fused_head_matches_1x1_conv():
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((1, 96, 8, 32, 32), generator=g, device="cuda").bfloat16().float()
    w = (torch.randn((96, 96, 3, 3, 3), generator=g, device="cuda") * 0.03).bfloat16().float()
    scale = torch.rand(96, generator=g, device="cuda") + 0.5
    bias = torch.randn(96, generator=g, device="cuda") * 0.2
    hw = torch.randn(96, generator=g, device="cuda") * 0.1
    y_c8, head = conv.conv3d_k3_c96(conv.to_c8(x), conv.pack_weights_k3(w), scale, bias, True, head_w=hw, head_b=0.25)
    _, head_only = conv.conv3d_k3_c96(
        conv.to_c8(x), conv.pack_weights_k3(w), scale, bias, True, head_w=hw, head_b=0.25, store_out=False
    )
    ref = _ref(x, w, scale, bias, True)
    href = (ref * hw.view(1, -1, 1, 1, 1)).sum(1) + 0.25  # fp32 activations, like the kernel's fused epilogue
    assert torch.equal(head, head_only)
    assert (head - href).abs().max().item() <= 1e-3 * max(1.0, href.abs().max().item())
    assert y_c8 is not None


@pytest.mark.parametrize("cin,res", [(33, 32), (96, 32), (96, 16), (192, 8)])
def test_fpn_lateral_matches_torch(cin, res):
    """pmnet_lateral_c96 on an fp32 NCDHW input: relu(bn(conv1x1(x))) + nearest-upsampled coarser level
    (fpn_decoder.py:100-111). Levels of >= 8192 voxels with C_in <= 96 run on the tensor cores (bf16 operands, fp32
    accumulation: error ~ 2^-9 of |x| . |w| summed over C_in), the others on the fp32 CUDA-core kernel."""
    import ctypes as C

    from pharmaconet_b200 import _lib

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(cin + res)
    B = 2
    x = torch.randn((B, cin, res, res, res), generator=g, device="cuda")
    w = torch.randn((96, cin), generator=g, device="cuda") * cin**-0.5
    scale = torch.rand(96, generator=g, device="cuda") + 0.5
    bias = torch.randn(96, generator=g, device="cuda") * 0.1
    up = torch.randn((B, 96, res // 2, res // 2, res // 2), generator=g, device="cuda").bfloat16()
    up_c8 = conv.to_c8(up.float())
    out = torch.empty((B, 12, res, res, res, 8), dtype=torch.bfloat16, device="cuda")
    w_t = w.t().contiguous()
    rc = L.pmnet_lateral_c96(
        x.data_ptr(), 0, cin, w_t.data_ptr(), scale.data_ptr(), bias.data_ptr(), 1, up_c8.data_ptr(), out.data_ptr(),
        B, res, res, res, C.c_void_p(torch.cuda.current_stream().cuda_stream),
    )  # fmt: skip
    _lib.check(rc, "pmnet_lateral_c96")
    torch.cuda.synchronize()
    tensor_cores = res**3 >= 8192 and cin <= 96
    xr, wr = (x.bfloat16().double(), w.bfloat16().double()) if tensor_cores else (x.double(), w.double())
    ref = torch.relu(torch.einsum("bcdhw,oc->bodhw", xr, wr) * scale.double().view(1, -1, 1, 1, 1) + bias.double().view(1, -1, 1, 1, 1))
    ref = ref + torch.nn.functional.interpolate(up.double(), scale_factor=2, mode="nearest")
    got = conv.from_c8(out).double()
    # the output itself is bf16: half an ulp of the largest value, plus fp32 accumulation noise
    assert float((got - ref).abs().max()) <= 2.0**-8 * float(ref.abs().max()) + 1e-3


@pytest.mark.parametrize("res", [16, 32])
def test_lateral_on_c8_input_matches_torch(res):
    """The mask head's shared laterals: 1x1 conv of a bf16 c8 activation without affine / ReLU / upsample
    (mask_head.py:170-196 folded, see pmnet_box_combine_c96). >= 8192 voxels: mma.sync with fragments read straight
    from the c8 tensor; below: the CUDA-core kernel. Same bound either way: the operands are bf16 already."""
    import ctypes as C

    from pharmaconet_b200 import _lib

    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(res)
    B = 2
    x = torch.randn((B, 96, res, res, res), generator=g, device="cuda").bfloat16()
    w = (torch.randn((96, 96), generator=g, device="cuda") * 96**-0.5).bfloat16().float()
    x_c8 = conv.to_c8(x.float())
    out = torch.empty((B, 12, res, res, res, 8), dtype=torch.bfloat16, device="cuda")
    w_t = w.t().contiguous()
    rc = L.pmnet_lateral_c96(
        x_c8.data_ptr(), 1, 96, w_t.data_ptr(), None, None, 0, None, out.data_ptr(), B, res, res, res,
        C.c_void_p(torch.cuda.current_stream().cuda_stream),
    )  # fmt: skip
    _lib.check(rc, "pmnet_lateral_c96")
    torch.cuda.synchronize()
    ref = torch.einsum("bcdhw,oc->bodhw", x.double(), w.double())
    got = conv.from_c8(out).double()
    assert float((got - ref).abs().max()) <= 2.0**-8 * float(ref.abs().max()) + 1e-3
