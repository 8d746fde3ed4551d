"""Developer probe: TFLOP/s of the tcgen05 3x3x3 conv vs torch/cuDNN on the same GPU (64^3 grid, 96 -> 96)."""
import argparse
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import conv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=64)
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--planes", type=int, default=0)
ap.add_argument("--skip-torch", action="store_true")
a = ap.parse_args()
B, S = a.batch, a.size
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn((B, 96, S, S, S), generator=g, device="cuda").bfloat16()
w = (torch.randn((96, 96, 3, 3, 3), generator=g, device="cuda") * 0.03).bfloat16()
scale = torch.rand(96, generator=g, device="cuda") + 0.5
bias = torch.randn(96, generator=g, device="cuda") * 0.2
flop = 2.0 * B * S**3 * 96 * 96 * 27


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


xc, wp = conv.to_c8(x), conv.pack_weights_k3(w)
ms = timeit(lambda: conv.conv3d_k3_c96(xc, wp, scale, bias, True, planes_per_item=a.planes), a.iters)
print(f"pmnet tcgen05 conv: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s  (B={B}, {S}^3)")
hw = torch.randn(96, device="cuda")
ms = timeit(lambda: conv.conv3d_k3_c96(xc, wp, scale, bias, True, head_w=hw, store_out=False, planes_per_item=a.planes), a.iters)
print(f"pmnet tcgen05 conv + fused head, no store: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s")
if not a.skip_torch:
    xl = x.contiguous(memory_format=torch.channels_last_3d)
    wl = w.contiguous(memory_format=torch.channels_last_3d)
    torch.backends.cudnn.benchmark = True
    ms = timeit(lambda: torch.relu(F.conv3d(xl, wl, padding=1) * scale.view(1, -1, 1, 1, 1).bfloat16() + bias.view(1, -1, 1, 1, 1).bfloat16()), a.iters)
    print(f"torch cuDNN bf16 channels_last_3d conv + scale/bias/relu: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s")
    ms = timeit(lambda: F.conv3d(xl, wl, padding=1), a.iters)
    print(f"torch cuDNN bf16 channels_last_3d conv only: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s")
    ms = timeit(lambda: F.conv3d(x, w, padding=1), a.iters)
    print(f"torch cuDNN bf16 NCDHW conv only: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s")
    xf, wf = x.float(), w.float()
    torch.backends.cudnn.allow_tf32 = True
    ms = timeit(lambda: F.conv3d(xf, wf, padding=1), max(2, a.iters // 3))
    print(f"torch cuDNN fp32 (TF32 allowed) NCDHW conv only: {ms:.3f} ms  {flop/ms/1e9:.1f} TFLOP/s")
