"""Developer probe for ncu: one launch of each CNN kernel at its benchmark shape (BASELINE configs[3]: 8 pockets of 64^3)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn, cnn_weights, conv, gemm  # noqa: E402

dev = "cuda:0"
g = torch.Generator(device=dev).manual_seed(0)
# (1) the 96 -> 96 3x3x3 convolution at 8 x 64^3
x = conv.to_c8(torch.randn((8, 96, 64, 64, 64), generator=g, device=dev))
w = conv.pack_weights_k3(torch.randn((96, 96, 3, 3, 3), generator=g, device=dev) * 0.03)
scale, bias = torch.rand(96, generator=g, device=dev) + 0.5, torch.randn(96, generator=g, device=dev) * 0.2
conv.conv3d_k3_c96(x, w, scale, bias, True)
# (2) the stage-0 fc1 GEMM (M = 8 * 32^3, N = 384, K = 96, GELU, bf16 split output): one pass and three passes
a = torch.randn((8 * 32768, 96), generator=g, device=dev)
wt = torch.randn((384, 96), generator=g, device=dev) * 0.1
b1 = torch.randn(384, generator=g, device=dev)
for split in (False, True):
    gemm.linear(gemm.Operand.from_float(a, split), gemm.Operand.from_float(wt, split), b1, gemm.ACT_GELU, want_f32=False, want_operand=True)
# (3) the network
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
model = cnn.PharmacoNetModel(cnn_weights.synth_state_dict(man, buf, 0), dev)
gc = torch.Generator().manual_seed(0)
images = torch.rand((8, 33, 64, 64, 64), generator=gc).to(dev)
tokens = torch.cat([torch.randint(0, 64, (200, 3), generator=gc), torch.randint(0, 10, (200, 1), generator=gc)], 1).long().to(dev)
for prec in (sys.argv[1:] or ["bf16"]):
    model.precision = prec
    feats = model.forward_feature(images, nchw=False)
    narrow, wide = model.forward_cavity_extraction(feats[-1])
    scores, tfeat = model.forward_token_prediction(feats[-1], [tokens] * 8)
    one = cnn.Features(f[:1] for f in feats)
    logits = model.forward_segmentation(one, [tokens[:32]], [tfeat[0][:32]], group_size=4)[0][0]
    maps = cnn.density_post(logits, tokens[:32], torch.ones((64, 64, 64), dtype=torch.bool), narrow[0, 0] > 0, 0.5)
torch.cuda.synchronize()
print("probe done", float(maps.sum()))
