"""Developer probe: kernels of one forward_feature + cavity + token pass (8 pockets) by total GPU time (torch profiler)."""
import json
import os
import sys
from collections import defaultdict

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn, cnn_weights  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
model = cnn.PharmacoNetModel(cnn_weights.synth_state_dict(man, buf, 0), "cuda:0")
model.precision = prec
g = torch.Generator().manual_seed(0)
image = torch.rand((8, 33, 64, 64, 64), generator=g).cuda()
for _ in range(3):
    f = model.forward_feature(image, nchw=False)
    model.forward_cavity_extraction(f[-1] if hasattr(f, "__getitem__") else f)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    f = model.forward_feature(image, nchw=False)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0.0, 0])
for e in prof.events():
    if e.device_time > 0:
        tot[e.name[:90]][0] += e.device_time
        tot[e.name[:90]][1] += 1
all_us = sum(v[0] for v in tot.values())
print(f"forward_feature, 8 pockets, {prec}: {all_us / 1e3:.2f} ms of kernels")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:28]:
    print(f"  {v[0] / 1e3:7.3f} ms {v[1]:4d} x  {100 * v[0] / all_us:5.1f}%  {k}")
