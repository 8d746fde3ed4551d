"""Developer probe: kernel-level time breakdown of the backbone (torch profiler)."""
import json, os, sys
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn, cnn_weights
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
model = cnn.PharmacoNetModel(cnn_weights.synth_state_dict(man, buf, 0), "cuda:0")
model.precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
what = sys.argv[2] if len(sys.argv) > 2 else "backbone"
x = torch.rand((8, 33, 64, 64, 64), device="cuda")
fn = (lambda: model.backbone.forward(x)) if what == "backbone" else (lambda: model.forward_feature(x))
for _ in range(2): fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fn(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
