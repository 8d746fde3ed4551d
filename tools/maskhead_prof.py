"""Developer probe: kernel-level time breakdown of the batched mask head (torch profiler)."""
import json, os, sys
import numpy as np, torch
from torch.profiler import ProfilerActivity, profile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn, cnn_weights
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
model = cnn.PharmacoNetModel(cnn_weights.synth_state_dict(man, buf, 0), "cuda:0")
model.backbone.precision = "bf16"
g = torch.Generator().manual_seed(0)
x = torch.rand((1, 33, 64, 64, 64), generator=g).cuda()
tokens = torch.cat([torch.randint(0, 64, (48, 3), generator=g), torch.randint(0, 10, (48, 1), generator=g)], 1).long().cuda()
feats = model.forward_feature(x, nchw=False)
_, tf = model.forward_token_prediction(feats[-1], [tokens])
fn = lambda: model.forward_segmentation(feats, [tokens], [tf[0]], group_size=4)
for _ in range(2): fn()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    fn(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=16, max_name_column_width=60))
