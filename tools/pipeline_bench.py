"""Developer probe / BASELINE configs[4] in miniature: P synthetic pockets -> pharmacophore models (batched CNN forward,
mask head, density maps, host graph construction) -> every model screened against ONE device-resident synthetic
library (`screening.screen_models`). Network weights are synthetic (no trained weights offline), so the models are
only structurally realistic; what is measured is the throughput of each stage."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import cnn_weights, screening, synthetic  # noqa: E402
from pharmaconet_b200.module import PharmacoNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--pockets", type=int, default=8)
ap.add_argument("--ligands", type=int, default=262144)
ap.add_argument("--conformers", type=int, default=32)
ap.add_argument("--chunk", type=int, default=8)
a = ap.parse_args()
dev = torch.device("cuda:0")
G = os.path.join(ROOT, "tests", "golden")
man = json.load(open(os.path.join(G, "cnn_manifest.json")))
buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
net = PharmacoNet("cuda:0", verbose=False, checkpoint=cnn_weights.synth_checkpoint(man, buf, 0))
gold = np.load(os.path.join(G, "cnn_pipeline_golden.npz"))
tokens = torch.from_numpy(gold["tokens"]).long()
token_pos = (tokens[:, :3].float() - 31.5) * 0.5
data = []
for p in range(a.pockets):
    image = torch.rand((33, 64, 64, 64), generator=torch.Generator().manual_seed(p))
    mask = torch.rand((64, 64, 64), generator=torch.Generator().manual_seed(1000 + p)) < 0.8
    data.append((image, mask, token_pos, tokens))


def sync():
    torch.cuda.synchronize(dev)


net.create_models(data[:1])  # warm-up (allocations, cuBLAS handles)
sync()
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel  # noqa: E402

t0 = time.perf_counter()
rs = []
for lo in range(0, a.pockets, a.chunk):
    rs += net._features_and_hotspots_batch(data[lo : lo + a.chunk], nchw=False)
sync()
t1 = time.perf_counter()
maps = [net._density_maps_of(r, sparse=True) for r in rs]
sync()
t2 = time.perf_counter()
models = [PharmacophoreModel.create("", (0.0, 0.0, 0.0), m) for m in maps]
t3 = time.perf_counter()
nh = [len(m) for m in maps]
print(f"{a.pockets} pockets, per pocket: features + token / cavity heads + hotspot filter {1e3 * (t1 - t0) / a.pockets:.1f} ms "
      f"(incl. H2D of the 33 x 64^3 grids), mask head + density post + sparse D2H {1e3 * (t2 - t1) / a.pockets:.1f} ms "
      f"({np.mean(nh):.0f} hotspots), model graph on the host {1e3 * (t3 - t2) / a.pockets:.1f} ms "
      f"({np.mean([len(m.nodes) for m in models]):.0f} nodes / {np.mean([len(m.node_clusters) for m in models]):.0f} clusters)")
sync()
t0 = time.perf_counter()
net.create_models(data, chunk=a.chunk)
sync()
print(f"create_models end to end: {1e3 * (time.perf_counter() - t0) / a.pockets:.1f} ms/pocket")
models = [m for m in models if len(m.nodes) > 0]
lib = synthetic.make_library_device(a.ligands, a.conformers, 1, dev, 4096)
screening.screen_models(models[:1], lib, k=100)  # warm-up
sync()
t0 = time.perf_counter()
res = screening.screen_models(models, lib, k=100)
sync()
dt = time.perf_counter() - t0
pairs = len(models) * lib.n_conformers_total
print(f"screen_models: {len(models)} models x {lib.n_ligands} ligands x {a.conformers} conformers in {dt:.3f} s = "
      f"{pairs / dt / 1e6:.1f} M model-conformer pairs/s; overflow re-runs {sum(r.n_overflow for r in res)}")
for i, r in enumerate(res[:3]):
    print(f"  model {i}: best {float(r.topk_scores[0]):.2f} (ligand {int(r.topk_ids[0])})")
