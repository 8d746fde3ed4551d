"""Developer probe (not the bench contract): time the scoring kernel on a replicated synthetic library."""
import argparse
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import scoring, synthetic  # noqa: E402
from pharmaconet_b200.packing import LigandBatch, PackedModel  # noqa: E402
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--unique", type=int, default=2048)
ap.add_argument("--rep", type=int, default=32)
ap.add_argument("--conf", type=int, default=32)
ap.add_argument("--iters", type=int, default=5)
ap.add_argument("--warps", type=int, default=0)
ap.add_argument("--blocks", type=int, default=0)
ap.add_argument("--rows", type=int, default=0)
ap.add_argument("--tiny", action="store_true")
ap.add_argument("--lpt", action="store_true", help="hand out ligands longest first (scoring.cost_order), like the bench")
ap.add_argument("--hotspots", type=int, default=0, help="build a synthetic model with this many hotspots instead of syn0")
a = ap.parse_args()

if a.hotspots:
    model = PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(seed=21, n_hotspots=a.hotspots))
    print(f"synthetic model: {len(model.nodes)} nodes, {len(model.node_clusters)} clusters", flush=True)
else:
    model = PharmacophoreModel.load(os.path.join(ROOT, "tests", "golden", "model_syn0.pm"))
dm = scoring.DeviceModel(PackedModel.from_model(model), "cuda:0")
t = time.time()
base = LigandBatch.from_typed(synthetic.make_ligands(a.unique, a.conf, seed=1))
print(f"generated {a.unique} ligands in {time.time()-t:.1f}s", flush=True)
idx = np.tile(np.arange(a.unique), a.rep)
big = base.select(idx) if a.rep > 1 else base
db = scoring.DeviceLigandBatch.from_host(big, "cuda:0")
if a.lpt:
    db.set_order(scoring.cost_order(dm, db))
cfg = scoring.ScoreConfig(a.warps, a.blocks, a.rows)
n = big.num_ligands
print(f"library: {n} ligands, {big.num_conformers_total} conformers, {db.nbytes()/1e6:.1f} MB", flush=True)
out = scoring.score_batch(dm, db, config=cfg, with_stats=True)
torch.cuda.synchronize()
st = out["status"].cpu().numpy()
stats = out["stats"].cpu().numpy().view(np.uint32)
print("status counts", np.bincount(st, minlength=4), "tree nodes mean", stats[:, 0].mean(), "rows mean/max", stats[:, 2].mean(), stats[:, 2].max(),
      "pairs mean/max", stats[:, 3].mean(), stats[:, 3].max())
if a.tiny:
    sys.exit(0)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(a.iters + 1)]
ev[0].record()
for i in range(a.iters):
    scoring.score_batch(dm, db, config=cfg, out_scores=out["scores"], out_status=out["status"])
    ev[i + 1].record()
torch.cuda.synchronize()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(a.iters)]
best = min(ms)
print(f"ms per launch: {['%.2f' % m for m in ms]}")
print(f"best: {n/best*1e3:.0f} ligands/s, {big.num_conformers_total/best*1e3/1e6:.2f} M conformers/s, "
      f"algorithmic {big.algorithmic_bytes()/best*1e3/1e9:.2f} GB/s")
