"""Developer tool (GPU box): randomized parity sweep of the scoring kernel against the CPU oracle on fresh inputs -
several models, conformer counts and seeds, tens of thousands of ligands each. Scores within 1e-5 relative and tree
shapes (node / leaf counts) identical, or the script exits non-zero."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as orc  # noqa: E402

from pharmaconet_b200 import scoring, synthetic  # noqa: E402
from pharmaconet_b200.packing import LigandBatch, PackedModel  # noqa: E402
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel  # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
CASES = [  # model, conformers, ligands, seed, extra make_ligands kwargs
    ("syn0", 32, 16000, 1001, {}),
    ("syn0", 8, 16000, 1002, {}),
    ("syn0", 13, 8000, 1003, {}),
    ("syn0", 48, 4000, 1004, {}),
    ("syn0", 100, 2000, 1005, {}),
    ("syn0", 5, 2000, 1006, dict(frag_range=(9, 15))),
    ("sparse", 16, 16000, 1007, {}),
    ("xbond", 4, 8000, 1008, {}),
    ("loose", 8, 1500, 1009, {}),
    ("hot70", 8, 2000, 1010, {}),
]
models = {n: PackedModel.from_model(PharmacophoreModel.load(os.path.join(G, f"model_{n}.pm"))) for n in ("syn0", "sparse", "xbond", "loose")}
models["hot70"] = PackedModel.from_model(
    PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(seed=21, n_hotspots=70))
)
# optional argument: a node budget for the task-parallel walk (PmScoreConfig.heavy_budget; default: the library's)
BUDGET = int(sys.argv[1]) if len(sys.argv) > 1 else 0
print(f"heavy_budget = {BUDGET} (0 = default)", flush=True)
bad = 0
for name, nconf, n, seed, kw in CASES:
    t0 = time.time()
    batch = LigandBatch.from_typed(synthetic.make_ligands(n, nconf, seed=seed, **kw))
    dm = scoring.DeviceModel(models[name], "cuda:0")
    out = scoring.score_library(dm, batch, config=scoring.ScoreConfig(heavy_budget=BUDGET), with_stats=True)
    ref = orc.score(models[name], batch)
    rel = np.abs(out["scores"] - ref["scores"]) / np.maximum(np.abs(ref["scores"]), 1e-12)
    same_status = np.array_equal(out["status"], ref["status"])
    same_tree = np.array_equal(out["stats"][:, :2], ref["stats"][:, :2].astype(np.uint32))
    ok = same_status and same_tree and rel.max() <= 1e-5
    bad += not ok
    print(f"{name:7s} C={nconf:3d} n={n:6d} nodes={models[name].num_nodes:3d}: max rel err {rel.max():.2e}, zero scores "
          f"{int((ref['scores'] == 0).sum()):6d}, mean tree nodes {ref['stats'][:, 0].mean():9.0f}, status equal {same_status}, "
          f"tree shapes equal {same_tree} -> {'OK' if ok else 'FAIL'} ({time.time() - t0:.0f} s)", flush=True)
sys.exit(1 if bad else 0)
