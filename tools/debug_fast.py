"""Developer probe: specialised vs general scoring kernel vs the CPU oracle on the golden cases (per-ligand diff)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle as orc  # noqa: E402
from golden_util import CASES, load_case  # noqa: E402

from pharmaconet_b200 import scoring  # noqa: E402

names = sys.argv[1:] or CASES
for name in names:
    c = load_case(name)
    dm = scoring.DeviceModel(c["model"], "cuda:0")
    db = scoring.DeviceLigandBatch.from_host(c["batch"], "cuda:0")
    from golden_util import weights_dict

    w = weights_dict(c["weights"])
    f = scoring.score_batch(dm, db, w, with_stats=True)
    g = scoring.score_batch(dm, db, w, scoring.ScoreConfig(32, 148, 8192), with_stats=True)
    o = orc.score(c["model"], c["batch"], c["weights"])
    fs, gs = f["scores"].cpu().numpy(), g["scores"].cpu().numpy()
    fst, gst = f["stats"].cpu().numpy().view(np.uint32), g["stats"].cpu().numpy().view(np.uint32)
    ref = c["ref"]
    relf = np.abs(fs - ref) / np.maximum(np.abs(ref), 1e-12)
    relg = np.abs(gs - ref) / np.maximum(np.abs(ref), 1e-12)
    nm = c["model"].num_nodes, c["model"].num_clusters
    print(f"== {name}: model {nm}, {len(ref)} ligands; default max rel {relf.max():.2e}; general-only max rel {relg.max():.2e}; "
          f"status default {np.bincount(f['status'].cpu().numpy(), minlength=5)} general {np.bincount(g['status'].cpu().numpy(), minlength=5)}")
    bad = np.nonzero((relf > 1e-5) | (relg > 1e-5))[0]
    for i in bad[:12]:
        print(f"  lig {i}: ref {ref[i]:.5f} default {fs[i]:.5f} general {gs[i]:.5f} | nodes/leaves/rows/pairs default {fst[i]} general {gst[i]} "
              f"oracle {o['stats'][i]}")
