"""Developer probe: the dense synthetic model of bench.py's `dense_model` leg (hotspot-rich pocket) under different
heavy-ligand budgets (PmScoreConfig.heavy_budget): launch time, ligands split, node statistics, result identity."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import scoring, synthetic  # noqa: E402
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ligands", type=int, default=32768)
ap.add_argument("--hotspots", type=int, default=60, help="0: the headline model (tests/golden/model_syn0.pm)")
ap.add_argument("--seed", type=int, default=0, help="library seed (bench.py uses 1)")
ap.add_argument("--iters", type=int, default=1)
ap.add_argument("--budgets", type=str, default="-1,0,65536,16384")
ap.add_argument("--general", action="store_true", help="general kernel alone (explicit launch shape)")
ap.add_argument("--timing", action="store_true", help="library built with -DPM_TIMING: stats[2] / [3] hold microseconds")
ap.add_argument("--profile", action="store_true", help="per-kernel times (torch profiler) and task sizes of the split")
a = ap.parse_args()

dev = torch.device("cuda:0")
lib = synthetic.make_library_device(a.ligands, 32, a.seed, dev, 4096)
if a.hotspots > 0:
    model = PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(seed=21, n_hotspots=a.hotspots))
else:  # the headline model of bench.py
    model = PharmacophoreModel.load(os.path.join(ROOT, "tests", "golden", "model_syn0.pm"))
dm = scoring.DeviceModel(model.packed, dev)
print(f"model: {len(model.nodes)} nodes / {len(model.node_clusters)} clusters; {lib.n_ligands} ligands", flush=True)
lib.set_order(scoring.cost_order(dm, lib))
ref = None
for b in [int(x) for x in a.budgets.split(",")]:
    cfg = scoring.ScoreConfig(16, 148, 8192, heavy_budget=b) if a.general else scoring.ScoreConfig(heavy_budget=b)
    ws = torch.zeros(scoring.workspace_bytes(dm, cfg, 32), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ms = 1e30
    for _ in range(a.iters):
        e0.record()
        out = scoring.score_batch(dm, lib, None, cfg, with_stats=True, workspace=ws)
        e1.record()
        torch.cuda.synchronize()
        ms = min(ms, e0.elapsed_time(e1))
    st = out["status"].cpu().numpy()
    stats = out["stats"].cpu().numpy().view(np.uint32).astype(np.float64)
    nh = int(ws[:8].view(torch.int32)[1].item())
    sc = out["scores"].cpu().numpy()
    same = "-" if ref is None else str(bool(np.array_equal(sc, ref[0]) and np.array_equal(stats, ref[1])))
    if ref is None:
        ref = (sc, stats)
        nodes = stats[:, 0]
        q = np.percentile(nodes, [50, 90, 99, 99.9, 100])
        print(f"tree nodes: mean {nodes.mean():.0f} p50 {q[0]:.0f} p90 {q[1]:.0f} p99 {q[2]:.0f} p99.9 {q[3]:.0f} max {q[4]:.0f}; "
              f"sum {nodes.sum():.3e}; > 2^18: {(nodes > 2**18).sum()} ligands holding {nodes[nodes > 2**18].sum() / nodes.sum():.2%}")
    hdr0 = ws[:256].view(torch.int32).cpu().numpy()
    print(f"budget {b:>7}: {ms:9.1f} ms  heavy {nh:4d}  deferred {hdr0[5]}  status {np.bincount(st, minlength=6).tolist()}  "
          f"{stats[:, 0].sum() / ms / 1e3:8.1f} M nodes/s  identical {same}", flush=True)
    if a.timing:
        t01, t2, nodes = stats[:, 2], stats[:, 3], stats[:, 0]
        print(f"  per-ligand us: phases 0-1 mean {t01.mean():.0f} max {t01.max():.0f}; DFS mean {t2.mean():.0f} max {t2.max():.0f}; "
              f"sum {t01.sum() / 1e6:.2f} s + {t2.sum() / 1e6:.2f} s of warp time; DFS nodes/us overall {nodes.sum() / max(t2.sum(), 1):.2f}")
        for i in np.argsort(-(t01 + t2))[:8]:
            print(f"    ligand {i}: phases 0-1 {t01[i]:.0f} us, DFS {t2[i]:.0f} us, nodes {nodes[i]:.0f}, leaves {stats[i, 1]:.0f}")
    if a.profile and nh > 0:
        hdr = ws[:256].view(torch.int32).cpu().numpy()
        print(f"  tasks donated {hdr[16]}, taken {hdr[24]}, replay failures {hdr[9]}, deferred {hdr[5]}")
        from torch.profiler import ProfilerActivity, profile

        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            scoring.score_batch(dm, lib, None, cfg, with_stats=True, workspace=ws)
            torch.cuda.synchronize()
        for ev in prof.events():
            if ev.device_time > 0:
                print(f"    {ev.name[:60]:60s} {ev.device_time / 1e3:9.2f} ms")
