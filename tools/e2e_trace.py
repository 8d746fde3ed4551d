"""Developer probe: timeline of one Screener.screen_host step (bench.py's e2e leg) - kernels and copies with their
start offsets, from the torch profiler (CUPTI)."""
import argparse
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pharmaconet_b200 import screening, synthetic  # noqa: E402
from pharmaconet_b200.packing import LigandBatch, PackedModel  # noqa: E402
from pharmaconet_b200.pharmacophore_model import PharmacophoreModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--ligands", type=int, default=1000000)
ap.add_argument("--block-ligands", type=int, default=262144)
ap.add_argument("--slots", type=int, default=3)
ap.add_argument("--no-ramp", action="store_true")
ap.add_argument("--min-us", type=float, default=300.0)
ap.add_argument("--ramp-cuts", type=str, default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
packed = PackedModel.from_model(PharmacophoreModel.load(os.path.join(ROOT, "tests", "golden", "model_syn0.pm")))
lib = synthetic.make_library_device(a.ligands, 32, 1, dev, 4096)
host = screening.pin_library(LigandBatch.from_arrays({k: v.cpu().numpy() for k, v in lib.tensors.items()}))
del lib
torch.cuda.empty_cache()
scr = screening.Screener(packed, dev, k=1000, block_ligands=a.block_ligands, n_slots=a.slots, ramp=not a.no_ramp)
if a.ramp_cuts:
    scr.ramp_cuts = tuple(float(x) for x in a.ramp_cuts.split(","))
for _ in range(2):
    scr.screen_host(host)
torch.cuda.synchronize()
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    scr.screen_host(host)
    ts.append((time.perf_counter() - t0) * 1e3)
print("wall ms per step:", [f"{t:.1f}" for t in ts], flush=True)
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    t0 = time.perf_counter()
    scr.screen_host(host)
    wall = (time.perf_counter() - t0) * 1e3
print(f"profiled step: {wall:.1f} ms wall")
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t_first = evs[0].time_range.start
busy_end = t_first
idle = 0.0
for e in evs:
    st, en = e.time_range.start, e.time_range.end
    if "Memcpy" not in e.name and "Memset" not in e.name:
        if st > busy_end:
            idle += st - busy_end
        busy_end = max(busy_end, en)
    if en - st >= a.min_us:
        print(f"  +{(st - t_first) / 1e3:8.2f} ms  {(en - st) / 1e3:8.2f} ms  {e.name[:70]}")
print(f"span {(evs[-1].time_range.end - t_first) / 1e3:.1f} ms, compute-idle gaps inside it {idle / 1e3:.1f} ms")
