#!/usr/bin/env python
"""bench.py - the driver's benchmark contract for the screening hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1]): a synthetic "6OIM-like" pharmacophore model (35 nodes / 26 clusters, built by
the reference's own PharmacophoreModel.create, shipped as tests/golden/model_syn0.pm) against a seeded synthetic
library of 1 M typed ligands x 32 conformers PER GPU (weak scaling). One "step" = one full pass of the scoring path
over the rank's library + per-rank top-k + (N > 1) one NCCL all-gather of the top-k and the merge.

  value  : conformers/s, whole job, library already resident in HBM        (CUDA events, max over ranks)
  e2e    : the same through Screener.screen_host with PINNED HOST buffers: every step copies the whole library
           host->device (double-buffered blocks) and the scores/status/top-k device->host inside the timed region
  roofline: HBM - algorithmic bytes of one scoring launch / its CUDA-event duration vs MEASURED_PEAKS.json hbm_gbs
  cpu_baseline: the UNMODIFIED reference (GraphMatcher.run + numba kernels under multiprocessing.Pool(all cores),
           /root/reference/screening.py:46-68; shipped to the box as the git-ignored copy oracle/_ref, kind "reference")
           on a bounded prefix of the same library, with the GPU scores of those ligands checked against it;
  cpu_port: the C restatement (oracle/pmnet_oracle.c) on all cores over a larger prefix - the wide parity check

`--impl reference` times the unmodified reference alone (rank 0 only), one bounded sample of the same library per step.
"""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "ligand_conformers_scored_per_sec"
UNIT = "conformers/s"
SAMPLE_LIGANDS = 4096  # ligands per reference step / cpu_baseline block


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ligands", type=int, default=1_000_000, help="ligands per GPU")
    ap.add_argument("--conformers", type=int, default=32)
    ap.add_argument("--templates", type=int, default=4096)
    ap.add_argument("--topk", type=int, default=1000)
    ap.add_argument("--block-ligands", type=int, default=262144)
    ap.add_argument("--slots", type=int, default=3, help="device staging slots of the streamed (e2e) leg")
    ap.add_argument("--no-lpt", action="store_true", help="process ligands in index order instead of longest first")
    ap.add_argument("--no-ramp", action="store_true", help="streamed leg: do not cut the first block into growing spans")
    ap.add_argument("--stream-warps", type=int, default=0, help="streamed leg: warps per CTA (0 = library default)")
    ap.add_argument("--stream-ctas", type=int, default=0, help="streamed leg: CTAs in the grid (0 = library default)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="budget of the cpu_port leg (C restatement)")
    ap.add_argument("--ref-seconds", type=float, default=15.0, help="budget of the cpu_baseline leg (real reference)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cnn", action="store_true", help="skip the CNN forward leg (BASELINE configs[3])")
    ap.add_argument("--no-dense", action="store_true", help="skip the dense-model scoring leg")
    ap.add_argument("--workload", default="screen", choices=["screen", "e2e"],
                    help="screen = BASELINE configs[1] (the headline); e2e = configs[4]: pockets -> models -> screening")
    ap.add_argument("--pockets-per-gpu", type=int, default=16, help="e2e: synthetic pockets per GPU (128 on 8 GPUs)")
    ap.add_argument("--e2e-ligands", type=int, default=131072, help="e2e: ligands per GPU")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16", "bf16x3"], help="e2e: CNN precision mode")
    return ap.parse_args()


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic():
    """dram bytes per scoring launch from the committed ncu --set full capture, if there is one."""
    path = os.path.join(ROOT, "profiles", "scoring_kernel_latest.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return None


class ClockSampler:
    QUERY = (
        "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
        "clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, gpu_index: int):
        self.proc = None
        self.path = f"/tmp/pmnet_clocks_{os.getpid()}.csv"
        self.gpu = gpu_index

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=self.f, stderr=subprocess.DEVNULL,
            )  # fmt: skip
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        with open(self.path) as f:
            for line in f:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 9:
                    continue
                try:
                    sm.append(float(parts[1]))
                    mx.append(float(parts[2]))
                except ValueError:
                    continue
                for name, val in zip(names, parts[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def workload_config(args, world):
    return {
        "workload": f"6OIM-like synthetic model (35 nodes/26 clusters) x {args.ligands} synthetic ligands x "
        f"{args.conformers} conformers per GPU (BASELINE configs[1])",
        "ligands_per_gpu": args.ligands,
        "conformers_per_ligand": args.conformers,
        "templates": args.templates,
        "topk": args.topk,
        "parallelism": f"ligand-sharded x{world}, one NCCL all-gather of top-k per step",
        "l2_policy": "inputs larger than L2 (library coordinates >> 126 MB per pass)",
    }


def conv_leg(dev, batch: int = 8, iters: int = 10):
    """Secondary measurement: the 96->96 3x3x3 convolution (+BN+ReLU) that is ~90 % of the CNN forward FLOPs
    (BASELINE configs[3] shape: 64^3 grids), against the measured bf16 tensor peak."""
    import torch

    from pharmaconet_b200 import conv

    g = torch.Generator(device=dev).manual_seed(0)
    x = conv.to_c8(torch.randn((batch, 96, 64, 64, 64), generator=g, device=dev))
    w = conv.pack_weights_k3(torch.randn((96, 96, 3, 3, 3), generator=g, device=dev) * 0.03)
    scale = torch.rand(96, generator=g, device=dev) + 0.5
    bias = torch.randn(96, generator=g, device=dev) * 0.2
    for _ in range(3):
        conv.conv3d_k3_c96(x, w, scale, bias, True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        conv.conv3d_k3_c96(x, w, scale, bias, True)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / iters
    flop = 2.0 * batch * 64**3 * 96 * 96 * 27
    peak, src = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peak, src = float(json.load(f)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    prof_path = os.path.join(ROOT, "profiles", "conv3d_r01_ncu_full.json")
    traffic = None
    if os.path.exists(prof_path):
        with open(prof_path) as f:
            m = json.load(f)["metrics"]
        unit = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        traffic = sum(float(m[k][0]) * unit[m[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
    return {
        "bound": "tensor", "achieved": flop / ms / 1e9, "peak": peak, "unit": "TFLOP/s", "frac": flop / ms / 1e9 / peak,
        "traffic": traffic, "kernel": "conv3d_k3_c96_kernel", "kernel_ms_avg": ms, "flop_per_launch": flop,
        "workload": f"{batch} x 96ch x 64^3 -> 96ch, 3x3x3, fused BN+ReLU, bf16 operands / fp32 accumulate",
        "peak_source": src,
    }  # fmt: skip


def cnn_forward_leg(dev, batch: int = 64, chunk: int = 8, iters: int = 2):
    """BASELINE configs[3]: the CNN forward (forward_feature + cavity extraction + token prediction) over `batch`
    synthetic 64^3 pockets in chunks of `chunk`, seeded synthetic weights (no trained checkpoint exists offline), in
    both precisions of this package, next to the unmodified reference nn.Modules (oracle/_ref) through torch / cuDNN on
    the same GPU. Algorithmic FLOPs per pocket: 479.9 G."""
    import numpy as np
    import torch

    from pharmaconet_b200 import cnn, cnn_weights

    G = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(G, "cnn_manifest.json")) as f:
        man = json.load(f)
    buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
    sd = cnn_weights.synth_state_dict(man, buf, 0)
    model = cnn.PharmacoNetModel(sd, dev)
    g = torch.Generator().manual_seed(0)
    images = torch.rand((chunk, 33, 64, 64, 64), generator=g).to(dev)
    tokens = torch.cat([torch.randint(0, 64, (200, 3), generator=g), torch.randint(0, 10, (200, 1), generator=g)], 1).long().to(dev)
    n_chunks = max(1, batch // chunk)
    flop = 479.9e9
    peak, src = 1450.0, "fallback (B200_PROFILING.md)"
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peak, src = float(json.load(f)["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained)"

    def timed(fn):
        fn()  # warm-up (also builds the cached weight operands)
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            for _ in range(n_chunks):
                fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / (iters * n_chunks * chunk)  # ms per pocket

    def ours():
        feats = model.forward_feature(images, nchw=False)
        model.forward_cavity_extraction(feats[-1])
        model.forward_token_prediction(feats[-1], [tokens] * chunk)

    out = {
        "workload": f"{batch} synthetic 64^3 x 33-channel pockets in chunks of {chunk}, 200 tokens each: forward_feature + "
                    "cavity extraction + token prediction (BASELINE configs[3]); seeded synthetic weights",
        "flop_per_pocket": flop, "peak": peak, "peak_unit": "TFLOP/s", "peak_source": src,
    }  # fmt: skip
    for prec in ("bf16x3", "bf16"):
        model.precision = prec
        ms = timed(ours)
        out[prec] = {"ms_per_pocket": ms, "tflops_algorithmic": flop / ms / 1e9, "frac_of_peak": flop / ms / 1e9 / peak}
    out["bf16x3"]["note"] = "two-term bf16 operands, 3 tensor-core passes per product: integer outputs follow the fp32 reference"
    out["bf16"]["note"] = "single pass on bf16 operands"
    del model
    # the reference's own modules on this GPU (torch / cuDNN; TF32 convolutions are torch's default)
    try:
        import ref_harness

        ref_harness.import_reference()
        from pmnet.network import build_model

        ref = build_model({}).eval()
        ref.load_state_dict(sd, strict=True)
        ref = ref.to(dev)

        def theirs():
            with torch.no_grad():
                feats = ref.forward_feature(images)
                ref.forward_cavity_extraction(feats[-1])
                ref.forward_token_prediction(feats[-1], [tokens] * chunk)

        ms = timed(theirs)
        out["reference_modules_same_gpu"] = {
            "ms_per_pocket": ms, "tflops_algorithmic": flop / ms / 1e9,
            "note": "unmodified pmnet.network modules (oracle/_ref) through torch/cuDNN, fp32 weights, TF32 convolutions",
        }  # fmt: skip
        del ref
    except Exception as e:  # noqa: BLE001
        out["reference_modules_same_gpu"] = {"error": repr(e)}
    torch.cuda.empty_cache()
    return out


def dense_model_leg(lib, dev, args, n_ligands: int = 32768, hotspots: int = 60):
    """Second scoring workload: a synthetic model of the size the CNN produces for hotspot-rich pockets (47 nodes / 27
    clusters instead of the headline's 35 / 26). Trees are an order of magnitude larger on average and heavy tailed
    (single ligands with 10^6 - 10^7 tree nodes): trees over PmScoreConfig.heavy_budget nodes are walked by many warps
    (the task rounds of pmnet_score_batch). One timed pass."""
    import torch

    from pharmaconet_b200 import scoring, synthetic
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    model = PharmacophoreModel.create("", (0.0, 0.0, 0.0), synthetic.make_hotspot_infos(seed=21, n_hotspots=hotspots))
    n = min(n_ligands, lib.n_ligands)
    sub = scoring.DeviceLigandBatch(lib.tensors, n, n * args.conformers, max_conformers=lib.max_conformers)
    dm = scoring.DeviceModel(model.packed, dev)
    sub.set_order(scoring.cost_order(dm, sub))
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = scoring.score_batch(dm, sub, with_stats=True)
    n_over = scoring.rescore_overflowed_device(dm, sub, out)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    stats = out["stats"].cpu().numpy().view("uint32").astype("float64")
    return {
        "workload": f"synthetic dense model ({len(model.nodes)} nodes / {len(model.node_clusters)} clusters, {hotspots} hotspots) x "
                    f"first {n} ligands x {args.conformers} conformers of the same library, resident, one pass",
        "value": n * args.conformers / (ms * 1e-3), "unit": UNIT, "ms_per_pass": ms,
        "tree_nodes_mean": float(stats[:, 0].mean()), "tree_nodes_max": float(stats[:, 0].max()),
        "tree_nodes_per_sec": float(stats[:, 0].sum() / (ms * 1e-3)),
        "pair_entries_mean": float(stats[:, 3].mean()), "n_overflow_rerun": int(n_over),
        "note": "trees over 65536 nodes are split over many warps (task rounds); the headline model has 1.6e3 tree nodes per ligand",
    }  # fmt: skip


def host_prefix(dev_lib, n):
    """First n ligands of a device library as a host LigandBatch (for the CPU port)."""
    import numpy as np

    from pharmaconet_b200.packing import LigandBatch

    t = dev_lib.tensors
    n = min(n, dev_lib.n_ligands)
    nn = int(t["lig_node_off"][n])
    nq = int(t["lig_cluster_off"][n])
    ncn = int(t["cluster_node_off"][nq])
    nx = int(t["coord_off"][n])
    return LigandBatch.from_arrays(
        dict(
            lig_node_off=t["lig_node_off"][: n + 1].cpu().numpy(),
            lig_cluster_off=t["lig_cluster_off"][: n + 1].cpu().numpy(),
            cluster_node_off=t["cluster_node_off"][: nq + 1].cpu().numpy(),
            cluster_nodes=t["cluster_nodes"][:ncn].cpu().numpy(),
            node_type_mask=t["node_type_mask"][:nn].cpu().numpy(),
            n_conf=t["n_conf"][:n].cpu().numpy(),
            coord_off=t["coord_off"][: n + 1].cpu().numpy(),
            coords=t["coords"][:nx].cpu().numpy(),
        )
    )


MODEL_PATH = os.path.join(ROOT, "tests", "golden", "model_syn0.pm")
REF_SECONDS_PER_LIGAND_CORE = 0.11  # reference, 32 conformers, one core (SURVEY section 6 probe); sizes the samples


def reference_available() -> bool:
    return os.path.isdir("/root/reference/src/pmnet") or os.path.isdir(os.path.join(ROOT, "oracle", "_ref", "pmnet"))


def run_real_reference(typed, steps: int, warmup: int, budget_s: float = 0.0) -> dict:
    """Score `typed` (list of TypedLigand) with the unmodified reference under Pool(os.cpu_count()) in a separate
    process (oracle/ref_pool.py); returns its JSON (step_seconds, scores, procs)."""
    import pickle
    import tempfile

    tmp = tempfile.mkdtemp(prefix="pmnet_ref_")
    lig_path, out_path = os.path.join(tmp, "ligands.pkl"), os.path.join(tmp, "out.json")
    with open(lig_path, "wb") as f:
        pickle.dump(list(typed), f)
    env = dict(os.environ)
    env.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "pmnet_numba_cache"))
    env["CUDA_VISIBLE_DEVICES"] = ""  # the reference's scoring path is CPU only
    cmd = [
        sys.executable, os.path.join(ROOT, "oracle", "ref_pool.py"), "--model", MODEL_PATH, "--ligands", lig_path,
        "--out", out_path, "--steps", str(steps), "--warmup", str(warmup), "--budget-s", str(budget_s),
    ]  # fmt: skip
    r = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"oracle/ref_pool.py failed:\n{r.stderr[-2000:]}")
    with open(out_path) as f:
        out = json.load(f)
    for pth in (lig_path, out_path):
        try:
            os.unlink(pth)
        except OSError:
            pass
    return out


def reference_sample_size(total_seconds: float, steps: int, conformers: int) -> int:
    cores = os.cpu_count() or 1
    per = REF_SECONDS_PER_LIGAND_CORE * max(1, conformers) / 32.0
    return int(max(2 * cores, min(2048, total_seconds * cores / per / max(1, steps))))


def run_reference(args, rank, world):
    """The reference's own CPU implementation of the path: GraphMatcher.run + numba under Pool(all host cores)."""
    if rank != 0:
        return
    if reference_available():
        return run_reference_real(args, world)
    return run_reference_port(args, world)


def run_reference_real(args, world):
    import torch

    from pharmaconet_b200 import synthetic

    n = reference_sample_size(150.0, args.steps, args.conformers)
    if torch.cuda.is_available():
        lib = synthetic.make_library_device(args.ligands, args.conformers, args.seed, "cuda:0", args.templates, keep_atoms=n)
        typed = lib.typed_prefix
        del lib
        torch.cuda.empty_cache()
        what = f"first {len(typed)} ligands x {args.conformers} conformers of the same library per step"
    else:
        typed = synthetic.make_ligands(n, args.conformers, args.seed)
        what = f"{len(typed)} ligands x {args.conformers} conformers of the same generator per step (no GPU here)"
    out = run_real_reference(typed, args.steps, args.warmup)
    dt = sum(out["step_seconds"])
    steps = len(out["step_seconds"])
    value = out["n_conformers"] * steps / dt
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt / steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": out["procs"], "kind": "reference",
            "sample": what + "; unmodified pmnet GraphMatcher.run + numba under multiprocessing.Pool",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def run_reference_port(args, world):
    """Fallback when neither /root/reference nor oracle/_ref exists: the C restatement on all host cores."""
    import numpy as np
    import torch

    import oracle as orc
    from pharmaconet_b200 import synthetic
    from pharmaconet_b200.packing import LigandBatch, PackedModel
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    packed = PackedModel.from_model(PharmacophoreModel.load(MODEL_PATH))
    if torch.cuda.is_available():
        lib = synthetic.make_library_device(args.ligands, args.conformers, args.seed, "cuda:0", args.templates)
        sample = host_prefix(lib, SAMPLE_LIGANDS)
        del lib
        torch.cuda.empty_cache()
    else:
        sample = LigandBatch.from_typed(synthetic.make_ligands(SAMPLE_LIGANDS, args.conformers, args.seed))
    cores = orc.max_threads()
    n_conf = sample.num_conformers_total
    for _ in range(args.warmup):
        orc.score(packed, sample, threads=cores, end=min(256, sample.num_ligands))
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.score(packed, sample, threads=cores)
    dt = time.perf_counter() - t0
    value = n_conf * args.steps / dt
    cfg = workload_config(args, world)
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {
            "value": value, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {sample.num_ligands} ligands x {args.conformers} conformers of the workload per step",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }  # fmt: skip
    print(json.dumps(line), flush=True)


def run_e2e(args, rank, local_rank, world):
    """BASELINE configs[4]: P synthetic pockets (P = pockets_per_gpu x N) through the CNN + model construction on their
    owner rank, one all-gather of the packed models, every rank screens its resident ligand shard against all models,
    one all-gather of the per-model top-k. Synthetic network weights (no trained checkpoint offline): the models are
    structurally realistic only; what is measured is throughput, stage by stage."""
    import numpy as np
    import torch
    import torch.distributed as dist

    from pharmaconet_b200 import cnn_weights, pipeline, synthetic
    from pharmaconet_b200.module import PharmacoNet

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    G = os.path.join(ROOT, "tests", "golden")
    with open(os.path.join(G, "cnn_manifest.json")) as f:
        man = json.load(f)
    buf = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(G, "cnn_buffers.npz")).items()}
    net = PharmacoNet(dev, verbose=False, checkpoint=cnn_weights.synth_checkpoint(man, buf, 0), precision=args.precision)
    gold = np.load(os.path.join(G, "cnn_pipeline_golden.npz"))
    tokens = torch.from_numpy(gold["tokens"]).long()
    token_pos = (tokens[:, :3].float() - 31.5) * 0.5
    n_pockets = args.pockets_per_gpu * world
    data = []
    for p in range(n_pockets):  # every rank describes all pockets; only its own are materialised as tensors
        if p % world == rank:
            image = torch.rand((33, 64, 64, 64), generator=torch.Generator().manual_seed(p))
            mask = torch.rand((64, 64, 64), generator=torch.Generator().manual_seed(1000 + p)) < 0.8
            data.append((image, mask, token_pos, tokens))
        else:
            data.append(None)
    shard = synthetic.make_library_device(args.e2e_ligands, args.conformers, args.seed, dev, args.templates,
                                          coord_seed=args.seed + rank)
    id_base = rank * args.e2e_ligands
    for _ in range(max(1, min(args.warmup, 1))):
        res = pipeline.model_and_screen(net, data, shard, id_base, args.topk, rank, world)
    barrier()
    t0 = time.perf_counter()
    stages = {"modeling": 0.0, "exchange": 0.0, "screening": 0.0}
    for _ in range(args.steps):
        res = pipeline.model_and_screen(net, data, shard, id_base, args.topk, rank, world)
        for k in stages:
            stages[k] += res.seconds[k]
    barrier()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    live = sum(m is not None for m in res.models)
    pairs = live * world * shard.n_conformers_total * args.steps
    if rank == 0:
        nodes = [m.num_nodes for m in res.models if m is not None]
        line = {
            "metric": "pocket_ligand_conformer_pairs_scored_per_sec", "value": pairs / dt, "unit": "pairs/s", "n_gpus": world,
            "steps": args.steps, "warmup": 1, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 scoring, " + args.precision + " CNN", "data": "synthetic",
            "config": {
                "workload": f"{n_pockets} synthetic 64^3 pockets -> pharmacophore models -> each against {args.e2e_ligands} "
                            f"ligands x {args.conformers} conformers per GPU (BASELINE configs[4] shape)",
                "pockets": n_pockets, "models_with_nodes": live, "model_nodes_mean": float(np.mean(nodes)) if nodes else 0.0,
                "ligands_per_gpu": args.e2e_ligands, "parallelism": f"pockets p -> rank p mod {world}; ligand shards x{world}; "
                "two small all-gathers (packed models, per-model top-k)",
            },
            "stage_seconds_per_step_rank0": {k: v / args.steps for k, v in stages.items()},
            "pockets_per_sec": n_pockets * args.steps / dt, "n_overflow_rerun": res.n_overflow,
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "e2e":
        run_e2e(args, rank, local_rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    from pharmaconet_b200 import screening, synthetic
    from pharmaconet_b200.packing import LigandBatch, PackedModel
    from pharmaconet_b200.pharmacophore_model import PharmacophoreModel

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the scoring path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = None
    if world > 1:
        # several ranks stream from host memory at once: keep this rank's pinned pages on its GPU's NUMA node
        from pharmaconet_b200.affinity import bind_to_gpu

        numa_cpus = bind_to_gpu(local_rank, world)
    if world > 1:
        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) out of it
        # (it can also come from /etc/nccl.conf, which never overrides an environment variable)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize(dev)

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    packed = PackedModel.from_model(PharmacophoreModel.load(MODEL_PATH))
    want_ref = rank == 0 and world == 1 and not args.no_cpu_baseline and reference_available()
    n_ref = reference_sample_size(args.ref_seconds, 1, args.conformers) if want_ref else 0
    # every rank: the same topology set, its own conformers (equal work per rank: weak scaling)
    lib = synthetic.make_library_device(
        args.ligands, args.conformers, args.seed, dev, args.templates, keep_atoms=n_ref, coord_seed=args.seed + rank
    )
    typed_prefix = lib.typed_prefix
    n_lig, n_conf = lib.n_ligands, lib.n_conformers_total
    alg_bytes = lib.nbytes() + 4 * n_lig  # every input array once + one fp32 score per ligand
    from pharmaconet_b200.scoring import ScoreConfig

    scfg = ScoreConfig(args.stream_warps, args.stream_ctas, 0) if (args.stream_warps or args.stream_ctas) else None
    scr = screening.Screener(packed, dev, k=args.topk, block_ligands=args.block_ligands, n_slots=args.slots,
                             ramp=not args.no_ramp, stream_config=scfg, lpt=not args.no_lpt)
    id_base = rank * n_lig

    # ---------------------------------------------------------------- leg 1: library resident in HBM
    for _ in range(args.warmup):
        res = scr.screen_device(lib, id_base)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    scr.record_kernel_events = True
    scr.kernel_events.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    e0.record()
    for _ in range(args.steps):
        lib.set_order(None)  # the longest-first order is recomputed inside every timed step (0.2 ms)
        res = scr.screen_device(lib, id_base)
        launches += res.launches
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    kernel_ms = [a.elapsed_time(b) for a, b in scr.kernel_events]
    scr.record_kernel_events = False
    kernel_ms_avg = max_over_ranks(sum(kernel_ms) / len(kernel_ms))
    value = world * n_conf * args.steps / (ms_total * 1e-3)
    gpu_scores = res.scores  # device tensor, this rank
    top_ids = res.topk_ids.cpu().numpy()

    # ---------------------------------------------------------------- leg 1b: a denser, CNN-sized model (rank 0, N = 1)
    dense = None
    if rank == 0 and world == 1 and not args.no_dense:
        try:
            dense = dense_model_leg(lib, dev, args)
        except Exception as e:  # noqa: BLE001 - the headline metric must still be printed
            dense = {"error": repr(e)}

    # ---------------------------------------------------------------- leg 2: end to end from pinned host memory
    e2e = None
    if not args.no_e2e:
        host = LigandBatch.from_arrays({k: v.cpu().numpy() for k, v in lib.tensors.items()})
        host = screening.pin_library(host)
        h2d = sum(v.nbytes for v in host.arrays().values())
        d2h = n_lig * 8 + args.topk * 12
        del lib
        torch.cuda.empty_cache()
        for _ in range(max(1, args.warmup)):
            r2 = scr.screen_host(host)
        barrier()
        scr.record_kernel_events = True
        scr.kernel_events.clear()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            r2 = scr.screen_host(host)
        e1.record()
        barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        scr.record_kernel_events = False
        e2e_kernel_ms = sum(a.elapsed_time(b) for a, b in scr.kernel_events) / args.steps
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        e2e = {
            "value": world * n_conf * args.steps / (ms_e2e * 1e-3), "unit": UNIT,
            "h2d_bytes_per_step": world * h2d, "d2h_bytes_per_step": world * d2h,
            "ms_per_step": ms_e2e / args.steps, "api": "pharmaconet_b200.screening.Screener.screen_host",
            "n_overflow_rerun": r2.n_overflow,
            # sum of the block launches' own durations (they overlap on two streams, so this may exceed the step)
            "kernel_ms_sum_per_step": e2e_kernel_ms, "block_ligands": args.block_ligands, "slots": args.slots,
            "cpus_bound_rank0": len(numa_cpus) if numa_cpus else None,
        }  # fmt: skip
        same = bool(np.array_equal(r2.scores, gpu_scores.cpu().numpy()))
        e2e["scores_identical_to_resident_leg"] = same
    else:
        host = None

    # ---------------------------------------------------------------- leg 3: CPU baselines on bounded prefixes (rank 0, N = 1)
    cpu_baseline = None
    cpu_port = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import oracle as orc

        g_all = gpu_scores.cpu().numpy().astype(np.float64)
        parity = {"tolerance": 1e-5}
        # (a) the unmodified reference: GraphMatcher.run + numba under Pool(all cores) on the first n_ref ligands
        if typed_prefix:
            try:
                out = run_real_reference(typed_prefix, steps=1, warmup=1)
                dt = sum(out["step_seconds"])
                ref = np.asarray(out["scores"], dtype=np.float64)
                cpu_baseline = {
                    "value": out["n_conformers"] * len(out["step_seconds"]) / dt, "unit": UNIT, "cores": out["procs"],
                    "kind": "reference",
                    "sample": f"first {len(typed_prefix)} ligands x {args.conformers} conformers of the same library "
                              f"({dt:.1f} s); unmodified pmnet GraphMatcher.run + numba under multiprocessing.Pool",
                }  # fmt: skip
                rel = np.abs(g_all[: len(ref)] - ref) / np.maximum(np.abs(ref), 1e-12)
                parity["checked_ligands_vs_reference"] = int(len(ref))
                parity["max_rel_err_vs_reference"] = float(rel.max())
            except Exception as e:  # noqa: BLE001 - the headline metric must still be printed
                cpu_baseline = {"error": repr(e), "kind": "reference"}
        # (b) the C restatement on all cores over a larger prefix: the wide parity check
        cores = orc.max_threads()
        src = host if host is not None else None
        if src is None:
            src = LigandBatch.from_arrays({k: v.cpu().numpy() for k, v in lib.tensors.items()})
        orc.score(packed, src, threads=cores, end=min(256, n_lig))  # warm-up
        done, t0 = 0, time.perf_counter()
        ref_scores = []
        while done < n_lig and time.perf_counter() - t0 < args.cpu_seconds:
            end = min(n_lig, done + SAMPLE_LIGANDS)
            ref_scores.append(orc.score(packed, src, threads=cores, begin=done, end=end)["scores"])
            done = end
        dt = time.perf_counter() - t0
        ref_scores = np.concatenate(ref_scores)
        cpu_port = {
            "value": done * args.conformers / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": f"first {done} ligands x {args.conformers} conformers of the same library ({dt:.1f} s)",
        }  # fmt: skip
        if cpu_baseline is None:
            cpu_baseline = cpu_port
        rel = np.abs(g_all[:done] - ref_scores) / np.maximum(np.abs(ref_scores), 1e-12)
        parity["checked_ligands"] = int(done)
        parity["max_rel_err_vs_cpu_port"] = float(rel.max())

    # ---------------------------------------------------------------- leg 4: the CNN's dominant kernel (tensor roofline)
    conv_roofline = None
    if rank == 0:
        try:
            conv_roofline = conv_leg(dev)
        except Exception as e:  # noqa: BLE001 - the headline metric must still be printed
            conv_roofline = {"error": repr(e)}

    cnn_forward = None
    if rank == 0 and world == 1 and not args.no_cnn:
        try:
            cnn_forward = cnn_forward_leg(dev)
        except Exception as e:  # noqa: BLE001
            cnn_forward = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = load_peaks()
        achieved = alg_bytes / (kernel_ms_avg * 1e-3) / 1e9
        prof = load_traffic()
        # the bound that actually limits the kernel: warp instructions issued vs 4 issue slots per SM per clock
        issue = None
        if prof and prof.get("warp_instructions_per_ligand") and clocks and clocks.get("sm_mhz"):
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            ach = prof["warp_instructions_per_ligand"] * n_lig / (kernel_ms_avg * 1e-3) / 1e12
            pk = n_sm * 4 * clocks["sm_mhz"] * 1e6 / 1e12
            issue = {
                "bound": "issue", "achieved": ach, "peak": pk, "unit": "T warp-inst/s", "frac": ach / pk,
                "warp_instructions_per_ligand": prof["warp_instructions_per_ligand"],
                "source": prof.get("source"), "peak_source": f"{n_sm} SMs x 4 schedulers x sampled SM clock",
            }  # fmt: skip
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "e2e": e2e, "gpu_launches": launches,
            "roofline": {
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                # ncu --set full capture of the same kernel on 131072 ligands of this workload, scaled per ligand
                "traffic": (prof["dram_bytes_per_ligand"] * n_lig) if prof else None,
                "traffic_source": (prof or {}).get("source"),
                # CUDA events around one pmnet_score_batch call on its stream: the specialised kernel is 98 % of it
                # (profiles/launches_r02_bench_summary.txt), the general / task / finish kernels of the call the rest
                "kernel": "pmnet_score_fast_kernel (+ the general, task-round and finish kernels of the same call)",
                "kernel_ms_avg": kernel_ms_avg,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "note": "the path is instruction-issue bound, not HBM bound (DESIGN.md section 5): the HBM fraction "
                        "is reported because BASELINE.json asks for it",
            },
            "issue_roofline": issue,
            "dense_model": dense,
            "cnn_conv3d_roofline": conv_roofline,
            "cnn_forward": cnn_forward,
            "cpu_baseline": cpu_baseline, "cpu_port": cpu_port, "parity": parity, "clocks": clocks,
            "top1": {"id": int(top_ids[0]), "score": float(res.topk_scores[0].item())},
        }  # fmt: skip
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
