"""TEST INFRASTRUCTURE - the UNMODIFIED reference's scoring path on host cores, the way its own CLI runs it.

    python oracle/ref_pool.py --model M.pm --ligands L.pkl --out O.json [--procs N] [--steps K] [--warmup W]

Mirrors `/root/reference/screening.py:46-68`: the model is loaded once, `multiprocessing.Pool(procs).map` hands one
ligand per task to `GraphMatcher(model, ligand, weights).run()` (graph_match.py:94-101; the numba kernels of
match_utils_numba.py). `L.pkl` holds a list of `pharmaconet_b200.ligand.TypedLigand` (typed atoms + conformer
coordinates); every worker builds the reference's own `LigandGraph` from them (ligand.py:110-259), i.e. node
positions, edge distances and clusters are the reference's arithmetic, not this package's.

It runs as a separate process so that the benchmark's CUDA context is never forked. The reference package is
imported from `/root/reference/src` when present, else from `oracle/_ref` (see oracle/make_ref.py).
Only bench.py (cpu_baseline / `--impl reference`) and tests may execute this file.
"""

from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import pickle
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)

_G = {}


def _init(model_path: str, weights):
    import ref_harness

    pm, graph_match, _, _ = ref_harness.import_reference()
    _G["model"] = pm.PharmacophoreModel.load(model_path)
    _G["gm"] = graph_match
    _G["weights"] = weights
    _G["RefLigand"] = ref_harness.RefLigand


def _score(typed) -> float:
    # screening.py:46-47 -> pharmacophore_model.py:101-106 (ligand construction included, as in scoring_file)
    lig = _G["RefLigand"](typed)
    return float(_G["gm"].GraphMatcher(_G["model"], lig, _G["weights"]).run())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", required=True)
    ap.add_argument("--ligands", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--procs", type=int, default=0)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--warmup-ligands", type=int, default=0, help="ligands per warm-up step (0 = 2 per process)")
    ap.add_argument("--budget-s", type=float, default=0.0, help="stop adding timed steps once this much time is spent")
    args = ap.parse_args()

    import ref_harness

    with open(args.ligands, "rb") as f:
        ligands = pickle.load(f)
    procs = args.procs or os.cpu_count() or 1
    ref_harness.import_reference()  # compiles / loads the numba kernels once; fork()ed workers inherit them
    ctx = mp.get_context("fork")
    with ctx.Pool(procs, initializer=_init, initargs=(args.model, None)) as pool:
        nw = args.warmup_ligands or min(len(ligands), 2 * procs)
        for _ in range(max(1, args.warmup)):
            pool.map(_score, ligands[:nw], chunksize=1)
        times, scores = [], None
        t_all = time.perf_counter()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            scores = pool.map(_score, ligands)  # screening.py:68 (default chunking)
            times.append(time.perf_counter() - t0)
            if args.budget_s and time.perf_counter() - t_all > args.budget_s:
                break
    n_conf = sum(int(l.atom_positions.shape[1]) for l in ligands)
    with open(args.out, "w") as f:
        json.dump(
            {
                "procs": procs, "n_ligands": len(ligands), "n_conformers": n_conf, "step_seconds": times,
                "scores": scores, "reference_src": ref_harness.REFERENCE_SRC,
            },
            f,
        )  # fmt: skip


if __name__ == "__main__":
    main()
