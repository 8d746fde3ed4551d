#!/usr/bin/env python
"""Virtual screening on GPUs - the reference's `screening.py` entry point (same flags, same CSV) on the B200 path.

    python screening.py -p model.pm -d library_dir -o result.csv [--gpus N] [weights...]
    torchrun --nproc-per-node N screening.py ...            # one process per GPU, library blocks interleaved

`-d` is either a directory of .sdf / .mol2 files (one ligand per file, every record a conformer - needs OpenBabel for
typing, like the reference) or a packed library `.npz` written by `pharmaconet_b200.packing.save_library` (pre-typed
ligands; no toolkit needed). `--cpus` sets the typing worker processes; scoring always runs on the GPU(s).
Reference: screening.py:9-75.
"""

from __future__ import annotations

import argparse
import multiprocessing
import os
from pathlib import Path

import numpy as np


def parse_args():
    p = argparse.ArgumentParser("scoring", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    cfg = p.add_argument_group("config")
    cfg.add_argument("-p", "--pharmacophore_model", type=str, required=True, help="path of pharmacophore model (.pm | .json)")
    cfg.add_argument("-d", "--library_dir", type=str, required=True, help="molecular library directory, a packed .npz, or a packed (memory-mapped) library directory")
    cfg.add_argument("-o", "--out", type=str, required=True, help="result file path")
    cfg.add_argument("--cpus", type=int, default=1, help="number of cpus (ligand typing workers)")
    cfg.add_argument("--num_conformers", type=int, default=None, help="use only the first N conformers of each file")
    par = p.add_argument_group("parameter")
    par.add_argument("--hydrophobic", type=float, default=1.0, help="weight for hydrophobic carbon")
    par.add_argument("--aromatic", type=float, default=4.0, help="weight for aromatic ring")
    par.add_argument("--hba", type=float, default=4.0, help="weight for hbond acceptor")
    par.add_argument("--hbd", type=float, default=4.0, help="weight for hbond donor")
    par.add_argument("--halogen", type=float, default=4.0, help="weight for halogen atom")
    par.add_argument("--anion", type=float, default=8.0, help="weight for anion")
    par.add_argument("--cation", type=float, default=8.0, help="weight for cation")
    return p.parse_args()


def _type_file(job):
    from pharmaconet_b200.ligand_typing import typed_ligand_from_file

    path, nconf = job
    return typed_ligand_from_file(path, nconf)


def load_library(args):
    from pharmaconet_b200.packing import LigandBatch, is_library_dir, load_library

    src = Path(args.library_dir)
    if (src.is_file() and src.suffix == ".npz") or is_library_dir(src):
        return load_library(src)  # (a packed directory is memory-mapped: blocks are read as they are streamed)
    files = sorted(src.rglob("*.sdf")) + sorted(src.rglob("*.mol2"))
    print(f"find {len(files)} molecules")
    jobs = [(str(f), args.num_conformers) for f in files]
    if args.cpus > 1:
        with multiprocessing.Pool(args.cpus) as pool:
            ligs = pool.map(_type_file, jobs, chunksize=64)
    else:
        ligs = [_type_file(j) for j in jobs]
    return LigandBatch.from_typed(ligs), [str(f) for f in files]


def main():
    args = parse_args()
    import torch
    import torch.distributed as dist

    from pharmaconet_b200 import PharmacophoreModel
    from pharmaconet_b200.screening import Screener, write_csv

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        from pharmaconet_b200.affinity import bind_to_gpu

        bind_to_gpu(local)  # pinned library pages on the GPU's NUMA node
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    model = PharmacophoreModel.load(args.pharmacophore_model)
    weights = dict(
        Cation=args.cation, Anion=args.anion, Aromatic=args.aromatic, HBond_donor=args.hbd, HBond_acceptor=args.hba,
        Halogen=args.halogen, Hydrophobic=args.hydrophobic,
    )  # fmt: skip
    library, names = load_library(args)
    scr = Screener(model, weights=weights, k=min(1000, max(1, library.num_ligands)))
    from pharmaconet_b200.packing import is_library_dir

    if library.coords.nbytes >= (64 << 20) and not is_library_dir(args.library_dir):
        from pharmaconet_b200.screening import pin_library

        library = pin_library(library)  # page-locked: block copies overlap the scoring kernel
    res = scr.screen_host(library, rank=rank, world=world, gather=world > 1)
    scores = np.zeros(library.num_ligands, dtype=np.float32)
    scores[res.ids] = res.scores
    if world > 1:  # every rank scored its blocks; sum of disjoint contributions = the full vector
        t = torch.from_numpy(scores).cuda()
        dist.all_reduce(t)
        scores = t.cpu().numpy()
    if rank == 0:
        write_csv(args.out, names, scores)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
