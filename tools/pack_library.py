#!/usr/bin/env python
"""Type a directory of molecule files once and write a packed library (`.npz`) that `screening.py -d lib.npz` streams
to the GPUs without touching a chemistry toolkit again (ligand featurisation -> packed
library format).

    python tools/pack_library.py -d library_dir -o library.npz [--cpus N] [--num_conformers K] [--perception auto]

One ligand per file, every record a conformer (the reference's convention, src/pmnet/scoring/ligand.py:63-84).
`.sdf` files are typed with OpenBabel when it is importable (the reference's perception), else with the built-in
approximate reader; `.mol2` / `.pdb` need OpenBabel. Files that fail to parse are reported and skipped.
The file holds the `LigandBatch` arrays (layouts: include/pmnet_b200.h, struct PmLigandBatch) plus `names`.
An output path that does not end in `.npz` becomes a DIRECTORY with one `.npy` per array: `screening.py -d that_dir`
memory-maps it, so a library larger than host memory is streamed from disk block by block.
"""

from __future__ import annotations

import argparse
import multiprocessing
import os
import sys
from pathlib import Path

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _type_file(job):
    from pharmaconet_b200.ligand_typing import typed_ligand_from_file

    path, nconf, perception = job
    try:
        return typed_ligand_from_file(path, nconf, perception=perception)
    except ImportError:
        raise
    except Exception as e:  # noqa: BLE001 - a broken file must not stop a library build
        return (path, repr(e))


def pack(library_dir, out, cpus: int = 1, num_conformers: int | None = None, perception: str = "auto"):
    from pharmaconet_b200.ligand import TypedLigand
    from pharmaconet_b200.packing import LigandBatch, save_library

    src = Path(library_dir)
    files = sorted(src.rglob("*.sdf")) + sorted(src.rglob("*.mol2")) + sorted(src.rglob("*.pdb"))
    jobs = [(str(f), num_conformers, perception) for f in files]
    if cpus > 1 and len(jobs) > 1:
        with multiprocessing.Pool(cpus) as pool:
            res = pool.map(_type_file, jobs, chunksize=64)
    else:
        res = [_type_file(j) for j in jobs]
    ligs = [r for r in res if isinstance(r, TypedLigand)]
    failed = [r for r in res if not isinstance(r, TypedLigand)]
    batch = LigandBatch.from_typed(ligs)
    save_library(out, batch, [lig.name for lig in ligs])
    return batch, failed


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-d", "--library_dir", required=True)
    ap.add_argument("-o", "--out", required=True)
    ap.add_argument("--cpus", type=int, default=1)
    ap.add_argument("--num_conformers", type=int, default=None)
    ap.add_argument("--perception", default="auto", choices=["auto", "openbabel", "builtin"])
    a = ap.parse_args()
    batch, failed = pack(a.library_dir, a.out, a.cpus, a.num_conformers, a.perception)
    for path, err in failed:
        print(f"skipped {path}: {err}", file=sys.stderr)
    print(f"{batch.num_ligands} ligands, {batch.num_conformers_total} conformers, "
          f"{sum(v.nbytes for v in batch.arrays().values()) / 1e6:.1f} MB -> {a.out} ({len(failed)} files skipped)")


if __name__ == "__main__":
    main()
