#!/usr/bin/env python
"""Protein-based pharmacophore modeling - the reference's `modeling.py` entry point on the B200 CNN path.

    python modeling.py --protein pocket.pdb --ref_ligand lig.sdf [--out_dir DIR] [--prefix NAME] [--suffix pm|json]
    python modeling.py --protein_data parsed.pt --center x y z ...      # pre-parsed (image, mask, token_pos, tokens)

Kept flags: -p/--protein, --ref_ligand, --center, --out_dir, --prefix, --suffix, --weight_path, --cuda, --force,
-v (modeling.py:17-57). RCSB download (--pdb / -l / -c / -a) and the PyMOL session export are out of scope
(SURVEY section 2, rows 19-20): they need the network and PyMOL. Protein parsing needs the reference's
`pmnet.data` (OpenBabel, biopython, molvoxel); with `--protein_data` a tensor tuple saved by torch.save is used.
"""

from __future__ import annotations

import argparse
import logging
from pathlib import Path


def parse_args():
    p = argparse.ArgumentParser("pharmacophore modeling script", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    cfg = p.add_argument_group("config")
    cfg.add_argument("-p", "--protein", type=str, help="custom path of protein pdb file (.pdb)")
    cfg.add_argument("--protein_data", type=str, help="torch-saved (image, mask, token_pos, tokens) tuple")
    cfg.add_argument("--out_dir", type=str, help="output directory. default: ./result/{prefix}")
    cfg.add_argument("--prefix", type=str, help="task name")
    cfg.add_argument("--suffix", choices=("pm", "json"), default="pm", help="extension of pharmacophore model")
    env = p.add_argument_group("environment")
    env.add_argument("--weight_path", type=str, required=True, help="pharmaconet weight path (model.tar)")
    env.add_argument("--cuda", action="store_true", help="accepted for compatibility: this path always runs on CUDA")
    env.add_argument("--force", action="store_true", help="force to save the pharmacophore model")
    env.add_argument("-v", "--verbose", action="store_true", help="verbose")
    adv = p.add_argument_group("Advanced Setting")
    adv.add_argument("--ref_ligand", type=str, help="path of ligand to define the center of box (.sdf, .pdb, .mol2)")
    adv.add_argument("--center", nargs="+", type=float, help="coordinate of the center")
    return p.parse_args()


def main():
    args = parse_args()
    logging.basicConfig(level=logging.DEBUG if args.verbose else logging.INFO)
    import torch

    from pharmaconet_b200.module import PharmacoNet

    assert args.protein or args.protein_data, "--protein or --protein_data is required"
    prefix = args.prefix or Path(args.protein or args.protein_data).stem
    out_dir = Path(args.out_dir) if args.out_dir else Path("./result") / prefix
    out_dir.mkdir(parents=True, exist_ok=True)
    model_path = out_dir / f"{prefix}_model.{args.suffix}"
    if model_path.exists() and not args.force:
        logging.info(f"Modeling Pass - {model_path} exists (use --force to overwrite)")
        return
    net = PharmacoNet("cuda", verbose=args.verbose, weight_path=args.weight_path)
    if args.protein_data:
        assert args.center is not None and len(args.center) == 3, "--center x y z is required with --protein_data"
        model = net.create_model(torch.load(args.protein_data), "", tuple(args.center))
    else:
        assert args.ref_ligand or args.center, "--ref_ligand or --center is required"
        model = net.run(args.protein, args.ref_ligand, tuple(args.center) if args.center else None)
    model.save(model_path)
    logging.info(f"Save Pharmacophore Model to {model_path} ({len(model.nodes)} nodes, {len(model.node_clusters)} clusters)")


if __name__ == "__main__":
    main()
