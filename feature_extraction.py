#!/usr/bin/env python
"""Feature extraction for downstream models - the reference's `feature_extraction.py` entry point (same flags).

    python feature_extraction.py -p pocket.pdb -o features.pt (--ref_ligand lig.sdf | --center x y z) --weight_path model.tar

Saves `[multi_scale_features, hotspot_infos]` with torch.save like feature_extraction.py:67-70. Parsing the PDB needs
the reference's `pmnet.data` (OpenBabel, biopython, molvoxel); `--protein_data` takes a pre-parsed tensor tuple.
"""

from __future__ import annotations

import argparse


def parse_args():
    p = argparse.ArgumentParser("PharmacoNet Feature Extraction Script", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    p.add_argument("-p", "--protein", type=str, help="custom path of protein pdb file (.pdb)")
    p.add_argument("--protein_data", type=str, help="torch-saved (image, mask, token_pos, tokens) tuple")
    p.add_argument("-o", "--out", type=str, required=True, help="save path of features (torch object)")
    p.add_argument("--ref_ligand", type=str, help="path of ligand to define the center of box (.sdf, .pdb, .mol2)")
    p.add_argument("--center", nargs="+", type=float, help="coordinate of the center")
    p.add_argument("--weight_path", type=str, required=True, help="pharmaconet weight path (model.tar)")
    p.add_argument("--cuda", action="store_true", help="accepted for compatibility: this path always runs on CUDA")
    return p.parse_args()


def main():
    args = parse_args()
    import torch

    from pharmaconet_b200.module import get_pmnet_dev

    net = get_pmnet_dev("cuda", weight_path=args.weight_path)
    if args.protein_data:
        out = net.run_extraction(torch.load(args.protein_data))
    else:
        assert args.protein and (args.ref_ligand or args.center)
        out = net.feature_extraction(args.protein, args.ref_ligand, tuple(args.center) if args.center else None)
    feats, infos = out
    torch.save([[f.cpu() for f in feats], [{k: (v.cpu() if hasattr(v, "cpu") else v) for k, v in i.items()} for i in infos]], args.out)


if __name__ == "__main__":
    main()
