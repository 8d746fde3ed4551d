#!/bin/bash
for so in pharmaconet_b200/_variants/libpmnet_b200_*.so; do
  echo "=== $so"
  PMNET_B200_SO=$PWD/$so timeout 600 python tools/dense_probe.py --hotspots 0 --seed 1 --ligands 262144 --iters 3 --budgets=0 2>&1 | tail -2
done
