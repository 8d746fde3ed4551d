#!/bin/bash
for so in pharmaconet_b200/_variants/libpmnet_b200_tf*.so; do
  echo "=== $so"
  PMNET_B200_SO=$PWD/$so timeout 600 python tools/dense_probe.py --budgets=0,131072 --iters 2 2>&1 | grep budget
done
