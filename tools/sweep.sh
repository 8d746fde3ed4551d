#!/bin/bash
# Developer tool (GPU box): time the scoring kernel for every variant library under pharmaconet_b200/_variants.
# usage: tools/sweep.sh [quick_bench args...]   (writes gpurun_out/sweep.log)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for so in pharmaconet_b200/_variants/libpmnet_b200_*.so; do
  tag=$(basename $so .so); tag=${tag#libpmnet_b200_}
  echo "=== $tag" | tee -a gpurun_out/sweep.log
  PMNET_B200_SO=$PWD/$so timeout 300 python tools/quick_bench.py "$@" 2>&1 | grep -E "status|best|Error|error" | tee -a gpurun_out/sweep.log
done
