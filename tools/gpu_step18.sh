#!/bin/bash
# ncu of the general kernel's passes on the dense model (main pass + task rounds), summarised on the box
mkdir -p gpurun_out
ncu --set full --clock-control none -k regex:pmnet_score_kernel -c 6 -o gpurun_out/score_r02_dense -f python tools/dense_probe.py --ligands 16384 --budgets=0 > gpurun_out/ncu_score_dense.log 2>&1; tail -2 gpurun_out/ncu_score_dense.log
python tools/ncu_kernels.py gpurun_out/score_r02_dense.ncu-rep gpurun_out/scoring_r02_dense_ncu.json "ncu --set full --clock-control none -k regex:pmnet_score_kernel -c 6 python tools/dense_probe.py --ligands 16384 --budgets=0" > gpurun_out/scoring_r02_dense_ncu.txt 2>&1
rm -f gpurun_out/score_r02_dense.ncu-rep
cat gpurun_out/scoring_r02_dense_ncu.txt
