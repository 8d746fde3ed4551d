#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scoring_gpu.py -x -q 2>&1 | tail -5 > gpurun_out/s14_tests.log; tail -3 gpurun_out/s14_tests.log
timeout 300 python tools/quick_bench.py --lpt --rep 64 2>&1 | tail -2
timeout 600 python tools/dense_probe.py --budgets=-1,0 2>&1 | tail -3
ncu --set full --clock-control none --import-source on -k regex:pmnet_score_fast -c 1 -o gpurun_out/score_r02_final -f python tools/quick_bench.py --unique 4096 --rep 32 --lpt --tiny > gpurun_out/ncu_score_final.log 2>&1; tail -1 gpurun_out/ncu_score_final.log
