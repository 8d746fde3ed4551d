#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scoring_gpu.py -x -q 2>&1 | tail -15 > gpurun_out/s12_tests.log
timeout 600 python tools/dense_probe.py --profile --budgets=-1,0,32768,16384 > gpurun_out/s12_dense.log 2>&1
tail -5 gpurun_out/s12_tests.log; cat gpurun_out/s12_dense.log
