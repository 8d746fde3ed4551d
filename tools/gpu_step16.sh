#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_scoring_gpu.py -x -q 2>&1 | tail -8 > gpurun_out/s16_tests.log; tail -4 gpurun_out/s16_tests.log
timeout 300 python tools/dense_probe.py --profile --budgets=-1,0,32768 > gpurun_out/s16_dense.log 2>&1; cat gpurun_out/s16_dense.log | grep -v "Warn\|Memset\|vectorized\|_warn"
timeout 200 python tools/quick_bench.py --lpt --rep 64 2>&1 | tail -1
