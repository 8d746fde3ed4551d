#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv3d_gpu.py tests/test_cnn_gpu.py -x -q 2>&1 | tail -8
timeout 600 python tools/cnn_bench.py 2>&1 | tail -25
