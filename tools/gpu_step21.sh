#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_scoring_gpu.py -q -x -k "task_parallel_walk_is_identical and syn0_c8" > gpurun_out/s13_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s13_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_scoring_gpu.py -q -x -k "task_parallel_walk_is_identical and True-syn0_c8" > gpurun_out/s13_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s13_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_conv3d_gpu.py -q -x -k "lateral" > gpurun_out/s13_memcheck_lat.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s13_memcheck_lat.log
tail -4 gpurun_out/s13_memcheck.log; tail -4 gpurun_out/s13_racecheck.log; tail -4 gpurun_out/s13_memcheck_lat.log
