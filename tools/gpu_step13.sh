#!/bin/bash
# round 2 final evidence: sanitizer on the task rounds, ncu of the shipped scoring kernels, launch lists
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_scoring_gpu.py -q -x -k "task_parallel_walk_is_identical and syn0_c8" > gpurun_out/s13_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/s13_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_scoring_gpu.py -q -x -k "task_parallel_walk_is_identical and True-syn0_c8" > gpurun_out/s13_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/s13_racecheck.log
tail -4 gpurun_out/s13_memcheck.log; tail -4 gpurun_out/s13_racecheck.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pmnet|ligand_cost' -c 60 --csv --log-file gpurun_out/launches_r02_quick.csv python tools/quick_bench.py --lpt --rep 64 --iters 2 > gpurun_out/s13_quick_ncu.log 2>&1
grep -c pmnet gpurun_out/launches_r02_quick.csv
ncu --set full --clock-control none --import-source on -k regex:pmnet_score_fast -c 1 -o gpurun_out/score_r02_final -f python tools/quick_bench.py --unique 4096 --rep 32 --lpt --tiny > gpurun_out/ncu_score_final.log 2>&1; tail -2 gpurun_out/ncu_score_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pmnet|iota_ids|topk_pad|DeviceRadixSort|ligand_cost' -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cnn --no-cpu-baseline --no-dense > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
ncu --set full --clock-control none -k regex:pmnet_score_kernel -c 3 -o gpurun_out/score_r02_dense -f python tools/dense_probe.py --ligands 8192 --budgets=0 > gpurun_out/ncu_score_dense.log 2>&1; tail -2 gpurun_out/ncu_score_dense.log
du -sh gpurun_out; ls -la gpurun_out | head -40
