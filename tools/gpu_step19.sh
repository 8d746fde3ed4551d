#!/bin/bash
mkdir -p gpurun_out
python bench.py --workload e2e --steps 1 > gpurun_out/r2_e2e_1gpu.json 2> gpurun_out/r2_e2e_1gpu.err; tail -c 900 gpurun_out/r2_e2e_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pmnet|iota_ids|topk_pad|DeviceRadixSort|ligand_cost' -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cnn --no-cpu-baseline --no-dense > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
