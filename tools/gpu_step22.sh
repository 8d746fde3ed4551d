#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:lateral_mma -c 2 -o gpurun_out/lat_mma -f python tools/cnn_trace.py bf16 > gpurun_out/ncu_lat.log 2>&1; tail -1 gpurun_out/ncu_lat.log
ncu -i gpurun_out/lat_mma.ncu-rep --page raw --csv > gpurun_out/lat_mma_raw.csv 2>/dev/null
ncu -i gpurun_out/lat_mma.ncu-rep --page source --print-source cuda,sass --csv > gpurun_out/lat_mma_src.csv 2>/dev/null
rm -f gpurun_out/lat_mma.ncu-rep
ls -la gpurun_out | grep lat_
