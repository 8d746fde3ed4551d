set -x
mkdir -p gpurun_out
timeout 600 python bench.py --workload e2e --pockets-per-gpu 4 --e2e-ligands 32768 --steps 1 > gpurun_out/r2_e2e_1gpu_small.json 2> gpurun_out/r2_e2e_1gpu_small.err; tail -c 1500 gpurun_out/r2_e2e_1gpu_small.json; tail -5 gpurun_out/r2_e2e_1gpu_small.err
timeout 900 python bench.py --workload e2e --steps 1 > gpurun_out/r2_e2e_1gpu.json 2> gpurun_out/r2_e2e_1gpu.err; tail -c 1500 gpurun_out/r2_e2e_1gpu.json; tail -5 gpurun_out/r2_e2e_1gpu.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'pmnet|iota_ids|topk_pad|DeviceRadixSort|ligand_cost' --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cnn --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-160
