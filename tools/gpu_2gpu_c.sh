set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload e2e --steps 1 --pockets-per-gpu 16 --e2e-ligands 1250000 > gpurun_out/r2_e2e_2gpu_full.json 2> gpurun_out/r2_e2e_2gpu_full.err; tail -c 1200 gpurun_out/r2_e2e_2gpu_full.json; tail -3 gpurun_out/r2_e2e_2gpu_full.err
