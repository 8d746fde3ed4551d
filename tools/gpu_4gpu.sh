set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2_bench_4gpu.json 2> gpurun_out/r2_bench_4gpu.err; tail -c 400 gpurun_out/r2_bench_4gpu.json; tail -3 gpurun_out/r2_bench_4gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err; tail -c 300 gpurun_out/r2_bench_2gpu.json
