set -x
mkdir -p gpurun_out
python -m pytest tests/test_scoring_gpu.py -q -m gpu -x > gpurun_out/r2_tests_scoring.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_scoring.log
tail -4 gpurun_out/r2_tests_scoring.log
python tools/quick_bench.py --unique 4096 --rep 32 > gpurun_out/r2_qb_fast.log 2>&1; tail -2 gpurun_out/r2_qb_fast.log
python -m pytest tests/test_cnn_gpu.py tests/test_conv3d_gpu.py -q -m gpu -s > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
grep -E "differ|identical|passed|failed|Error|error|assert" gpurun_out/r2_tests_cnn.log | tail -30
python bench.py --steps 3 --warmup 3 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 3000 gpurun_out/r2_bench_a.json; tail -5 gpurun_out/r2_bench_a.err
