set -x
mkdir -p gpurun_out
python -m pytest tests/test_scoring_gpu.py -q -m gpu -x > gpurun_out/r2_tests_scoring.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_scoring.log
tail -4 gpurun_out/r2_tests_scoring.log
python tools/quick_bench.py --unique 4096 --rep 64 --lpt > gpurun_out/r2_qb_fast.log 2>&1; tail -2 gpurun_out/r2_qb_fast.log
