set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/r2_tests_gemm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_gemm.log
tail -25 gpurun_out/r2_tests_gemm.log
timeout 900 python -m pytest tests/test_cnn_gpu.py -q -m gpu -s > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
grep -E "differ|identical|passed|failed|Error|error|assert" gpurun_out/r2_tests_cnn.log | tail -30
rm -f gpurun_out/sweep.log; bash tools/sweep.sh --unique 4096 --rep 32 > /dev/null 2>&1; cat gpurun_out/sweep.log
