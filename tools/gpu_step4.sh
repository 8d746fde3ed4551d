set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_gemm_gpu.py tests/test_conv3d_gpu.py -q -m gpu -s -x > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
grep -E "differ|identical|passed|failed|Error|error|assert|rc=" gpurun_out/r2_tests_cnn.log | tail -30
tail -30 gpurun_out/r2_tests_cnn.log
timeout 600 python tools/cnn_bench.py --reference > gpurun_out/r2_cnn_bench.log 2>&1; tail -40 gpurun_out/r2_cnn_bench.log
rm -f gpurun_out/sweep.log; bash tools/sweep.sh --unique 4096 --rep 64 --lpt > /dev/null 2>&1; cat gpurun_out/sweep.log
