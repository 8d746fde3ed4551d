set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cnn_gpu.py tests/test_gemm_gpu.py -q -m gpu -x > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
tail -5 gpurun_out/r2_tests_cnn.log
python tools/backbone_prof.py bf16x3 backbone > gpurun_out/r2_bbprof_x3.log 2>&1; head -16 gpurun_out/r2_bbprof_x3.log | cut -c1-70,130-200
timeout 600 python tools/cnn_bench.py > gpurun_out/r2_cnn_bench.log 2>&1; tail -12 gpurun_out/r2_cnn_bench.log
rm -f gpurun_out/sweep.log; bash tools/sweep.sh --unique 4096 --rep 64 --lpt > /dev/null 2>&1; cat gpurun_out/sweep.log
