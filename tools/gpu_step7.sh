set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_cnn_gpu.py -q -m gpu -x > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
tail -5 gpurun_out/r2_tests_cnn.log
python tools/backbone_prof.py bf16x3 backbone > gpurun_out/r2_bbprof_x3.log 2>&1; head -16 gpurun_out/r2_bbprof_x3.log | cut -c1-70,130-200
python tools/backbone_prof.py bf16 backbone > gpurun_out/r2_bbprof_bf16.log 2>&1; head -14 gpurun_out/r2_bbprof_bf16.log | cut -c1-70,130-200
timeout 600 python tools/cnn_bench.py > gpurun_out/r2_cnn_bench.log 2>&1; tail -8 gpurun_out/r2_cnn_bench.log
