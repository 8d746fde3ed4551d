set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_cnn_gpu.py -q -m gpu -x > gpurun_out/r2_tests_cnn.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_cnn.log
tail -4 gpurun_out/r2_tests_cnn.log
timeout 600 python tools/cnn_bench.py > gpurun_out/r2_cnn_bench.log 2>&1; tail -8 gpurun_out/r2_cnn_bench.log
ncu --set full --clock-control none -k regex:'lateral_kernel|box_combine|density_post' -c 6 -o gpurun_out/cnn_r02_c -f python tools/cnn_probe.py bf16 > gpurun_out/ncu_cnn_c.log 2>&1; tail -2 gpurun_out/ncu_cnn_c.log
python tools/ncu_kernels.py gpurun_out/cnn_r02_c.ncu-rep gpurun_out/cnn_r02_kernels_c.json "ncu --set full --clock-control none -k regex:lateral_kernel|box_combine|density_post -c 6 python tools/cnn_probe.py bf16" > gpurun_out/cnn_r02_kernels_c.txt 2>&1
rm -f gpurun_out/cnn_r02_c.ncu-rep; cat gpurun_out/cnn_r02_kernels_c.txt | cut -c1-180
