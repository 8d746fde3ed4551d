set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:pmnet_score_fast -c 1 -o gpurun_out/score_r02_final -f python tools/quick_bench.py --unique 4096 --rep 32 --lpt --tiny > gpurun_out/ncu_score_final.log 2>&1; tail -2 gpurun_out/ncu_score_final.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_bench.csv python bench.py --steps 2 --warmup 1 --no-cnn --no-cpu-baseline --no-e2e > gpurun_out/launches_bench.log 2>&1; tail -1 gpurun_out/launches_bench.log | cut -c1-200
ncu --set full --clock-control none -k regex:'conv3d_k3|gemm_kernel|window_attention|ln_residual' -c 14 -o gpurun_out/cnn_r02_a -f python tools/cnn_probe.py bf16 > gpurun_out/ncu_cnn_a.log 2>&1; tail -2 gpurun_out/ncu_cnn_a.log
python tools/ncu_kernels.py gpurun_out/cnn_r02_a.ncu-rep gpurun_out/cnn_r02_kernels_a.json "ncu --set full --clock-control none -k regex:conv3d_k3|gemm_kernel|window_attention|ln_residual -c 14 python tools/cnn_probe.py bf16" > gpurun_out/cnn_r02_kernels_a.txt 2>&1
rm -f gpurun_out/cnn_r02_a.ncu-rep
ncu --set full --clock-control none -k regex:'lateral_kernel|box_combine|density_post|window_attention_kernel' -c 8 -o gpurun_out/cnn_r02_b -f python tools/cnn_probe.py bf16x3 > gpurun_out/ncu_cnn_b.log 2>&1; tail -2 gpurun_out/ncu_cnn_b.log
python tools/ncu_kernels.py gpurun_out/cnn_r02_b.ncu-rep gpurun_out/cnn_r02_kernels_b.json "ncu --set full --clock-control none -k regex:lateral_kernel|box_combine|density_post|window_attention_kernel -c 8 python tools/cnn_probe.py bf16x3" > gpurun_out/cnn_r02_kernels_b.txt 2>&1
rm -f gpurun_out/cnn_r02_b.ncu-rep
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_cnn.csv python tools/cnn_probe.py bf16 > gpurun_out/launches_cnn.log 2>&1; tail -1 gpurun_out/launches_cnn.log
du -sh gpurun_out; ls -la gpurun_out
cat gpurun_out/cnn_r02_kernels_a.txt gpurun_out/cnn_r02_kernels_b.txt | cut -c1-200
