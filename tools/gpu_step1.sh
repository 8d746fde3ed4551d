set -x
mkdir -p gpurun_out
python tools/debug_fast.py examples_loose syn0_c32 syn0_c8 loose_c8 > gpurun_out/debug_fast.log 2>&1; grep "==" gpurun_out/debug_fast.log
python -m pytest tests/test_scoring_gpu.py tests/test_abi.py -q -m gpu > gpurun_out/r2_tests1.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests1.log
tail -12 gpurun_out/r2_tests1.log
python tools/quick_bench.py --unique 4096 --rep 32 > gpurun_out/r2_qb_fast.log 2>&1; tail -3 gpurun_out/r2_qb_fast.log
ncu --set full --clock-control none --import-source on -k regex:pmnet_score_fast -c 1 -o gpurun_out/score_r02b -f python tools/quick_bench.py --unique 4096 --rep 32 --tiny > gpurun_out/ncu_r02b.log 2>&1; tail -2 gpurun_out/ncu_r02b.log
