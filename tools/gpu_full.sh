set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_tests_full.log
tail -6 gpurun_out/r2_tests_full.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -3 gpurun_out/r2_smoke.log
python bench.py > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err; tail -c 4500 gpurun_out/r2_bench_b.json; tail -3 gpurun_out/r2_bench_b.err
