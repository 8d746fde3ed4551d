set -x
mkdir -p gpurun_out
python tools/backbone_prof.py bf16x3 backbone > gpurun_out/r2_bbprof_x3.log 2>&1; head -40 gpurun_out/r2_bbprof_x3.log
python tools/backbone_prof.py bf16 backbone > gpurun_out/r2_bbprof_bf16.log 2>&1; head -34 gpurun_out/r2_bbprof_bf16.log
rm -f gpurun_out/sweep.log; bash tools/sweep.sh --unique 4096 --rep 64 --lpt > /dev/null 2>&1; cat gpurun_out/sweep.log
