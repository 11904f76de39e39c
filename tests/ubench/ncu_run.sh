set -x
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --cuda-profiler-range"
$CMD > gpurun_out/ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r01_launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline --no-e2e --cuda-profiler-range"
$CMD2 > gpurun_out/ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'group_bwd_accum|query_kernel|group_fwd_kernel|fps_cluster|interp_bwd|interp_fwd_kernel' -c 40 -o gpurun_out/r01_prof_full $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
ls -la gpurun_out/
