# usage: bash tests/ubench/ncu_dram.sh <tag>  -- DRAM bytes + duration of every launch of one bench step
set -x
TAG=$1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-gpu-baseline --no-configs --no-prefetch --cuda-profiler-range"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_dram.csv $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"
