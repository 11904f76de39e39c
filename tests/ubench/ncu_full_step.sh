# usage: bash tests/ubench/ncu_full_step.sh <tag> [kernel regex] [count]  -- ncu --set full of the step's main kernels (batch 32,
# one step, one stream).  The report is converted to the raw-page CSV on the box and removed: with more than ~25 launches it
# exceeds gpurun's 64 MiB return limit.  Single-stream launch order: sampling chain and forward groupings first, then the 18
# group backwards (last grouping first), then the interpolation chain.
set -x
TAG=$1
REGEX=${2:-'scatter_private_kernel|seg_dense_kernel|group_fwd_kernel|fps_cluster_kernel|grid_query_kernel|interp_fwd_kernel'}
COUNT=${3:-24}
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-gpu-baseline --no-configs --no-prefetch --no-overlap --cuda-profiler-range"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off -k regex:"$REGEX" \
    -c $COUNT -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"
ncu -i gpurun_out/${TAG}.ncu-rep --page raw --csv > gpurun_out/${TAG}_raw.csv 2> gpurun_out/${TAG}_raw.err
rm -f gpurun_out/${TAG}.ncu-rep
