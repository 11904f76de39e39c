# (no --import-source and <= 36 launches: the report must stay well under gpurun's 64 MiB return limit)
# usage: bash tests/ubench/ncu_full_step.sh <tag>  -- ncu --set full of the step's main kernels (batch 32, one step)
set -x
TAG=$1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-gpu-baseline --no-configs --no-prefetch --no-overlap --cuda-profiler-range"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --set full --clock-control none --profile-from-start off \
    -k regex:'scatter_private_kernel|seg_dense_kernel|group_fwd_kernel|fps_cluster_kernel|grid_query_kernel|interp_fwd_kernel' \
    -c 30 -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu.log 2>&1
echo "rc=$?"
