CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-gpu-baseline --no-configs --no-prefetch --no-overlap --cuda-profiler-range"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'grid_query_kernel|group_xyz_kernel|three_nn_grid_kernel|collision_kernel|query_kernel' -c 40 -o gpurun_out/r02u_queries $CMD > gpurun_out/r02u_queries_ncu.log 2>&1
echo rc=$?
