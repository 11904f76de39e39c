# final measurement of the round-2 build (one gpurun call, one B200): default bench line, ncu launch list / DRAM bytes / full set
# of the step's main kernels (converted to CSV on the box: the report itself exceeds gpurun's 64 MiB return limit), per-shape timings
set -x
( time python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err ) 2> gpurun_out/r02w_bench_time.txt
bash tests/ubench/ncu_dram.sh r02w
bash tests/ubench/ncu_full_step.sh r02w_full_fwd 'group_fwd_kernel|fps_cluster_kernel|grid_query_kernel' 14
bash tests/ubench/ncu_full_step.sh r02w_full_bwd 'scatter_private_kernel|seg_dense_kernel|interp_fwd_kernel' 24
python tests/ubench/bwd_defaults.py > gpurun_out/r02w_bwd_defaults.json 2>/dev/null
python tests/ubench/fwd_shapes.py > gpurun_out/r02w_fwd_shapes.json 2>/dev/null
python tests/ubench/interp_bwd.py > gpurun_out/r02w_interp_bwd.json 2>/dev/null
cat gpurun_out/r02w_bench_time.txt
du -sh gpurun_out; ls -la gpurun_out | tail -20
