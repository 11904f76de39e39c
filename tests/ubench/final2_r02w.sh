# last call of the round: grouper / backward / pipeline suites, smoke(), the default bench line and the reference arm
set -x
python -m pytest tests/test_groupers_gpu.py tests/test_backward_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02w_pytest.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02w_smoke.txt 2>&1
python bench.py > gpurun_out/r02w_bench.json 2> gpurun_out/r02w_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02w_ref.json 2> gpurun_out/r02w_ref.err
cat gpurun_out/r02w_pytest.txt; tail -2 gpurun_out/r02w_smoke.txt; tail -c 400 gpurun_out/r02w_bench.err; cat gpurun_out/r02w_ref.json
