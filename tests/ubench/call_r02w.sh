set -x
python -m pytest tests/test_groupers_gpu.py tests/test_backward_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02w_pytest.txt
python tests/ubench/interp_bwd.py > gpurun_out/r02w_interp_new.json 2> gpurun_out/r02w_interp_new.err
GBOPS_LIB=.variants/libgbops_oldsort.so python tests/ubench/interp_bwd.py > gpurun_out/r02w_interp_old.json 2> gpurun_out/r02w_interp_old.err
for v in fold nofold fold2 nofold2; do
  T=""; case $v in nofold*) T="--tune group_mode=32";; esac
  python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-configs --no-cpu-baseline --no-strong --no-e2e $T > gpurun_out/r02w_bench_$v.json 2> gpurun_out/r02w_bench_$v.err
done
cat gpurun_out/r02w_pytest.txt gpurun_out/r02w_interp_new.json gpurun_out/r02w_interp_old.json
python - <<'PY'
import json
for v in ("fold","nofold","fold2","nofold2"):
    try:
        d=json.loads(open(f"gpurun_out/r02w_bench_{v}.json").read().strip().splitlines()[-1])
        po={p['kernel']:(round(p['ms_per_step'],3), p['launches_per_step']) for p in d['per_op']}
        print(v, round(d['value'],1), round(d['ms_per_step'],3), 'nopf', round(d['no_prefetch']['ms_per_step'],3), 'prof', round(d['roofline_pass_ms_per_step'],3), po.get('gb_group_fwd'), po.get('gb_group_xyz'), d['gpu_launches'])
    except Exception as e:
        print(v, 'ERR', e)
PY
