run() { name=$1; shift; env "$@" > gpurun_out/prio2_$name.json 2> gpurun_out/prio2_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/prio2_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d['value'],1), round(d['ms_per_step'],3), 'nopf', round(d['no_prefetch']['ms_per_step'],3), 'strong', d['strong'] and round(d['strong']['ms_per_step'],3), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],3), d['config']['device_allocs_in_timed_region'], d['config']['step_diagnostics'])
except Exception as e:
    print("$name", 'ERR', e)
PY
}
python -m pytest tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -3
B="python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-configs --no-cpu-baseline"
run hi $B
run nohi GB_PRIO_MAIN=0 $B
run hi2 GB_PRIO_MAIN=-2 GB_PRIO_AUX=-1 $B
run hi3 $B
