run() { name=$1; shift; env "$@" > gpurun_out/prio_$name.json 2> gpurun_out/prio_$name.err; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/prio_$name.json").read().strip().splitlines()[-1])
    print("$name", round(d['value'],1), round(d['ms_per_step'],3), 'nopf', round(d['no_prefetch']['ms_per_step'],3), 'strong', d['strong'] and round(d['strong']['ms_per_step'],3))
except Exception as e:
    print("$name", 'ERR', e)
PY
}
B="python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-configs --no-cpu-baseline --no-e2e"
run base $B
run mainhi $B --main-priority -1
run auxhi GB_PRIO_AUX=-1 $B
run fpshi GB_PRIO_FPS=-1 $B
run mainhi_fpshi GB_PRIO_FPS=-2 $B --main-priority -1
run base2 $B
