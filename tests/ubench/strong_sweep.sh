for t in 4 8 16; do for pf in ""; do
python bench.py --steps 10 --warmup 3 --no-gpu-baseline --no-configs --no-e2e --no-cpu-baseline --strong-total $t $pf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['strong']; print('total',s['scenes_total'],'prefetch',s['prefetch'],'ms',round(s['ms_per_step'],3),'scenes/s',round(s['value'],1))"
done; done
