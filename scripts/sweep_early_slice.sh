for early in 0 6 10 14 18; do
QS_SETTLE_SLICE_EARLY=$early python bench.py --steps 200 --warmup 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('early $early', round(d['ms_per_step'],4), round(d['value']/1e6,2))"
done
