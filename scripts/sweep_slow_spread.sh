for sp in 32 16 8 4 2; do for early in 10 24; do
QS_SLOW_SPREAD=$sp QS_SETTLE_SLICE_EARLY=$early python bench.py --steps 200 --warmup 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('spread $sp early $early', round(d['ms_per_step'],4), round(d['value']/1e6,2), [ (k.get('kernel'), round(k.get('kernel_ms',0),3)) for k in [d['roofline']]+d['roofline'].get('other_kernels',[])])"
done; done
