#!/bin/bash
# bench lines of several configurations: bash scripts/ab_cfg.sh "3 2 5"
for c in $1; do
  python bench.py --config $c --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('config $c', round(d['value']/1e6,3), 'M/s', round(d['ms_per_step'],4), 'ms; slices', round(r['kernel_ms'],4) if r['kernel']=='k_settle_slice' else round(r['other_kernels'][0]['kernel_ms'],4), 'slow', round(r.get('k_step_slow_ms',0),3), 'e2e', round(d['e2e']['value']/1e6,2))"
done
