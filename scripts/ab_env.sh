#!/bin/bash
# bench under sets of environment knobs: bash scripts/ab_env.sh "QS_FILL=90,QS_RING=20 QS_FILL=96,QS_RING=16 ..."
for spec in $1; do
  env $(echo $spec | tr ',' ' ') python bench.py --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$spec', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'ms; slices', round(d['roofline']['kernel_ms'],4), 'slow', round(d['roofline']['k_step_slow_ms'],3), 'step kernels', round(d['roofline']['other_kernels'][0]['kernel_ms'],4), 'urgent', d['steady_state']['urgent_settles_last_step'], 'e2e', round(d['e2e']['value']/1e6,2))"
done
