#!/bin/bash
# A/B of the two-chain step (QS_DUAL_CHAIN) on one box: bench lines, then the invariance tests under the new schedule
for dc in 0 1 0 1; do
QS_DUAL_CHAIN=$dc python bench.py --steps 200 --warmup 60 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('dual $dc', round(d['ms_per_step'],4), round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), [ (k.get('kernel'), round(k.get('kernel_ms',0),3)) for k in [d['roofline']]+d['roofline'].get('other_kernels',[])])"
done
QS_DUAL_CHAIN=1 python -m pytest tests -m gpu -x -q -k "conveyor or determinism or everything or urgent or rollout_free or landing or full_size or vecenv" 2>&1 | tail -5
