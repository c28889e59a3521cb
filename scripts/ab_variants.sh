#!/bin/bash
# A/B of library variants and knobs on one box: bash scripts/ab_variants.sh "packed:90 scalar:90 packed:100 scalar:100"
# (variant = packed | scalar -> libqs_b200.so | libqs_b200_scalar.so built with -DQS_PACK_LEGS=0; number = QS_FILL)
cp quadruped_springs_b200/csrc/libqs_b200.so /tmp/packed.so
for spec in $1; do
  v=${spec%%:*}; f=${spec##*:}
  if [ $v = scalar ]; then cp quadruped_springs_b200/csrc/libqs_b200_scalar.so quadruped_springs_b200/csrc/libqs_b200.so; else cp /tmp/packed.so quadruped_springs_b200/csrc/libqs_b200.so; fi
  touch quadruped_springs_b200/csrc/libqs_b200.so
  QS_FILL=$f python bench.py --no-cpu-baseline --steps 200 --warmup 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$spec', round(d['value']/1e6,2), 'M/s', round(d['ms_per_step'],4), 'ms; slices', round(d['roofline']['kernel_ms'],4), 'slow', round(d['roofline']['k_step_slow_ms'],3), 'step kernels', round(d['roofline']['other_kernels'][0]['kernel_ms'],4), 'urgent', d['steady_state']['urgent_settles_last_step'], 'e2e', round(d['e2e']['value']/1e6,2))"
done
cp /tmp/packed.so quadruped_springs_b200/csrc/libqs_b200.so
