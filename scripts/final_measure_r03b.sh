#!/bin/bash
# Bench lines and ncu captures of the round's final binary (the full pass incl. sanitizers is scripts/final_measure_r03.sh)
TAG=${TAG:-r03}
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_20_5.json 2>/dev/null
for c in 2 4 5; do python bench.py --config $c > gpurun_out/${TAG}_bench_c$c.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/${TAG}_launches_steady.csv \
    python bench.py --steps 40 --warmup 10 --min-preroll 200 --max-preroll 200 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pre|k_step|k_settle_slice|k_finish" -s 1400 -c 8 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 40 --warmup 10 --min-preroll 200 --max-preroll 200 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
for f in c3 c3_20_5 c2 c4 c5; do python -c "
import json; d=json.load(open('gpurun_out/${TAG}_bench_$f.json')); print('$f', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])"; done
