#!/bin/bash
# One pass of everything the round's numbers come from (run under gpurun on one B200).
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r01f_bench_reference.json 2> gpurun_out/r01f_bench_reference.err
tail -c 600 gpurun_out/r01f_bench_reference.json
python bench.py > gpurun_out/r01f_bench_1gpu.json 2> gpurun_out/r01f_bench_1gpu.err
tail -c 300 gpurun_out/r01f_bench_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 650 --csv --log-file gpurun_out/r01f_launches_steady.csv \
    python bench.py --steps 120 --warmup 120 --no-cpu-baseline > gpurun_out/r01f_ncu_list.log 2>&1
tail -2 gpurun_out/r01f_ncu_list.log
ncu --set full --clock-control none --import-source on -k "regex:k_pre|k_step|k_settle_slice" -s 1050 -c 7 -f -o gpurun_out/r01f_full \
    python bench.py --steps 120 --warmup 120 --no-cpu-baseline > gpurun_out/r01f_ncu_full.log 2>&1
tail -2 gpurun_out/r01f_ncu_full.log
ls -la gpurun_out/r01f*
