#!/bin/bash
# One pass of everything round 2, second session's numbers come from (run under gpurun on one B200): TAG=r02x bash scripts/final_measure_r02.sh
TAG=${TAG:-r03}
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
python bench.py > gpurun_out/${TAG}_bench_c3.json 2> gpurun_out/${TAG}_bench_c3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_c3_20_5.json 2> gpurun_out/${TAG}_bench_c3_20_5.err
for c in 2 4 5; do python bench.py --config $c > gpurun_out/${TAG}_bench_c$c.json 2> gpurun_out/${TAG}_bench_c$c.err; done
# launch list of steady-state steps (graph nodes are profiled one by one; cold-cache and serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv --log-file gpurun_out/${TAG}_launches_steady.csv \
    python bench.py --steps 40 --warmup 10 --min-preroll 200 --max-preroll 200 --no-cpu-baseline > gpurun_out/${TAG}_ncu_list.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_list.log
ncu --set full --clock-control none --import-source on -k "regex:k_pre|k_step|k_settle_slice|k_finish" -s 1400 -c 8 -f -o gpurun_out/${TAG}_full \
    python bench.py --steps 40 --warmup 10 --min-preroll 200 --max-preroll 200 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log
ls -la gpurun_out/${TAG}*
# compute-sanitizer on the new tick (shared-memory parking, packed loads): memcheck at full settle length, racecheck on a
# shortened workload with the urgent path forced (racecheck slows the tick kernels ~1000x)
(time compute-sanitizer --tool memcheck python scripts/sanitize_run.py 40) > gpurun_out/${TAG}_memcheck.log 2>&1; tail -6 gpurun_out/${TAG}_memcheck.log
(time env QS_SAN_ENVS=256 QS_SAN_SETTLE=30 QS_SETTLE_SLICE_MAX=8 compute-sanitizer --tool racecheck python scripts/sanitize_run.py 30) > gpurun_out/${TAG}_racecheck.log 2>&1; tail -6 gpurun_out/${TAG}_racecheck.log
tools/ffma2_latency > gpurun_out/${TAG}_ffma2_latency.jsonl 2>&1
