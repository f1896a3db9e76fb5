#!/bin/bash
# ncu evidence for one round: run under gpurun, outputs to gpurun_out/ (copy the summaries into profiles/).
#   bash profiles/collect.sh r01
R=${1:-r01}
mkdir -p gpurun_out
# every launch with its device time (cold cache, serialised: compare SHARES, not absolutes); 2 timed steps after 1 warm-up
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$R.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_$R.log 2>&1
# the two hot kernels, full set, full-size workload (third step)
ncu --set full --clock-control none --import-source on -k 'regex:k_cell_(decide|scatter)' -s 4 -c 2 -o gpurun_out/prof_$R \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu_full_$R.log 2>&1
# clocks while a plain run is going on
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_$R.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err
kill $SMI
tail -1 gpurun_out/bench_$R.json
