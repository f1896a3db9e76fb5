#!/bin/bash
# One GPU session of a round (run under gpurun): parity suite first, then the plain bench, then the ncu evidence.
#   bash profiles/run_round.sh r01 [fast]      "fast": tests + bench only
R=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_$R.txt 2>&1
( time timeout 720 python -m pytest tests -m gpu -q --timeout 300 -rf -p no:cacheprovider ) > gpurun_out/pytest_gpu_$R.log 2>&1
tail -15 gpurun_out/pytest_gpu_$R.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$R.log 2>&1; tail -2 gpurun_out/smoke_$R.log
if [ "$2" == "fast" ]; then
  timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$R.json 2> gpurun_out/bench_$R.err; tail -1 gpurun_out/bench_$R.json
  exit 0
fi
timeout 600 bash profiles/collect.sh $R 2>&1 | tail -3
timeout 300 python profiles/bench_configs.py > gpurun_out/bench_configs_$R.json 2> gpurun_out/bench_configs_$R.err; tail -1 gpurun_out/bench_configs_$R.json
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$R.json 2>&1; tail -1 gpurun_out/bench_ref_$R.json
