#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu_r02.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider ) > gpurun_out/pytest_gpu_r02_final.log 2>&1
tail -6 gpurun_out/pytest_gpu_r02_final.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02.log 2>&1; tail -1 gpurun_out/smoke_r02.log
bash profiles/collect_r02.sh r02
