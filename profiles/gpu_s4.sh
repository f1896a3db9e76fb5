#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider ) > gpurun_out/pytest_gpu_r02k.log 2>&1
tail -12 gpurun_out/pytest_gpu_r02k.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
