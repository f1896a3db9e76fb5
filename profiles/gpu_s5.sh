#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -k "genetic or genom or ooa" ) > gpurun_out/pytest_gpu_r02l.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02l.log
CONFIGS="C4 C2" bash profiles/try_libs.sh 2>&1 | tail -12
bash profiles/collect_r02.sh r02
