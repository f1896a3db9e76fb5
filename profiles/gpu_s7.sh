#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02n.log 2>&1
tail -6 gpurun_out/pytest_gpu_r02n.log
CONFIGS="C4 C2 C3 C5" bash profiles/try_libs.sh 2>&1 | tail -18
