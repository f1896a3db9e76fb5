#!/bin/bash
# memcheck over the whole GPU parity suite but the 1e7-agent and the 32-seed tests
mkdir -p gpurun_out
K="not full_scale and not statistical and not genotype_distributions and not two_gpu"
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 10 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "$K" > gpurun_out/memcheck_r02b.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|at .*qhg" gpurun_out/memcheck_r02b.log | head -12
