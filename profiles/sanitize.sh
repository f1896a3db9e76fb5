#!/bin/bash
# memory checks of the CUDA library on a subset of the parity tests (slow: run under gpurun with a generous timeout)
mkdir -p gpurun_out
K=${1:-"two_bit or genetic_population or confined or partheno or run_queues or run_recovers or env_interpolation"}
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -p no:cacheprovider -k "$K" > gpurun_out/memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck.log | head -20
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 3 --print-limit 30 python -m pytest tests/test_parity_gpu.py -x -q -p no:cacheprovider -k "two_bit or trajectory or run_queues" > gpurun_out/initcheck.log 2>&1
echo "initcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Uninitialized|at .*qhg" gpurun_out/initcheck.log | head -40
