#!/bin/bash
# memcheck + racecheck (shared memory hazards inside a warp's slice) of the kernels rewritten in round 2, on small parity tests
mkdir -p gpurun_out
K="trajectory_bit_exact_vs_oracle or genetic_populations_on_the_fast_path or navigate_on_the_fast_path or cond_weighted or move_stats or count_mirror"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -p no:cacheprovider -k "$K" > gpurun_out/memcheck_r02.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/memcheck_r02.log | head -12
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_parity_gpu.py -x -q -p no:cacheprovider -k "trajectory_bit_exact_vs_oracle and tiled and not cell" > gpurun_out/racecheck_r02.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard|Race" gpurun_out/racecheck_r02.log | head -12
