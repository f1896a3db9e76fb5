#!/bin/bash
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02q.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02q.log
show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], {k: v for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 0.05}, d["checksum"]["cell_counts_sha1"])'
for c in C4 C2; do python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"; done
