#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02p.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02p.log
show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d["roofline"]["kernels_ms_per_step"], d["checksum"]["cell_counts_sha1"])'
for sc in two one; do for c in C2 C4; do
  echo "== $c QHG_SCAN=$sc"
  QHG_SCAN=$sc python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done; done
