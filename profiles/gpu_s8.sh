#!/bin/bash
mkdir -p gpurun_out
( QHG_SEG_SB=16 timeout 900 python -m pytest tests -m gpu -q --timeout 600 -rf -p no:cacheprovider -x ) > gpurun_out/pytest_gpu_r02o.log 2>&1
tail -4 gpurun_out/pytest_gpu_r02o.log
show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {k: v for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 0.05}, d["checksum"]["cell_counts_sha1"])'
for sb in 8 16; do for c in C2 C3 C5; do
  echo "== $c QHG_SEG_SB=$sb"
  QHG_SEG_SB=$sb python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
done; done
