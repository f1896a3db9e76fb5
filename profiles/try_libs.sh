#!/bin/bash
# A/B of compile-time knobs: run the bench with every library variant under build/libs
for f in build/libs/*.so; do
  cp "$f" qhg4_b200/libqhg_b200.so
  echo "== $f"
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernels_ms_per_step'])"
done
