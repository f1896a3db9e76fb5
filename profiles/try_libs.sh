#!/bin/bash
# run the small profile workload with every library variant under build/libs (A/B of compile-time knobs)
for f in build/libs/*.so; do
  cp "$f" qhg4_b200/libqhg_b200.so
  echo "== $f"
  python -m pytest tests -m gpu -q -x 2>&1 | tail -1
  python profiles/prof_small.py 2>&1 | tail -1 | python -c "import sys,ast; d=ast.literal_eval(sys.stdin.read()); print(d['frac'], d['kernels_ms_per_step'])"
done
