#!/bin/bash
# A/B of compile-time knobs: run the bench with every library variant under build/libs (C4 and C2), kernel times per step
cp qhg4_b200/libqhg_b200.so /tmp/lib_keep.so
show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {k: v for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 0.05})'
for f in build/libs/*.so; do
  cp "$f" qhg4_b200/libqhg_b200.so
  for c in ${CONFIGS:-C4 C2}; do
    echo "== $c $f"
    python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  done
done
cp /tmp/lib_keep.so qhg4_b200/libqhg_b200.so
