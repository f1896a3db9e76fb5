#!/bin/bash
# A/B of compile-time knobs / older commits: run the bench with every library variant under build/libs (C4 twice, C2 once)
cp qhg4_b200/libqhg_b200.so /tmp/lib_keep.so
show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {k: v for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 0.05})'
for rep in 1; do
for f in build/libs/*.so; do
  cp "$f" qhg4_b200/libqhg_b200.so
  echo "== C4 $f"
  QHG_AB_OLD_LIB=1 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  if [ $rep == 1 ]; then
    echo "== C2 $f"
    QHG_AB_OLD_LIB=1 python bench.py --agents 10000000 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  fi
done
done
cp /tmp/lib_keep.so qhg4_b200/libqhg_b200.so
