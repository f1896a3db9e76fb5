show='import sys,json; d=json.loads(sys.stdin.read()); print("%.4g" % d["value"], "%.4f ms" % d["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], {k: v for k, v in d["roofline"]["kernels_ms_per_step"].items() if v > 0.05}, d["checksum"]["id_xor"])'
for t in 250 450 600 900 1500 4000; do
  for c in C4 C2; do
    echo "== $c QHG_SEG_AGENTS=$t"
    QHG_SEG_AGENTS=$t python bench.py --config $c --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "$show"
  done
done
