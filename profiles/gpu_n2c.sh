#!/bin/bash
N=2
mkdir -p gpurun_out
for c in C5; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --config $c --steps 20 --warmup 5 > gpurun_out/bench_${c}_r02i_n$N.json 2> gpurun_out/bench_${c}_r02i_n$N.err
  tail -2 gpurun_out/bench_${c}_r02i_n$N.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${c}_r02i_n$N.json').read().strip().splitlines()[-1])
print('$c', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['checksum']['cell_counts_sha1'], d['e2e'].get('mirror_equals_readback'))
print(d['roofline']['per_rank_busy_ms_per_step'])
PY
done
