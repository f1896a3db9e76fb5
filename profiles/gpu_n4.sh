#!/bin/bash
N=4
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --config C4 --steps 20 --warmup 5 > gpurun_out/bench_C4_r02_n$N.json 2> gpurun_out/bench_C4_r02_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_C4_r02_n$N.json').read().strip().splitlines()[-1])
print('C4', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['checksum']['cell_counts_sha1'], d['config']['migrations_per_step'])
print(d['roofline']['kernels_ms_per_step'])
print(d['roofline']['per_rank_busy_ms_per_step'])
PY
