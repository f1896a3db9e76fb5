#!/bin/bash
# 8 GPUs: sharded parity (every mode over peer memory; the tutorial population and the big cells over NCCL as well), then the
# bench lines of C4 and C5 at 1e8 agents
N=${1:-8}; R=${2:-r02f}
mkdir -p gpurun_out
OUT=gpurun_out/mgpu_check_${R}_n$N.txt
: > $OUT
run() { echo "== $N GPUs, mgpu_check.py $1, QHG_P2P=$2" >> $OUT
  QHG_P2P=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py $1 2>&1 | grep -v "^W[0-9]\|OMP_NUM_THREADS\|^\*\*\*\*" | tail -4 >> $OUT; }
run "" 1; run genetic 1; run rebalance 1; run rebalance-genetic 1; run bigcell 1; run "" 0; run bigcell 0
cat $OUT
for c in C4 C5; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --config $c --steps 20 --warmup 5 > gpurun_out/bench_${c}_${R}_n$N.json 2> gpurun_out/bench_${c}_${R}_n$N.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${c}_${R}_n$N.json').read().strip().splitlines()[-1])
print('$c', '%.4g' % d['value'], '%.4f ms' % d['ms_per_step'], 'e2e %.4g' % d['e2e']['value'], d['checksum']['cell_counts_sha1'], d['e2e'].get('mirror_equals_readback'))
print(d['roofline']['kernels_ms_per_step'])
print(d['roofline']['per_rank_busy_ms_per_step'])
PY
done
