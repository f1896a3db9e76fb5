#!/bin/bash
# sharded parity + scaling on N GPUs of one box (run under `gpurun --gpus N`): bash profiles/mgpu_round.sh N tag
N=${1:-2}; R=${2:-r02}
mkdir -p gpurun_out
OUT=gpurun_out/mgpu_check_${R}_n$N.txt
: > $OUT
for what in ${WHAT:-tutorial genetic rebalance rebalance-genetic bigcell}; do
  [ "$what" == "tutorial" ] && what=""
  for p2p in 1 0; do
    echo "== $N GPUs, mgpu_check.py $what, QHG_P2P=$p2p" >> $OUT
    QHG_P2P=$p2p timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py $what 2>&1 | grep -v "^W[0-9]\|OMP_NUM_THREADS\|^\*\*\*\*" | tail -6 >> $OUT
  done
done
cat $OUT
for c in ${CONFIGS-C4 C5}; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus $N --config $c --steps 10 --warmup 3 > gpurun_out/bench_${c}_${R}_n$N.json 2> gpurun_out/bench_${c}_${R}_n$N.err
  tail -c 600 gpurun_out/bench_${c}_${R}_n$N.json
done
